mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stencil_gen.py -q -x -k "lns or stencil or device" > $O/lns_tests.log 2>&1; tail -5 $O/lns_tests.log
run() { echo "== $*"; timeout 300 env $1 python bench.py --only lns3d --no-cpu --steps 3 $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['lns3d']; print(d.get('error') or (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac_model'], d['roofline']['frac_layout']))"; }
(
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2"
run RSVDT_TILED_ORDER=0 "--option lns:tiled_cv=2"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2 --option lns:tiled_flags=0"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2 --option lns:tiled_flags=35"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=1"
run RSVDT_TILED_ORDER=4 "--option lns:tiled_cv=2"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2 --option lns:tiled_flags=11"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2 --option lns:tiled_flags=19"
run RSVDT_TILED_ORDER=8 "--option lns:tiled_cv=2 --option lns:tiled_flags=27"
) > $O/lns_timing.log 2>&1
cat $O/lns_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage_ -s 8 -c 4 -o $O/tiled_v4 python bench.py --only lns3d --no-cpu --steps 1 --option lns:tiled_cv=2 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:rk4_stage_ -s 8 -c 2 -o $O/tiled_v4_nocompute python bench.py --only lns3d --no-cpu --steps 1 --option lns:tiled_cv=2 --option lns:tiled_flags=11 > $O/ncu2.log 2>&1
