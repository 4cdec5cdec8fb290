mkdir -p gpurun_out/r2k; O=gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stencil_gen.py tests/test_ref_golden.py -q -x -m gpu -k "lns or stencil or device or tiled or svd or full_solve" > $O/lns_tests.log 2>&1; tail -5 $O/lns_tests.log
run() { echo "== $*"; timeout 300 python bench.py --only ${SEC:-lns3d} --no-cpu --steps 3 $1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['${SEC:-lns3d}']; print(d.get('error') or (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac_model'], d['roofline']['frac_layout']))"; }
(
run ""
run "--option lns:tiled_smr=0"
run "--option lns:tiled_xslots=3"
run "--option lns:tiled_stages=4"
run "--option lns:tiled_flags=11"
run "--option lns:tiled_flags=19"
run "--option lns:tiled_flags=27"
run "--option lns:tiled_flags=3"
SEC=lns2d run ""
SEC=lns3d_strong run ""
) > $O/lns_timing.log 2>&1
cat $O/lns_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage_ -s 8 -c 4 -o $O/tiled_v7 python bench.py --only lns3d --no-cpu --steps 1 > $O/ncu1.log 2>&1
tail -3 $O/ncu1.log
