mkdir -p gpurun_out/r2h; O=gpurun_out/r2h
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "qr or svd or full_solve or tutorial" > $O/qr_tests.log 2>&1; tail -5 $O/qr_tests.log
timeout 600 python tools/time_dense_phases.py > $O/dense_phases.json 2> $O/dense_phases.err; head -40 $O/dense_phases.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64_op_dmma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gram|apply|chol|cross_svd|small_matmul" --csv --log-file $O/dense_launches.csv python tools/time_dense_phases.py --reps 1 > $O/ncu_dense.log 2>&1
run() { echo "== $*"; timeout 300 python bench.py --only lns3d --no-cpu --steps 3 $1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())['lns3d']; print(d.get('error') or (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac_model'], d['roofline']['frac_layout']))"; }
(
run ""
run "--option lns:tiled_ctas=3"
run "--option lns:tiled_ctas=3 --option lns:tiled_stages=2"
run "--option lns:tiled_stages=2"
) > $O/lns_timing.log 2>&1
cat $O/lns_timing.log
