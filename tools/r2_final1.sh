O=gpurun_out/r2final; mkdir -p $O
( time timeout 900 python -m pytest tests -q -x -m gpu ) > $O/gpu_suite.log 2>&1; tail -4 $O/gpu_suite.log
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/bench_1gpu.log 2> $O/bench_1gpu.err; tail -3 $O/bench_1gpu.err; tail -1 $O/bench_1gpu.log | cut -c1-300
timeout 900 python tools/capture_traffic.py $O/traffic_new.json > $O/traffic.log 2>&1; tail -5 $O/traffic.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage_tbcsr -s 40 -c 4 -o $O/tiled_final python bench.py --only lns3d --no-cpu --no-parity --steps 1 > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity > $O/launches_bench.log 2>&1; tail -1 $O/launches_bench.log | cut -c1-200
