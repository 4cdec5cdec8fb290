O=gpurun_out/r2final2; mkdir -p $O
timeout 900 python tools/capture_traffic.py $O/traffic_new.json > $O/traffic.log 2>&1; tail -4 $O/traffic.log | cut -c1-200
cp $O/traffic_new.json profiles/traffic.json
( time timeout 300 python -c "import __graft_entry__ as e; e.build(); e.smoke()" ) > $O/smoke.log 2>&1; tail -8 $O/smoke.log
( time timeout 900 python -m pytest tests -q -x -m gpu ) > $O/gpu_suite.log 2>&1; tail -4 $O/gpu_suite.log
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/bench_1gpu.log 2> $O/bench_1gpu.err; tail -3 $O/bench_1gpu.err; tail -1 $O/bench_1gpu.log | cut -c1-300
( time timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/bench_reference.log 2>&1; tail -1 $O/bench_reference.log | cut -c1-300
