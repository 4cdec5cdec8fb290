P=${1:-2}; O=gpurun_out/r2multi${P}c; mkdir -p $O
( time timeout 1200 python -m pytest tests -q -x -m gpu -k "two_gpus or ngpus2 or 2gpu or partition" ) > $O/gpu_suite_multi.log 2>&1; tail -4 $O/gpu_suite_multi.log
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $P --steps 20 --warmup 5 --only gl,parity ) > $O/bench_${P}gpu.log 2>&1; tail -2 $O/bench_${P}gpu.log | cut -c1-400
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus $P --steps 20 --warmup 5 --only gl --no-e2e --option use_overlap1=0 ) > $O/bench_${P}gpu_nooverlap.log 2>&1; tail -2 $O/bench_${P}gpu_nooverlap.log | cut -c1-400
