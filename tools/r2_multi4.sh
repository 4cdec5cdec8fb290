P=4; O=gpurun_out/r2multi4c; mkdir -p $O
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $P --steps 20 --warmup 5 ) > $O/bench_${P}gpu.log 2>&1; tail -2 $O/bench_${P}gpu.log | cut -c1-300
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus $P --steps 20 --warmup 5 --only gl --no-e2e --option use_overlap1=0 ) > $O/bench_${P}gpu_nooverlap.log 2>&1; tail -2 $O/bench_${P}gpu_nooverlap.log | cut -c1-300
