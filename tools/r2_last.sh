O=gpurun_out/r2last; mkdir -p $O
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 --only gl,parity --no-e2e ) > $O/bench_2gpu.log 2>&1; tail -1 $O/bench_2gpu.log | cut -c1-250
