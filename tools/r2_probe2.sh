mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -x > $O/gputests_2gpu.log 2>&1; tail -8 $O/gputests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.log 2>&1; tail -1 $O/bench_2gpu.log | cut -c1-6000
