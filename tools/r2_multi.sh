P=${1:-2}; O=gpurun_out/r2multi$P; mkdir -p $O
if [ "$P" = "2" ]; then ( time timeout 1200 python -m pytest tests -q -x -m gpu -k "two_gpus or ngpus2 or 2gpu or partition" ) > $O/gpu_suite_multi.log 2>&1; tail -4 $O/gpu_suite_multi.log; fi
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $P --steps 20 --warmup 5 ) > $O/bench_${P}gpu.log 2>&1; tail -4 $O/bench_${P}gpu.log | cut -c1-600
( time timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > $O/bench_reference.log 2>&1; tail -4 $O/bench_reference.log | cut -c1-400
