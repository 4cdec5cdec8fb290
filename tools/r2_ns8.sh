# 8 x B200: (1) the north-star solve at a size the reference executable finished (prepared on CPU), compared with it;
# (2) the north-star solve itself (2e7 DOF, k = 8, Nw = 32, TR, q = 1); (3) configs[4] (weights); (4) the bench line at N = 8
P=${1:-8}
O=gpurun_out/r2ns$P; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python tools/north_star_parity.py --gpus $P --prepared --out oracle/_ref/ns_parity_case --json $O/parity_${P}gpu.json > $O/parity_${P}gpu.log 2>&1; echo "parity rc=$?"; tail -25 $O/parity_${P}gpu.log
timeout 420 bash tools/north_star.sh /tmp/ns $P > $O/north_star_${P}gpu.log 2>&1; echo "north star rc=$?"; tail -45 $O/north_star_${P}gpu.log; cp /tmp/ns/timings.json $O/north_star_${P}gpu_timings.json
timeout 420 bash tools/north_star.sh /tmp/nsw $P 100 200 200 --weights > $O/north_star_weights_${P}gpu.log 2>&1; echo "weights rc=$?"; tail -30 $O/north_star_weights_${P}gpu.log; cp /tmp/nsw/timings.json $O/north_star_weights_${P}gpu_timings.json
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $P --steps 3 --warmup 3 > $O/bench_${P}gpu.log 2>&1; echo "bench rc=$?"; tail -1 $O/bench_${P}gpu.log | cut -c1-1500
