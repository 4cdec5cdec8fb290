mkdir -p gpurun_out/r2b; O=gpurun_out/r2b
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_stencil_gen.py -q -k "lns or stencil or device" > $O/lns_tests.log 2>&1; tail -15 $O/lns_tests.log
for opt in "use_tiled=0" "tiled_ctas=2" "tiled_ctas=3" "tiled_ctas=1"; do
  echo "== $opt"; timeout 300 python tools/bench_lns_dist.py --shape 10,200,200 --k 8 --Nw 32 --steps 8 --passes 2 --no-adjoint --option $opt 2>&1 | tail -1
done > $O/lns_timing.log 2>&1
cat $O/lns_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage_ -s 8 -c 4 -o $O/tiled_v1 python tools/bench_lns_dist.py --shape 10,200,200 --k 8 --Nw 32 --steps 4 --passes 1 --no-adjoint > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage_ -s 8 -c 4 -o $O/bcsr_r1final python tools/bench_lns_dist.py --shape 10,200,200 --k 8 --Nw 32 --steps 4 --passes 1 --no-adjoint --option use_tiled=0 > $O/ncu2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -3 $O/gputests.log
