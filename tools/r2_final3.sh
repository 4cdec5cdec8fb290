O=gpurun_out/r2final3; mkdir -p $O
( time timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "transient_removal or qr_" ) > $O/tr_tests.log 2>&1; tail -6 $O/tr_tests.log
( time timeout 900 python -m pytest tests -q -x -m gpu ) > $O/gpu_suite.log 2>&1; tail -4 $O/gpu_suite.log
( time timeout 200 python -c "import __graft_entry__ as e; e.build(); e.smoke()" ) > $O/smoke.log 2>&1; tail -4 $O/smoke.log
