# partitioned PETSc-binary I/O at a size where it shows: 2e6-DOF 3-D LNS operator as a 6.3 GB MATMPIAIJ file, modes written (2 x 2 GB)
O=gpurun_out/r2io; mkdir -p $O
W=/tmp/io_case
( time python tools/make_case.py lns3d --shape 40 100 100 --k 8 --Nw 8 --w 2.0 --dt 0.003 --TL 0.5 --q 0 --opt 2 --out $W ) > $O/make_case.log 2>&1; tail -4 $O/make_case.log
ls -la $W > $O/case_files.txt
cd $W
( time RSVDT_MPI_TIMINGS=$W/timings.json python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 /root/repo/rsvd-delta-t_b200/rsvdt_mpi.py -inputs variables.yaml ) > /root/repo/$O/run.log 2>&1
cd /root/repo
tail -30 $O/run.log | cut -c1-200
cp $W/timings.json $O/timings.json 2>/dev/null
ls -la $W/results/ResolventModes_0 | head -30 > $O/result_files.txt; du -sh $W/results/ResolventModes_0 >> $O/result_files.txt; tail -3 $O/result_files.txt
