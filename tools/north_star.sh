#!/bin/bash
# The BASELINE.json north-star run as a WHOLE SOLVE (transient removal, power iteration q = 1, adjoint pass, SVDs) through the
# multi-GPU drop-in executable: configs[3], 3-D block-stencil LNS operator, 100 x 200 x 200 points x 5 variables = 2e7 DOF,
# k = 8, Nw = 32, rows partitioned over P GPUs.  The operator is generated on the device (OperatorDir: synthetic:..., the same
# entries tools/make_case.py would write to a 65 GB file), the forcing by CreateRandomMat on the device, and the modes
# (2 x 82 GB) are not written (RSVDT_MPI_SAVE=gains).  With --weights the case is configs[4] (B, C on sub-boxes, diagonal
# weights).  Writes <workdir>/timings.json (per-phase seconds) and prints the wall time of the executable.
#   usage: tools/north_star.sh <workdir> [P=8] [nz ny nx = 100 200 200] [--weights]
set -e
W=${1:?workdir}; P=${2:-8}; NZ=${3:-100}; NY=${4:-200}; NX=${5:-200}; EXTRA=${6:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
python "$ROOT/tools/make_case.py" lns3d --shape $NZ $NY $NX --k 8 --Nw 32 --w 2.0 --dt 0.003 --TL 0.5 --TR --q 1 --synthetic $EXTRA --out "$W"
cd "$W"
T0=$(date +%s%N)
RSVDT_MPI_SAVE=gains RSVDT_MPI_TIMINGS="$W/timings.json" python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29650 \
    "$ROOT/rsvd-delta-t_b200/rsvdt_mpi.py" -inputs variables.yaml | tail -60
T1=$(date +%s%N)
echo "north-star whole solve on $P GPUs ($NZ x $NY x $NX x 5 DOF): $(( (T1 - T0) / 1000000 )) ms wall (includes process start, NCCL setup, operator generation)"
cat "$W/timings.json"
