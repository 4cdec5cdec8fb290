#!/usr/bin/env python
"""Row-partitioned LNS-type operator (BASELINE.json configs[3]/[4] shapes) on 1..8 GPUs: times a window of forced RK4
steps with the block-CSR stage kernel and one halo exchange per stage (layout 0).

    python tools/bench_lns_dist.py --shape 25,200,200 --k 8 --Nw 32 --steps 16           # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 ... tools/bench_lns_dist.py --shape 200,200,200 ...

--shape is the GLOBAL grid (slowest dimension first); the grid is split along the slowest dimension in row order, which is
the PETSC_DECIDE row split whenever the number of grid planes divides evenly."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="16,100,100")
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--Nw", type=int, default=32)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--option", action="append", default=[])
ap.add_argument("--no-adjoint", action="store_true", help="do not build A^H (direct-only timing, halves the setup memory)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
pkg = entry.load_package()
shape = tuple(int(x) for x in args.shape.split(","))
nvar = 5
N = int(np.prod(shape)) * nvar
plane = int(np.prod(shape[1:])) * nvar
bw = plane + (shape[-1] + 1) * nvar + nvar          # |col - row| bound of the 3^d-point block stencil
dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(0.05, args.Nw, 0.003, 10.0)
grid = dict(k=args.k, Nw=args.Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=0.05)


def gen_rows(a, b):
    a5, b5 = a - a % nvar, b + (-b) % nvar
    _, rpp, cc, vv = pkg.synthetic.lns_stencil(shape, nvar=nvar, row_range=(a5, b5))
    s, e = rpp[a - a5], rpp[b - a5]
    return rpp[a - a5: b - a5 + 1] - s, cc[s:e], vv[s:e]


t0 = time.time()
if world == 1:
    n, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=nvar)
    r = pkg.capi.RSVDt(lrank)
    r.LNSStructure(N, rp, ci, va)
    r.set_grid(args.k, args.Nw, Ns, Nt, RR, dt, 0.05)
    n_own, nnz_own = N, int(rp[-1])
else:
    r, plan = pkg.dist.setup_rank(pkg, lrank, N, gen_rows, bw, 0, None, grid, with_adjoint=not args.no_adjoint)
    n_own = plan.n_own
    nnz_own = None
for opt in args.option:
    name, value = opt.split("=")
    r.set_option(name, int(value))
setup_s = time.time() - t0
rng = np.random.default_rng(14 + rank)
F = np.empty((n_own, args.k * args.Nw), np.complex128, order="F")
c0 = rng.random(n_own) + 1j * rng.random(n_own)
for j in range(F.shape[1]):
    F[:, j] = np.roll(c0, 11 * j)
r.set_Y_hat(F)
nnz_row = 27 * nvar if len(shape) == 3 else 9 * nvar
model_bytes = 4 * (nnz_row * 20 + 4) / args.k + 20 * 16      # SURVEY.md 8(d), interior rows
best = None
for _ in range(args.passes):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = r.time_forced_steps(args.steps)
    st, n = r.last_kernel_ms("stage")
    t = torch.tensor([ms, st], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, st = float(t[0]), float(t[1])
    best = (ms, st, n) if best is None or ms < best[0] else best
ms, st, n = best
if rank == 0:
    units = float(N) * args.k * args.steps
    out = {"operator": f"LNS block stencil {shape} x {nvar} vars", "N": N, "k": args.k, "Nw": args.Nw, "n_gpus": world,
           "rk4_steps": args.steps, "ms_total": ms, "ms_rk4_stage_kernels_incl_halo": st, "stage_launches": n,
           "units_per_s": units / (ms * 1e-3),
           "stage_algorithmic_GBps_per_gpu": model_bytes * units / world / (st * 1e-3) / 1e9,
           "stage_frac_of_measured_hbm_peak": model_bytes * units / world / (st * 1e-3) / 1e9 / 6539.2,
           "setup_s": setup_s}
    print(json.dumps(out), flush=True)
r.close()
if world > 1:
    dist.destroy_process_group()
