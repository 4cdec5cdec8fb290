#!/usr/bin/env python
"""Small driver for ncu: sets up the bench workload (BASELINE.json configs[1]) and runs a few forced RK4 steps.
Usage under ncu (one GPU):   ncu ... python tools/profile_step.py --steps 32 [--option name=value]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=32)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--N", type=int, default=bench.WORKLOAD["N"])
ap.add_argument("--option", action="append", default=[])
ap.add_argument("--adjoint", action="store_true")
ap.add_argument("--op", default="gl", help="gl | lns2d | lns3d")
ap.add_argument("--shape", default="", help="grid shape for lns operators, e.g. 1000,1000 or 40,100,100")
ap.add_argument("--k", type=int, default=0)
ap.add_argument("--Nw", type=int, default=0)
args = ap.parse_args()

pkg = entry.load_package()
c = dict(bench.WORKLOAD)
c["N"] = args.N
if args.k:
    c["k"] = args.k
if args.Nw:
    c["Nw"] = args.Nw
if args.op == "gl":
    N, rowptr, col, val = pkg.synthetic.ginzburg_landau(c["N"])
else:
    shape = tuple(int(x) for x in args.shape.split(","))
    N, rowptr, col, val = pkg.synthetic.lns_stencil(shape, nvar=5)
    print(f"{args.op}: N={N} nnz={rowptr[-1]} nnz/row={rowptr[-1] / N:.1f}", flush=True)
dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(c["w"], c["Nw"], c["dt_in"], c["TransientLength"])
r = pkg.capi.RSVDt(0)
r.LNSStructure(N, rowptr, col, val)
r.set_grid(c["k"], c["Nw"], Ns, Nt, RR, dt, c["w"], False)
for opt in args.option:
    name, value = opt.split("=")
    r.set_option(name, int(value))
rng = np.random.default_rng(14)
F = np.empty((N, c["k"] * c["Nw"]), np.complex128, order="F")
col0 = rng.random(N) + 1j * rng.random(N)
for j in range(F.shape[1]):
    F[:, j] = np.roll(col0, 7 * j)
nnz_row = rowptr[-1] / N
model_bytes = 4 * (nnz_row * 20 + 4) / c["k"] + 20 * 16
r.set_Y_hat(F)
for _ in range(args.passes):
    ms = r.time_forced_steps(args.steps, direct=not args.adjoint)
    st, n = r.last_kernel_ms("stage")
    idf, m = r.last_kernel_ms("idft")
    print(f"{args.steps} steps: {ms:.3f} ms total, stage {st:.3f} ms / {n} launches, idft {idf:.3f} ms / {m} launches")
    gbs = model_bytes * N * c["k"] * args.steps / (st * 1e-3) / 1e9
    print(f"   RK4 kernels: {model_bytes:.1f} algorithmic B/unit/step -> {gbs:.0f} GB/s algorithmic "
          f"({gbs / 6539.2:.2f} of the measured HBM peak); {N * c['k'] * args.steps / (ms * 1e-3):.3e} units/s overall")
r.close()
