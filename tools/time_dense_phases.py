#!/usr/bin/env python
"""Times the dense phases between the actions -- QRAllFreqs, SVDAllFreqsBeforeAdjoint, StoreU, SVDAllFreqs -- at the per-GPU
size of BASELINE.json configs[3] (2.5e6 rows, k = 8, Nw = 32: Y_hat is 10.2 GB) and compares the Gram / QR variants:

    default            CholeskyQR2 in 5 passes over Y_hat (first apply fused with the second Gram), FP64 tensor cores (DMMA)
    qr_fused=0         6 passes (Gram, apply, Gram, apply), DMMA Gram
    gram_variant=0     ... with the FMA Gram kernel of round 1

    python tools/time_dense_phases.py [--rows 2500000] [--k 8] [--Nw 32]
Under ncu (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,...`) the same script gives the per-kernel evidence kept in
profiles/r2_dense_phases_ncu.txt."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_500_000)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--Nw", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", default="")
args = ap.parse_args()

import torch  # noqa: E402

pkg = entry.load_package()
entry.build()
nz = max(1, args.rows // (200 * 200 * 5))
shape = (nz, 200, 200)
N = nz * 200 * 200 * 5
dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(0.05, args.Nw, 0.003, 10.0)
r = pkg.capi.RSVDt(0)
r.set_operator_stencil(shape, 5, with_adjoint=False)
r.set_grid(args.k, args.Nw, Ns, Nt, RR, dt, 0.05, False)
pass_gb = N * args.k * args.Nw * 16 / 1e9
out = {"rows": N, "k": args.k, "Nw": args.Nw, "Y_hat_GB": pass_gb, "phases": {}}


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.time()
    fn()
    torch.cuda.synchronize()
    return 1e3 * (time.time() - t0)


variants = [("qr 5-pass: apply fused with the second Gram, all DMMA (default)", {}), ("qr 6-pass, DMMA Gram", dict(qr_fused=0)),
            ("qr 6-pass, FMA Gram (round 1)", dict(qr_fused=0, gram_variant=0))]
for name, opts in variants:
    if args.only and args.only not in name:
        continue
    for o in ("qr_fused", "gram_variant"):
        r.set_option(o, opts.get(o, 1))
    best = None
    for _ in range(args.reps):
        r.CreateRandomMat(seed=5)
        ms = timed(r.QRAllFreqs)
        best = ms if best is None else min(best, ms)
    npass = 5 if not opts else 6
    out["phases"][name] = {"ms": best, "passes_over_Y_hat": npass, "GBps_if_hbm_bound": npass * pass_gb / (best * 1e-3)}
for o in ("qr_fused", "gram_variant"):
    r.set_option(o, 1)
if not args.only:
    r.CreateRandomMat(seed=5)
    out["phases"]["SVDAllFreqsBeforeAdjoint (Gram + Jacobi + apply: 3 passes)"] = {"ms": timed(r.SVDAllFreqsBeforeAdjoint)}
    out["phases"]["StoreU (copy: 2 passes)"] = {"ms": timed(r.StoreU)}
    out["phases"]["SVDAllFreqs (Gram + Jacobi + apply: 3 passes)"] = {"ms": timed(r.SVDAllFreqs)}
print(json.dumps(out, indent=1))
r.close()
