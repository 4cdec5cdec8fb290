#!/usr/bin/env python
"""Times QRAllFreqs / SVDAllFreqs at the bench size (N=1e6, k=10, Nw=64) -- the phases between the actions."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

pkg = entry.load_package()
c = bench.WORKLOAD
N, rowptr, col, val, dt, Ns, Nt, RR = bench.build_workload(pkg)
r = pkg.capi.RSVDt(0)
r.LNSStructure(N, rowptr, col, val)
r.set_grid(c["k"], c["Nw"], Ns, Nt, RR, dt, c["w"], False)
rng = np.random.default_rng(1)
F = np.empty((N, c["k"] * c["Nw"]), np.complex128, order="F")
c0 = rng.random(N) + 1j * rng.random(N)
for j in range(F.shape[1]):
    F[:, j] = np.roll(c0, 13 * j) * (1 + 0.01 * j)
r.set_Y_hat(F)
for name, fn in (("QRAllFreqs", r.QRAllFreqs), ("QRAllFreqs", r.QRAllFreqs), ("SVDAllFreqsBeforeAdjoint", r.SVDAllFreqsBeforeAdjoint),
                 ("StoreU", r.StoreU), ("SVDAllFreqs", r.SVDAllFreqs)):
    t0 = time.time()
    fn()
    print(f"{name}: {1e3 * (time.time() - t0):.1f} ms", flush=True)
r.close()
