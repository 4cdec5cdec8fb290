#!/usr/bin/env python
"""Times the actions of the tutorial configuration (A_GL, k=10, Nw=42, w=0.05, TransientRemoval on) for a few option sets."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
M, N, rp, ci, va = pkg.petsc_io.read_mat(os.path.join(ROOT, "tests", "golden", "A_GL"))
k, Nw, w = 10, 42, 0.05
dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(w, Nw, 0.003, 10.0)
F = pkg.synthetic.random_forcing(N, k, Nw, seed=14)
for opts in ({}, {"use_graph": 0}, {"use_banded": 0}, {"use_banded": 0, "use_graph": 0}, {"use_small": 2}):
    r = pkg.capi.RSVDt(0)
    r.LNSStructure(N, rp, ci, va)
    r.set_grid(k, Nw, Ns, Nt, RR, dt, w, True)
    for a, b in opts.items():
        r.set_option(a, b)
    r.set_Y_hat(F)
    r.DirectActionRK4()          # warm-up (allocations, graph instantiation)
    ts = []
    for direct in (False, True, False):
        r.QRAllFreqs()           # PowerIterationRK4.c:20-29 order: QR between the actions
        t0 = time.time()
        (r.DirectActionRK4 if direct else r.AdjointActionRK4)()
        ts.append(time.time() - t0)
    print(opts, "action seconds:", " ".join(f"{t:.3f}" for t in ts), "launches", r.kernel_launches(), flush=True)
    r.close()
