#!/usr/bin/env python
"""One direct action WITH transient removal on the per-GPU share of BASELINE.json configs[3] (3-D LNS block stencil,
12 x 200 x 200 x 5 = 2.4e6 rows, k = 8, Nw = 32), short period, for an ncu launch list of everything that is not the RK4
stage kernel: snapshot handling, DFT, the transient-removal algebra (TransientRemovalStrategy.c:30-40, CreateMtilde.c:41-54,
GalerkinProj.c:26-74), e^{A dT} on the snapshot slots.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        -k regex:'^(?!.*rk4_stage_tbcsr).*' --csv --log-file out.csv python tools/profile_tr_action.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
shape, nvar, k, Nw = (12, 200, 200), 5, 8, 32
w = 20.0                                   # short period: T = 0.314 -> 128 RK4 steps per period at dt = 0.003
dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(w, Nw, 0.003, 0.1)
r = pkg.capi.RSVDt(0)
r.set_operator_stencil(shape, nvar, with_adjoint=False)
r.set_grid(k, Nw, Ns, Nt, RR, dt, w, True)
N = 12 * 200 * 200 * 5
r.CreateRandomMat(N, 0, seed=14)
t0 = time.time()
r.DirectActionRK4()
t1 = time.time()
print(json.dumps({"rows": N, "k": k, "Nw": Nw, "Ns": Ns, "Nt": Nt, "ResRatio": RR, "dt": dt, "action_s": t1 - t0,
                  "launches": r.kernel_launches()}))
r.close()
