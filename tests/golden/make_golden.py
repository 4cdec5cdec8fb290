#!/usr/bin/env python
"""Regenerates tests/golden/gl_small_golden.npz: inputs and outputs of the CPU oracle (oracle/rsvdt_oracle.c) for a small
Ginzburg-Landau problem.  The reference itself cannot be run here (PETSc/SLEPc/MPI absent), so these vectors pin the
ORACLE (and, through the GPU tests, the CUDA path) against silent changes; they are not outputs of the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import oracle as O  # noqa: E402

pkg = entry.load_package()
M, N, rp, ci, va = pkg.petsc_io.read_mat(os.path.join(HERE, "A_GL"))
A = O.Csr(M, N, rp, ci, va).shifted(0.3)
k, Nw, w = 3, 6, 0.6
out = {"k": k, "Nw": Nw, "w": w, "beta": 0.3, "dt_in": 0.003, "TransientLength": 1.0}
rng = np.random.default_rng(2026)
F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
out["F_hat"] = F
for TR in (False, True):
    dt, Ns, Nt, RR = O.setup_time_freq_grid(w, Nw, 0.003, 1.0)
    v = O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w, TransientRemoval=TR)
    dft, idft = O.create_dft_idft_mats(Nw, Nw, Ns, False, False)
    tag = "tr" if TR else "notr"
    out[f"Y_direct_{tag}"] = O.ts_action_rk4(A, v, dft, idft, F)
    va_ = O.Vars(**{**v.__dict__, "DirAdj": False})
    out[f"Y_adjoint_{tag}"] = O.ts_action_rk4(A.hermitian_transpose(), va_, dft, idft, F)
    S, U, V = O.full_solve(A, v, 1, F)
    out[f"S_{tag}"] = S
out["grid"] = np.array([dt, Ns, Nt, RR])
np.savez_compressed(os.path.join(HERE, "gl_small_golden.npz"), **out)
print("written", os.path.join(HERE, "gl_small_golden.npz"))
