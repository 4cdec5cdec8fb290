"""TransientRemovalEst.c, restated for the host drivers (a diagnostic of the transient run, `TransRemovalEst: true`).

The transient run (TransientRunRK4.c:174-184) keeps the last Nw snapshots of the decaying solution, one every
Nt_period / Nw steps, in the N x Nw matrix Q_all and, at the end of every period, asks how well the transient could be
removed from a steady-state response at each non-negative frequency w_i = i * w:

  * "initial" norms  (TransientRemovalEst.c:58-71):   || Q_all * dft[:, i] / Nw ||_2,  dft = CreateDFTiDFTMats.c:74-84
  * "updated" norms  (:75-119): with synthetic snapshots  y_s = q_ss e^{i w_i s dT} + Q_all[:, s]  (q_ss = -q0, the negated
    initial condition; in the first period the leading snapshot is skipped), U = orth(Y):  || q_ss - U U^H q_ss || / || q_ss ||

The snapshots come from the device through the second callback cadence of `rsvdt_transient_run_rk4_est`; the arithmetic
here is host work on Nw vectors per period (numpy).  Rows may be partitioned over ranks: `allreduce(x)` sums an array over the
ranks and `allgather(x)` returns the list of every rank's array; the orthonormal basis is then a TSQR (local QR, QR of the
stacked R factors), which is as accurate as the Gram-Schmidt with refinement BVOrthogonalize performs (QRDecomposition.c:15-19).
"""
from __future__ import annotations

import math

import numpy as np


def dft_matrix(Nw: int, Nw_eff: int) -> np.ndarray:
    """CreateDFTiDFTMats.c:74-84: dft[s][i] = exp(-2 pi i s i / Nw), Nw x Nw_eff (Nw_eff = Nw / 2 for a real operator)"""
    s = np.arange(Nw, dtype=np.float64)[:, None]
    i = np.arange(Nw_eff, dtype=np.float64)[None, :]
    return np.exp(-2j * math.pi * s * i / Nw)


def _orth(Y, allgather):
    """orthonormal basis of the columns of the row-partitioned Y (TSQR)"""
    if allgather is None:
        return np.linalg.qr(Y)[0]
    n = Y.shape[1]
    if Y.shape[0] >= n:
        Q1, R1 = np.linalg.qr(Y)
    else:                                  # fewer local rows than columns: pad the R factor so that every rank sends n x n
        Q1, R1 = np.linalg.qr(np.vstack([Y, np.zeros((n - Y.shape[0], n), Y.dtype)]))
        Q1 = Q1[:Y.shape[0]]
    Rs = allgather(np.ascontiguousarray(R1))
    Q2 = np.linalg.qr(np.vstack(Rs))[0]
    me = allgather.rank
    return Q1 @ Q2[me * n:(me + 1) * n]


def transient_removal_est(Q_all, qss, period_index, w, Nw_eff, log, allreduce=None, allgather=None):
    """-> (initial_norms, updated_norms), each Nw // 2 doubles.  Q_all: local rows x Nw, qss: local rows."""
    red = (lambda x: x) if allreduce is None else allreduce
    Nw = Q_all.shape[1]
    deltaT = 2 * math.pi / w / Nw                                                  # TransientRemovalEst.c:26
    log(f"**** deltaT = {deltaT:g} ****")
    skip = 1 if period_index == 1 else 0                                           # :28 (the first snapshot of the first period)
    Ns = Nw - skip
    log(f"**** Subspace dim = {Ns} ****")
    log(f"**** period index = {period_index} ****")
    log("**** Displaying transient norms for (positive) frequencies ****")
    nhalf = Nw // 2
    Qhat = (Q_all @ dft_matrix(Nw, Nw_eff)) / Nw                                   # :58-59
    init = np.sqrt(red(np.sum(np.abs(Qhat[:, :nhalf]) ** 2, axis=0)))
    for i in range(nhalf):
        log(f"**** Norm of the ratio norm(qt)/norm(qs): {init[i]:g} @ w = {w * i:g} ****")
    log("**** Displaying updated transient norms for (positive) frequencies ****")
    qn = math.sqrt(float(red(np.array([np.vdot(qss, qss).real]))[0]))
    upd = np.zeros(nhalf)
    s = np.arange(Ns, dtype=np.float64)
    for i in range(nhalf):
        wi = w * i
        Y = Q_all[:, skip:skip + Ns] + qss[:, None] * np.exp(1j * wi * s * deltaT)[None, :]     # :86-98
        U = _orth(Y, allgather)                                                    # :106
        b = red(U.conj().T @ qss)                                                  # :107
        d = qss - U @ b                                                            # :108-109
        upd[i] = math.sqrt(float(red(np.array([np.vdot(d, d).real]))[0])) / qn     # :110-111
        log(f"**** Norm of the ratio norm(qt_up)/norm(qs): {upd[i]:g} @ w = {wi:g} ****")
    return init, upd
