"""ctypes binding of the CPU oracle (oracle/rsvdt_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/rsvdt_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

All dense matrices are numpy complex128 arrays in Fortran (column-major) order, i.e. exactly the
memory image of a one-rank PETSc MATDENSE, so that the comparisons with the CUDA path happen at the
reference's own boundary (Y_hat: N x (k*Nw_eff), frequency-major columns, TSActionRK4.c:49-50).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librsvdt_oracle.so")
_lib = None

c128 = np.complex128
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "rsvdt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Csr(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("rowptr", _vp), ("col", _vp), ("val", _vp)]


class _Vars(C.Structure):
    _fields_ = [("N", C.c_int64), ("k", C.c_int64), ("Nw", C.c_int64), ("Nw_eff", C.c_int64),
                ("Ns", C.c_int64), ("Nt", C.c_int64), ("ResRatio", C.c_int64),
                ("dt", C.c_double), ("w", C.c_double),
                ("RealOperator", C.c_int), ("TransientRemoval", C.c_int), ("DirAdj", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_time_forced_steps.restype = C.c_double
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


class Csr:
    """CSR operator with 64-bit indices and complex128 values (a one-rank MATMPIAIJ)."""

    def __init__(self, nrows, ncols, rowptr, col, val):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = np.ascontiguousarray(val, dtype=c128)
        self._c = _Csr(self.nrows, self.ncols, _p(self.rowptr), _p(self.col), _p(self.val))

    @property
    def c(self):
        return C.byref(self._c)

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    def hermitian_transpose(self) -> "Csr":
        tr = np.empty(self.ncols + 1, np.int64)
        tc = np.empty(max(self.nnz, 1), np.int64)
        tv = np.empty(max(self.nnz, 1), c128)
        lib().orc_csr_hermitian_transpose(self.c, _p(tr), _p(tc), _p(tv))
        return Csr(self.ncols, self.nrows, tr, tc[: self.nnz], tv[: self.nnz])

    def shifted(self, beta: float) -> "Csr":
        """A - beta*I  (LNSStructure.c:53-56, MatShift); requires a stored diagonal."""
        val = self.val.copy()
        for r in range(self.nrows):
            s, e = self.rowptr[r], self.rowptr[r + 1]
            hit = np.nonzero(self.col[s:e] == r)[0]
            if hit.size == 0:
                raise ValueError("diagonal entry missing")
            val[s + hit[0]] -= beta
        return Csr(self.nrows, self.ncols, self.rowptr, self.col, val)

    def to_dense(self) -> np.ndarray:
        A = np.zeros((self.nrows, self.ncols), c128)
        for r in range(self.nrows):
            s, e = self.rowptr[r], self.rowptr[r + 1]
            np.add.at(A[r], self.col[s:e], self.val[s:e])
        return A


@dataclass
class Vars:
    N: int
    k: int
    Nw: int
    Nw_eff: int
    Ns: int
    Nt: int
    ResRatio: int
    dt: float
    w: float
    RealOperator: bool = False
    TransientRemoval: bool = False
    DirAdj: bool = True

    def c(self):
        return _Vars(self.N, self.k, self.Nw, self.Nw_eff, self.Ns, self.Nt, self.ResRatio, self.dt, self.w,
                     int(self.RealOperator), int(self.TransientRemoval), int(self.DirAdj))


def setup_time_freq_grid(w, Nw, dt_in, TransientLength):
    """SetupTimeFreqGrid.c:16-21 -> (dt, Ns, Nt, ResRatio)"""
    dt = C.c_double()
    Ns, Nt, RR = C.c_int64(), C.c_int64(), C.c_int64()
    lib().orc_setup_time_freq_grid(C.c_double(w), C.c_int64(Nw), C.c_double(dt_in), C.c_double(TransientLength),
                                   C.byref(dt), C.byref(Ns), C.byref(Nt), C.byref(RR))
    return dt.value, Ns.value, Nt.value, RR.value


def create_dft_idft_mats(Nw, Nw_eff, Ns, real_operator=False, faithful=False):
    """CreateDFTiDFTMats.c -> (dft [Nw x Nw_eff], idft [Nw_eff x 2Ns]) column-major"""
    dft = np.empty((Nw, Nw_eff), c128, order="F")
    idft = np.empty((Nw_eff, 2 * Ns), c128, order="F")
    lib().orc_create_dft_idft_mats(C.c_int64(Nw), C.c_int64(Nw_eff), C.c_int64(Ns), C.c_int(int(real_operator)),
                                   C.c_int(int(faithful)), _p(dft), _p(idft))
    return dft, idft


def _F(a):
    return np.asfortranarray(a, dtype=c128)


def ts_action_rk4(Op: Csr, v: Vars, dft, idft, Y_hat_in, want_snapshots=False, max_steps=0):
    """TSActionRK4.c:38-148.  Op must already be A (DirAdj) or A^H (adjoint)."""
    Yin = _F(Y_hat_in)
    assert Yin.shape == (v.N, v.k * v.Nw_eff)
    Yout = np.empty_like(Yin, order="F")
    Nstore = v.Nw + 1 if v.TransientRemoval else v.Nw
    snaps = np.empty((v.N, Nstore * v.k), c128, order="F") if want_snapshots else None
    vc = v.c()
    err = lib().orc_ts_action_rk4(Op.c, C.byref(vc), _p(_F(dft)), _p(_F(idft)), _p(Yin), _p(Yout),
                                  _p(snaps) if snaps is not None else None, C.c_int64(max_steps))
    if err:
        raise RuntimeError(f"orc_ts_action_rk4 failed: {err}")
    return (Yout, snaps) if want_snapshots else Yout


def qr_all_freqs(N, k, Nw_eff, Y_hat):
    Y = _F(Y_hat).copy(order="F")
    err = lib().orc_qr_all_freqs(C.c_int64(N), C.c_int64(k), C.c_int64(Nw_eff), _p(Y))
    if err:
        raise RuntimeError("orc_qr_all_freqs: rank deficient block")
    return Y


def qr_decomposition(V):
    Q = _F(V).copy(order="F")
    err = lib().orc_qr_decomposition(C.c_int64(Q.shape[0]), C.c_int64(Q.shape[1]), _p(Q))
    if err:
        raise RuntimeError("orc_qr_decomposition: rank deficient")
    return Q


def svd_all_freqs(N, k, Nw_eff, Y_hat):
    """-> (U [N x k*Nw_eff], S [k x Nw_eff])"""
    Y = _F(Y_hat)
    U = np.empty_like(Y, order="F")
    S = np.empty((k, Nw_eff), np.float64, order="F")
    lib().orc_svd_all_freqs(C.c_int64(N), C.c_int64(k), C.c_int64(Nw_eff), _p(Y), _p(U), _p(S))
    return U, S


def ts_exp_a_deltat(Op: Csr, v: Vars, U):
    Uc = _F(U).copy(order="F")
    vc = v.c()
    lib().orc_ts_exp_a_deltat(Op.c, C.byref(vc), C.c_int64(Uc.shape[1]), _p(Uc))
    return Uc


def transient_removal_strategy(Op: Csr, v: Vars, dft, Y_all):
    Y = _F(Y_all)
    out = np.empty((v.N, v.k * v.Nw_eff), c128, order="F")
    vc = v.c()
    err = lib().orc_transient_removal_strategy(Op.c, C.byref(vc), _p(_F(dft)), _p(Y), _p(out))
    if err:
        raise RuntimeError("orc_transient_removal_strategy failed")
    return out


def dft_action(v: Vars, dft, Y_all):
    out = np.empty((v.N, v.k * v.Nw_eff), c128, order="F")
    lib().orc_dft(C.c_int64(v.N), C.c_int64(v.k), C.c_int64(v.Nw), C.c_int64(v.Nw_eff), C.c_int64(v.ResRatio),
                  _p(_F(dft)), _p(_F(Y_all)), _p(out))
    return out


def transient_run_rk4(Op: Csr, dt, rend, TransSaveMod, TransDivVal, TransConVal, q0):
    """TransientRunRK4.c:128-158 -> (q_last, norms, steps_done, reason)"""
    q = np.ascontiguousarray(q0, dtype=c128).copy()
    norms = np.zeros(rend // TransSaveMod + 1, np.float64)
    steps, reason = C.c_int64(), C.c_int()
    lib().orc_transient_run_rk4(Op.c, C.c_double(dt), C.c_int64(rend), C.c_int64(TransSaveMod),
                                C.c_double(TransDivVal), C.c_double(TransConVal), _p(q), _p(norms),
                                C.byref(steps), C.byref(reason))
    return q, norms, steps.value, reason.value


def mat_mat_mult(M: Csr, X, hermitian_transpose=False):
    Xc = _F(X)
    rows = M.ncols if hermitian_transpose else M.nrows
    Y = np.empty((rows, Xc.shape[1]), c128, order="F")
    lib().orc_mat_mat_mult(M.c, C.c_int(int(hermitian_transpose)), C.c_int64(Xc.shape[1]), _p(Xc), _p(Y))
    return Y


def time_forced_steps(Op: Csr, v: Vars, idft, Y_hat_in, ncols, nsteps) -> float:
    vc = v.c()
    return lib().orc_time_forced_steps(Op.c, C.byref(vc), _p(_F(idft)), _p(_F(Y_hat_in)), C.c_int64(ncols),
                                       C.c_int64(nsteps))


def forced_state(Op: Csr, v: Vars, idft, Y_hat_in, nsteps, ncols=None):
    """y after `nsteps` forced RK4 steps of TSActionRK4.c:116-122 for test vectors 0..ncols-1 -> N x ncols"""
    ncols = v.k if ncols is None else ncols
    y = np.empty((v.N, ncols), c128, order="F")
    vc = v.c()
    lib().orc_forced_state(Op.c, C.byref(vc), _p(_F(idft)), _p(_F(Y_hat_in)), C.c_int64(ncols), C.c_int64(nsteps), _p(y))
    return y


def num_threads() -> int:
    return lib().orc_num_threads()


# ----------------------------------------------------------------------------------------------------
# Full solve, main.c:134-146 + PowerIterationRK4.c:20-29 (weights off).
# ----------------------------------------------------------------------------------------------------
def full_solve(A: Csr, v: Vars, q: int, F_hat, faithful_twiddles=False):
    """-> (S [k x Nw_eff], U [N x k*Nw_eff], V [N x k*Nw_eff]) in the reference's frequency-major layout."""
    AH = A.hermitian_transpose()
    dft, idft = create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, v.RealOperator, faithful_twiddles)

    def act(direct, Y):
        vv = Vars(**{**v.__dict__, "DirAdj": direct})
        return ts_action_rk4(A if direct else AH, vv, dft, idft, Y)

    Y = act(True, F_hat)                                   # main.c:134
    for _ in range(q):                                     # PowerIterationRK4.c:20-29
        Y = qr_all_freqs(v.N, v.k, v.Nw_eff, Y)
        Y = act(False, Y)
        Y = qr_all_freqs(v.N, v.k, v.Nw_eff, Y)
        Y = act(True, Y)
    U, _ = svd_all_freqs(v.N, v.k, v.Nw_eff, Y)            # main.c:138-140
    Y = act(False, U)                                      # main.c:142
    V, S = svd_all_freqs(v.N, v.k, v.Nw_eff, Y)            # main.c:144
    return S, U, V
