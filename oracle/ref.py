"""ctypes binding of oracle/_ref/librsvdt_ref.so: the reference's OWN source files (SourceCode/*.c), compiled unmodified
from /root/reference by oracle/Makefile.ref against the PETSc/SLEPc API shim (oracle/petsc_shim/), behind the small
harness oracle/ref_harness.c.

TEST INFRASTRUCTURE ONLY.  Only tests/, tests/golden/make_ref_golden.py, __graft_entry__ (build) and bench.py's
cpu_baseline / --impl reference legs may import this module.  The sources exist only in the build container
(/root/reference is absent on the GPU box); the built _ref/ files travel with the snapshot.

The calling conventions mirror oracle/oracle.py so a test can run the same inputs through the oracle restatement, the
reference code and the CUDA path.  One difference: functions here take the forward operator A and select A or A^H through
Vars.DirAdj, exactly as the reference does (TSRK4.c:18-19).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import Csr, Vars, _F, _p, _vp, c128

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "librsvdt_ref.so")
EXE_PATH = os.path.join(REF_DIR, "RSVDt_ref")
REFERENCE_ROOT = os.environ.get("RSVDT_REFERENCE_ROOT", "/root/reference")
_lib = None


class _RefCsr(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("rowptr", _vp), ("col", _vp), ("val", _vp)]


def sources_present() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "SourceCode"))


def available() -> bool:
    return os.path.exists(LIB_PATH) and os.path.exists(EXE_PATH)


def build(force: bool = False) -> bool:
    """Compile the reference sources where they lie (never copied).  Returns False when they are not present."""
    if not sources_present():
        return available()
    if force:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", "clean"])
    subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", f"REF={REFERENCE_ROOT}"])
    return True


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/librsvdt_ref.so is missing: run oracle.ref.build() where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _check(err, what):
    if err:
        raise RuntimeError(f"reference {what} returned PetscErrorCode {err}")


def num_threads() -> int:
    return lib().ref_num_threads()


def setup_time_freq_grid(w, Nw, dt_in, TransientLength):
    """SetupTimeFreqGrid.c -> (dt, Ns, Nt, ResRatio)"""
    dt = C.c_double()
    Ns, Nt, RR = C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib().ref_setup_time_freq_grid(C.c_double(w), C.c_int64(Nw), C.c_double(dt_in), C.c_double(TransientLength),
                                          C.byref(dt), C.byref(Ns), C.byref(Nt), C.byref(RR)), "SetupTimeFreqGrid")
    return dt.value, Ns.value, Nt.value, RR.value


def create_dft_idft_mats(v: Vars):
    """CreateDFTiDFTMats.c -> (dft [Nw x Nw_eff], idft [Nw_eff x 2Ns])"""
    dft = np.empty((v.Nw, v.Nw_eff), c128, order="F")
    idft = np.empty((v.Nw_eff, 2 * v.Ns), c128, order="F")
    vc = v.c()
    _check(lib().ref_create_dft_idft_mats(C.byref(vc), _p(dft), _p(idft)), "CreateDFTiDFTMats")
    return dft, idft


def ts_action_rk4(A: Csr, v: Vars, Y_hat_in):
    """TSActionRK4.c with the reference's own DFT / transient-removal tail."""
    Yin = _F(Y_hat_in)
    assert Yin.shape == (v.N, v.k * v.Nw_eff)
    Yout = np.empty_like(Yin, order="F")
    vc = v.c()
    _check(lib().ref_ts_action_rk4(C.byref(vc), _p(A.rowptr), _p(A.col), _p(A.val), _p(Yin), _p(Yout)), "TSActionRK4")
    return Yout


def _rc(M):
    if M is None:
        return None, None
    s = _RefCsr(M.nrows, M.ncols, _p(M.rowptr), _p(M.col), _p(M.val))
    return s, C.byref(s)


def weighted_action(A: Csr, v: Vars, Y_hat_in, B=None, Cm=None, Wf_inv=None, Wq=None):
    """DirectActionRK4.c / AdjointActionRK4.c (v.DirAdj) with ApplyWeightMats.c on both sides."""
    Nb = B.ncols if B is not None else v.N
    Nc = Cm.nrows if Cm is not None else v.N
    rows_in, rows_out = (Nb, Nc) if v.DirAdj else (Nc, Nb)
    Yin = _F(Y_hat_in)
    assert Yin.shape == (rows_in, v.k * v.Nw_eff)
    Yout = np.empty((rows_out, v.k * v.Nw_eff), c128, order="F")
    keep = [_rc(B), _rc(Cm), _rc(Wf_inv), _rc(Wq)]
    vc = v.c()
    _check(lib().ref_weighted_action(C.byref(vc), _p(A.rowptr), _p(A.col), _p(A.val), keep[0][1], keep[1][1], keep[2][1],
                                     keep[3][1], C.c_int64(rows_in), _p(Yin), C.c_int64(rows_out), _p(Yout)),
           "Direct/AdjointActionRK4")
    return Yout


def qr_all_freqs(N, k, Nw_eff, Y_hat):
    Y = _F(Y_hat).copy(order="F")
    _check(lib().ref_qr_all_freqs(C.c_int64(N), C.c_int64(k), C.c_int64(Nw_eff), _p(Y)), "QRAllFreqs")
    return Y


def qr_decomposition(V):
    Q = _F(V).copy(order="F")
    _check(lib().ref_qr_decomposition(C.c_int64(Q.shape[0]), C.c_int64(Q.shape[1]), _p(Q)), "QRDecomposition")
    return Q


def svd_before_adjoint(N, k, Nw_eff, Y_hat):
    Y = _F(Y_hat).copy(order="F")
    _check(lib().ref_svd_before_adjoint(C.c_int64(N), C.c_int64(k), C.c_int64(Nw_eff), _p(Y)), "SVDAllFreqsBeforeAdjoint")
    return Y


def svd_all_freqs(N, k, Nw_eff, Y_hat):
    """-> (left singular vectors [N x k*Nw_eff], S [k x Nw_eff])"""
    Y = _F(Y_hat)
    U = np.empty_like(Y, order="F")
    S = np.empty((k, Nw_eff), np.float64, order="F")
    _check(lib().ref_svd_all_freqs(C.c_int64(N), C.c_int64(k), C.c_int64(Nw_eff), _p(Y), _p(U), _p(S)), "SVDAllFreqs")
    return U, S


def ts_exp_a_deltat(A: Csr, v: Vars, U):
    Uc = _F(U).copy(order="F")
    vc = v.c()
    _check(lib().ref_ts_exp_a_deltat(C.byref(vc), _p(A.rowptr), _p(A.col), _p(A.val), C.c_int64(Uc.shape[1]), _p(Uc)),
           "TSExpADeltaT")
    return Uc


def permute_mat(F, k, reverse=False):
    Fc = _F(F).copy(order="F")
    N, cols = Fc.shape
    _check(lib().ref_permute_mat(C.c_int64(N), C.c_int64(k), C.c_int64(cols // k), C.c_int(int(reverse)), _p(Fc)), "PermuteMat")
    return Fc


def trans_steps(A: Csr, v: Vars, nsteps, q0):
    q = np.ascontiguousarray(q0, dtype=c128).copy()
    vc = v.c()
    _check(lib().ref_trans_steps(C.byref(vc), _p(A.rowptr), _p(A.col), _p(A.val), C.c_int64(nsteps), _p(q)), "TSTransRK4")
    return q


def time_forced_steps(A: Csr, v: Vars, F_hat_k, ncols, warm, nsteps, want_state=False):
    """seconds spent in `nsteps` calls of TSRK4.c on each of `ncols` test vectors (run one after the other);
    want_state: also the vector y the reference's loop holds after step warm + nsteps (of the last test vector)"""
    Fk = _F(F_hat_k)
    assert Fk.shape == (v.N, v.Nw_eff)
    sec = C.c_double()
    vc = v.c()
    y = np.empty(v.N, c128) if want_state else None
    _check(lib().ref_time_forced_steps(C.byref(vc), _p(A.rowptr), _p(A.col), _p(A.val), _p(Fk), C.c_int64(ncols),
                                       C.c_int64(warm), C.c_int64(nsteps), C.byref(sec), _p(y) if want_state else None),
           "TSRK4 timing")
    return (sec.value, y) if want_state else sec.value


def run_executable(workdir, yaml_name="variables.yaml", timeout=3600, threads=None, autoinit=False):
    """`RSVDt_ref -inputs variables.yaml` in workdir (README.md:221) -> CompletedProcess.
    autoinit=True runs the build with -ftrivial-auto-var-init=pattern (see Makefile.ref: main.c never initialises
    RSVDt.TS.DirAdj on the TransRun path, so which of A / A^H the transient run integrates is undefined behaviour)."""
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    return subprocess.run([EXE_PATH + ("_autoinit" if autoinit else ""), "-inputs", yaml_name], cwd=workdir, capture_output=True, text=True, timeout=timeout, env=env)
