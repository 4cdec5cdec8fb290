"""ctypes binding of the C ABI (include/rsvdt.h -> librsvdt_b200.so).

Mirrors the reference's call structure: the `RSVDt` class below has one method per reference function
(DirectActionRK4, AdjointActionRK4, QRAllFreqs, PowerIterationRK4, SVDAllFreqsBeforeAdjoint, StoreU, SVDAllFreqs,
TSExpADeltaT, TransientRunRK4; SourceCode/main.c:113-146).  There is no CPU fallback: if the shared library is
missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsvdt_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rsvdt.h")
_lib = None

_vp = C.c_void_p
c128 = np.complex128


class RsvdtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rsvdt error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    """rsvdt_grid (include/rsvdt.h), i.e. the RSVD_vars / TS_vars members the hot path reads (Variables.h:5-28)."""
    _fields_ = [("k", C.c_int64), ("Nw", C.c_int64), ("Nw_eff", C.c_int64), ("Ns", C.c_int64), ("Nt", C.c_int64),
                ("ResRatio", C.c_int64), ("dt", C.c_double), ("w", C.c_double), ("RealOperator", C.c_int32),
                ("TransientRemoval", C.c_int32)]


SNAPSHOT_CB = C.CFUNCTYPE(C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int64, _vp)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (csrc/Makefile), in-tree."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in ("kernels.cu", "stage_tiled.cu", "stencil_gen.cu", "dense_k4.cu", "capi.cu", "group.inc", "kernels.cuh", "stage_tiled.cuh")] + [HEADER_PATH]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + ([] if verbose else ["-s"])
        subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RsvdtError(-1, f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                                 "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.rsvdt_last_error.restype = C.c_char_p
        L.rsvdt_version.restype = C.c_char_p
        L.rsvdt_yhat_rows.restype = C.c_int64
        L.rsvdt_kernel_launches.restype = C.c_int64
        L.rsvdt_last_error.argtypes = [_vp]
        L.rsvdt_kernel_launches.argtypes = [_vp, C.c_int]
        L.rsvdt_yhat_rows.argtypes = [_vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def plan_partition(N, nranks, rank):
    """PETSC_DECIDE row split through the C ABI -> (row_begin, row_end)"""
    b, e = C.c_int64(), C.c_int64()
    rc = lib().rsvdt_plan_partition(C.c_int64(N), C.c_int(nranks), C.c_int(rank), C.byref(b), C.byref(e))
    if rc:
        raise RsvdtError(rc, "rsvdt_plan_partition: bad arguments")
    return b.value, e.value


def setup_time_freq_grid(w, Nw, dt_in, TransientLength):
    """SetupTimeFreqGrid.c:16-21 through the C ABI -> (dt, Ns, Nt, ResRatio)"""
    dt = C.c_double()
    Ns, Nt, RR = C.c_int64(), C.c_int64(), C.c_int64()
    rc = lib().rsvdt_setup_time_freq_grid(C.c_double(w), C.c_int64(Nw), C.c_double(dt_in), C.c_double(TransientLength),
                                          C.byref(dt), C.byref(Ns), C.byref(Nt), C.byref(RR))
    if rc:
        raise RsvdtError(rc, "rsvdt_setup_time_freq_grid: bad arguments")
    return dt.value, Ns.value, Nt.value, RR.value


def _F(a):
    return np.asfortranarray(a, dtype=c128)


class RSVDt:
    """One handle = the reference's set of structs for one run (Variables.h)."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        rc = lib().rsvdt_create(C.c_int(device), C.byref(self._h))
        if rc:
            raise RsvdtError(rc, "rsvdt_create failed (no CUDA device? there is no CPU fallback)")
        self.N = 0
        self.grid = None
        self.real_operator = None

    def close(self):
        if self._h:
            lib().rsvdt_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RsvdtError(rc, lib().rsvdt_last_error(self._h).decode())

    # ---- pre-processing -----------------------------------------------------------------------------
    def LNSStructure(self, N, rowptr, col, val, DiscFlg=False, beta=0.0):
        """LNSStructure.c:23-56 -> RealOperator flag"""
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        col = np.ascontiguousarray(col, np.int64)
        val = np.ascontiguousarray(val, c128)
        ro = C.c_int()
        self._ck(lib().rsvdt_set_operator_csr(self._h, C.c_int64(N), _p(rowptr), _p(col), _p(val),
                                              C.c_int(int(DiscFlg)), C.c_double(beta), C.byref(ro)))
        self.N = int(N)
        self.real_operator = bool(ro.value)
        return self.real_operator

    def dist_init(self, rank, nranks, layout):
        """row partition over GPUs (MATMPIAIJ / VECMPI of the reference); see dist.py"""
        self._ck(lib().rsvdt_dist_init(self._h, C.c_int(rank), C.c_int(nranks), C.c_int(layout)))

    def dist_layout_supported(self, nvec) -> bool:
        ok = C.c_int()
        self._ck(lib().rsvdt_dist_layout_supported(self._h, C.c_int(int(nvec)), C.byref(ok)))
        return bool(ok.value)

    def LNSStructure_local(self, N_global, n_own, own0, n_ext, A, AH, real_operator, DiscFlg=False, beta=0.0):
        """LNSStructure.c:23-56 for one rank.  A, AH: (rowptr, col_local, val) of the stored local rows"""
        arrs = []
        for rp, ci, va in (A, AH):
            arrs += [np.ascontiguousarray(rp, np.int64), np.ascontiguousarray(ci, np.int64), np.ascontiguousarray(va, c128)]
        n_rows = arrs[0].size - 1
        self._ck(lib().rsvdt_set_operator_csr_local(self._h, C.c_int64(N_global), C.c_int64(n_own), C.c_int64(own0),
                                                    C.c_int64(n_ext), C.c_int64(n_rows), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
                                                    _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), C.c_int(int(DiscFlg)),
                                                    C.c_double(beta), C.c_int(int(real_operator))))
        self.N = int(n_own)
        self.real_operator = bool(real_operator)

    def set_operator_stencil(self, shape, nvar=5, pt_range=None, ghost_pts=None, with_adjoint=True, complex_coupling=True,
                             DiscFlg=False, beta=0.0, **gen):
        """synthetic.lns_stencil(shape, nvar, ...) generated on the device in the tiled block-CSR layout (include/rsvdt.h);
        pt_range = (pt0, pt1) grid points owned by this rank, ghost_pts = sorted grid points outside it that its rows touch"""
        from . import synthetic
        shape = tuple(int(x) for x in shape)
        blocks, mod_tab, diag_tab = synthetic.stencil_tables(len(shape), nvar, complex_coupling, **gen)
        npts = int(np.prod(shape))
        pt0, pt1 = (0, npts) if pt_range is None else pt_range
        gh = np.ascontiguousarray(ghost_pts if ghost_pts is not None else np.zeros(0, np.int64), np.int64)
        shp = np.array(shape, np.int64)
        self._ck(lib().rsvdt_set_operator_stencil(self._h, C.c_int(len(shape)), _p(shp), C.c_int(nvar),
                                                  _p(np.ascontiguousarray(blocks, c128)), _p(np.ascontiguousarray(mod_tab, np.float64)),
                                                  _p(np.ascontiguousarray(diag_tab, c128)), C.c_int64(pt0), C.c_int64(pt1),
                                                  _p(gh), C.c_int64(gh.size), C.c_int(int(with_adjoint)), C.c_int(int(DiscFlg)),
                                                  C.c_double(beta), C.c_int(int(not complex_coupling))))
        self.N = int((pt1 - pt0) * nvar)
        self.real_operator = not complex_coupling
        return self.real_operator

    def set_grid(self, k, Nw, Ns, Nt, ResRatio, dt, w, TransientRemoval=False, RealOperator=None):
        ro = self.real_operator if RealOperator is None else RealOperator
        g = Grid(k, Nw, Nw // 2 if ro else Nw, Ns, Nt, ResRatio, dt, w, int(bool(ro)), int(bool(TransientRemoval)))
        self._ck(lib().rsvdt_set_grid(self._h, C.byref(g)))
        self.grid = g
        return g

    def ReadWeightMats(self, which, nrows, ncols, rowptr, col, val):
        idx = {"W_f_sqrt_inv": 0, "B": 1, "W_q_sqrt_inv": 2, "W_q_sqrt": 3, "C": 4}[which]
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        col = np.ascontiguousarray(col, np.int64)
        val = np.ascontiguousarray(val, c128)
        self._ck(lib().rsvdt_set_weight_csr(self._h, C.c_int(idx), C.c_int64(nrows), C.c_int64(ncols), _p(rowptr),
                                            _p(col), _p(val)))

    def ReadWeightMats_local(self, which, nrows_global, ncols_global, M_rows, MH_rows):
        """row-partitioned ReadWeightMats: M_rows / MH_rows = (rowptr, col_global, val) of the rows of M / M^H this rank owns
        (PETSC_DECIDE splits of M's row and column spaces); see dist.local_weight"""
        idx = {"W_f_sqrt_inv": 0, "B": 1, "W_q_sqrt_inv": 2, "W_q_sqrt": 3, "C": 4}[which]
        a = [np.ascontiguousarray(M_rows[0], np.int64), np.ascontiguousarray(M_rows[1], np.int64), np.ascontiguousarray(M_rows[2], c128),
             np.ascontiguousarray(MH_rows[0], np.int64), np.ascontiguousarray(MH_rows[1], np.int64), np.ascontiguousarray(MH_rows[2], c128)]
        self._ck(lib().rsvdt_set_weight_csr_local(self._h, C.c_int(idx), C.c_int64(nrows_global), C.c_int64(ncols_global),
                                                  C.c_int64(a[0].size - 1), _p(a[0]), _p(a[1]), _p(a[2]),
                                                  C.c_int64(a[3].size - 1), _p(a[3]), _p(a[4]), _p(a[5])))

    # ---- Y_hat ---------------------------------------------------------------------------------------
    def set_Y_hat(self, Y):
        Y = _F(Y)
        assert Y.shape[1] == self.grid.k * self.grid.Nw_eff
        self._ck(lib().rsvdt_set_yhat(self._h, _p(Y), C.c_int64(Y.shape[0])))

    def CreateRandomMat(self, nrows=None, global_row0=0, seed=1373):
        """CreateRandomMat.c:36-39 on the device (uniform [0,1) + i[0,1)); RandSeed default of ReadUserInput.c:121-125"""
        nrows = self.N if nrows is None else nrows
        self._ck(lib().rsvdt_create_random_mat(self._h, C.c_int64(nrows), C.c_int64(global_row0), C.c_uint64(seed)))

    def get_state(self):
        """the RK4 state left by the last stepping call: N x k"""
        y = np.empty((self.N, self.grid.k), c128, order="F")
        self._ck(lib().rsvdt_get_state(self._h, _p(y), C.c_int64(self.N)))
        return y

    def get_Y_hat(self):
        rows = lib().rsvdt_yhat_rows(self._h)
        Y = np.empty((rows, self.grid.k * self.grid.Nw_eff), c128, order="F")
        self._ck(lib().rsvdt_get_yhat(self._h, _p(Y), C.c_int64(rows)))
        return Y

    # ---- the hot path ----------------------------------------------------------------------------------
    def DirectActionRK4(self):
        self._ck(lib().rsvdt_direct_action_rk4(self._h))

    def AdjointActionRK4(self):
        self._ck(lib().rsvdt_adjoint_action_rk4(self._h))

    def TSActionRK4(self, F_hat, direct=True, out=None):
        """end-to-end form with host buffers (upload, action, download)"""
        F = _F(F_hat)
        Y = out if out is not None else np.empty_like(F, order="F")
        self._ck(lib().rsvdt_ts_action_rk4(self._h, C.c_int(int(direct)), _p(F), _p(Y)))
        return Y

    def QRAllFreqs(self):
        self._ck(lib().rsvdt_qr_all_freqs(self._h))

    def PowerIterationRK4(self, q):
        self._ck(lib().rsvdt_power_iteration_rk4(self._h, C.c_int64(q)))

    def SVDAllFreqsBeforeAdjoint(self):
        self._ck(lib().rsvdt_svd_all_freqs_before_adjoint(self._h))

    def StoreU(self):
        self._ck(lib().rsvdt_store_u(self._h))

    def SVDAllFreqs(self):
        self._ck(lib().rsvdt_svd_all_freqs(self._h))

    def get_U_hat(self, rows=None):
        rows = self.N if rows is None else rows
        U = np.empty((rows, self.grid.k * self.grid.Nw_eff), c128, order="F")
        self._ck(lib().rsvdt_get_u_hat(self._h, _p(U)))
        return U

    def get_V_hat(self, rows=None):
        rows = self.N if rows is None else rows
        V = np.empty((rows, self.grid.k * self.grid.Nw_eff), c128, order="F")
        self._ck(lib().rsvdt_get_v_hat(self._h, _p(V)))
        return V

    def get_S_hat(self):
        S = np.empty((self.grid.k, self.grid.Nw_eff), np.float64, order="F")
        self._ck(lib().rsvdt_get_s_hat(self._h, _p(S)))
        return S

    def TSExpADeltaT(self, U, direct=True):
        Uc = _F(U).copy(order="F")
        self._ck(lib().rsvdt_ts_exp_a_deltat(self._h, C.c_int(int(direct)), _p(Uc), C.c_int64(Uc.shape[1])))
        return Uc

    def TransientRunRK4(self, q0, rend, TransSaveMod, TransDivVal=1e3, TransConVal=1e-8, direct=True, on_snapshot=None,
                        est_mod=0, on_est=None):
        """on_snapshot(pos, q) every TransSaveMod steps; on_est(pos, q) every est_mod steps (TransientRemovalEst.c's cadence)"""
        q = np.ascontiguousarray(q0, c128).copy()
        norms = np.zeros(rend // TransSaveMod + 1, np.float64)
        steps, reason = C.c_int64(), C.c_int()

        def wrap(fn):
            def _cb(pos, ptr, n, user):
                try:
                    fn(int(pos), np.ctypeslib.as_array(ptr, shape=(2 * n,)).view(c128).copy())
                    return 0
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return 1
            return SNAPSHOT_CB(_cb)

        cb = wrap(on_snapshot) if on_snapshot is not None else C.cast(None, SNAPSHOT_CB)
        ecb = wrap(on_est) if (on_est is not None and est_mod > 0) else C.cast(None, SNAPSHOT_CB)
        self._ck(lib().rsvdt_transient_run_rk4_est(self._h, C.c_int(int(direct)), _p(q), C.c_int64(rend),
                                                   C.c_int64(TransSaveMod), C.c_double(TransDivVal),
                                                   C.c_double(TransConVal), _p(norms), C.byref(steps), C.byref(reason),
                                                   cb, None, C.c_int64(est_mod), ecb, None))
        return q, norms, steps.value, reason.value

    # ---- test / measurement hooks -------------------------------------------------------------------------
    def keep_snapshots(self, keep=True):
        self._ck(lib().rsvdt_keep_snapshots(self._h, C.c_int(int(keep))))

    def get_snapshots(self):
        g = self.grid
        Nstore = g.Nw + 1 if g.TransientRemoval else g.Nw
        Y = np.empty((self.N, Nstore * g.k), c128, order="F")
        self._ck(lib().rsvdt_get_snapshots(self._h, _p(Y)))
        return Y

    def time_forced_steps(self, nsteps, direct=True) -> float:
        ms = C.c_float()
        self._ck(lib().rsvdt_time_forced_steps(self._h, C.c_int(int(direct)), C.c_int64(nsteps), C.byref(ms)))
        return ms.value

    def operator_info(self, direct=True, nvec=None):
        out = np.zeros(4, np.int64)
        nvec = self.grid.k if nvec is None else nvec
        self._ck(lib().rsvdt_operator_info(self._h, C.c_int(int(direct)), C.c_int(int(nvec)), _p(out)))
        return {"format": ("csr", "bcsr", "tiled-bcsr")[int(out[0])], "bytes_per_stage": int(out[1]), "stored_entries": int(out[2]),
                "rows": int(out[3])}

    def kernel_launches(self, reset=False) -> int:
        return lib().rsvdt_kernel_launches(self._h, C.c_int(int(reset)))

    def set_option(self, name, value):
        self._ck(lib().rsvdt_set_option(self._h, name.encode(), C.c_int64(value)))

    def last_kernel_ms(self, which):
        ms, n = C.c_float(), C.c_int64()
        self._ck(lib().rsvdt_last_kernel_ms(self._h, which.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value


def exported_symbols_from_header():
    """names of the functions declared in include/rsvdt.h (used by the CPU-side ABI test)"""
    import re
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rsvdt_[a-z0-9_]+)\s*\(", txt)) - {"rsvdt_snapshot_cb"})
