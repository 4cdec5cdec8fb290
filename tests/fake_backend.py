"""A CPU stand-in for the row-partitioned CUDA handle, used ONLY by tests/test_mpi_driver.py to drive the host logic of
rsvd-delta-t_b200/rsvdt_mpi.py (input rules, partitioned reads, collective result writers, weight splitting) with gloo ranks.
Every rank gathers the row blocks it is handed, lets the CPU oracle compute the global result (test infrastructure, allowed
here) and keeps its own PETSC_DECIDE rows -- so what the driver writes to disk can be compared with the reference
executable's files.  It mirrors the methods of capi.RSVDt that the driver calls; it is not a fallback of the product."""
from types import SimpleNamespace

import numpy as np
import torch.distributed as dist


def _gather(x):
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, x)
    return parts


class FakeHandle:
    def __init__(self, pkg, O, N, gen_rows, grid, DiscFlg, beta):
        self.pkg, self.O, self.D = pkg, O, pkg.dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.Nglob = N
        r0, r1 = self.D.split_ownership(N, self.world, self.rank)
        rp, cg, va = gen_rows(r0, r1)
        parts = _gather((np.asarray(rp), np.asarray(cg), np.asarray(va)))
        rowptr = np.concatenate([[0]] + [p[0][1:] + sum(int(q[0][-1]) for q in parts[:i]) for i, p in enumerate(parts)])
        A = O.Csr(N, N, rowptr, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
        self.real_operator = bool(np.sqrt(np.sum(A.val.imag ** 2)) < 1e-12)
        self.A = A.shifted(beta) if DiscFlg else A
        self.AH = self.A.hermitian_transpose()
        self.N = r1 - r0
        self.W, self.WH = {}, {}
        self.closed = False
        self.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"],
                      grid.get("TransientRemoval", False))

    # ---- what rsvdt_mpi.py calls ------------------------------------------------------------------------------------
    def dist_layout_supported(self, nvec):
        return self.rank != self.world - 1          # one rank says no: the ranks must agree on layout 0 together

    def set_grid(self, k, Nw, Ns, Nt, RR, dt, w, TR=False):
        Nw_eff = Nw // 2 if self.real_operator else Nw
        self.v = self.O.Vars(N=self.Nglob, k=k, Nw=Nw, Nw_eff=Nw_eff, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w,
                             RealOperator=self.real_operator, TransientRemoval=bool(TR))
        self.grid = SimpleNamespace(k=k, Nw=Nw, Nw_eff=Nw_eff)
        self.dft, self.idft = self.O.create_dft_idft_mats(Nw, Nw_eff, Ns, self.real_operator, True)

    def ReadWeightMats_local(self, which, M, N, Mr, MHr):
        def assemble(rows_parts, nrows, ncols):
            rowptr = np.concatenate([[0]] + [np.asarray(p[0])[1:] + sum(int(q[0][-1]) for q in rows_parts[:i])
                                             for i, p in enumerate(rows_parts)])
            return self.O.Csr(nrows, ncols, rowptr, np.concatenate([p[1] for p in rows_parts]),
                              np.concatenate([p[2] for p in rows_parts]))
        own_r = self.D.split_ownership(M, self.world, self.rank)
        own_c = self.D.split_ownership(N, self.world, self.rank)
        assert len(Mr[0]) - 1 == own_r[1] - own_r[0] and len(MHr[0]) - 1 == own_c[1] - own_c[0]   # the C ABI's check
        Mg = assemble(_gather(tuple(np.asarray(a) for a in Mr)), M, N)
        MHg = assemble(_gather(tuple(np.asarray(a) for a in MHr)), N, M)
        assert np.allclose(MHg.to_dense(), Mg.to_dense().conj().T)       # the rows of M^H really are those of the transpose
        self.W[which], self.WH[which] = Mg, MHg

    def _local(self, Yg):
        r0, r1 = self.D.split_ownership(Yg.shape[0], self.world, self.rank)
        return np.asfortranarray(Yg[r0:r1])

    def set_Y_hat(self, Yl):
        self.Y = np.asfortranarray(np.vstack(_gather(np.ascontiguousarray(Yl))))

    def _act(self, direct):
        O, mm, W, WH = self.O, self.O.mat_mat_mult, self.W, self.W
        v = O.Vars(**{**self.v.__dict__, "DirAdj": direct})
        Y = self.Y
        if direct:                                                       # ApplyWeightMats.c:18-41
            if "W_f_sqrt_inv" in W:
                Y = mm(W["W_f_sqrt_inv"], Y)
            if "B" in W:
                Y = mm(W["B"], Y)
            Y = O.ts_action_rk4(self.A, v, self.dft, self.idft, Y)
            if "C" in W:
                Y = mm(W["C"], Y)
            if "W_q_sqrt" in W:
                Y = mm(W["W_q_sqrt"], Y)
        else:                                                            # :43-78
            if "W_q_sqrt" in W:
                Y = mm(W["W_q_sqrt"], Y, True)
            if "C" in W:
                Y = mm(W["C"], Y, True)
            Y = O.ts_action_rk4(self.AH, v, self.dft, self.idft, Y)
            if "B" in W:
                Y = mm(W["B"], Y, True)
            if "W_f_sqrt_inv" in W:
                Y = mm(W["W_f_sqrt_inv"], Y, True)
        self.Y = Y

    def DirectActionRK4(self):
        self._act(True)

    def AdjointActionRK4(self):
        self._act(False)

    def QRAllFreqs(self):
        self.Y = self.O.qr_all_freqs(self.Y.shape[0], self.v.k, self.v.Nw_eff, self.Y)

    def SVDAllFreqsBeforeAdjoint(self):
        self.Y, _ = self.O.svd_all_freqs(self.Y.shape[0], self.v.k, self.v.Nw_eff, self.Y)

    def StoreU(self):
        self.U = self.Y.copy()

    def SVDAllFreqs(self):
        self.V, self.S = self.O.svd_all_freqs(self.Y.shape[0], self.v.k, self.v.Nw_eff, self.Y)
        if "W_q_sqrt_inv" in self.W:                                     # SVDAllFreqs.c:63-64
            self.U = self.O.mat_mat_mult(self.W["W_q_sqrt_inv"], self.U)
        if "W_f_sqrt_inv" in self.W:
            self.V = self.O.mat_mat_mult(self.W["W_f_sqrt_inv"], self.V)

    def get_S_hat(self):
        return self.S

    def get_U_hat(self, rows=None):
        out = self._local(self.U)
        assert rows is None or rows == out.shape[0]
        return out

    def get_V_hat(self, rows=None):
        out = self._local(self.V)
        assert rows is None or rows == out.shape[0]
        return out

    def TransientRunRK4(self, q0, rend, mod, div, con, direct=True, on_snapshot=None, est_mod=0, on_est=None):
        q = np.concatenate(_gather(np.asarray(q0)))
        Op = self.A if direct else self.AH
        norms = np.zeros(rend // mod + 1)
        norms[0] = np.linalg.norm(q)
        i, reason = 0, 0
        while i < rend and not reason:
            stop = min(rend, (i // mod + 1) * mod)
            if est_mod > 0 and on_est is not None:
                stop = min(stop, (i // est_mod + 1) * est_mod)
            q, _, _, _ = self.O.transient_run_rk4(Op, self.v.dt, stop - i, 10 ** 9, 1e300, 0.0, q)
            i = stop
            if i % mod == 0:
                nr = np.linalg.norm(q)
                p = i // mod
                reason = 1 if np.isnan(nr) else (2 if nr > div else (3 if nr < con else 0))
                if reason:
                    break
                if on_snapshot is not None:
                    on_snapshot(p, self._local(q[:, None])[:, 0])
                norms[p] = nr
            if est_mod > 0 and on_est is not None and i % est_mod == 0:
                on_est(i // est_mod, self._local(q[:, None])[:, 0])
        return self._local(q[:, None])[:, 0], norms, i, reason

    def close(self):
        self.closed = True


def make_backend(O):
    """-> a function with the signature of dist.setup_rank"""
    def setup_rank(pkg, device, N, gen_rows, max_bw, layout, band, grid, group=None, DiscFlg=False, beta=0.0, with_adjoint=True):
        r0, r1 = pkg.dist.split_ownership(N, dist.get_world_size(), dist.get_rank())
        plan = SimpleNamespace(r0=r0, r1=r1, n_own=r1 - r0, layout=layout)
        return FakeHandle(pkg, O, N, gen_rows, grid, DiscFlg, beta), plan
    return setup_rank
