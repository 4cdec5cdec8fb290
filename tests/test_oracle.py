"""CPU tests of the oracle (oracle/rsvdt_oracle.c): the reference ships no golden vectors ("parity unpinned"), so the
oracle is pinned by (i) the reference's own derived workload numbers, (ii) the analytic resolvent identity of
SURVEY.md A.6, (iii) exact resolvent gains of the shipped tutorial operator A_GL and (iv) uniqueness properties."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_vars, principal_angle, rel_err


def test_time_freq_grid_tutorial_values(O):
    # variables.yaml:16-39 through SetupTimeFreqGrid.c:16-21
    dt, Ns, Nt, RR = O.setup_time_freq_grid(0.05, 42, 0.003, 10.0)
    assert dt == pytest.approx(0.0029979889813816136, rel=1e-15)
    assert (Ns, Nt, RR) == (41916, 3336, 998)
    dt, Ns, Nt, RR = O.setup_time_freq_grid(0.05, 64, 0.003, 10.0)
    assert (Ns, RR) == (41920, 655)


def test_twiddles_faithful_vs_exact(O):
    Nw, Ns = 8, 1200
    dft_f, idft_f = O.create_dft_idft_mats(Nw, Nw, Ns, False, True)
    dft_e, idft_e = O.create_dft_idft_mats(Nw, Nw, Ns, False, False)
    j = np.arange(2 * Ns)
    for iw in range(Nw):
        m = iw if iw <= Nw // 2 - 1 else iw - Nw
        exact = np.exp(2j * np.pi * ((m * j) % (2 * Ns)) / (2 * Ns)) / Ns
        assert np.abs(idft_e[iw] - exact).max() < 1e-18 * Ns + 3e-19
        assert np.abs(idft_f[iw] - exact).max() < 1e-11 / Ns * 10
    jj = np.arange(Nw)
    for iw in range(Nw):
        assert np.abs(dft_e[:, iw] - np.exp(-2j * np.pi * ((iw * jj) % Nw) / Nw)).max() < 1e-15
        assert np.abs(dft_f[:, iw] - dft_e[:, iw]).max() < 1e-13
    # real operator: Nw_eff = Nw/2, only non-negative harmonics (CreateDFTiDFTMats.c:44)
    dft_r, idft_r = O.create_dft_idft_mats(Nw, Nw // 2, Ns, True, False)
    assert dft_r.shape == (Nw, Nw // 2) and idft_r.shape == (Nw // 2, 2 * Ns)
    assert np.abs(idft_r[3] - np.exp(2j * np.pi * ((3 * j) % (2 * Ns)) / (2 * Ns)) / Ns).max() < 1e-17


def test_petsc_io_roundtrip_and_tutorial_operator(pkg, O, tmp_path, A_GL):
    io = pkg.petsc_io
    assert A_GL.nnz == 2000 and set(np.diff(A_GL.rowptr)) == {4}
    # bandwidth 3: columns r-2..r+1 in the interior
    r = np.repeat(np.arange(500), 4)
    assert (A_GL.col - r).min() == -3 and (A_GL.col - r).max() == 3
    p = str(tmp_path / "A")
    for int64 in (True, False):
        io.write_mat_aij(p, 500, 500, A_GL.rowptr, A_GL.col, A_GL.val, int64=int64)
        M, N, rp, ci, va = io.read_mat(p)
        assert (M, N) == (500, 500) and np.array_equal(rp, A_GL.rowptr) and np.array_equal(ci, A_GL.col)
        assert np.array_equal(va, A_GL.val)
    rng = np.random.default_rng(1)
    D = rng.standard_normal((7, 5)) + 1j * rng.standard_normal((7, 5))
    io.write_mat_dense(p, D)
    assert np.array_equal(io.read_mat_dense(p), D)
    v = rng.standard_normal(9) + 1j * rng.standard_normal(9)
    io.write_vec(p, v)
    assert np.array_equal(io.read_vec(p), v)


def test_exact_gains_of_tutorial_operator(A_GL):
    # dense SVD of (i w I - A)^-1 on the shipped A_GL
    A = A_GL.to_dense()
    for w, s in ((0.0, 27.71065986), (-0.65, 1672.02190959), (1.0, 3.85121426)):
        sv = np.linalg.svd(np.linalg.inv(1j * w * np.eye(500) - A), compute_uv=False)
        assert sv[0] == pytest.approx(s, rel=2e-8)


def test_hermitian_transpose_and_permutes(O):
    rng = np.random.default_rng(0)
    N = 37
    D = (rng.random((N, N)) < 0.2) * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)))
    rp = np.concatenate([[0], np.cumsum((D != 0).sum(1))])
    ci = np.concatenate([np.nonzero(D[r])[0] for r in range(N)])
    va = np.concatenate([D[r][np.nonzero(D[r])[0]] for r in range(N)])
    A = O.Csr(N, N, rp, ci, va)
    assert np.array_equal(A.to_dense(), D)
    assert np.array_equal(A.hermitian_transpose().to_dense(), D.conj().T)


def _analytic_check(O, pkg, A, beta, k, Nw, w, TL, tol):
    Ash = A.shifted(beta)
    Ad = Ash.to_dense()
    N = A.nrows
    rng = np.random.default_rng(0)
    F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
    v = make_vars(O, pkg, N, k, Nw, w, 0.003, TL)
    dft, idft = O.create_dft_idft_mats(Nw, Nw, v.Ns, False, False)
    Yd = O.ts_action_rk4(Ash, v, dft, idft, F)
    va = O.Vars(**{**v.__dict__, "DirAdj": False})
    Ya = O.ts_action_rk4(Ash.hermitian_transpose(), va, dft, idft, F)
    T = 2 * np.pi / w
    for iw in range(Nw):
        m = iw if iw <= Nw // 2 - 1 else iw - Nw
        om = m * w
        R = np.linalg.inv(1j * om * np.eye(N) - Ad)
        blk = slice(iw * k, (iw + 1) * k)
        exd = np.exp(1j * om * (v.Nt * v.dt + T / Nw)) * (R @ F[:, blk])
        exa = np.exp(-1j * om * (v.Nt * v.dt)) * (R.conj().T @ F[:, blk])
        assert rel_err(Yd[:, blk], exd) < tol
        assert rel_err(Ya[:, blk], exa) < tol


def test_action_matches_exact_resolvent(O, pkg, A_GL):
    # damped tutorial operator: transient fully decayed -> only the RK4 truncation error remains
    _analytic_check(O, pkg, A_GL, beta=1.0, k=2, Nw=6, w=0.6, TL=25.0, tol=2e-8)


def test_qr_is_the_unique_thin_qr(O):
    rng = np.random.default_rng(2)
    N, k, Nw = 300, 5, 3
    Y = np.asfortranarray(rng.standard_normal((N, k * Nw)) + 1j * rng.standard_normal((N, k * Nw)))
    Y[:, 1] = Y[:, 0] * (1 + 1e-6) + 1e-6 * Y[:, 1]  # ill-conditioned block
    Q = O.qr_all_freqs(N, k, Nw, Y)
    for iw in range(Nw):
        blk = slice(iw * k, (iw + 1) * k)
        q, r = np.linalg.qr(Y[:, blk])
        ph = np.diag(r) / np.abs(np.diag(r))
        q = q * ph  # make diag(R) positive
        assert np.abs(Q[:, blk].conj().T @ Q[:, blk] - np.eye(k)).max() < 1e-13
        assert rel_err(Q[:, blk], q) < 1e-8


def test_svd_matches_lapack(O):
    rng = np.random.default_rng(3)
    N, k, Nw = 200, 4, 2
    Y = np.asfortranarray(rng.standard_normal((N, k * Nw)) + 1j * rng.standard_normal((N, k * Nw)))
    Y[:, :k] *= np.array([100.0, 10.0, 1.0, 0.1])
    U, S = O.svd_all_freqs(N, k, Nw, Y)
    for iw in range(Nw):
        blk = slice(iw * k, (iw + 1) * k)
        u, s, _ = np.linalg.svd(Y[:, blk], full_matrices=False)
        assert np.allclose(S[:, iw], s, rtol=1e-10)
        for m in range(k):
            assert abs(abs(np.vdot(u[:, m], U[:, iw * k + m])) - 1) < 1e-8


def test_transient_run_decay_rate(O, A_GL):
    # Tutorial README: ||q|| decays at the least-stable eigenvalue rate (max Re lambda = -0.008003)
    rng = np.random.default_rng(4)
    q0 = rng.standard_normal(500) + 1j * rng.standard_normal(500)
    q0 /= np.linalg.norm(q0)
    q, norms, steps, reason = O.transient_run_rk4(A_GL, 0.003, 4000, 500, 1e3, 1e-8, q0)
    assert reason == 0 and steps == 4000 and norms[0] == pytest.approx(1.0)
    Ad = A_GL.to_dense()
    import scipy.linalg as sl
    ex = sl.expm(Ad * 0.003 * 4000) @ q0
    assert rel_err(q, ex) < 1e-8
    assert norms[-1] == pytest.approx(np.linalg.norm(ex), rel=1e-8)


def test_transient_removal_recovers_steady_state(O, pkg, A_GL):
    # tutorial settings (variables.yaml:16-39, k=1): the transient is still O(1) after TransientLength=10; the Galerkin
    # projection (TransientRemovalStrategy.c) must bring the action within ~1e-4..1e-2 of the exact resolvent
    # (SURVEY.md A.6 measured 2e-4..2e-3 with the NumPy restatement)
    k, Nw, w, TL = 1, 42, 0.05, 10.0
    N = 500
    Ad = A_GL.to_dense()
    rng = np.random.default_rng(5)
    F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
    errs = {}
    for TR in (False, True):
        v = make_vars(O, pkg, N, k, Nw, w, 0.003, TL, TR=TR)
        assert (v.Ns, v.Nt, v.ResRatio) == (41916, 3336, 998)
        dft, idft = O.create_dft_idft_mats(Nw, Nw, v.Ns, False, False)
        Y = O.ts_action_rk4(A_GL, v, dft, idft, F)
        e = []
        for iw in range(Nw):
            m = iw if iw <= Nw // 2 - 1 else iw - Nw
            om = m * w
            R = np.linalg.inv(1j * om * np.eye(N) - Ad)
            blk = slice(iw * k, (iw + 1) * k)
            ex = np.exp(1j * om * (v.Nt * v.dt + 2 * np.pi / w / Nw)) * (R @ F[:, blk])
            e.append(rel_err(Y[:, blk], ex))
        errs[TR] = (min(e), max(e))
    assert errs[False][0] > 0.05          # without removal the transient pollutes every frequency
    assert errs[True][1] < 0.05 and errs[True][0] < 1e-3
    assert errs[True][1] < 0.02 * errs[False][1]


def test_full_solve_approximates_exact_gains(O, pkg, A_GL):
    # RSVD with k=4, q=1 on the damped tutorial operator approximates the leading exact gain from below
    k, Nw, w = 4, 4, 0.65
    A = A_GL.shifted(0.5)
    v = make_vars(O, pkg, 500, k, Nw, w, 0.004, 20.0)
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=14)
    S, U, V = O.full_solve(A, v, 1, F)
    Ad = A.to_dense()
    for iw in range(Nw):
        m = iw if iw <= Nw // 2 - 1 else iw - Nw
        sv = np.linalg.svd(np.linalg.inv(1j * m * w * np.eye(500) - Ad), compute_uv=False)
        assert S[0, iw] == pytest.approx(sv[0], rel=0.1)    # randomized estimate (k=4, q=1, weak gain separation)
        assert S[0, iw] <= sv[0] * (1 + 1e-6)               # and never above the exact optimal gain


def test_oracle_reproduces_committed_golden_vectors(O, pkg, A_GL):
    # tests/golden/gl_small_golden.npz (made by tests/golden/make_golden.py): pins the oracle against silent changes
    g = np.load(os.path.join(GOLDEN, "gl_small_golden.npz"))
    k, Nw, w = int(g["k"]), int(g["Nw"]), float(g["w"])
    A = A_GL.shifted(float(g["beta"]))
    v = make_vars(O, pkg, 500, k, Nw, w, float(g["dt_in"]), float(g["TransientLength"]))
    assert np.allclose([v.dt, v.Ns, v.Nt, v.ResRatio], g["grid"], rtol=0, atol=0)
    dft, idft = O.create_dft_idft_mats(Nw, Nw, v.Ns, False, False)
    assert rel_err(O.ts_action_rk4(A, v, dft, idft, g["F_hat"]), g["Y_direct_notr"]) < 1e-13
    vt = O.Vars(**{**v.__dict__, "TransientRemoval": True, "DirAdj": False})
    assert rel_err(O.ts_action_rk4(A.hermitian_transpose(), vt, dft, idft, g["F_hat"]), g["Y_adjoint_tr"]) < 1e-11
