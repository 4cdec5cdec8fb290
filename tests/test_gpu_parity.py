"""GPU parity tests: the CUDA path (through the C ABI, include/rsvdt.h) against the CPU oracle on identical inputs.

Tolerances: the arithmetic is complex double on both sides; the two paths differ only in FMA contraction and in the
summation order of reductions, so block results agree to ~1e-13 relative.  BASELINE.json's north_star asks for gains
to <= 1e-8 relative and mode subspaces to principal angle <= 1e-6; those are asserted on the full solve."""
import numpy as np
import pytest
import torch

from conftest import make_vars, principal_angle, rel_err

pytestmark = pytest.mark.gpu

ACTION_TOL = 1e-11


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (there is no CPU fallback)")


def _handle(pkg, A, v, DiscFlg=False, beta=0.0, small=False):
    # small=False switches the single-CTA persistent kernel off so that the test operators (a few hundred rows) exercise
    # the kernels large problems use; test_small_persistent_kernel and the tutorial test cover the persistent path
    r = pkg.capi.RSVDt(0)
    real = r.LNSStructure(A.nrows, A.rowptr, A.col, A.val, DiscFlg, beta)
    assert real == v.RealOperator
    r.set_grid(v.k, v.Nw, v.Ns, v.Nt, v.ResRatio, v.dt, v.w, v.TransientRemoval)
    r.set_option("use_small", 2 if small else 0)
    return r


def _forcing(N, cols, seed):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.random((N, cols)) + 1j * rng.random((N, cols)))


def _oracle_action(O, A, v, F, direct, snaps=False):
    vv = O.Vars(**{**v.__dict__, "DirAdj": direct})
    dft, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, v.RealOperator, False)
    Op = A if direct else A.hermitian_transpose()
    return O.ts_action_rk4(Op, vv, dft, idft, F, want_snapshots=snaps)


@pytest.mark.parametrize("direct", [True, False])
def test_action_tutorial_operator(pkg, O, A_GL, direct):
    v = make_vars(O, pkg, 500, 3, 6, 0.6, 0.003, 2.0)
    F = _forcing(500, 18, 1)
    ref, snaps_ref = _oracle_action(O, A_GL, v, F, direct, snaps=True)
    r = _handle(pkg, A_GL, v)
    r.keep_snapshots(True)
    r.set_Y_hat(F)
    (r.DirectActionRK4 if direct else r.AdjointActionRK4)()
    got = r.get_Y_hat()
    assert rel_err(r.get_snapshots(), snaps_ref) < ACTION_TOL
    assert rel_err(got, ref) < ACTION_TOL
    # the end-to-end host-buffer entry point gives the same bits as the device-resident sequence
    got2 = r.TSActionRK4(F, direct=direct)
    assert np.array_equal(got, got2)
    r.close()


def test_action_faithful_twiddles_within_reference_tolerance(pkg, O, A_GL):
    # the reference evaluates exp() of un-reduced phases (CreateDFTiDFTMats.c:52-53): its tables carry ~1e-11;
    # the CUDA path reduces phases exactly.  The two must still agree far inside the 1e-8 parity budget.
    v = make_vars(O, pkg, 500, 2, 6, 0.6, 0.003, 2.0)
    F = _forcing(500, 12, 2)
    dft, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, False, True)
    ref = O.ts_action_rk4(A_GL, v, dft, idft, F)
    r = _handle(pkg, A_GL, v)
    assert rel_err(r.TSActionRK4(F), ref) < 1e-9
    r.close()


@pytest.mark.parametrize("shape,cplx,nvar,k,opts", [((9, 11), True, 5, 4, {}), ((7, 8), False, 5, 4, {}),
                                                      ((4, 5, 6), True, 5, 4, {}), ((9, 11), True, 5, 3, {}),
                                                      ((9, 11), True, 5, 4, dict(use_bcsr=0)),
                                                      ((9, 11), True, 5, 4, dict(stage_variant=1)),
                                                      ((9, 11), True, 5, 4, dict(use_bcsr=0, stage_variant=1)),
                                                      ((9, 11), True, 5, 3, dict(use_bcsr=0)),
                                                      ((6, 7), True, 4, 6, {}), ((5, 6), True, 3, 2, {}),
                                                      ((9, 11), True, 5, 4, dict(use_tiled=0)),
                                                      ((4, 5, 6), True, 5, 8, dict(use_tiled=0)),
                                                      ((4, 5, 6), True, 5, 8, {}), ((4, 5, 7), True, 5, 10, {}),
                                                      ((4, 5, 6), True, 5, 8, dict(tiled_ctas=3)),
                                                      ((4, 5, 6), True, 5, 8, dict(tiled_ctas=1)),
                                                      ((6, 7), True, 4, 6, dict(use_tiled=0)), ((5, 6), True, 3, 2, dict(use_tiled=0)),
                                                      ((5, 6), True, 3, 16, {}), ((3, 3, 3), True, 5, 12, {})])
def test_action_lns_stencil(pkg, O, shape, cplx, nvar, k, opts):
    # block-stencil operators: the TMA-staged tiled block-CSR stage kernel (default for an even number of test vectors), the
    # block-CSR stage kernel (B = 5, 4, 3; one or two test vectors per thread) and the plain CSR kernel against the oracle
    N, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=nvar, complex_coupling=cplx)
    A = O.Csr(N, N, rp, ci, va)
    Nw = 8
    v = make_vars(O, pkg, N, k, Nw, 2.0, 0.003, 0.5, real=not cplx)
    F = _forcing(N, k * v.Nw_eff, 3)
    r = _handle(pkg, A, v)
    for name, val in opts.items():
        r.set_option(name, val)
    for direct in (True, False):
        ref = _oracle_action(O, A, v, F, direct)
        assert rel_err(r.TSActionRK4(F, direct=direct), ref) < ACTION_TOL
    r.close()


def test_lns_3d_million_rows_against_oracle(pkg, O):
    # BASELINE.json configs[3] kind of operator at a bandwidth-bound size (10^6 rows, 1.3e8 stored entries, k = 8: the tiled
    # stage kernel in the variant the 2e7-DOF run uses, 6250 tiles, 64-bit offsets into the 2.5 GB of chunks): the state after
    # three forced RK4 steps against the oracle for two of the eight test vectors, every row
    shape, nvar, k, Nw = (20, 100, 100), 5, 8, 4
    N, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=nvar)
    assert N == 1_000_000
    v = make_vars(O, pkg, N, k, Nw, 2.0, 0.003, 0.25)
    rng = np.random.default_rng(3)
    F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
    _, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, v.RealOperator, False)
    ref = O.forced_state(O.Csr(N, N, rp, ci, va), v, idft, F, 3, ncols=2)
    for dev in (True, False):       # operator generated on the device / uploaded as CSR (what MatLoad feeds)
        r = pkg.capi.RSVDt(0)
        if dev:
            r.set_operator_stencil(shape, nvar, with_adjoint=False)
        else:
            r.LNSStructure(N, rp, ci, va)
        r.set_grid(v.k, v.Nw, v.Ns, v.Nt, v.ResRatio, v.dt, v.w, False)
        assert r.operator_info(True, k)["format"] == "tiled-bcsr"
        r.set_Y_hat(F)
        r.time_forced_steps(3)
        y = r.get_state()
        r.close()
        assert rel_err(y[:, :2], ref) < 1e-12
        assert np.isfinite(y).all() and np.abs(y[-5:]).max() > 0      # the last tile did its rows


def test_lns_stage_kernels_bit_identical(pkg, O):
    # CSR, block-CSR and tiled block-CSR stage kernels sum every row in the same order: identical bits, both directions,
    # on a grid whose block rows do not fill the last tile and whose boundary rows have fewer blocks (padded chunks)
    shape, k, Nw = (5, 6, 7), 8, 4
    N, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=5)
    A = O.Csr(N, N, rp, ci, va)
    v = make_vars(O, pkg, N, k, Nw, 2.0, 0.003, 0.3)
    F = _forcing(N, k * v.Nw_eff, 11)
    r = _handle(pkg, A, v)
    outs = {}
    for name, opts in (("tiled", {}), ("bcsr", dict(use_tiled=0)), ("csr", dict(use_tiled=0, use_bcsr=0))):
        for o in ("use_tiled", "use_bcsr"):
            r.set_option(o, opts.get(o, 1))
        outs[name] = [r.TSActionRK4(F, direct=d) for d in (True, False)]
    for d in (0, 1):
        assert np.array_equal(outs["tiled"][d], outs["bcsr"][d])
        assert np.array_equal(outs["tiled"][d], outs["csr"][d])
    assert rel_err(outs["tiled"][0], _oracle_action(O, A, v, F, True)) < ACTION_TOL
    r.close()


@pytest.mark.parametrize("N,k,Nw,chunk", [(5, 1, 2, 1), (33, 1, 4, 7), (257, 10, 6, 16), (64, 7, 5, 3), (100, 32, 4, 16)])
def test_action_edge_shapes(pkg, O, N, k, Nw, chunk):
    # tiny / ragged sizes, k = 1 and k = 32 (max), odd Nw, chunk lengths that do not divide the step count
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    A = O.Csr(N, N, rp, ci, va)
    v = make_vars(O, pkg, N, k, Nw, 3.0, 0.004, 0.3)
    F = _forcing(N, k * Nw, 4)
    r = _handle(pkg, A, v)
    r.set_option("chunk_steps", chunk)
    for direct in (True, False):
        assert rel_err(r.TSActionRK4(F, direct=direct), _oracle_action(O, A, v, F, direct)) < ACTION_TOL
    r.close()


def test_discounting(pkg, O, A_GL):
    v = make_vars(O, pkg, 500, 2, 4, 1.0, 0.003, 1.0)
    F = _forcing(500, 8, 5)
    ref = _oracle_action(O, A_GL.shifted(0.25), v, F, True)
    r = _handle(pkg, A_GL, v, DiscFlg=True, beta=0.25)
    assert rel_err(r.TSActionRK4(F), ref) < ACTION_TOL
    r.close()


def test_discounting_inserts_missing_diagonal_entries(pkg, O, A_GL):
    # MatShift(A, -beta) (LNSStructure.c:53-56) creates diagonal entries the file does not store: rows 7, 8 and 499 lose theirs here
    keep = np.ones(A_GL.val.size, dtype=bool)
    for r in (7, 8, 499):
        for p in range(A_GL.rowptr[r], A_GL.rowptr[r + 1]):
            if A_GL.col[p] == r:
                keep[p] = False
    rowptr = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), A_GL.rowptr[:-1]))]).astype(np.int64)
    A0 = O.Csr(500, 500, rowptr, A_GL.col[keep].copy(), A_GL.val[keep].copy())
    dense = np.zeros((500, 500), dtype=np.complex128)
    for r in range(500):
        dense[r, A0.col[A0.rowptr[r]:A0.rowptr[r + 1]]] = A0.val[A0.rowptr[r]:A0.rowptr[r + 1]]
    dense -= 0.25 * np.eye(500)
    rr, cc = np.nonzero(dense)
    rp = np.concatenate([[0], np.cumsum(np.bincount(rr, minlength=500))]).astype(np.int64)
    shifted = O.Csr(500, 500, rp, cc.astype(np.int64), dense[rr, cc])
    v = make_vars(O, pkg, 500, 2, 4, 1.0, 0.003, 1.0)
    F = _forcing(500, 8, 5)
    r = _handle(pkg, A0, v, DiscFlg=True, beta=0.25)
    for direct in (True, False):
        assert rel_err(r.TSActionRK4(F, direct=direct), _oracle_action(O, shifted, v, F, direct)) < ACTION_TOL
    r.close()


@pytest.mark.parametrize("k,opts", [(10, {}), (8, {}), (3, {}), (16, {}), (20, {}), (10, dict(qr_fused=0)),
                                    (10, dict(qr_fused=0, gram_variant=0)), (8, dict(gram_variant=0)), (32, {})])
def test_qr_all_freqs(pkg, O, k, opts):
    # CholeskyQR2 in four passes (second Gram of Y*M1 on the fly, FP64 tensor cores) / in six passes / with the FMA Gram kernel
    rng = np.random.default_rng(6)
    N, Nw = 3001, 6
    Y = np.asfortranarray(rng.standard_normal((N, k * Nw)) + 1j * rng.standard_normal((N, k * Nw)))
    Y[:, 3 % k] = Y[:, 2 % k] + 1e-4 * Y[:, 3 % k]   # cond ~ 1e4 block
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    v = make_vars(O, pkg, N, k, Nw, 3.0, 0.004, 0.3)
    r = _handle(pkg, O.Csr(N, N, rp, ci, va), v)
    for name, val in opts.items():
        r.set_option(name, val)
    r.set_Y_hat(Y)
    r.QRAllFreqs()
    Q = r.get_Y_hat()
    Qref = O.qr_all_freqs(N, k, Nw, Y)
    for iw in range(Nw):
        blk = slice(iw * k, (iw + 1) * k)
        assert np.abs(Q[:, blk].conj().T @ Q[:, blk] - np.eye(k)).max() < 1e-13
    assert rel_err(Q, Qref) < 1e-10
    r.close()


def test_svd_all_freqs(pkg, O):
    rng = np.random.default_rng(7)
    N, k, Nw = 2000, 6, 4
    Y = np.asfortranarray(rng.standard_normal((N, k * Nw)) + 1j * rng.standard_normal((N, k * Nw)))
    Y[:, :k] *= np.array([300.0, 50.0, 9.0, 2.0, 1.0, 0.3])
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    v = make_vars(O, pkg, N, k, Nw, 3.0, 0.004, 0.3)
    r = _handle(pkg, O.Csr(N, N, rp, ci, va), v)
    r.set_Y_hat(Y)
    r.SVDAllFreqsBeforeAdjoint()
    U = r.get_Y_hat()
    r.set_Y_hat(Y)
    r.SVDAllFreqs()
    V, S = r.get_V_hat(), r.get_S_hat()
    Uref, Sref = O.svd_all_freqs(N, k, Nw, Y)
    assert np.allclose(S, Sref, rtol=1e-10, atol=0)
    for c in range(k * Nw):
        # singular vectors are unique up to a unit phase
        assert abs(abs(np.vdot(Uref[:, c], U[:, c])) - 1) < 1e-8
        assert abs(abs(np.vdot(Uref[:, c], V[:, c])) - 1) < 1e-8
    r.close()


def test_exp_a_deltat_and_transient_run(pkg, O, A_GL):
    v = make_vars(O, pkg, 500, 2, 6, 0.6, 0.003, 2.0)
    r = _handle(pkg, A_GL, v)
    U = _forcing(500, 7, 8)
    for direct in (True, False):
        Op = A_GL if direct else A_GL.hermitian_transpose()
        assert rel_err(r.TSExpADeltaT(U, direct=direct), O.ts_exp_a_deltat(Op, v, U)) < 1e-12
    q0 = _forcing(500, 1, 9)[:, 0]
    q0 /= np.linalg.norm(q0)
    q, norms, steps, reason = r.TransientRunRK4(q0, 1500, 250, 1e3, 1e-8)
    qr, nr, sr, rr = O.transient_run_rk4(A_GL, v.dt, 1500, 250, 1e3, 1e-8, q0)
    assert (steps, reason) == (sr, rr) == (1500, 0)
    assert rel_err(q, qr) < 1e-12 and np.allclose(norms, nr, rtol=1e-12)
    # stop tests (TransientRunRK4.c:143-158): convergence threshold reached at the first check
    q, norms, steps, reason = r.TransientRunRK4(q0, 1500, 250, 1e3, 0.999)
    qr, nr, sr, rr = O.transient_run_rk4(A_GL, v.dt, 1500, 250, 1e3, 0.999, q0)
    assert (steps, reason) == (sr, rr) and reason == 3
    r.close()


def test_full_solve_gains_and_subspaces(pkg, O, A_GL):
    # main.c:134-146 with q = 1 on the damped tutorial operator; north_star tolerances
    k, Nw, w = 4, 4, 0.65
    A = A_GL.shifted(0.5)
    v = make_vars(O, pkg, 500, k, Nw, w, 0.004, 20.0)
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=14)
    S_ref, U_ref, V_ref = O.full_solve(A, v, 1, F)
    r = _handle(pkg, A_GL, v, DiscFlg=True, beta=0.5)
    r.set_Y_hat(F)
    r.DirectActionRK4()
    r.PowerIterationRK4(1)
    r.SVDAllFreqsBeforeAdjoint()
    r.StoreU()
    r.AdjointActionRK4()
    r.SVDAllFreqs()
    S, U, V = r.get_S_hat(), r.get_U_hat(), r.get_V_hat()
    assert np.abs(S / S_ref - 1).max() <= 1e-8
    for iw in range(Nw):
        blk = slice(iw * k, (iw + 1) * k)
        assert principal_angle(U[:, blk], U_ref[:, blk]) <= 1e-6
        assert principal_angle(V[:, blk], V_ref[:, blk]) <= 1e-6
        # leading mode individually (well separated)
        assert principal_angle(U[:, iw * k:iw * k + 1], U_ref[:, iw * k:iw * k + 1]) <= 1e-6
        assert principal_angle(V[:, iw * k:iw * k + 1], V_ref[:, iw * k:iw * k + 1]) <= 1e-6
    r.close()


def test_error_behaviour(pkg, O, A_GL):
    r = pkg.capi.RSVDt(0)
    with pytest.raises(pkg.capi.RsvdtError) as e:
        r.DirectActionRK4()                       # action before the operator is set
    assert e.value.code == 73
    r.LNSStructure(500, A_GL.rowptr, A_GL.col, A_GL.val)
    with pytest.raises(pkg.capi.RsvdtError):
        r.set_grid(33, 4, 100, 10, 25, 0.01, 1.0)  # k > 32
    with pytest.raises(pkg.capi.RsvdtError):
        bad = A_GL.col.copy()
        bad[5] = 500
        r.LNSStructure(500, A_GL.rowptr, bad, A_GL.val)
    with pytest.raises(pkg.capi.RsvdtError):
        r.LNSStructure(500, A_GL.rowptr, A_GL.col, A_GL.val, True, -1.0)  # beta must be positive (ReadUserInput.c:133)
    r.close()


def test_full_size_properties(pkg, O):
    # BASELINE.json config 2 size (N = 1e6, k = 10), short period: size-independent properties
    #   linearity of the action in F_hat, and orthonormality + span preservation of QRAllFreqs
    N, k, Nw = 1_000_000, 10, 8
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(8.0, Nw, 0.003, 0.1)
    r = pkg.capi.RSVDt(0)
    r.LNSStructure(N, rp, ci, va)
    r.set_grid(k, Nw, Ns, Nt, RR, dt, 8.0)
    F1 = pkg.synthetic.random_forcing(N, k, Nw, seed=1)
    F2 = pkg.synthetic.random_forcing(N, k, Nw, seed=2)
    a, b = 0.7 - 0.2j, -1.3 + 0.4j
    Y1 = r.TSActionRK4(F1)
    Y2 = r.TSActionRK4(F2)
    Y12 = r.TSActionRK4(a * F1 + b * F2)
    assert rel_err(Y12, a * Y1 + b * Y2) < 1e-12
    r.set_Y_hat(Y1)
    r.QRAllFreqs()
    Q = r.get_Y_hat()
    for iw in (0, Nw - 1):
        blk = slice(iw * k, (iw + 1) * k)
        G = Q[:, blk].conj().T @ Q[:, blk]
        assert np.abs(G - np.eye(k)).max() < 1e-12
        # span(Q) == span(Y): projecting Y on Q loses nothing
        P = Q[:, blk] @ (Q[:, blk].conj().T @ Y1[:, blk])
        assert rel_err(P, Y1[:, blk]) < 1e-10
    r.close()


@pytest.mark.parametrize("opts", [dict(use_banded=0, idft_variant=10), dict(use_banded=0, idft_variant=0),
                                  dict(use_banded=1, idft_variant=10), dict(use_banded=1, stage_variant=1),
                                  dict(use_banded=1, stage_variant=2), dict(use_banded=1, stage_variant=3),
                                  dict(use_banded=1, chunk_steps=64), dict(use_banded=0, idft_variant=11, chunk_steps=5),
                                  dict(use_banded=2), dict(use_banded=2, chunk_steps=7, stage_variant=2),
                                  dict(use_banded=3), dict(use_banded=3, tma_variant=1)])
def test_kernel_variants_agree_with_oracle(pkg, O, A_GL, opts):
    # every code path (stage-by-stage vs fused banded step; direct vs Taylor-moment forcing synthesis) gives the
    # oracle's answer; Ns is large enough here (w = 0.05) for the Taylor path to be selected
    k, Nw = 4, 6
    dt, Ns, Nt, RR = O.setup_time_freq_grid(0.05, Nw, 0.02, 3.0)
    v = O.Vars(N=500, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=0.05)
    F = _forcing(500, k * Nw, 11)
    A = A_GL.shifted(2.0)
    r = _handle(pkg, A, v)
    for name, val in opts.items():
        r.set_option(name, val)
    for direct in (True, False):
        ref, snaps_ref = _oracle_action(O, A, v, F, direct, snaps=True)
        r.keep_snapshots(True)
        got = r.TSActionRK4(F, direct=direct)
        assert rel_err(r.get_snapshots(), snaps_ref) < ACTION_TOL
        assert rel_err(got, ref) < ACTION_TOL
    r.close()


def test_row_partition_two_gpus():
    # the row-partitioned path (halo exchange by peer stores, both local layouts) against the single-GPU path;
    # needs 2 GPUs -- on a 1-GPU box the same check is run by tools/dist_check.py under `gpurun --gpus 2`
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("direct", [True, False])
def test_transient_removal_action(pkg, O, A_GL, direct):
    # TransientRemovalStrategy.c on the device (shifted CholeskyQR3 basis, batched Galerkin solves) against the oracle's
    # Gram-Schmidt / LU restatement; the projection depends only on span(V), so the two agree to ~cond(Y_all) * eps
    k, Nw, w = 3, 8, 0.6
    A = A_GL.shifted(0.3)
    v = make_vars(O, pkg, 500, k, Nw, w, 0.003, 1.0, TR=True)
    F = _forcing(500, k * Nw, 12)
    ref, snaps_ref = _oracle_action(O, A, v, F, direct, snaps=True)
    r = _handle(pkg, A, v)
    r.keep_snapshots(True)
    got = r.TSActionRK4(F, direct=direct)
    assert rel_err(r.get_snapshots(), snaps_ref) < ACTION_TOL
    err = rel_err(got, ref)
    print(f"transient-removal action ({'direct' if direct else 'adjoint'}): rel err vs oracle = {err:.2e}")
    assert err < 1e-8
    r.close()


def test_transient_removal_when_the_transient_has_already_decayed(pkg, O, A_GL):
    # strong discounting + a long TransientLength: the snapshots are periodic to rounding, snapshot Nw + 1 repeats snapshot 1 and
    # the snapshot matrices lose a rank.  SLEPc's Gram-Schmidt (and the oracle's) carries on; the CUDA path leaves the dependent
    # column out of its CholeskyQR basis (it used to stop with "QR breakdown") -- the projection only depends on the span
    k, Nw, w = 2, 8, 2.0
    A = A_GL.shifted(12.0)
    v = make_vars(O, pkg, 500, k, Nw, w, 0.003, 4.0, TR=True)
    F = _forcing(500, k * Nw, 12)
    ref = _oracle_action(O, A, v, F, True)
    r = _handle(pkg, A, v)
    got = r.TSActionRK4(F, direct=True)
    r.close()
    assert np.isfinite(got).all() and rel_err(got, ref) < 1e-9


def test_tutorial_config_full_solve_with_transient_removal(pkg, O, A_GL):
    # BASELINE.json configs[0]: the tutorial as shipped (variables.yaml: w = 0.05, Nw = 42, dt = 0.003, TransientLength = 10,
    # TransientRemoval = true, q = 1), with k = 3 instead of 10 to keep the oracle run short; F_hat fed to both sides
    k, Nw, w = 3, 42, 0.05
    v = make_vars(O, pkg, 500, k, Nw, w, 0.003, 10.0, TR=True)
    assert (v.Ns, v.Nt, v.ResRatio) == (41916, 3336, 998)
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=14)
    S_ref, U_ref, V_ref = O.full_solve(A_GL, v, 1, F)
    r = _handle(pkg, A_GL, v, small=True)   # the latency path this configuration is meant for
    r.set_Y_hat(F)
    r.DirectActionRK4()
    r.PowerIterationRK4(1)
    r.SVDAllFreqsBeforeAdjoint()
    r.StoreU()
    r.AdjointActionRK4()
    r.SVDAllFreqs()
    S, U, V = r.get_S_hat(), r.get_U_hat(), r.get_V_hat()
    gerr = np.abs(S / S_ref - 1)
    print("tutorial full solve: max gain rel err (leading mode) =", gerr[0].max(), " all modes =", gerr.max())
    # leading gains to the north-star tolerance; the peak of the tutorial (sigma_1 ~ 1.67e3 at w = -0.65) included
    assert gerr[0].max() <= 1e-8
    iw_peak = Nw - 13                       # -0.65 = -13 w  ->  slot Nw - 13 (FFT ordering)
    assert S[0, iw_peak] == pytest.approx(1672.0219, rel=2e-3)   # exact resolvent gain
    ang = max(principal_angle(U[:, i * k:i * k + 1], U_ref[:, i * k:i * k + 1]) for i in range(Nw))
    angv = max(principal_angle(V[:, i * k:i * k + 1], V_ref[:, i * k:i * k + 1]) for i in range(Nw))
    print("tutorial full solve: max principal angle of the leading modes U, V =", ang, angv)
    assert ang <= 1e-6 and angv <= 1e-6
    r.close()


def _rand_sparse(rng, nrows, ncols, per_row, diag=False):
    rp = [0]
    ci, va = [], []
    for r in range(nrows):
        cols = np.sort(rng.choice(ncols, size=min(per_row, ncols), replace=False))
        if diag and r < ncols and r not in cols:
            cols = np.sort(np.append(cols[:-1], r))
        ci.append(cols)
        va.append(rng.standard_normal(cols.size) + 1j * rng.standard_normal(cols.size))
        rp.append(rp[-1] + cols.size)
    return np.array(rp, np.int64), np.concatenate(ci).astype(np.int64), np.concatenate(va)


def test_weighted_actions(pkg, O, A_GL):
    # DirectActionRK4.c:24-28 / AdjointActionRK4.c:24-28 with ApplyWeightMats.c:16-80: W_q^{1/2} C R B W_f^{-1/2} and its
    # Hermitian transpose, B: N x Nb, C: Nc x N (Nb, Nc < N), plus the final W^{-1/2} scalings of SVDAllFreqs.c:63-64
    rng = np.random.default_rng(17)
    N, Nb, Nc, k, Nw = 500, 120, 90, 3, 4
    v = make_vars(O, pkg, N, k, Nw, 1.0, 0.003, 1.0)
    mats = {"B": _rand_sparse(rng, N, Nb, 3), "C": _rand_sparse(rng, Nc, N, 4),
            "W_f_sqrt_inv": _rand_sparse(rng, Nb, Nb, 2, diag=True), "W_q_sqrt": _rand_sparse(rng, Nc, Nc, 2, diag=True),
            "W_q_sqrt_inv": _rand_sparse(rng, Nc, Nc, 2, diag=True)}
    shapes = {"B": (N, Nb), "C": (Nc, N), "W_f_sqrt_inv": (Nb, Nb), "W_q_sqrt": (Nc, Nc), "W_q_sqrt_inv": (Nc, Nc)}
    Oc = {n: O.Csr(shapes[n][0], shapes[n][1], *mats[n]) for n in mats}
    r = _handle(pkg, A_GL, v)
    for n in mats:
        r.ReadWeightMats(n, shapes[n][0], shapes[n][1], *mats[n])
    dft, idft = O.create_dft_idft_mats(Nw, Nw, v.Ns, False, False)
    # direct: Y <- W_q_sqrt C R (B W_f_sqrt_inv Y)
    F = _forcing(Nb, k * Nw, 18)
    X = O.mat_mat_mult(Oc["B"], O.mat_mat_mult(Oc["W_f_sqrt_inv"], F))
    Y = O.ts_action_rk4(A_GL, v, dft, idft, X)
    ref = O.mat_mat_mult(Oc["W_q_sqrt"], O.mat_mat_mult(Oc["C"], Y))
    r.set_Y_hat(F)
    r.DirectActionRK4()
    got = r.get_Y_hat()
    assert got.shape == (Nc, k * Nw) and rel_err(got, ref) < ACTION_TOL
    # adjoint: Y <- W_f_sqrt_inv^H B^H R^H (C^H W_q_sqrt^H Y)
    G = _forcing(Nc, k * Nw, 19)
    X = O.mat_mat_mult(Oc["C"], O.mat_mat_mult(Oc["W_q_sqrt"], G, True), True)
    va = O.Vars(**{**v.__dict__, "DirAdj": False})
    Y = O.ts_action_rk4(A_GL.hermitian_transpose(), va, dft, idft, X)
    ref = O.mat_mat_mult(Oc["W_f_sqrt_inv"], O.mat_mat_mult(Oc["B"], Y, True), True)
    r.set_Y_hat(G)
    r.AdjointActionRK4()
    got = r.get_Y_hat()
    assert got.shape == (Nb, k * Nw) and rel_err(got, ref) < ACTION_TOL
    # dimension checks of ReadWeightMats.c:34-98
    with pytest.raises(pkg.capi.RsvdtError):
        r.ReadWeightMats("B", N - 1, Nb, *_rand_sparse(rng, N - 1, Nb, 2))
    with pytest.raises(pkg.capi.RsvdtError):
        r.ReadWeightMats("W_q_sqrt", Nc, Nc + 1, *_rand_sparse(rng, Nc, Nc + 1, 2))
    r.close()


@pytest.mark.parametrize("TR", [False, True])
def test_small_persistent_kernel(pkg, O, A_GL, TR):
    # K1s: one persistent CTA advances whole chunks of RK4 steps (forced, with in-kernel snapshots) and, for the transient
    # removal, all snapshot slots at once; also the unforced entry points
    k, Nw, w = 10, 8, 0.6
    A = A_GL.shifted(0.3)
    v = make_vars(O, pkg, 500, k, Nw, w, 0.003, 1.0, TR=TR)
    F = _forcing(500, k * Nw, 31)
    r = _handle(pkg, A, v, small=True)
    r.keep_snapshots(True)
    for direct in (True, False):
        ref, snaps_ref = _oracle_action(O, A, v, F, direct, snaps=True)
        n0 = r.kernel_launches(reset=True)
        got = r.TSActionRK4(F, direct=direct)
        launches = r.kernel_launches()
        assert rel_err(r.get_snapshots(), snaps_ref) < ACTION_TOL
        assert rel_err(got, ref) < (1e-8 if TR else ACTION_TOL)
        rend = v.Nt + v.Ns + (v.ResRatio if TR else 0)
        assert launches < rend / 8          # far fewer launches than RK4 steps
    U = _forcing(500, 7, 32)
    assert rel_err(r.TSExpADeltaT(U), O.ts_exp_a_deltat(A, v, U)) < 1e-12
    q0 = _forcing(500, 1, 33)[:, 0]
    q, norms, steps, reason = r.TransientRunRK4(q0, 900, 300, 1e3, 1e-12)
    qr, nr, sr, rr = O.transient_run_rk4(A, v.dt, 900, 300, 1e3, 1e-12, q0)
    assert rel_err(q, qr) < 1e-12 and np.allclose(norms, nr, rtol=1e-12)
    r.close()


@pytest.mark.parametrize("TR", [False, True])
def test_committed_golden_vectors(pkg, O, A_GL, TR):
    # the CUDA path against the committed fixture tests/golden/gl_small_golden.npz (no oracle call at test time)
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "gl_small_golden.npz"))
    k, Nw, w = int(g["k"]), int(g["Nw"]), float(g["w"])
    dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(w, Nw, float(g["dt_in"]), float(g["TransientLength"]))
    assert np.array_equal([dt, Ns, Nt, RR], g["grid"])
    tag = "tr" if TR else "notr"
    r = pkg.capi.RSVDt(0)
    r.LNSStructure(500, A_GL.rowptr, A_GL.col, A_GL.val, True, float(g["beta"]))
    r.set_grid(k, Nw, Ns, Nt, RR, dt, w, TR)
    tol = 1e-9 if TR else ACTION_TOL
    assert rel_err(r.TSActionRK4(g["F_hat"], direct=True), g[f"Y_direct_{tag}"]) < tol
    assert rel_err(r.TSActionRK4(g["F_hat"], direct=False), g[f"Y_adjoint_{tag}"]) < tol
    r.set_Y_hat(g["F_hat"])
    r.DirectActionRK4(); r.PowerIterationRK4(1); r.SVDAllFreqsBeforeAdjoint(); r.StoreU(); r.AdjointActionRK4(); r.SVDAllFreqs()
    assert np.abs(r.get_S_hat() / g[f"S_{tag}"] - 1).max() <= 1e-8
    r.close()


@pytest.mark.parametrize("opts", [dict(use_banded=3), dict(use_banded=1), dict(use_banded=0)])
def test_cuda_graph_replay_is_bit_identical(pkg, O, A_GL, opts):
    # chunks of RK4 steps without snapshot steps are replayed as CUDA graphs on launch-bound problems; same bits either way
    k, Nw = 4, 6
    dt, Ns, Nt, RR = O.setup_time_freq_grid(0.05, Nw, 0.02, 3.0)
    v = O.Vars(N=500, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=0.05)
    F = _forcing(500, k * Nw, 41)
    A = A_GL.shifted(2.0)
    res = {}
    for ug in (0, 1):
        r = _handle(pkg, A, v)
        for name, val in opts.items():
            r.set_option(name, val)
        r.set_option("use_graph", ug)
        r.set_option("chunk_steps", 16)
        r.keep_snapshots(True)
        res[ug] = (r.TSActionRK4(F, direct=True), r.TSActionRK4(F, direct=False), r.get_snapshots())
        r.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert rel_err(res[1][0], _oracle_action(O, A, v, F, True)) < ACTION_TOL


def test_qr_ill_conditioned_block_falls_back_to_shifted_cholqr3(pkg, O):
    # a block with condition number ~1e11 breaks plain CholeskyQR2; the shifted restart must still return the thin-QR factor
    rng = np.random.default_rng(51)
    N, k, Nw = 4000, 6, 3
    Y = np.asfortranarray(rng.standard_normal((N, k * Nw)) + 1j * rng.standard_normal((N, k * Nw)))
    Y[:, k + 2] = Y[:, k + 1] + 1e-11 * Y[:, k + 2]
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    v = make_vars(O, pkg, N, k, Nw, 3.0, 0.004, 0.3)
    r = _handle(pkg, O.Csr(N, N, rp, ci, va), v)
    r.set_Y_hat(Y)
    r.QRAllFreqs()
    Q = r.get_Y_hat()
    for iw in range(Nw):
        blk = slice(iw * k, (iw + 1) * k)
        assert np.abs(Q[:, blk].conj().T @ Q[:, blk] - np.eye(k)).max() < 1e-12
        # same column space: Y = Q (Q^H Y)
        P = Q[:, blk] @ (Q[:, blk].conj().T @ Y[:, blk])
        assert rel_err(P, Y[:, blk]) < 1e-10
    Qref = O.qr_all_freqs(N, k, Nw, Y)
    assert rel_err(Q[:, :k], Qref[:, :k]) < 1e-10       # well-conditioned blocks agree with Gram-Schmidt
    r.close()


def test_pipelined_forcing_synthesis(pkg, O):
    # block vectors above the launch-bound threshold: the forcing of chunk c+1 is synthesised on a low-priority stream
    # while chunk c steps; the result must be bit-identical to the serial order and equal to the oracle
    N, k, Nw = 30011, 10, 8
    n, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
    A = O.Csr(N, N, rp, ci, va)
    v = make_vars(O, pkg, N, k, Nw, 8.0, 0.003, 0.2)
    F = _forcing(N, k * Nw, 61)
    res = {}
    for pipe in (1, 0):
        r = _handle(pkg, A, v)
        r.set_option("use_pipeline", pipe)
        r.set_option("chunk_steps", 16)
        r.keep_snapshots(True)
        res[pipe] = (r.TSActionRK4(F, direct=True), r.TSActionRK4(F, direct=False), r.get_snapshots())
        r.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert rel_err(res[1][0], _oracle_action(O, A, v, F, True)) < ACTION_TOL
    assert rel_err(res[1][1], _oracle_action(O, A, v, F, False)) < ACTION_TOL
