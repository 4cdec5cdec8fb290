"""tools/make_case.py: BASELINE.json configurations written as case directories (PETSc-binary A, B, C, W_*, F_hat and a
variables.yaml) that the reference executable and both drop-in executables run unchanged.  The committed fixture
tests/golden/ref_lns_case.npz is what the reference executable (oracle/_ref/RSVDt_ref) produced for the miniature config-5
case below: 3-D block stencil, selection-type B and C with N_b, N_c < N, diagonal weights, transient removal, q = 1."""
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, principal_angle

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_case  # noqa: E402

EXE = os.path.join(ROOT, "rsvd-delta-t_b200", "RSVDt")
ARGS = dict(shape=(4, 5, 6), k=3, Nw=4, w=2.0, TL=0.5, TR=True, q=1, opt=2, weights=True, forcing=True)


def _read_modes(io, res, k, Nw_eff):
    U = np.stack([io.read_mat_dense(f"{res}/U_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
    V = np.stack([io.read_mat_dense(f"{res}/V_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
    return U, V


def test_case_files_are_what_the_generators_define(pkg, O, tmp_path):
    io = pkg.petsc_io
    tmp = str(tmp_path)
    info = make_case.write_case(pkg, tmp, "lns3d", chunk_points=37, **ARGS)        # several row chunks through the stream writer
    g = np.load(os.path.join(GOLDEN, "ref_lns_case.npz"))
    assert hashlib.sha256(open(os.path.join(tmp, "A"), "rb").read()).hexdigest() == str(g["A_sha256"])
    n, rp, ci, va = pkg.synthetic.lns_stencil((4, 5, 6), nvar=5, complex_coupling=True)
    M, N, rp2, ci2, va2 = io.read_mat(os.path.join(tmp, "A"))
    assert (M, N) == (n, n) == (600, 600) and np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(va, va2)
    B, Cm = io.read_mat_dense(os.path.join(tmp, "B")), io.read_mat_dense(os.path.join(tmp, "C"))
    assert B.shape == (600, info["Nb"]) and Cm.shape == (info["Nc"], 600) and (info["Nb"], info["Nc"]) == (int(g["Nb"]), int(g["Nc"]))
    assert np.array_equal(B.conj().T @ B, np.eye(info["Nb"])) and np.array_equal(Cm @ Cm.conj().T, np.eye(info["Nc"]))   # selections
    Wq, Wqi = io.read_mat_dense(os.path.join(tmp, "W_q_sqrt")), io.read_mat_dense(os.path.join(tmp, "W_q_sqrt_inv"))
    assert np.allclose(Wq @ Wqi, np.eye(info["Nc"]), atol=1e-15) and (np.diag(Wq).real > 0).all()
    assert io.read_mat_dense(os.path.join(tmp, "F_hat")).shape == (info["Nb"], 3 * 4)
    drv = importlib.import_module("rsvdt_b200.rsvdt_mpi")                          # the YAML obeys the reference's input rules
    v = drv.read_user_input(["RSVDt", "-inputs", os.path.join(tmp, "variables.yaml")], lambda m: None)
    assert v["TransientRemoval"] and v["InputMatrixFlg"] and v["OutputWeightFlg"] and v["OperatorDir"] == "A" and v["q"] == 1
    # a Ginzburg-Landau case streams the same rows synthetic.ginzburg_landau gives
    gl = str(tmp_path / "gl")
    make_case.write_case(pkg, gl, "gl", n=3000, k=2, Nw=4)
    M, N, rp2, ci2, va2 = io.read_mat(os.path.join(gl, "A"))
    _, rp, ci, va = pkg.synthetic.ginzburg_landau(3000)
    assert M == 3000 and np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(va, va2)


def test_oracle_weighted_solve_matches_reference_executable_on_the_case(pkg, O, tmp_path):
    # main.c:134-146 with ApplyWeightMats.c on both sides of every action and SVDAllFreqs.c:63-64 at the end, restated with the
    # oracle's pieces, against what the reference executable wrote for the same files
    io = pkg.petsc_io
    tmp = str(tmp_path)
    make_case.write_case(pkg, tmp, "lns3d", **ARGS)
    g = np.load(os.path.join(GOLDEN, "ref_lns_case.npz"))
    rd = lambda name: O.Csr(*io.read_mat(os.path.join(tmp, name)))  # noqa: E731
    A, B, Cm, Wf, Wq, Wqi = rd("A"), rd("B"), rd("C"), rd("W_f_sqrt_inv"), rd("W_q_sqrt"), rd("W_q_sqrt_inv")
    F = io.read_mat_dense(os.path.join(tmp, "F_hat"))
    k, Nw = 3, 4
    dt, Ns, Nt, RR = O.setup_time_freq_grid(2.0, Nw, 0.003, 0.5)
    v = O.Vars(N=600, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=2.0, TransientRemoval=True)
    dft, idft = O.create_dft_idft_mats(Nw, Nw, Ns, False, True)
    AH = A.hermitian_transpose()
    mm = O.mat_mat_mult

    def direct(Y):
        Y = O.ts_action_rk4(A, v, dft, idft, mm(B, mm(Wf, Y)))
        return mm(Wq, mm(Cm, Y))

    def adjoint(Y):
        va = O.Vars(**{**v.__dict__, "DirAdj": False})
        Y = O.ts_action_rk4(AH, va, dft, idft, mm(Cm, mm(Wq, Y, True), True))
        return mm(Wf, mm(B, Y, True), True)

    Nb, Nc = B.ncols, Cm.nrows
    Y = direct(F)
    Y = O.qr_all_freqs(Nc, k, Nw, Y)
    Y = adjoint(Y)
    Y = O.qr_all_freqs(Nb, k, Nw, Y)
    Y = direct(Y)
    U, _ = O.svd_all_freqs(Nc, k, Nw, Y)
    V, S = O.svd_all_freqs(Nb, k, Nw, adjoint(U))
    U, V = mm(Wqi, U), mm(Wf, V)
    assert np.abs(S / g["S_hat"].real - 1).max() < 1e-9
    for iw in range(Nw):
        assert principal_angle(g["U_hat"][:, iw, :1], U[:, iw * k:iw * k + 1]) <= 1e-6
        assert principal_angle(g["V_hat"][:, iw, :1], V[:, iw * k:iw * k + 1]) <= 1e-6


@pytest.mark.gpu
def test_driver_runs_the_case_like_the_reference_executable(pkg, tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    io = pkg.petsc_io
    tmp = str(tmp_path)
    make_case.write_case(pkg, tmp, "lns3d", **ARGS)
    g = np.load(os.path.join(GOLDEN, "ref_lns_case.npz"))
    out = subprocess.run([EXE, "-inputs", "variables.yaml"], cwd=tmp, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = os.path.join(tmp, "results", "ResolventModes_0")
    S = io.read_mat_dense(f"{res}/S_hat")
    assert S.shape == (3, 4) and np.abs(S.real / g["S_hat"].real - 1).max() <= 1e-8
    U, V = _read_modes(io, res, 3, 4)
    assert U.shape == g["U_hat"].shape and V.shape == g["V_hat"].shape
    for iw in range(4):
        assert principal_angle(U[:, iw, :1], g["U_hat"][:, iw, :1]) <= 1e-6
        assert principal_angle(V[:, iw, :1], g["V_hat"][:, iw, :1]) <= 1e-6
