#!/usr/bin/env python
"""Regenerates tests/golden/ref_*.npz: OUTPUTS OF THE REFERENCE'S OWN CODE.

The reference's SourceCode/*.c is compiled unmodified from /root/reference into oracle/_ref/ (oracle/Makefile.ref) against
the one-rank PETSc/SLEPc API shim oracle/petsc_shim/ (the real PETSc/SLEPc/MPI are not installed in this image), and run
here (a) as the executable `RSVDt_ref -inputs variables.yaml`, exactly the reference's process interface, and (b) function
by function through oracle/ref_harness.c.  /root/reference does not exist on the GPU box, so the outputs are committed
as small fixtures; tests/test_ref_golden.py checks the oracle against them on CPU and the CUDA path against them on GPU.

    python tests/golden/make_ref_golden.py            # needs /root/reference; a few minutes
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402
from ref_cases import (TUTORIAL_YAML, SMALL_YAML, csr_of, real_operator, weight_mats, write_case_files)  # noqa: E402

pkg = entry.load_package()
io = pkg.petsc_io
assert R.build(), "the reference sources are needed to regenerate the fixtures"

M, N, rp, ci, va = io.read_mat(os.path.join(HERE, "A_GL"))
A0 = O.Csr(M, N, rp, ci, va)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def read_modes(res, k, Nw_eff, opt):
    if opt == 1:   # k files N x Nw_eff  -> [N, Nw_eff, k]
        U = np.stack([io.read_mat_dense(f"{res}/U_hat_k{j + 1}_allFreqs") for j in range(k)], axis=2)
        V = np.stack([io.read_mat_dense(f"{res}/V_hat_k{j + 1}_allFreqs") for j in range(k)], axis=2)
    else:          # Nw_eff files N x k
        U = np.stack([io.read_mat_dense(f"{res}/U_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
        V = np.stack([io.read_mat_dense(f"{res}/V_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
    return U, V


def run_exe(yaml_text, files, autoinit=False, **fmt):
    tmp = tempfile.mkdtemp(prefix="refgold_")
    os.makedirs(os.path.join(tmp, "results"))
    write_case_files(io, tmp, files)
    with open(os.path.join(tmp, "variables.yaml"), "w") as f:
        f.write(yaml_text.format(root=tmp, **fmt))
    out = R.run_executable(tmp, autoinit=autoinit)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return tmp, out.stdout


# ---------------------------------------------------------------------------------------------- 1. small GL, functions
def small_functions():
    k, Nw, w, beta = 3, 6, 0.6, 0.3
    A = A0.shifted(beta)
    out = {"k": k, "Nw": Nw, "w": w, "beta": beta, "dt_in": 0.003, "TransientLength": 1.0}
    rng = np.random.default_rng(2026)
    F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
    out["F_hat"] = F
    dt, Ns, Nt, RR = R.setup_time_freq_grid(w, Nw, 0.003, 1.0)
    out["grid"] = np.array([dt, Ns, Nt, RR])
    for TR in (False, True):
        tag = "tr" if TR else "notr"
        for direct in (True, False):
            v = O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w, TransientRemoval=TR, DirAdj=direct)
            out[f"Y_{'direct' if direct else 'adjoint'}_{tag}"] = R.ts_action_rk4(A, v, F)
    Y = out["Y_direct_notr"]
    out["Q"] = R.qr_all_freqs(N, k, Nw, Y)
    out["U_svd"], out["S_svd"] = R.svd_all_freqs(N, k, Nw, Y)
    v = O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w)
    out["expAdT"] = R.ts_exp_a_deltat(A, v, F[:, :4])
    out["q_trans_300"] = R.trans_steps(A, v, 300, F[:, 0])
    dftm, idftm = R.create_dft_idft_mats(v)
    out["dft"] = dftm
    out["idft_cols"] = idftm[:, ::997].copy()          # a sample of the 2Ns columns (the whole table is 2 MB)
    out["idft_sha256"] = sha(idftm)
    # real operator: Nw_eff = Nw/2 (LNSStructure.c:40-44), other twiddle tables (CreateDFTiDFTMats.c:35-36,61)
    Ar = real_operator(O, A0)
    kr, Nwr = 2, 8
    Fr = np.asfortranarray(rng.random((N, kr * Nwr // 2)) + 1j * rng.random((N, kr * Nwr // 2)))
    dt, Ns, Nt, RR = R.setup_time_freq_grid(0.8, Nwr, 0.003, 1.0)
    out["real_F_hat"] = Fr
    out["real_grid"] = np.array([dt, Ns, Nt, RR])
    for TR in (False, True):
        for direct in (True, False):
            v = O.Vars(N=N, k=kr, Nw=Nwr, Nw_eff=Nwr // 2, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=0.8, RealOperator=True,
                       TransientRemoval=TR, DirAdj=direct)
            out[f"real_Y_{'direct' if direct else 'adjoint'}_{'tr' if TR else 'notr'}"] = R.ts_action_rk4(Ar, v, Fr)
    # edge shapes: k = 1, Nw = 2 (the smallest grid), odd Nw
    for (k_, Nw_) in ((1, 2), (2, 5)):
        dt, Ns, Nt, RR = R.setup_time_freq_grid(1.1, Nw_, 0.003, 0.5)
        Fe = np.asfortranarray(rng.random((N, k_ * Nw_)) + 1j * rng.random((N, k_ * Nw_)))
        v = O.Vars(N=N, k=k_, Nw=Nw_, Nw_eff=Nw_, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=1.1)
        out[f"edge_{k_}_{Nw_}_F"] = Fe
        out[f"edge_{k_}_{Nw_}_grid"] = np.array([dt, Ns, Nt, RR])
        out[f"edge_{k_}_{Nw_}_Y"] = R.ts_action_rk4(A, v, Fe)
    np.savez_compressed(os.path.join(HERE, "ref_gl_functions.npz"), **out)
    print("written ref_gl_functions.npz")


# ---------------------------------------------------------------------------------------------- 2. weighted actions
def weighted():
    k, Nw = 3, 4
    mats, shapes = weight_mats(N)
    A = A0
    dt, Ns, Nt, RR = R.setup_time_freq_grid(1.0, Nw, 0.003, 1.0)
    rng = np.random.default_rng(18)
    Nb, Nc = shapes["B"][1], shapes["C"][0]
    F = np.asfortranarray(rng.random((Nb, k * Nw)) + 1j * rng.random((Nb, k * Nw)))
    G = np.asfortranarray(rng.random((Nc, k * Nw)) + 1j * rng.random((Nc, k * Nw)))
    Oc = {n: csr_of(O, shapes[n], mats[n]) for n in mats}
    out = {"k": k, "Nw": Nw, "w": 1.0, "F": F, "G": G, "grid": np.array([dt, Ns, Nt, RR])}
    v = O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=1.0, DirAdj=True)
    out["Y_direct"] = R.weighted_action(A, v, F, B=Oc["B"], Cm=Oc["C"], Wf_inv=Oc["W_f_sqrt_inv"], Wq=Oc["W_q_sqrt"])
    v = O.Vars(**{**v.__dict__, "DirAdj": False})
    out["Y_adjoint"] = R.weighted_action(A, v, G, B=Oc["B"], Cm=Oc["C"], Wf_inv=Oc["W_f_sqrt_inv"], Wq=Oc["W_q_sqrt"])
    # the whole weighted solve through the executable (ReadWeightMats.c, ApplyWeightMats.c, SVDAllFreqs.c:63-64)
    files = {"A_GL": ("copy", os.path.join(HERE, "A_GL")), "F_hat": ("dense", F)}
    for n in mats:
        files[n] = ("csr", shapes[n], mats[n])
    tmp, _ = run_exe(SMALL_YAML, files, Nw=Nw, w=1.0, TL=1.0, TR="false", opt=2, transrun="false", q=1, disc="false",
                     weights="true")
    res = os.path.join(tmp, "results", "ResolventModes_0")
    out["S_hat"] = io.read_mat_dense(f"{res}/S_hat")
    out["U_hat"], out["V_hat"] = read_modes(res, k, Nw, 2)
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_gl_weighted.npz"), **out)
    print("written ref_gl_weighted.npz")


# ---------------------------------------------------------------------------------------------- 3. executable runs
def executable():
    out = {}
    k, Nw, w, TL = 3, 8, 0.6, 1.0
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=3)
    out["small_F_sha256"] = sha(F)
    for TR, opt in (("true", 1), ("false", 2)):
        files = {"A_GL": ("copy", os.path.join(HERE, "A_GL")), "F_hat": ("dense", F)}
        tmp, stdout = run_exe(SMALL_YAML, files, Nw=Nw, w=w, TL=TL, TR=TR, opt=opt, transrun="false", q=1, disc="true",
                              weights="false")
        res = os.path.join(tmp, "results", "ResolventModes_0")
        tag = "tr" if TR == "true" else "notr"
        out[f"small_S_{tag}"] = io.read_mat_dense(f"{res}/S_hat")
        out[f"small_U_{tag}"], out[f"small_V_{tag}"] = read_modes(res, k, Nw, opt)
        out[f"small_files_{tag}"] = np.array(sorted(os.listdir(res)))
        out[f"small_stdout_{tag}"] = np.array(stdout)
        shutil.rmtree(tmp)
    # transient run (TransientRunRK4.c): norms, snapshots, last snapshot
    rng = np.random.default_rng(8)
    q0 = rng.standard_normal(500) + 1j * rng.standard_normal(500)
    q0 /= np.linalg.norm(q0)
    files = {"A_GL": ("copy", os.path.join(HERE, "A_GL")), "F_hat": ("dense", np.zeros((500, 8), complex)),
             "results/q0": ("vec", q0)}
    # main.c never sets RSVDt.TS.DirAdj before the transient run (uninitialised stack variable, TSTransRK4.c:15 branches
    # on it): "direct" is the run with the flag reading true (build with -ftrivial-auto-var-init=pattern), "plain" is
    # whatever the plain build's stack held (zero here, i.e. the A^H system).  Both are kept; the drop-in integrates A.
    out["trans_q0"] = q0
    for tag, autoinit in (("", True), ("plain_", False)):
        tmp, stdout = run_exe(SMALL_YAML, files, autoinit=autoinit, Nw=8, w=0.6, TL=1.0, TR="false", opt=1,
                              transrun="true", q=1, disc="true", weights="false")
        res = os.path.join(tmp, "results", "TransientSnapshots_0")
        out[f"trans_{tag}norms"] = io.read_vec(f"{res}/q_transient_norms").real
        out[f"trans_{tag}last"] = io.read_vec(f"{res}/q_transient_last_snapshot")
        out[f"trans_{tag}q2"] = io.read_vec(f"{res}/q_transient_2")
        out[f"trans_{tag}files"] = np.array(sorted(os.listdir(res)))
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_gl_executable.npz"), **out)
    print("written ref_gl_executable.npz")


# ---------------------------------------------------------------------------------------------- 4. the tutorial as shipped
def tutorial():
    """BASELINE.json configs[0]: variables.yaml as shipped (k=10, q=1, Nw=42, w=0.05, dt=0.003, TransientLength=10,
    TransientRemoval on, SaveResultsOpt=1) on Tutorial/Ginzburg-Landau/A_GL; only InputForcingFlg is switched on so that
    both sides read the same F_hat (CreateRandomMat.c ignores RandSeed)."""
    k, Nw = 10, 42
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=14)
    files = {"A_GL": ("copy", os.path.join(HERE, "A_GL")), "F_hat": ("dense", F)}
    tmp, stdout = run_exe(TUTORIAL_YAML, files)
    res = os.path.join(tmp, "results", "ResolventModes_0")
    U, V = read_modes(res, k, Nw, 1)
    out = {"F_sha256": sha(F), "S_hat": io.read_mat_dense(f"{res}/S_hat"),
           "U_lead": U[:, :, 0].copy(), "V_lead": V[:, :, 0].copy(), "U_second": U[:, :, 1].copy(),
           "stdout_head": np.array(stdout[:3000])}
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_gl_tutorial.npz"), **out)
    print("written ref_gl_tutorial.npz")


# ---------------------------------------------------------------------------------------------- 5. LNS-type operators
def lns():
    """Block-stencil operators of the BASELINE.json configs[2..4] kind (5 variables per grid point, 9- / 27-point stencil, dense
    5x5 coupling blocks; the operators the block-CSR stage kernel is for), small enough for the reference code: actions in
    both directions, with and without transient removal, and a whole solve through the executable for the 3-D one."""
    out = {}
    for tag, shape in (("2d", (9, 11)), ("3d", (4, 5, 6))):
        n, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=5, complex_coupling=True)
        A = O.Csr(n, n, rp, ci, va)
        k, Nw, w = 4, 8, 2.0
        dt, Ns, Nt, RR = R.setup_time_freq_grid(w, Nw, 0.003, 0.5)
        F = pkg.synthetic.random_forcing(n, k, Nw, seed=3)
        out[f"{tag}_shape"] = np.array(shape)
        out[f"{tag}_val_sha256"] = sha(va)
        out[f"{tag}_F_sha256"] = sha(F)
        out[f"{tag}_grid"] = np.array([dt, Ns, Nt, RR])
        for TR in (False, True):
            for direct in (True, False):
                v = O.Vars(N=n, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w, TransientRemoval=TR, DirAdj=direct)
                out[f"{tag}_Y_{'direct' if direct else 'adjoint'}_{'tr' if TR else 'notr'}"] = R.ts_action_rk4(A, v, F)
        if tag == "3d":
            tmp = tempfile.mkdtemp(prefix="refgold_")
            os.makedirs(os.path.join(tmp, "results"))
            io.write_mat_aij(os.path.join(tmp, "A_GL"), n, n, rp, ci, va)          # SMALL_YAML names the operator file A_GL
            io.write_mat_dense(os.path.join(tmp, "F_hat"), F)
            with open(os.path.join(tmp, "variables.yaml"), "w") as f:
                f.write(SMALL_YAML.format(root=tmp, Nw=Nw, w=w, TL=0.5, TR="true", opt=2, transrun="false", q=1, disc="false",
                                          weights="false"))
            res = R.run_executable(tmp)
            assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
            rd = os.path.join(tmp, "results", "ResolventModes_0")
            out["3d_S_hat"] = io.read_mat_dense(f"{rd}/S_hat")
            U, V = read_modes(rd, k, Nw, 2)
            out["3d_U_lead"], out["3d_V_lead"] = U[:, :, 0].copy(), V[:, :, 0].copy()
            shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_lns_functions.npz"), **out)
    print("written ref_lns_functions.npz")


# ---------------------------------------------------------------------------------------------- 6. a generated case directory
def case():
    """tools/make_case.py (config 5 in miniature: 3-D block stencil, selection-type B and C, diagonal weights, transient
    removal, q = 1) run by the reference executable exactly as written to disk."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_case
    tmp = tempfile.mkdtemp(prefix="refgold_")
    info = make_case.write_case(pkg, tmp, "lns3d", shape=(4, 5, 6), k=3, Nw=4, w=2.0, TL=0.5, TR=True, q=1, opt=2, weights=True,
                                forcing=True)
    res = R.run_executable(tmp)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    rd = os.path.join(tmp, "results", "ResolventModes_0")
    U, V = read_modes(rd, 3, 4, 2)
    out = {"S_hat": io.read_mat_dense(f"{rd}/S_hat"), "U_hat": U, "V_hat": V, "Nb": info["Nb"], "Nc": info["Nc"],
           "A_sha256": hashlib.sha256(open(os.path.join(tmp, "A"), "rb").read()).hexdigest()}
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_lns_case.npz"), **out)
    print("written ref_lns_case.npz")


# ---------------------------------------------------------------------------------------------- 7. TransientRemovalEst.c
def transient_est():
    """the transient run with `TransRemovalEst: true` (TransientRunRK4.c:56-65,96-114,174-184 + TransientRemovalEst.c) through the
    reference executable: three periods on the tutorial operator, given initial vector; norms per period and the banner lines"""
    import re
    rng = np.random.default_rng(8)
    q0 = rng.standard_normal(500) + 1j * rng.standard_normal(500)
    q0 /= np.linalg.norm(q0)
    files = {"A_GL": ("copy", os.path.join(HERE, "A_GL")), "F_hat": ("dense", np.zeros((500, 8), complex)), "results/q0": ("vec", q0)}
    text = SMALL_YAML.replace("TransPeriods:       1", "TransPeriods:       3") + "TransRemovalEst:    true\n"
    assert "TransPeriods:       3" in text
    tmp, stdout = run_exe(text, files, autoinit=True, Nw=8, w=0.6, TL=1.0, TR="false", opt=1, transrun="true", q=1, disc="true",
                          weights="false")
    res = os.path.join(tmp, "results", "TransientSnapshots_0")
    out = {"q0": q0, "Nw": 8, "w": 0.6, "periods": 3, "files": np.array(sorted(os.listdir(res))),
           "norms": io.read_vec(f"{res}/q_transient_norms").real, "last": io.read_vec(f"{res}/q_transient_last_snapshot"),
           "stdout": np.array([ln for ln in stdout.splitlines() if ln.startswith("****") or ln.startswith("Norm of")])}
    for p in (1, 2, 3):
        out[f"initial_{p}"] = io.read_vec(f"{res}/Initial_transient_norm_period_{p}").real
        out[f"updated_{p}"] = io.read_vec(f"{res}/Updated_transient_norm_period_{p}").real
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_gl_transient_est.npz"), **out)
    print("written ref_gl_transient_est.npz")


if __name__ == "__main__":
    which = sys.argv[1:] or ["functions", "weighted", "executable", "tutorial", "lns", "case", "transient_est"]
    if "transient_est" in which:
        transient_est()
    if "functions" in which:
        small_functions()
    if "weighted" in which:
        weighted()
    if "executable" in which:
        executable()
    if "tutorial" in which:
        tutorial()
    if "lns" in which:
        lns()
    if "case" in which:
        case()
