"""The multi-GPU drop-in `rsvd-delta-t_b200/rsvdt_mpi.py` (stands for `mpiexec -n P RSVDt -inputs variables.yaml`).

CPU tests (gloo, world_size 2): the host logic -- YAML rules of ReadUserInput.c, partitioned PETSc-binary readers and the
collective writers (every rank its own PETSC_DECIDE rows), rank-count-independent random forcing, row/column splitting of
the weight matrices and the gathered-index product the CUDA library forms.
GPU test (2 GPUs): the driver under torchrun against the committed outputs of the reference executable
(tests/golden/ref_gl_executable.npz, ref_gl_weighted.npz) -- same variables.yaml, same input files."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, principal_angle, rel_err
from ref_cases import SMALL_YAML, weight_mats, write_case_files


@pytest.fixture(scope="module")
def drv(pkg):
    return importlib.import_module("rsvdt_b200.rsvdt_mpi")


# ------------------------------------------------------------------------------------------------ YAML / input rules
def _yaml(tmp_path, text):
    p = tmp_path / "variables.yaml"
    p.write_text(text)
    return ["RSVDt", "-inputs", str(p)]


BASE = """RootDir: /data/   # trailing comment
ResultsDir: 'res/'
OperatorDir: "A"
Nw: 8
w: 0.5
dt: 0.003
TransientLength: 10
"""


def test_read_user_input_defaults_and_required_keys(drv, tmp_path):
    logs = []
    v = drv.read_user_input(_yaml(tmp_path, BASE), logs.append)
    # defaults of ReadUserInput.c:32-170 and their warnings
    assert (v["k"], v["q"], v["RandSeed"], v["Display"], v["SaveResultsOpt"]) == (3, 0, 1373, 2, 1)
    assert not any(v[f] for f in ("InputForcingFlg", "TwoPI", "TransRun", "TransientRemoval", "DiscFlg"))
    assert (v["RootDir"], v["ResultsDir"], v["OperatorDir"]) == ("/data/", "res/", "A")      # comment / quotes stripped
    assert "Warning: 'k' variable not found. Setting 'k' to default value: 3" in logs
    assert "Warning: 'InvInputWeightFlg' variable not found. Setting 'InputWeightFlg' to default value: 0" in logs  # sic
    assert not any("TransSaveMod" in m for m in logs)                 # transient-run warnings only when TransRun (:146-170)
    for key, msg in (("TransientLength", "Must specify 'TransientLength'"), ("dt", "Must specify 'dt'"),
                     ("w", "Must specify 'dt'"), ("Nw", "Must specify 'Nw'"), ("RootDir", "Must specify 'RootDir'")):
        text = "\n".join(line for line in BASE.splitlines() if not line.startswith(key + ":")) + "\n"
        with pytest.raises(drv.UserError, match=msg):
            drv.read_user_input(_yaml(tmp_path, text), lambda m: None)
    with pytest.raises(drv.UserError, match="-inputs"):
        drv.read_user_input(["RSVDt"], lambda m: None)


def test_read_user_input_conditionals(drv, tmp_path):
    quiet = lambda m: None  # noqa: E731
    with pytest.raises(drv.UserError, match="Must specify 'InputForcingDir'"):
        drv.read_user_input(_yaml(tmp_path, BASE + "InputForcingFlg: true\n"), quiet)
    with pytest.raises(drv.UserError, match="Discounting flag is on"):
        drv.read_user_input(_yaml(tmp_path, BASE + "DiscFlg: true\n"), quiet)
    with pytest.raises(drv.UserError, match="'beta' must be positive"):
        drv.read_user_input(_yaml(tmp_path, BASE + "DiscFlg: true\nbeta: -1\n"), quiet)
    with pytest.raises(drv.UserError, match="Unknown logical value"):
        drv.read_user_input(_yaml(tmp_path, BASE + "TwoPI: maybe\n"), quiet)
    v = drv.read_user_input(_yaml(tmp_path, BASE + "TwoPI: yes\nTransRun: on\nTransICFlg: false\n"), quiet)
    assert v["w"] == pytest.approx(0.5 * 2 * np.pi) and v["TransRun"] and v["TransSaveMod"] == 100 and v["TransConVal"] == 1e-8
    with pytest.raises(drv.UserError, match="Must specify 'TransICDir'"):
        drv.read_user_input(_yaml(tmp_path, BASE + "TransRun: true\nTransICFlg: true\n"), quiet)


def test_row_cache_serves_a_containing_slab_from_the_block_already_read(drv, pkg):
    N, rp, ci, va = pkg.synthetic.lns_stencil((6, 7, 8), nvar=5)
    calls = []

    def fn(a, b):
        calls.append((a, b))
        return pkg.dist.slice_rows(rp, ci, va, 0, a, b)
    for a, b, c, d in ((100, 500, 40, 700), (0, 300, 0, 600), (300, N, 200, N), (100, 200, 100, 200), (100, 500, 200, 400)):
        rc = drv.RowCache(fn)
        rc.prime(a, b)
        calls.clear()
        got = rc(c, d)
        assert all(np.array_equal(x, y) for x, y in zip(got, pkg.dist.slice_rows(rp, ci, va, 0, c, d)))
        if c <= a and b <= d:
            assert rc.have is None and all(not (x < b and y > a) for x, y in calls)      # the owned block was not read again
        else:
            assert calls == [(c, d)]


def test_random_rows_do_not_depend_on_the_partition(drv):
    full = drv.random_rows(14, 0, 1000, 3, block=128)
    parts = [drv.random_rows(14, a, b, 3, block=128) for a, b in ((0, 77), (77, 640), (640, 1000))]
    assert np.array_equal(np.vstack(parts), full)
    assert 0 <= full.real.min() and full.real.max() < 1 and 0 <= full.imag.min() and full.imag.max() < 1
    assert not np.array_equal(full[:, 0], full[:, 1]) and not np.array_equal(full, drv.random_rows(15, 0, 1000, 3, block=128))


# ------------------------------------------------------------------------------------- partitioned I/O, gloo world-2
def _io_worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    pkg = entry.load_package()
    io, D = pkg.petsc_io, pkg.dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # MatLoad of a MATMPIAIJ: every rank its PETSC_DECIDE rows of the tutorial operator (64-bit file) and of a 32-bit copy
        M, N, rp, ci, va = io.read_mat(os.path.join(GOLDEN, "A_GL"))
        r0, r1 = D.split_ownership(N, world, rank)
        for path in (os.path.join(GOLDEN, "A_GL"), os.path.join(tmp, "A32")):
            a, b, c = io.MatRowReader(path).rows(r0, r1)
            assert np.array_equal(a, rp[r0:r1 + 1] - rp[r0]) and np.array_equal(b, ci[rp[r0]:rp[r1]])
            assert np.array_equal(c, va[rp[r0]:rp[r1]])
        # the operator as dist.local_operator consumes it (both layouts) == rows of the global matrix
        rd = io.MatRowReader(os.path.join(GOLDEN, "A_GL"))
        plan, A, AH, im2 = D.local_operator(N, world, rank, rd.rows, 2, 0)
        assert plan.n_own == r1 - r0 and A[0][-1] == rp[r1] - rp[r0]
        # F_hat rows (dense record) and the initial vector
        F = np.load(os.path.join(tmp, "F.npy"))
        assert np.array_equal(io.MatRowReader(os.path.join(tmp, "F_hat")).dense_rows(r0, r1), F[r0:r1])
        n, qv = io.read_vec_rows(os.path.join(tmp, "q0"), r0, r1)
        assert n == N and np.array_equal(qv, F[r0:r1, 0])
        # collective MatView / VecView: byte-identical to the one-rank writers
        w = io.DenseMatWriter(os.path.join(tmp, "F_out"), N, F.shape[1])
        if rank == 0:
            w.create()
        dist.barrier()
        w.write_rows(r0, F[r0:r1])
        vw = io.VecWriter(os.path.join(tmp, "q_out"), N)
        if rank == 0:
            vw.create()
        dist.barrier()
        vw.write_rows(r0, F[r0:r1, 0])
        dist.barrier()
        if rank == 0:
            assert open(os.path.join(tmp, "F_out"), "rb").read() == open(os.path.join(tmp, "F_hat"), "rb").read()
            assert open(os.path.join(tmp, "q_out"), "rb").read() == open(os.path.join(tmp, "q0"), "rb").read()
        # weight matrices: local rows of M and of M^H; the product the library forms on the all-gathered block
        from ref_cases import weight_mats
        mats, shapes = weight_mats(N)
        rng = np.random.default_rng(3)
        for name, (m, n_) in shapes.items():
            Mr, MHr = D.local_weight(m, n_, *mats[name], world, rank)
            X = rng.standard_normal((n_, 2)) + 1j * rng.standard_normal((n_, 2))      # same on both ranks (same seed)
            c0, c1 = D.split_ownership(n_, world, rank)
            big = -(-n_ // world)
            send = np.zeros((big, 2), complex)
            send[: c1 - c0] = X[c0:c1]
            parts = [None] * world
            dist.all_gather_object(parts, send)
            G = np.vstack(parts)                                                       # [rank][gather_rows] x k
            owner = D.owner_of(Mr[1], D.all_ranges(n_, world))
            starts = np.array([a for a, _ in D.all_ranges(n_, world)])
            idx = owner * big + (Mr[1] - starts[owner])                                # remap_gathered of csrc/capi.cu
            y = np.zeros((Mr[0].size - 1, 2), complex)
            rows = np.repeat(np.arange(Mr[0].size - 1), np.diff(Mr[0]))
            np.add.at(y, rows, Mr[2][:, None] * G[idx])
            full = np.zeros((m, n_), complex)
            frp, fci, fva = mats[name]
            full[np.repeat(np.arange(m), np.diff(frp)), fci] = fva
            a0, a1 = D.split_ownership(m, world, rank)
            assert np.allclose(y, (full @ X)[a0:a1])
            dH = np.zeros((MHr[0].size - 1, m), complex)
            dH[np.repeat(np.arange(MHr[0].size - 1), np.diff(MHr[0])), MHr[1]] = MHr[2]
            assert np.allclose(dH, full.conj().T[c0:c1])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_partitioned_io_and_weight_split_world2(pkg, tmp_path):
    io = pkg.petsc_io
    M, N, rp, ci, va = io.read_mat(os.path.join(GOLDEN, "A_GL"))
    tmp = str(tmp_path)
    io.write_mat_aij(os.path.join(tmp, "A32"), M, N, rp, ci, va, int64=False)
    rng = np.random.default_rng(1)
    F = np.asfortranarray(rng.random((N, 6)) + 1j * rng.random((N, 6)))
    np.save(os.path.join(tmp, "F.npy"), F)
    io.write_mat_dense(os.path.join(tmp, "F_hat"), F)
    io.write_vec(os.path.join(tmp, "q0"), F[:, 0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_io_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_driver_host_logic_dry_run_world2(pkg, tmp_path):
    # the executable itself under torchrun with 2 gloo ranks: inputs parsed, operator partitioned, band / layout / k decided
    # collectively -- and a loud stop where the first CUDA handle would be created (there is no CPU fallback)
    io = pkg.petsc_io
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "results"))
    F = pkg.synthetic.random_forcing(500, 3, 8, seed=3)
    write_case_files(io, tmp, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", F)})
    open(os.path.join(tmp, "variables.yaml"), "w").write(SMALL_YAML.format(
        root=tmp, Nw=8, w=0.6, TL=1.0, TR="true", opt=1, transrun="false", q=1, disc="true", weights="false"))
    env = dict(os.environ, RSVDT_MPI_HOSTONLY="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
                          os.path.join(ROOT, "rsvd-delta-t_b200", "rsvdt_mpi.py"), "-inputs", "variables.yaml"],
                         cwd=tmp, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode != 0
    assert "@ Ns              = 3496" in out.stdout and "Transient removal strategy is requested" in out.stdout
    assert "host-only dry run: N = 500, rows [0, 250) on rank 0, band 3/3, layout 1, real = False, k = 3" in out.stdout
    assert "host-only dry run stops before rsvdt_create (no CPU fallback)" in out.stderr
    assert not os.listdir(os.path.join(tmp, "results"))              # nothing was computed or written


# ------------------------------------------------------------- the whole executable on CPU ranks, oracle as the engine
def _flow_worker(rank, world, port, tmp, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), RSVDT_MPI_HOSTONLY="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import __graft_entry__ as entry
        entry.load_package()
        from oracle import oracle as O
        import fake_backend
        drv = importlib.import_module("rsvdt_b200.rsvdt_mpi")
        # the test seam lives here, not in the product: the oracle-backed stand-in replaces dist.setup_rank for this process
        drv_pkg = sys.modules["rsvdt_b200"]
        drv_pkg.dist.setup_rank = fake_backend.make_backend(O)
        os.chdir(tmp)
        if rank != 0:
            sys.stdout = open(os.devnull, "w")
        rc = drv.main(["rsvdt_mpi.py", "-inputs", "variables.yaml"])
        q.put((rank, rc))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))


def _run_flow(tmp, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() * 7 + world) % 150
    procs = [ctx.Process(target=_flow_worker, args=(r, world, port, tmp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 0 for r in res), res


def _case_dir(tmp_path, name, pkg, files, **fmt):
    tmp = str(tmp_path / name)
    os.makedirs(os.path.join(tmp, "results"))
    write_case_files(pkg.petsc_io, tmp, files)
    open(os.path.join(tmp, "variables.yaml"), "w").write(SMALL_YAML.format(root=tmp, **fmt))
    return tmp


@pytest.mark.parametrize("tag,opt,world", [("tr", 1, 3), ("notr", 2, 2)])
def test_executable_flow_on_cpu_ranks_matches_reference_files(pkg, tmp_path, tag, opt, world):
    # rsvdt_mpi.py end to end on gloo ranks (uneven PETSC_DECIDE splits for 3 ranks), the oracle standing in for the CUDA
    # library: input rules, partitioned reads, layout agreement, collective writers -> the reference executable's files
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_executable.npz"))
    k, Nw = 3, 8
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=3)
    tmp = _case_dir(tmp_path, tag, pkg, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", F)}, Nw=Nw, w=0.6, TL=1.0,
                    TR="true" if tag == "tr" else "false", opt=opt, transrun="false", q=1, disc="true", weights="false")
    _run_flow(tmp, world)
    res = os.path.join(tmp, "results", "ResolventModes_0")
    assert set(os.listdir(res)) == set(str(f) for f in g[f"small_files_{tag}"])
    S = io.read_mat_dense(f"{res}/S_hat")
    assert np.abs(S.imag).max() == 0 and np.abs(S.real / g[f"small_S_{tag}"].real - 1).max() <= 1e-10
    U, V = _read_modes(io, res, k, Nw, opt)
    assert U.shape == g[f"small_U_{tag}"].shape
    for iw in range(Nw):
        for j in range(k):
            assert principal_angle(U[:, iw, j:j + 1], g[f"small_U_{tag}"][:, iw, j:j + 1]) <= 1e-6
            assert principal_angle(V[:, iw, j:j + 1], g[f"small_V_{tag}"][:, iw, j:j + 1]) <= 1e-6


def test_executable_flow_transient_run_and_weights_on_cpu_ranks(pkg, tmp_path):
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_executable.npz"))
    tmp = _case_dir(tmp_path, "trans", pkg, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")),
                                             "F_hat": ("dense", np.zeros((500, 8), complex)), "results/q0": ("vec", g["trans_q0"])},
                    Nw=8, w=0.6, TL=1.0, TR="false", opt=1, transrun="true", q=1, disc="true", weights="false")
    _run_flow(tmp, 3)
    res = os.path.join(tmp, "results", "TransientSnapshots_0")
    assert set(os.listdir(res)) == set(str(f) for f in g["trans_files"])
    assert np.allclose(io.read_vec(f"{res}/q_transient_norms").real, g["trans_norms"], rtol=1e-12)
    assert rel_err(io.read_vec(f"{res}/q_transient_last_snapshot"), g["trans_last"]) < 1e-12
    assert rel_err(io.read_vec(f"{res}/q_transient_2"), g["trans_q2"]) < 1e-12
    # weighted solve, SaveResultsOpt 1 this time (the fixture was written with 2): B 500 x 120, C 90 x 500 over 3 ranks
    gw = np.load(os.path.join(GOLDEN, "ref_gl_weighted.npz"))
    kw, Nww = int(gw["k"]), int(gw["Nw"])
    mats, shapes = weight_mats(500)
    files = {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", gw["F"])}
    for n in mats:
        files[n] = ("csr", shapes[n], mats[n])
    tmp = _case_dir(tmp_path, "weighted", pkg, files, Nw=Nww, w=1.0, TL=1.0, TR="false", opt=1, transrun="false", q=1, disc="false",
                    weights="true")
    _run_flow(tmp, 3)
    res = os.path.join(tmp, "results", "ResolventModes_0")
    S = io.read_mat_dense(f"{res}/S_hat")
    assert np.abs(S.real / gw["S_hat"].real - 1).max() <= 1e-10
    U, V = _read_modes(io, res, kw, Nww, 1)
    assert U.shape == gw["U_hat"].shape and V.shape == gw["V_hat"].shape          # 90 x Nw x k, 120 x Nw x k
    for iw in range(Nww):
        assert principal_angle(U[:, iw, :1], gw["U_hat"][:, iw, :1]) <= 1e-6
        assert principal_angle(V[:, iw, :1], gw["V_hat"][:, iw, :1]) <= 1e-6


def test_executable_flow_transient_removal_estimate_on_cpu_ranks(pkg, tmp_path):
    # `TransRemovalEst: true` (TransientRemovalEst.c) through rsvdt_mpi.py on 3 gloo ranks: the estimate's files against the
    # reference executable's (norms summed over the ranks, orthonormal bases by TSQR over the row blocks)
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_transient_est.npz"))
    tmp = _case_dir(tmp_path, "est", pkg, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")),
                                           "F_hat": ("dense", np.zeros((500, 8), complex)), "results/q0": ("vec", g["q0"])},
                    Nw=8, w=0.6, TL=1.0, TR="false", opt=1, transrun="true", q=1, disc="true", weights="false")
    y = os.path.join(tmp, "variables.yaml")
    text = open(y).read().replace("TransPeriods:       1", "TransPeriods:       3") + "TransRemovalEst:    true\n"
    open(y, "w").write(text)
    _run_flow(tmp, 3)
    res = os.path.join(tmp, "results", "TransientSnapshots_0")
    assert set(os.listdir(res)) == set(str(f) for f in g["files"])
    assert np.allclose(io.read_vec(f"{res}/q_transient_norms").real, g["norms"], rtol=1e-12)
    for p in (1, 2, 3):
        assert np.allclose(io.read_vec(f"{res}/Initial_transient_norm_period_{p}").real, g[f"initial_{p}"], rtol=1e-10)
        assert np.allclose(io.read_vec(f"{res}/Updated_transient_norm_period_{p}").real, g[f"updated_{p}"], rtol=1e-6, atol=2e-14)


def test_executable_random_forcing_does_not_depend_on_the_rank_count(pkg, tmp_path):
    # InputForcingFlg: false -> RandSeed-seeded forcing drawn per global row block: 2 and 3 ranks give the same gains
    io = pkg.petsc_io
    S = []
    for world in (2, 3):
        tmp = _case_dir(tmp_path, f"rand{world}", pkg, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL"))}, Nw=4, w=1.0, TL=0.5,
                        TR="false", opt=2, transrun="false", q=0, disc="true", weights="false")
        y = os.path.join(tmp, "variables.yaml")
        text = open(y).read().replace("InputForcingFlg:    true", "InputForcingFlg:    false").replace("k:                  7", "k:                  2")
        open(y, "w").write(text)
        _run_flow(tmp, world)
        S.append(io.read_mat_dense(os.path.join(tmp, "results", "ResolventModes_0", "S_hat")).real)
    assert S[0].shape == (2, 4) and np.allclose(S[0], S[1], rtol=1e-11)


def test_synthetic_operator_rows_equal_the_case_file(pkg, tmp_path):
    # `OperatorDir: synthetic:lns3d:4x5x6` generates exactly the rows tools/make_case.py streams to the PETSc-binary file,
    # for arbitrary (unaligned) row ranges; and the executable produces the same gains from either source
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_case
    io = pkg.petsc_io
    case = str(tmp_path / "case")
    make_case.write_case(pkg, case, "lns3d", shape=(4, 5, 6), k=2, Nw=4, w=2.0, TL=0.5, TR=False, q=0, opt=2, forcing=True)
    rd, sy = io.MatRowReader(os.path.join(case, "A")), pkg.synthetic.SyntheticRows("lns3d:4x5x6")
    assert (rd.M, rd.N) == (sy.M, sy.N) == (600, 600)
    for a, b in ((0, 600), (0, 1), (3, 7), (197, 403), (598, 600), (250, 250)):
        for x, y in zip(rd.rows(a, b), sy.rows(a, b)):
            assert np.array_equal(x, y)
    gl = pkg.synthetic.SyntheticRows("gl:1000")
    _, rp, ci, va = pkg.synthetic.ginzburg_landau(1000)
    r = gl.rows(123, 777)
    assert np.array_equal(r[1], ci[rp[123]:rp[777]]) and np.array_equal(r[2], va[rp[123]:rp[777]])
    with pytest.raises(ValueError):
        pkg.synthetic.SyntheticRows("lns3d:4x5")
    _run_flow(case, 2)
    S_file = io.read_mat_dense(os.path.join(case, "results", "ResolventModes_0", "S_hat")).real
    syn = str(tmp_path / "syn")                 # the same case without an operator file; gains only (benchmark mode)
    make_case.write_case(pkg, syn, "lns3d", shape=(4, 5, 6), k=2, Nw=4, w=2.0, TL=0.5, TR=False, q=0, opt=2, forcing=True,
                         synthetic=True)
    assert not os.path.exists(os.path.join(syn, "A")) and "synthetic:lns3d:4x5x6" in open(os.path.join(syn, "variables.yaml")).read()
    os.environ["RSVDT_MPI_SAVE"] = "gains"
    try:
        _run_flow(syn, 3)
    finally:
        del os.environ["RSVDT_MPI_SAVE"]
    res = os.path.join(syn, "results", "ResolventModes_0")
    assert sorted(os.listdir(res)) == ["S_hat", "variables.yaml"]
    S_syn = io.read_mat_dense(os.path.join(res, "S_hat")).real
    assert S_file.shape == (2, 4) and np.allclose(S_file, S_syn, rtol=1e-11)


# ------------------------------------------------------------------------------------------------------ GPU, 2 ranks
def _torchrun(tmp, nproc=2, port=29541, layout=None):
    env = dict(os.environ)
    if layout is not None:
        env["RSVDT_LAYOUT"] = str(layout)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                           "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(ROOT, "rsvd-delta-t_b200", "rsvdt_mpi.py"), "-inputs", "variables.yaml"],
                          cwd=tmp, capture_output=True, text=True, timeout=900, env=env)


def _diag(out):
    """the part of a failed torchrun's output that matters: the first Python traceback of a rank + the driver's own log"""
    err = out.stderr
    i = err.find("Traceback")
    return "\n--- stdout tail ---\n" + out.stdout[-1500:] + "\n--- first traceback ---\n" + (err[i:i + 4000] if i >= 0 else err[-4000:])


def _read_modes(io, res, k, Nw_eff, opt):
    if opt == 1:
        U = np.stack([io.read_mat_dense(f"{res}/U_hat_k{j + 1}_allFreqs") for j in range(k)], axis=2)
        V = np.stack([io.read_mat_dense(f"{res}/V_hat_k{j + 1}_allFreqs") for j in range(k)], axis=2)
    else:
        U = np.stack([io.read_mat_dense(f"{res}/U_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
        V = np.stack([io.read_mat_dense(f"{res}/V_hat_Freq{iw}_allK") for iw in range(Nw_eff)], axis=1)
    return U, V


def _need_two_gpus():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")


@pytest.mark.gpu
@pytest.mark.parametrize("tag,opt,layout", [("tr", 1, None), ("notr", 2, 0)])
def test_mpi_driver_two_gpus_full_solve_matches_reference_executable(pkg, tmp_path, tag, opt, layout):
    # whole solves, TR on (SaveResultsOpt 1; automatic layout: the band of A_GL suggests layout 1, its 250 rows per rank do not
    # qualify for the fused band kernel, so the ranks agree on layout 0) and off (SaveResultsOpt 2, layout 0 forced)
    _need_two_gpus()
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_executable.npz"))
    k, Nw = 3, 8
    F = pkg.synthetic.random_forcing(500, k, Nw, seed=3)
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "results"))
    write_case_files(io, tmp, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", F)})
    open(os.path.join(tmp, "variables.yaml"), "w").write(SMALL_YAML.format(
        root=tmp, Nw=Nw, w=0.6, TL=1.0, TR="true" if tag == "tr" else "false", opt=opt, transrun="false", q=1,
        disc="true", weights="false"))
    out = _torchrun(tmp, layout=layout, port=29541 + opt)
    assert out.returncode == 0, _diag(out)
    assert "Based on the input forcing: k = 3" in out.stdout and "Rows partitioned over 2 GPUs" in out.stdout
    assert "local layout 0" in out.stdout
    if layout is None:
        assert "Local layout 1 (fused band kernel on halo rows) is not available" in out.stdout
    res = os.path.join(tmp, "results", "ResolventModes_0")
    assert set(os.listdir(res)) == set(str(f) for f in g[f"small_files_{tag}"])
    S = io.read_mat_dense(f"{res}/S_hat")
    assert np.abs(S.imag).max() == 0 and np.abs(S.real / g[f"small_S_{tag}"].real - 1).max() <= 1e-8
    U, V = _read_modes(io, res, k, Nw, opt)
    for iw in range(Nw):
        for j in range(k):
            assert principal_angle(U[:, iw, j:j + 1], g[f"small_U_{tag}"][:, iw, j:j + 1]) <= 1e-6
            assert principal_angle(V[:, iw, j:j + 1], g[f"small_V_{tag}"][:, iw, j:j + 1]) <= 1e-6


@pytest.mark.gpu
def test_mpi_driver_two_gpus_transient_run_matches_reference_executable(pkg, tmp_path):
    _need_two_gpus()
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_executable.npz"))
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "results"))
    write_case_files(io, tmp, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", np.zeros((500, 8), complex)),
                               "results/q0": ("vec", g["trans_q0"])})
    open(os.path.join(tmp, "variables.yaml"), "w").write(SMALL_YAML.format(
        root=tmp, Nw=8, w=0.6, TL=1.0, TR="false", opt=1, transrun="true", q=1, disc="true", weights="false"))
    out = _torchrun(tmp, port=29544)
    assert out.returncode == 0, _diag(out)
    res = os.path.join(tmp, "results", "TransientSnapshots_0")
    assert set(os.listdir(res)) == set(str(f) for f in g["trans_files"])
    assert np.allclose(io.read_vec(f"{res}/q_transient_norms").real, g["trans_norms"], rtol=1e-11)
    assert rel_err(io.read_vec(f"{res}/q_transient_last_snapshot"), g["trans_last"]) < 1e-11
    assert rel_err(io.read_vec(f"{res}/q_transient_2"), g["trans_q2"]) < 1e-11


@pytest.mark.gpu
def test_mpi_driver_two_gpus_transient_removal_estimate_matches_reference_executable(pkg, tmp_path):
    # TransRemovalEst: true through rsvdt_mpi.py on 2 GPUs: sums over the ranks, TSQR bases, against the reference executable
    _need_two_gpus()
    io = pkg.petsc_io
    g = np.load(os.path.join(GOLDEN, "ref_gl_transient_est.npz"))
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "results"))
    write_case_files(io, tmp, {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", np.zeros((500, 8), complex)),
                               "results/q0": ("vec", g["q0"])})
    text = SMALL_YAML.replace("TransPeriods:       1", "TransPeriods:       3") + "TransRemovalEst:    true\n"
    open(os.path.join(tmp, "variables.yaml"), "w").write(text.format(
        root=tmp, Nw=8, w=0.6, TL=1.0, TR="false", opt=1, transrun="true", q=1, disc="true", weights="false"))
    out = _torchrun(tmp, port=29546)
    assert out.returncode == 0, _diag(out)
    res = os.path.join(tmp, "results", "TransientSnapshots_0")
    assert set(os.listdir(res)) == set(str(f) for f in g["files"])
    assert np.allclose(io.read_vec(f"{res}/q_transient_norms").real, g["norms"], rtol=1e-11)
    for p in (1, 2, 3):
        assert np.allclose(io.read_vec(f"{res}/Initial_transient_norm_period_{p}").real, g[f"initial_{p}"], rtol=1e-9)
        assert np.allclose(io.read_vec(f"{res}/Updated_transient_norm_period_{p}").real, g[f"updated_{p}"], rtol=1e-5, atol=5e-14)


@pytest.mark.gpu
def test_mpi_driver_two_gpus_weighted_solve_matches_reference_executable(pkg, tmp_path):
    # B, C, W_f^{-1/2}, W_q^{1/2}, W_q^{-1/2} row-partitioned (rsvdt_set_weight_csr_local, all-gather per frequency block)
    _need_two_gpus()
    io = pkg.petsc_io
    gw = np.load(os.path.join(GOLDEN, "ref_gl_weighted.npz"))
    kw, Nww = int(gw["k"]), int(gw["Nw"])
    mats, shapes = weight_mats(500)
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "results"))
    files = {"A_GL": ("copy", os.path.join(GOLDEN, "A_GL")), "F_hat": ("dense", gw["F"])}
    for n in mats:
        files[n] = ("csr", shapes[n], mats[n])
    write_case_files(io, tmp, files)
    open(os.path.join(tmp, "variables.yaml"), "w").write(SMALL_YAML.format(
        root=tmp, Nw=Nww, w=1.0, TL=1.0, TR="false", opt=2, transrun="false", q=1, disc="false", weights="true"))
    out = _torchrun(tmp, port=29545)
    assert out.returncode == 0, _diag(out)
    res = os.path.join(tmp, "results", "ResolventModes_0")
    S = io.read_mat_dense(f"{res}/S_hat")
    assert np.abs(S.real / gw["S_hat"].real - 1).max() <= 1e-8
    U, V = _read_modes(io, res, kw, Nww, 2)
    assert U.shape == gw["U_hat"].shape and V.shape == gw["V_hat"].shape
    for iw in range(Nww):
        assert principal_angle(U[:, iw, :1], gw["U_hat"][:, iw, :1]) <= 1e-6
        assert principal_angle(V[:, iw, :1], gw["V_hat"][:, iw, :1]) <= 1e-6


# --------------------------------------------------------------- input rules against the reference executable itself (live)
@pytest.mark.parametrize("drop,add,expect", [
    ("TransientLength", "", "Must specify 'TransientLength'"),
    ("dt", "", "Must specify 'dt'"),
    ("w", "", "Must specify 'dt'"),                                    # sic (ReadUserInput.c:28)
    ("Nw", "", "Must specify 'Nw'"),
    ("RootDir", "", "Must specify 'RootDir'"),
    ("ResultsDir", "", "Must specify 'ResultsDir'"),
    ("OperatorDir", "", "Must specify 'OperatorDir'"),
    (None, "InputForcingFlg: true\n", "Must specify 'InputForcingDir'"),
    (None, "OutputMatrixFlg: true\n", "Must specify 'OutputMatrixDir'"),
    (None, "DiscFlg: true\n", "Discounting flag is on! Either set Discounting to zero or specify 'beta'"),
    (None, "DiscFlg: true\nbeta: -2\n", "'beta' must be positive, current value: -2"),
])
def test_input_errors_match_the_reference_executable(drv, tmp_path, drop, add, expect):
    # the same broken variables.yaml given to the reference executable (oracle/_ref/RSVDt_ref, the reference's ReadUserInput.c)
    # and to the drop-in's input reader: same complaint, non-zero exit
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    text = BASE + "TransICDir: none\n"
    if drop:
        text = "\n".join(line for line in text.splitlines() if not line.startswith(drop + ":")) + "\n"
    (tmp_path / "variables.yaml").write_text(text + add)
    out = R.run_executable(str(tmp_path))
    assert out.returncode != 0 and expect in out.stderr, out.stderr[-500:]
    with pytest.raises(drv.UserError) as ei:
        drv.read_user_input(["RSVDt", "-inputs", str(tmp_path / "variables.yaml")], lambda m: None)
    assert str(ei.value) == expect


def test_input_defaults_match_the_reference_executable(drv, tmp_path):
    # the warnings the reference prints for missing optional keys (its defaults) are the ones the drop-in prints
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    (tmp_path / "variables.yaml").write_text(BASE + "TransICDir: none\n")
    out = R.run_executable(str(tmp_path))                              # stops later, at the missing operator file
    ref_warn = [ln for ln in out.stdout.splitlines() if ln.startswith("Warning:")]
    logs = []
    drv.read_user_input(["RSVDt", "-inputs", str(tmp_path / "variables.yaml")], logs.append)
    mine = [ln for ln in logs if ln.startswith("Warning:")]
    assert len(ref_warn) >= 12 and mine == ref_warn
