"""Self-tests of the PETSc/SLEPc API shim (oracle/petsc_shim/) that stands under the reference's own sources in oracle/_ref/.

The fixtures tests/golden/ref_*.npz are only as good as the primitives the shim restates, so each is checked here against
an independent implementation: numpy / LAPACK for the algebra (MatMult*, MatMatMult, MatTransposeMatMult, in-place
MatHermitianTranspose, Vec*AXPY, KSPSolve, BVOrthogonalize through QRDecomposition.c, SVDSolve through SVDAllFreqs.c) and
rsvd-delta-t_b200/petsc_io.py for the PETSc binary records and the YAML options reader.  Skipped when oracle/_ref was not
built (it needs /root/reference at build time; the built files travel to the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

c128 = np.complex128


@pytest.fixture(scope="module")
def R():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    return ref


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rand_csr(rng, m, n, per_row):
    rp, ci, va = [0], [], []
    for _ in range(m):
        cols = np.sort(rng.choice(n, size=min(per_row, n), replace=False))
        ci.append(cols)
        va.append(rng.standard_normal(cols.size) + 1j * rng.standard_normal(cols.size))
        rp.append(rp[-1] + cols.size)
    rp, ci, va = np.array(rp, np.int64), np.concatenate(ci).astype(np.int64), np.concatenate(va).astype(c128)
    D = np.zeros((m, n), c128)
    D[np.repeat(np.arange(m), np.diff(rp)), ci] = va
    return rp, ci, va, D


def _cv(rng, *shape):
    return np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(c128)


@pytest.mark.parametrize("dense", [0, 1])
def test_mat_mult_family(R, dense):
    rng = np.random.default_rng(0)
    m, n = 37, 23
    rp, ci, va, D = _rand_csr(rng, m, n, 5)
    L = R.lib()
    for mode, (A, nx, ny) in enumerate(((D, n, m), (D, n, m), (D.conj().T, m, n), (D.conj().T, m, n))):
        x, add, y = _cv(rng, nx), _cv(rng, ny), np.empty(ny, c128)
        assert L.shim_mat_mult(C.c_int64(m), C.c_int64(n), _p(rp), _p(ci), _p(va), C.c_int(dense), C.c_int(mode), _p(x),
                               _p(add) if mode % 2 else None, _p(y)) == 0
        want = A @ x + (add if mode % 2 else 0)
        assert np.allclose(y, want, rtol=1e-13, atol=1e-13), mode


def test_mat_mat_products_and_views(R):
    rng = np.random.default_rng(1)
    m, n, c = 19, 11, 4
    rp, ci, va, D = _rand_csr(rng, m, n, 3)
    L = R.lib()
    B = _cv(rng, n, c)
    for op in (0, 1):   # MAT_INITIAL_MATRIX, and MAT_REUSE_MATRIX into a MatDenseGetSubMatrix view (DFT.c:29)
        out = np.empty((m, c), c128, order="F")
        assert L.shim_mat_mat(C.c_int(op), C.c_int64(m), C.c_int64(n), _p(rp), _p(ci), _p(va), C.c_int64(c), _p(B), _p(out)) == 0
        assert np.allclose(out, D @ B, rtol=1e-13, atol=1e-13)
    B2 = _cv(rng, m, c)
    out = np.empty((n, c), c128, order="F")
    assert L.shim_mat_mat(C.c_int(2), C.c_int64(m), C.c_int64(n), _p(rp), _p(ci), _p(va), C.c_int64(c), _p(B2), _p(out)) == 0
    assert np.allclose(out, D.T @ B2, rtol=1e-13, atol=1e-13)          # MatTransposeMatMult: plain transpose (CreateMtilde.c:51)
    assert L.shim_mat_mat(C.c_int(3), C.c_int64(m), C.c_int64(n), _p(rp), _p(ci), _p(va), C.c_int64(c), _p(B2), _p(out)) == 0
    assert np.allclose(out, D.conj().T @ B2, rtol=1e-13, atol=1e-13)   # in-place MatHermitianTranspose (ApplyWeightMats.c:44)


def test_vec_ops_and_ksp(R):
    rng = np.random.default_rng(2)
    n = 101
    x, y = _cv(rng, n), _cv(rng, n)
    a, b, g = 0.3 - 0.2j, -1.1 + 0.4j, 0.7 + 0.9j
    y2, w, nrm = y.copy(), np.empty(n, c128), C.c_double()
    assert R.lib().shim_vec_ops(C.c_int64(n), _p(x), _p(y2), _p(w), C.c_double(a.real), C.c_double(a.imag), C.c_double(b.real),
                                C.c_double(b.imag), C.c_double(g.real), C.c_double(g.imag), C.byref(nrm)) == 0
    w1 = a * x + y              # VecWAXPY
    yy = y + b * w1             # VecAXPY
    w2 = x + g * w1             # VecAYPX
    assert np.allclose(y2, yy, rtol=1e-14) and np.allclose(w, w2, rtol=1e-14) and nrm.value == pytest.approx(np.linalg.norm(w2), rel=1e-14)
    for nn in (1, 7, 43):        # GalerkinProj.c:56-62: (Nw+1)^2 systems, solved in place
        A = np.asfortranarray(_cv(rng, nn, nn) + 3 * np.eye(nn))
        rhs = _cv(rng, nn)
        sol = rhs.copy()
        assert R.lib().shim_ksp_solve(C.c_int64(nn), _p(A), _p(sol)) == 0
        assert np.allclose(sol, np.linalg.solve(A, rhs), rtol=1e-11)
    A = np.asfortranarray(np.zeros((3, 3), c128))
    assert R.lib().shim_ksp_solve(C.c_int64(3), _p(A), _p(_cv(rng, 3))) != 0      # singular -> error code


@pytest.mark.parametrize("cond", [1e1, 1e7, 1e11])
def test_bv_orthogonalize_is_the_unique_thin_qr(R, cond):
    # QRDecomposition.c (BVCreateFromMat / BVOrthogonalize / BVCreateMat) against LAPACK, also for the ill-conditioned
    # snapshot matrices of the transient-removal step (cond ~ 1e7)
    rng = np.random.default_rng(3)
    n, k = 200, 9
    U, _ = np.linalg.qr(_cv(rng, n, k))
    V, _ = np.linalg.qr(_cv(rng, k, k))
    A = (U * np.logspace(0, -np.log10(cond), k)) @ V.conj().T
    Q = R.qr_decomposition(A)
    assert np.abs(Q.conj().T @ Q - np.eye(k)).max() < 1e-13 * max(1.0, cond * 1e-9)
    Rm = Q.conj().T @ A
    assert np.abs(np.tril(Rm, -1)).max() < 1e-12 and (np.diag(Rm).real > 0).all() and np.abs(np.diag(Rm).imag).max() < 1e-13
    Ql, Rl = np.linalg.qr(A)
    ph = np.diag(Rl) / np.abs(np.diag(Rl))
    assert np.abs(Q - Ql * ph).max() < 1e-15 * cond + 1e-12            # same factor as LAPACK up to the sign convention


def test_svd_solve_matches_lapack(R):
    rng = np.random.default_rng(4)
    n, k, Nw = 150, 5, 3
    Y = _cv(rng, n, k * Nw)
    U, S = R.svd_all_freqs(n, k, Nw, Y)
    for iw in range(Nw):
        blk = Y[:, iw * k:(iw + 1) * k]
        u, s, vh = np.linalg.svd(blk, full_matrices=False)
        assert np.allclose(S[:, iw], s, rtol=1e-13)                     # sorted, sigma = sqrt(eig(Y^H Y))
        for j in range(k):
            ip = np.vdot(u[:, j], U[:, iw * k + j])
            assert abs(abs(ip) - 1) < 1e-12                             # same left vectors up to a unit phase
        Ub = U[:, iw * k:(iw + 1) * k]
        assert np.abs(Ub.conj().T @ Ub - np.eye(k)).max() < 1e-13


def test_binary_records_roundtrip_against_petsc_io(R, pkg, tmp_path):
    io = pkg.petsc_io
    rng = np.random.default_rng(5)
    m, n = 31, 17
    rp, ci, va, D = _rand_csr(rng, m, n, 4)
    L = R.lib()
    for int64 in (True, False):                                         # both index widths are detected on input
        src = str(tmp_path / f"A{int(int64)}")
        io.write_mat_aij(src, m, n, rp, ci, va, int64=int64)
        dst = str(tmp_path / f"A{int(int64)}_out")
        assert L.shim_mat_roundtrip(src.encode(), dst.encode(), C.c_int(0), C.c_double(0.0)) == 0
        M2, N2, rp2, ci2, va2 = io.read_mat(dst)
        assert (M2, N2) == (m, n) and np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(va2, va)
        dstd = str(tmp_path / f"A{int(int64)}_dense")                   # MatLoad into MATDENSE, MatView(dense): every entry
        assert L.shim_mat_roundtrip(src.encode(), dstd.encode(), C.c_int(1), C.c_double(0.0)) == 0
        assert np.array_equal(io.read_mat_dense(dstd), D) and io.read_mat(dstd)[2][-1] == m * n
    Fd = _cv(rng, 9, 6)
    io.write_mat_dense(str(tmp_path / "F"), Fd)                          # what CreateRandomMat.c:23-31 loads
    assert L.shim_mat_roundtrip(str(tmp_path / "F").encode(), str(tmp_path / "F_out").encode(), C.c_int(1), C.c_double(0.0)) == 0
    assert open(tmp_path / "F", "rb").read() == open(tmp_path / "F_out", "rb").read()
    # MatShift on an AIJ matrix whose diagonal is partly missing (LNSStructure.c:54)
    sq_rp, sq_ci, sq_va, Dsq = _rand_csr(rng, 12, 12, 3)
    io.write_mat_aij(str(tmp_path / "S"), 12, 12, sq_rp, sq_ci, sq_va)
    assert L.shim_mat_roundtrip(str(tmp_path / "S").encode(), str(tmp_path / "S_out").encode(), C.c_int(0), C.c_double(-0.25)) == 0
    assert np.allclose(io.read_mat_dense(str(tmp_path / "S_out")), Dsq - 0.25 * np.eye(12))
    v = _cv(rng, 13)
    io.write_vec(str(tmp_path / "v"), v)
    assert L.shim_vec_roundtrip(str(tmp_path / "v").encode(), str(tmp_path / "v_out").encode()) == 0
    assert open(tmp_path / "v", "rb").read() == open(tmp_path / "v_out", "rb").read()
    # the tutorial operator as shipped (written by MATLAB's PetscBinaryWrite, trailing records ignored like MatLoad does)
    a_gl = os.path.join(os.path.dirname(__file__), "golden", "A_GL")
    assert L.shim_mat_roundtrip(a_gl.encode(), str(tmp_path / "A_GL_out").encode(), C.c_int(0), C.c_double(0.0)) == 0
    A1, A2 = io.read_mat(a_gl), io.read_mat(str(tmp_path / "A_GL_out"))
    assert all(np.array_equal(a, b) for a, b in zip(A1, A2))


def test_yaml_options_reader(R, tmp_path):
    y = tmp_path / "variables.yaml"
    y.write_text("# comment line\nRootDir:   /a/b/   # trailing comment\nk:  10\nw: 0.05\nTwoPI: false\nTransientRemoval:   true\n"
                 "TransDivVal:        1e3\nquoted: 'x y'\nInputForcingDir:    /path/to/F_hat # expected a matrix\n")
    L = R.lib()

    def look(key, typ):
        s, buf = C.c_int(), C.create_string_buffer(256)
        assert L.shim_option_lookup(str(y).encode(), key.encode(), C.c_int(typ), C.byref(s), buf, C.c_int64(256)) == 0
        return bool(s.value), buf.value.decode()

    assert look("-RootDir", 0) == (True, "/a/b/") and look("-InputForcingDir", 0) == (True, "/path/to/F_hat")
    assert look("-quoted", 0) == (True, "x y") and look("-missing", 0)[0] is False
    assert look("-k", 2) == (True, "10") and look("-TwoPI", 1) == (True, "0") and look("-TransientRemoval", 1) == (True, "1")
    assert float(look("-w", 3)[1]) == 0.05 and float(look("-TransDivVal", 3)[1]) == 1000.0
