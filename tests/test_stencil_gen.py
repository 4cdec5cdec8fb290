"""The device-side generator of the synthetic block-stencil operators (rsvdt_set_operator_stencil) against the host generator
synthetic.lns_stencil: host logic on CPU (tables, ghost points = what dist.plan_general finds from the actual columns),
entries and actions on the GPU."""
import itertools

import numpy as np
import pytest
import torch

from conftest import make_vars, rel_err


def _rebuild_from_tables(S, shape, nvar, cc):
    """lns_stencil's entries recomputed the way the device kernel does: blocks[b][v][u] * mod[pt % 997], diag[pt % 499]"""
    blocks, mod_tab, diag_tab = S.stencil_tables(len(shape), nvar, cc)
    d = len(shape)
    npts = int(np.prod(shape))
    offsets = list(itertools.product((-1, 0, 1), repeat=d))
    strides = [int(np.prod(shape[i + 1:])) for i in range(d)]
    centre = offsets.index((0,) * d)
    cols, vals, rp = [], [], [0]
    for q in range(npts):
        c, rem = [], q
        for s in reversed(shape):
            c.append(rem % s)
            rem //= s
        c = c[::-1]
        nb = [(b, q + sum(o * st for o, st in zip(off, strides))) for b, off in enumerate(offsets)
              if all(0 <= ci + oi < si for ci, oi, si in zip(c, off, shape))]
        for v in range(nvar):
            for b, p in nb:
                for u in range(nvar):
                    cols.append(p * nvar + u)
                    vals.append(diag_tab[q % 499] if (b == centre and u == v) else blocks[b, v, u] * mod_tab[q % 997])
            rp.append(len(cols))
    return np.array(rp), np.array(cols), np.array(vals)


@pytest.mark.parametrize("shape,nvar,cc", [((4, 5, 6), 5, True), ((7, 8), 4, True), ((5, 4), 3, False)])
def test_tables_reproduce_the_host_generator(pkg, shape, nvar, cc):
    S = pkg.synthetic
    N, rp, ci, va = S.lns_stencil(shape, nvar=nvar, complex_coupling=cc)
    rp2, ci2, va2 = _rebuild_from_tables(S, shape, nvar, cc)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    assert np.array_equal(va, va2)          # bit for bit


@pytest.mark.parametrize("shape,nvar,world", [((8, 5, 6), 5, 2), ((9, 5, 4), 5, 3), ((12, 7), 4, 4), ((16, 3, 3), 5, 8)])
def test_stencil_plan_equals_the_plan_from_the_columns(pkg, shape, nvar, world):
    S, D = pkg.synthetic, pkg.dist
    N = int(np.prod(shape)) * nvar
    for rank in range(world):
        r0, r1 = D.split_ownership(N, world, rank)
        if r0 % nvar or r1 % nvar:
            with pytest.raises(ValueError):
                D.plan_stencil(shape, nvar, world, rank)
            continue
        _, rp, cg, va = S.lns_stencil(shape, nvar=nvar, row_range=(r0, r1))
        ref = D.plan_general(N, world, rank, cg)
        got = D.plan_stencil(shape, nvar, world, rank)
        assert got.n_ext == ref.n_ext and np.array_equal(got.ghost_global, ref.ghost_global)
        assert sorted(got.requests) == sorted(ref.requests)
        for q in ref.requests:
            assert np.array_equal(got.requests[q][0], ref.requests[q][0]) and got.requests[q][1] == ref.requests[q][1]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nvar,k,cc", [((5, 6, 7), 5, 8, True), ((9, 11), 5, 10, True), ((6, 7), 4, 6, True), ((5, 6), 3, 2, False)])
def test_device_generated_operator_is_the_host_operator(pkg, O, shape, nvar, k, cc):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    N, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=nvar, complex_coupling=cc)
    v = make_vars(O, pkg, N, k, 4, 2.0, 0.003, 0.3, real=not cc)
    rng = np.random.default_rng(5)
    F = np.asfortranarray(rng.random((N, k * v.Nw_eff)) + 1j * rng.random((N, k * v.Nw_eff)))
    outs = []
    for dev in (False, True):
        r = pkg.capi.RSVDt(0)
        if dev:
            assert r.set_operator_stencil(shape, nvar, complex_coupling=cc) == (not cc)
        else:
            r.LNSStructure(N, rp, ci, va)
        r.set_grid(v.k, v.Nw, v.Ns, v.Nt, v.ResRatio, v.dt, v.w, False)
        r.set_option("use_small", 0)
        outs.append([r.TSActionRK4(F, direct=d) for d in (True, False)])
        r.close()
    for d in (0, 1):
        assert np.array_equal(outs[0][d], outs[1][d])      # same entries, same kernel: same bits (direct and adjoint)
    A = O.Csr(N, N, rp, ci, va)
    vv = O.Vars(**{**v.__dict__, "DirAdj": True})
    dft, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, v.RealOperator, False)
    assert rel_err(outs[1][0], O.ts_action_rk4(A, vv, dft, idft, F)) < 1e-11


@pytest.mark.gpu
def test_device_random_mat_and_state_readback(pkg, O):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    shape, nvar, k = (4, 5, 6), 5, 4
    N = int(np.prod(shape)) * nvar
    v = make_vars(O, pkg, N, k, 4, 2.0, 0.003, 0.3)
    r = pkg.capi.RSVDt(0)
    r.set_operator_stencil(shape, nvar)
    r.set_grid(v.k, v.Nw, v.Ns, v.Nt, v.ResRatio, v.dt, v.w, False)
    r.CreateRandomMat(seed=7)
    Y = r.get_Y_hat()
    assert Y.shape == (N, k * v.Nw_eff) and 0.0 <= Y.real.min() and Y.real.max() < 1.0 and 0.0 <= Y.imag.min() and Y.imag.max() < 1.0
    assert abs(Y.real.mean() - 0.5) < 0.02 and abs(Y.imag.mean() - 0.5) < 0.02 and np.unique(Y).size == Y.size
    r.CreateRandomMat(seed=7)
    assert np.array_equal(Y, r.get_Y_hat())
    r.CreateRandomMat(seed=8)
    assert not np.array_equal(Y, r.get_Y_hat())
    # the state after n forced steps equals the oracle's (one test vector at a time, TSActionRK4.c:116-122)
    r.set_Y_hat(Y)
    n = 6
    r.time_forced_steps(n)
    y = r.get_state()
    _, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=nvar)
    A = O.Csr(N, N, rp, ci, va)
    _, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, False, False)
    ref = O.forced_state(A, v, idft, Y, n)
    assert rel_err(y, ref) < 1e-12
    r.close()
