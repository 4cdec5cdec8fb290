"""CPU tests (gloo, world_size 2) of the row-partition planner: the halo plans of both local layouts reproduce the
global sparse mat-vec / a global 4-stage band recomputation when the ghost rows are exchanged as planned."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dense_rows(rp, ci, va, ncols):
    M = np.zeros((rp.size - 1, ncols), complex)
    rows = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    np.add.at(M, (rows, ci), va)
    return M


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    D = pkg.dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- layout 0 on a 2-D block-stencil operator (45 nnz/row) ---------------------------------------------------
        shape = (6, 7)
        N, rp, ci, va = pkg.synthetic.lns_stencil(shape, nvar=5)
        Ad = _dense_rows(rp, ci, va, N)
        gen = lambda a, b: pkg.synthetic.lns_stencil(shape, nvar=5, row_range=(a - a % 5, b + (-b) % 5))  # noqa: E731

        def gen_rows(a, b):  # arbitrary row windows from the 5-aligned generator
            a5, b5 = a - a % 5, b + (-b) % 5
            _, rpp, cc, vv = pkg.synthetic.lns_stencil(shape, nvar=5, row_range=(a5, b5))
            s, e = rpp[a - a5], rpp[b - a5]
            return rpp[a - a5: b - a5 + 1] - s, cc[s:e], vv[s:e]

        bw = int(np.abs(ci - np.repeat(np.arange(N), np.diff(rp))).max())
        plan, A, AH, im2 = D.local_operator(N, world, rank, gen_rows, bw, 0)
        assert (plan.r0, plan.r1) == pkg.capi.plan_partition(N, world, rank) == D.split_ownership(N, world, rank)
        allreq = [None] * world
        dist.all_gather_object(allreq, plan.requests)
        out, inc = D.build_messages(plan, allreq)
        rng = np.random.default_rng(5)
        xg = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        xl = np.zeros(plan.n_ext, complex)
        xl[: plan.n_own] = xg[plan.r0: plan.r1]
        # emulate the exchange: every rank publishes the rows it sends, receivers place them at dest_row0
        sent = [None] * world
        dist.all_gather_object(sent, {p: (xl[rows], d0) for p, rows, d0 in out})
        for qrank, cnt, d0 in inc:
            vals, dd = sent[qrank][rank]
            assert dd == d0 and len(vals) == cnt
            xl[d0: d0 + cnt] = vals
        for (lrp, lci, lva), ref in ((A, Ad), (AH, Ad.conj().T)):
            y = _dense_rows(lrp, lci, lva, plan.n_ext) @ xl
            assert np.allclose(y, (ref @ xg)[plan.r0: plan.r1], rtol=1e-13, atol=1e-13)
        # ---- layout 1 on the Ginzburg-Landau band (lo = 2, hi = 1) ---------------------------------------------------
        N = 211
        _, rp, ci, va = pkg.synthetic.ginzburg_landau(N)
        Ad = _dense_rows(rp, ci, va, N)
        gl = lambda a, b: pkg.synthetic.ginzburg_landau(N, row_range=(a, b))[1:]  # noqa: E731
        plan, A, AH, _ = D.local_operator(N, world, rank, gl, 3, 1, band=(2, 1))
        assert plan.own0 == plan.r0 - plan.ext0 and plan.n_ext <= plan.n_own + 16
        allreq = [None] * world
        dist.all_gather_object(allreq, plan.requests)
        out, inc = D.build_messages(plan, allreq)
        yg = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        yl = np.zeros(plan.n_ext, complex)
        yl[plan.own0: plan.own0 + plan.n_own] = yg[plan.r0: plan.r1]
        sent = [None] * world
        dist.all_gather_object(sent, {p: (yl[rows], d0) for p, rows, d0 in out})
        for qrank, cnt, d0 in inc:
            vals, dd = sent[qrank][rank]
            yl[d0: d0 + cnt] = vals
        assert np.array_equal(yl, yg[plan.ext0: plan.ext0 + plan.n_ext])
        # four nested band products on the halo-extended rows reproduce A^4 y on the owned rows
        for (lrp, lci, lva), ref in ((A, Ad), (AH, Ad.conj().T)):
            Ml = _dense_rows(lrp, lci, lva, plan.n_ext)
            z = yl.copy()
            for _ in range(4):
                z = Ml @ z
            want = np.linalg.matrix_power(ref, 4) @ yg
            own = slice(plan.own0, plan.own0 + plan.n_own)
            assert np.allclose(z[own], want[plan.r0: plan.r1], rtol=1e-12, atol=1e-9 * np.abs(want).max())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise e
    finally:
        dist.destroy_process_group()


def test_partition_plans_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_split_ownership_matches_petsc_decide():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    D = entry.load_package().dist
    for N, P in ((10, 3), (500, 8), (7, 7), (20_000_000, 8), (5, 8)):
        rs = D.all_ranges(N, P)
        assert rs[0][0] == 0 and rs[-1][1] == N
        for p in range(P):
            assert rs[p][1] - rs[p][0] == N // P + (1 if p < N % P else 0)
            if p:
                assert rs[p][0] == rs[p - 1][1]


@pytest.mark.parametrize("nranks", [1, 2, 3, 5])
def test_local_operator_rows_equal_the_global_matrix(nranks):
    # dist.local_operator (one pass over the row source: owned rows sliced out of the slab that also yields the rows of A^H by a
    # stable sort) against the global matrices, both local layouts, every rank of 1 / 2 / 3 / 5
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    D = pkg.dist
    calls = []
    for name, src, bw, band in (("lns", pkg.synthetic.SyntheticRows("lns2d:6x7"), None, None),
                                ("gl", pkg.synthetic.SyntheticRows("gl:211"), 2, (2, 1))):
        N = src.M
        rp, ci, va = src.rows(0, N)
        Ad = _dense_rows(rp, ci, va, N)
        AHd = Ad.conj().T
        if bw is None:
            bw = int(np.abs(ci - np.repeat(np.arange(N), np.diff(rp))).max())

        def gen(a, b):
            calls.append((a, b))
            return src.rows(a, b)

        for rank in range(nranks):
            calls.clear()
            plan, A, AH, im2 = D.local_operator(N, nranks, rank, gen, bw, 0)
            assert len(calls) == 1                                            # the source is read once
            assert im2 == pytest.approx(float(np.sum(Ad[plan.r0:plan.r1].imag ** 2)))
            for (lrp, lci, lva), ref in ((A, Ad), (AH, AHd)):
                cg = np.where(lci < plan.n_own, lci + plan.r0, plan.ghost_global[np.maximum(lci - plan.n_own, 0)] if plan.ghost_global.size else 0)
                assert np.array_equal(_dense_rows(lrp, cg, lva, N), ref[plan.r0:plan.r1])
                for r in range(lrp.size - 1):
                    assert (np.diff(cg[lrp[r]:lrp[r + 1]]) > 0).all()           # ascending global columns within a row
            if band is not None and N // nranks > 8 * max(band):
                calls.clear()
                plan, A, AH, im2 = D.local_operator(N, nranks, rank, gen, bw, 1, band)
                assert len(calls) == 1
                for (lrp, lci, lva), ref in ((A, Ad), (AH, AHd)):
                    loc = _dense_rows(lrp, lci, lva, plan.n_ext)
                    own = slice(plan.own0, plan.own0 + plan.n_own)
                    want = ref[plan.r0:plan.r1, plan.ext0:plan.ext0 + plan.n_ext]
                    assert np.array_equal(loc[own], want)
                    assert np.abs(ref[plan.r0:plan.r1]).sum() == pytest.approx(np.abs(want).sum())   # nothing of an owned row was dropped
