#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank runs the row-partitioned path (both local layouts); rank 0 also runs the single-GPU path and the CPU oracle on
the same inputs and compares: direct / adjoint actions, QRAllFreqs, the full solve (gains, subspaces)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gather_rows(local, plan_ranges, world):
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return np.concatenate(parts, axis=0)


def run_case(pkg, name, N, gen_rows, full_csr, max_bw, layout, band, k, Nw, w, dt_in, TL, rank, world, dev, TR=False, options=None):
    dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(w, Nw, dt_in, TL)
    grid = dict(k=k, Nw=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w, TransientRemoval=TR)
    r, plan = pkg.dist.setup_rank(pkg, dev, N, gen_rows, max_bw, layout, band, grid)
    for oname, oval in (options or {}).items():
        r.set_option(oname, oval)
    Nw_eff = r.grid.Nw_eff
    rng = np.random.default_rng(21)
    F = np.asfortranarray(rng.random((N, k * Nw_eff)) + 1j * rng.random((N, k * Nw_eff)))
    Fl = np.asfortranarray(F[plan.r0:plan.r1])
    out = {}
    for direct in (True, False):
        r.set_Y_hat(Fl)
        (r.DirectActionRK4 if direct else r.AdjointActionRK4)()
        out["direct" if direct else "adjoint"] = gather_rows(np.ascontiguousarray(r.get_Y_hat()), None, world)
    r.set_Y_hat(Fl)
    r.QRAllFreqs()
    out["qr"] = gather_rows(np.ascontiguousarray(r.get_Y_hat()), None, world)
    # full solve, q = 1 (main.c:134-146)
    r.set_Y_hat(Fl)
    r.DirectActionRK4(); r.PowerIterationRK4(1); r.SVDAllFreqsBeforeAdjoint(); r.StoreU(); r.AdjointActionRK4(); r.SVDAllFreqs()
    out["S"] = r.get_S_hat()
    out["U"] = gather_rows(np.ascontiguousarray(r.get_U_hat()), None, world)
    out["V"] = gather_rows(np.ascontiguousarray(r.get_V_hat()), None, world)
    launches = r.kernel_launches()
    r.close()
    ok = True
    if rank == 0:
        from conftest import principal_angle
        n, rp, ci, va = full_csr
        s = pkg.capi.RSVDt(dev)
        s.LNSStructure(N, rp, ci, va)
        s.set_grid(k, Nw, Ns, Nt, RR, dt, w, TR)
        errs = {}
        for direct in (True, False):
            errs["direct" if direct else "adjoint"] = rel(out["direct" if direct else "adjoint"], s.TSActionRK4(F, direct=direct))
        s.set_Y_hat(F); s.QRAllFreqs(); errs["qr"] = rel(out["qr"], s.get_Y_hat())
        s.set_Y_hat(F)
        s.DirectActionRK4(); s.PowerIterationRK4(1); s.SVDAllFreqsBeforeAdjoint(); s.StoreU(); s.AdjointActionRK4(); s.SVDAllFreqs()
        S1, U1, V1 = s.get_S_hat(), s.get_U_hat(), s.get_V_hat()
        errs["gains"] = float(np.abs(out["S"] / S1 - 1).max())
        errs["angleU"] = max(principal_angle(out["U"][:, i * k:(i + 1) * k], U1[:, i * k:(i + 1) * k]) for i in range(Nw_eff))
        errs["angleV"] = max(principal_angle(out["V"][:, i * k:(i + 1) * k], V1[:, i * k:(i + 1) * k]) for i in range(Nw_eff))
        s.close()
        tol = 1e-9 if TR else 1e-11   # with transient removal the sums of the Gram products are split across ranks
        ok = errs["direct"] < tol and errs["adjoint"] < tol and errs["qr"] < 1e-10 and errs["gains"] < 1e-8 \
            and errs["angleU"] < 1e-6 and errs["angleV"] < 1e-6
        print(f"[dist_check] {name}{' +TR' if TR else ''}: world={world} layout={layout} N={N} launches={launches} " +
              " ".join(f"{a}={b:.2e}" for a, b in errs.items()) + ("  OK" if ok else "  FAIL"), flush=True)
    return ok


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    pkg = entry.load_package()
    ok = True
    # Ginzburg-Landau band: layout 1 (fused step, one exchange per step) and layout 0 (stage kernels, exchange per stage)
    N = 20011
    full = pkg.synthetic.ginzburg_landau(N)
    gl = lambda a, b: pkg.synthetic.ginzburg_landau(N, row_range=(a, b))[1:]  # noqa: E731
    for layout in (1, 0):
        ok &= run_case(pkg, "ginzburg-landau", N, gl, full, 3, layout, (2, 1), 6, 6, 2.0, 0.003, 0.4, rank, world, lrank)
    # transient removal (Galerkin projection) on the partitioned rows, both layouts
    for layout in (1, 0):
        ok &= run_case(pkg, "ginzburg-landau", N, gl, full, 3, layout, (2, 1), 4, 8, 2.0, 0.003, 0.4, rank, world, lrank, TR=True)
    # 2-D LNS-like block stencil (45 nnz/row): layout 0
    shape = (24, 20)
    fullL = pkg.synthetic.lns_stencil(shape, nvar=5)
    NL = fullL[0]

    def lns(a, b):
        a5, b5 = a - a % 5, b + (-b) % 5
        _, rpp, cc, vv = pkg.synthetic.lns_stencil(shape, nvar=5, row_range=(a5, b5))
        s, e = rpp[a - a5], rpp[b - a5]
        return rpp[a - a5: b - a5 + 1] - s, cc[s:e], vv[s:e]

    ok &= run_case(pkg, "lns-2d", NL, lns, fullL, 5 * (20 + 2), 0, None, 4, 4, 3.0, 0.003, 0.3, rank, world, lrank)
    ok &= run_case(pkg, "lns-2d overlap", NL, lns, fullL, 5 * (20 + 2), 0, None, 4, 4, 3.0, 0.003, 0.3, rank, world, lrank, options={"use_overlap": 1})
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("[dist_check] ALL OK" if int(t.item()) else "[dist_check] FAILED", flush=True)
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
