#!/usr/bin/env python
"""Writes a BASELINE.json configuration as a case directory the reference executable -- and therefore the drop-ins
rsvd-delta-t_b200/RSVDt (one GPU) and rsvd-delta-t_b200/rsvdt_mpi.py (P GPUs) -- can run as is:

    <out>/A                        the operator, PETSc-binary AIJ (64-bit indices, complex scalars), streamed in row chunks
    <out>/{B,C,W_f_sqrt_inv,W_q_sqrt,W_q_sqrt_inv}     with --weights: selection-type B (N x Nb) and C (Nc x N) on a sub-box of
                                   the grid, diagonal positive weights
    <out>/F_hat                    with --forcing: the N_b x (k*Nw_eff) forcing every implementation should read
    <out>/variables.yaml           the reference's input file (variables.yaml:1-129)
    <out>/results/

Configurations:  gl      complex Ginzburg-Landau, N rows (config 2: --n 1000000, k 10, Nw 64)
                                        lns2d   5 variables x 9-point stencil  (config 3: --shape 1000 1000)
                                        lns3d   5 variables x 27-point stencil (configs 4, 5: --shape 100 200 200)

    python tools/make_case.py lns3d --shape 100 200 200 --k 8 --Nw 32 --w 2.0 --TR --q 1 --weights --out /scratch/case4
    python -m torch.distributed.run --nproc-per-node 8 rsvd-delta-t_b200/rsvdt_mpi.py -inputs variables.yaml   (cwd = out)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

YAML = """# written by tools/make_case.py: {desc}
RootDir:            {root}/
ResultsDir:         results/
OperatorDir:        {operator}
k:                  {k}
q:                  {q}
Nw:                 {Nw}
w:                  {w}
TwoPI:              false
TransientLength:    {TL}
dt:                 {dt}
TransientRemoval:   {TR}
Display:            1
DiscFlg:            false
beta:               0.0
RandSeed:           14
SaveResultsOpt:     {opt}
InputForcingFlg:    {forcing}
InputForcingDir:    F_hat
InvInputWeightFlg:  {weights}
InvInputWeightDir:  W_f_sqrt_inv
OutputWeightFlg:    {weights}
OutputWeightDir:    W_q_sqrt
InvOutputWeightFlg: {weights}
InvOutputWeightDir: W_q_sqrt_inv
InputMatrixFlg:     {weights}
InputMatrixDir:     B
OutputMatrixFlg:    {weights}
OutputMatrixDir:    C
TransRun:           false
TransPeriods:       1
TransSave:          false
TransSaveMod:       500
TransRemovalEst:    false
TransDivVal:        1e3
TransConVal:        1e-8
TransICFlg:         false
TransICDir:         none
"""


def selection_rows(shape, nvar, box):
    """global DOF indices (sorted) of all variables at the grid points inside `box` = [(lo, hi), ...] per dimension"""
    grids = np.meshgrid(*[np.arange(lo, hi) for lo, hi in box], indexing="ij")
    pt = np.zeros_like(grids[0])
    for i, g in enumerate(grids):
        pt = pt * shape[i] + g
    dof = (pt.reshape(-1)[:, None] * nvar + np.arange(nvar)[None, :]).reshape(-1)
    return np.sort(dof.astype(np.int64))


def write_case(pkg, out, config, n=None, shape=None, k=8, Nw=32, w=2.0, dt=0.003, TL=0.5, TR=False, q=1, opt=1, weights=False,
               forcing=False, chunk_points=200_000, nvar=5, synthetic=False):
    io, syn = pkg.petsc_io, pkg.synthetic
    os.makedirs(os.path.join(out, "results"), exist_ok=True)
    if config == "gl":
        N = int(n)
        gen = lambda a, b: syn.ginzburg_landau(N, row_range=(a, b))[1:]  # noqa: E731
        step, real = 1_000_000, False
        desc = f"complex Ginzburg-Landau, N = {N}"
        shape_t = (N,)
        nv = 1
    else:
        shape_t = tuple(int(s) for s in shape)
        assert len(shape_t) == (2 if config == "lns2d" else 3)
        N = int(np.prod(shape_t)) * nvar
        gen = lambda a, b: syn.lns_stencil(shape_t, nvar=nvar, complex_coupling=True, row_range=(a, b))[1:]  # noqa: E731
        step, real = chunk_points * nvar, False
        desc = f"{nvar} variables x {3 ** len(shape_t)}-point stencil on {'x'.join(map(str, shape_t))}, N = {N}"
        nv = nvar
    operator = "A"
    if synthetic:   # rsvdt_mpi.py generates the same rows per rank (synthetic.SyntheticRows); no operator file is written
        operator = "synthetic:" + (f"gl:{N}" if config == "gl" else f"{config}:" + "x".join(map(str, shape_t)))
    else:
        wtr = io.AijStreamWriter(os.path.join(out, "A"), N, N)
        for a in range(0, N, step):
            wtr.append_rows(*gen(a, min(N, a + step)))
        wtr.close()
    Nb = Nc = N
    if weights:
        # B / C select a sub-box (the middle half of every dimension for the forcing, the downstream half of the slowest one
        # for the response); the weights are positive diagonals (quadrature-like, smoothly varying)
        boxb = [(s // 4, max(s // 4 + 1, 3 * s // 4)) for s in shape_t]
        boxc = [(shape_t[0] // 2, shape_t[0])] + [(0, s) for s in shape_t[1:]]
        rb, rc = selection_rows(shape_t, nv, boxb), selection_rows(shape_t, nv, boxc)
        Nb, Nc = rb.size, rc.size
        one = lambda m: np.ones(m, np.complex128)  # noqa: E731
        # B: N x Nb, column j has its single entry in row rb[j] -> as CSR by rows of N
        rowlen = np.zeros(N, np.int64)
        rowlen[rb] = 1
        rp = np.zeros(N + 1, np.int64)
        np.cumsum(rowlen, out=rp[1:])
        io.write_mat_aij(os.path.join(out, "B"), N, Nb, rp, np.arange(Nb, dtype=np.int64), one(Nb))
        io.write_mat_aij(os.path.join(out, "C"), Nc, N, np.arange(Nc + 1, dtype=np.int64), rc, one(Nc))
        wf = 1.0 + 0.5 * np.cos(np.arange(Nb) * (2 * np.pi / 977.0))
        wq = 1.0 + 0.5 * np.sin(np.arange(Nc) * (2 * np.pi / 1013.0))
        dia = lambda m: (np.arange(m + 1, dtype=np.int64), np.arange(m, dtype=np.int64))  # noqa: E731
        io.write_mat_aij(os.path.join(out, "W_f_sqrt_inv"), Nb, Nb, *dia(Nb), (1.0 / np.sqrt(wf)).astype(np.complex128))
        io.write_mat_aij(os.path.join(out, "W_q_sqrt"), Nc, Nc, *dia(Nc), np.sqrt(wq).astype(np.complex128))
        io.write_mat_aij(os.path.join(out, "W_q_sqrt_inv"), Nc, Nc, *dia(Nc), (1.0 / np.sqrt(wq)).astype(np.complex128))
    Nw_eff = Nw // 2 if real else Nw
    if forcing:
        io.write_mat_dense(os.path.join(out, "F_hat"), syn.random_forcing(Nb, k, Nw_eff, seed=14))
    b = lambda x: "true" if x else "false"  # noqa: E731
    with open(os.path.join(out, "variables.yaml"), "w") as f:
        f.write(YAML.format(desc=desc, root=os.path.abspath(out), k=k, q=q, Nw=Nw, w=w, TL=TL, dt=dt, TR=b(TR), opt=opt,
                            forcing=b(forcing), weights=b(weights), operator=operator))
    return dict(N=N, Nb=Nb, Nc=Nc, Nw_eff=Nw_eff, desc=desc)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("config", choices=["gl", "lns2d", "lns3d"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--n", type=int, default=1_000_000, help="rows of the Ginzburg-Landau operator")
    ap.add_argument("--shape", type=int, nargs="+", help="grid, slowest dimension first")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--Nw", type=int, default=32)
    ap.add_argument("--w", type=float, default=2.0)
    ap.add_argument("--dt", type=float, default=0.003)
    ap.add_argument("--TL", type=float, default=0.5, help="TransientLength")
    ap.add_argument("--TR", action="store_true", help="TransientRemoval")
    ap.add_argument("--q", type=int, default=1)
    ap.add_argument("--opt", type=int, default=1, help="SaveResultsOpt")
    ap.add_argument("--weights", action="store_true")
    ap.add_argument("--forcing", action="store_true", help="write F_hat and set InputForcingFlg")
    ap.add_argument("--synthetic", action="store_true",
                    help="do not write A: OperatorDir becomes synthetic:<spec>, which rsvdt_mpi.py generates per rank")
    a = ap.parse_args()
    pkg = entry.load_package()
    info = write_case(pkg, a.out, a.config, n=a.n, shape=a.shape, k=a.k, Nw=a.Nw, w=a.w, dt=a.dt, TL=a.TL, TR=a.TR, q=a.q,
                      opt=a.opt, weights=a.weights, forcing=a.forcing, synthetic=a.synthetic)
    print(f"{a.out}: {info['desc']}; N_b = {info['Nb']}, N_c = {info['Nc']}, Nw_eff = {info['Nw_eff']}")


if __name__ == "__main__":
    main()
