"""Synthetic operators and forcings for the configurations named in BASELINE.json.

The reference ships only the 500x500 tutorial operator (Tutorial/Ginzburg-Landau/A_GL); every larger
configuration must be generated.  All builders are deterministic, state their coefficients, and return
CSR with 64-bit indices and complex128 values -- the memory image of the PETSc-binary AIJ record the
reference loads through MatLoad (LNSStructure.c:22-27).

Operators
  ginzburg_landau(N)   complex Ginzburg-Landau, 4 nnz/row (columns r-2..r+1, same pattern as A_GL)
  lns_stencil(shape, nvar, complex_coupling)
                       nvar variables per grid point, full 3^d-point stencil with dense nvar x nvar
                       blocks: 2-D -> 45 nnz/row, 3-D -> 135 nnz/row (interior); strictly diagonally
                       dominant with negative real diagonal, hence stable (Gershgorin).
"""
from __future__ import annotations

import itertools

import numpy as np

# --------------------------------------------------------------------------------------------------
# Ginzburg-Landau:  A = -nu d/dx + gamma d2/dx2 + mu(x)
#   nu = U + i*cu = 2 + 0.2i, gamma = 1 - i  (the classical linear GL parameters),
#   mu(x) = mu_max + mu2 * xi^2 / 2 with xi the coordinate rescaled to [-100, 100] over the whole domain
#   (the tutorial's extent), mu2 = -0.01, mu_max = -0.02 so that mu < 0 everywhere: with the dissipative
#   upwind first derivative the numerical abscissa stays slightly negative, the analogue of the tutorial's
#   max Re(lambda) = -0.008.  Grid spacing dx = 200/499 as in the tutorial, so the spectral radius stays
#   ~50-80 and dt = 0.003 is far inside the RK4 stability region for any N.
#   d/dx: second-order upwind (3 u_i - 4 u_{i-1} + u_{i-2}) / (2 dx); d2/dx2: central; Dirichlet ends.
# --------------------------------------------------------------------------------------------------
GL_PARAMS = dict(nu=2.0 + 0.2j, gamma=1.0 - 1.0j, mu2=-0.01, mu_max=-0.02, dx=200.0 / 499.0)


def ginzburg_landau(N: int, row_range=None, **over):
    """row_range=(r0, r1) builds only those rows (global column indices), for per-GPU generation"""
    p = {**GL_PARAMS, **over}
    nu, gamma, mu2, mu_max, dx = p["nu"], p["gamma"], p["mu2"], p["mu_max"], p["dx"]
    if row_range is not None:
        r0, r1 = row_range
        Nf, rp, ci, va = _gl_rows(N, np.arange(r0, r1, dtype=np.int64), nu, gamma, mu2, mu_max, dx)
        return Nf, rp, ci, va
    r = np.arange(N, dtype=np.int64)
    return _gl_rows(N, r, nu, gamma, mu2, mu_max, dx)


def _gl_rows(N, r, nu, gamma, mu2, mu_max, dx):
    n = r.size
    xi = -100.0 + 200.0 * r / max(N - 1, 1)
    mu = mu_max + mu2 * xi * xi / 2.0
    # stencil coefficients for columns r-2, r-1, r, r+1
    c_m2 = np.full(n, -nu * (1.0 / (2 * dx)), np.complex128)
    c_m1 = np.full(n, -nu * (-4.0 / (2 * dx)) + gamma / dx**2, np.complex128)
    c_0 = -nu * (3.0 / (2 * dx)) - 2.0 * gamma / dx**2 + mu.astype(np.complex128)
    c_p1 = np.full(n, gamma / dx**2, np.complex128)
    # every row keeps exactly 4 stored entries (as A_GL does): the window is shifted at the ends
    # and the out-of-domain coefficients are dropped (Dirichlet) by storing explicit zeros.
    start = np.clip(r - 2, 0, N - 4)
    col = (start[:, None] + np.arange(4)[None, :]).astype(np.int64)
    val = np.zeros((n, 4), np.complex128)
    idx = np.arange(n)
    for off, c in ((-2, c_m2), (-1, c_m1), (0, c_0), (1, c_p1)):
        tgt = r + off
        ok = (tgt >= 0) & (tgt < N)
        slot = tgt - start
        val[idx[ok], slot[ok]] = c[ok]
    rowptr = np.arange(n + 1, dtype=np.int64) * 4
    return N, rowptr, col.reshape(-1), val.reshape(-1)


# --------------------------------------------------------------------------------------------------
# LNS-like block stencil operator
# --------------------------------------------------------------------------------------------------
def lns_stencil(shape, nvar: int = 5, complex_coupling: bool = True, seed: int = 14, diag: float = -40.0,
                row_range=None, chunk_pts: int = 1 << 15):
    """Block-stencil operator on a structured grid (x fastest), nvar unknowns per point (fastest index).

    Interior rows have 3^d * nvar entries.  Off-diagonal entries are fixed pseudo-random complex numbers
    (one nvar x nvar block per stencil offset, modulated smoothly in space so that rows are not all equal),
    scaled so that sum_j |a_ij| (j != i) <= 0.45 |diag|; the diagonal is `diag` + a small imaginary part.
    Hence every Gershgorin disc lies in Re < 0 and |lambda| <= 1.45 |diag| (dt = 0.003 -> dt*rho < 0.18).

    row_range=(r0, r1) builds only rows [r0, r1) (for per-GPU generation); columns stay global.
    Returns (N, rowptr, col, val).  The rows are produced chunk by chunk as dense (point, variable, block, variable)
    arrays and compacted with the validity mask, which is the CSR order (ascending columns) -- the same entries as
    _lns_stencil_loops, ~20x faster at the 1e8-entry sizes of the BASELINE.json configurations.
    """
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    npts = int(np.prod(shape))
    N = npts * nvar
    r0, r1 = (0, N) if row_range is None else row_range
    assert r0 % nvar == 0 and r1 % nvar == 0
    p0, p1 = r0 // nvar, r1 // nvar
    offsets = list(itertools.product((-1, 0, 1), repeat=d))  # slowest dim first
    nb = len(offsets)
    rng = np.random.default_rng(seed)
    blocks = rng.standard_normal((nb, nvar, nvar)) + (1j * rng.standard_normal((nb, nvar, nvar))
                                                      if complex_coupling else 0.0)
    centre = offsets.index((0,) * d)
    absrow = np.abs(blocks).sum(axis=(0, 2)) - np.abs(blocks[centre][np.arange(nvar), np.arange(nvar)])
    blocks = (blocks * (0.45 * abs(diag) / 1.25 / absrow.max())).astype(np.complex128)
    strides = [int(np.prod(shape[i + 1:])) for i in range(d)]
    blocks_vbu = np.ascontiguousarray(blocks.transpose(1, 0, 2))          # [v][b][u]
    uu = np.arange(nvar, dtype=np.int64)
    dia = np.arange(nvar)

    npl = p1 - p0
    rowptr = np.zeros(npl * nvar + 1, np.int64)
    cols, vals = [], []
    filled = 0
    for c0 in range(p0, p1, chunk_pts):
        pts = np.arange(c0, min(p1, c0 + chunk_pts), dtype=np.int64)
        coords = []
        rem = pts.copy()
        for s in reversed(shape):
            coords.append(rem % s)
            rem //= s
        coords = coords[::-1]
        mod = 1.0 + 0.25 * np.sin(2 * np.pi * (pts % 997) / 997.0)
        valid = np.ones((pts.size, nb), bool)
        nbr = np.empty((pts.size, nb), np.int64)
        for b, off in enumerate(offsets):
            idx = pts.copy()
            for i in range(d):
                c = coords[i] + off[i]
                valid[:, b] &= (c >= 0) & (c < shape[i])
                idx = idx + off[i] * strides[i]
            nbr[:, b] = idx
        v4 = blocks_vbu[None, :, :, :] * mod[:, None, None, None]              # [pt][v][b][u]
        dval = diag + 0.5j * np.sin(2 * np.pi * (pts % 499) / 499.0) * (1.0 if complex_coupling else 0.0)
        v4[:, dia, centre, dia] = dval[:, None]
        c4 = np.broadcast_to((nbr[:, None, :, None] * nvar + uu[None, None, None, :]), v4.shape)
        cnt = valid.sum(axis=1) * nvar
        lo = (c0 - p0) * nvar
        np.cumsum(np.repeat(cnt, nvar), out=rowptr[lo + 1: lo + pts.size * nvar + 1])
        rowptr[lo + 1: lo + pts.size * nvar + 1] += filled
        filled = int(rowptr[lo + pts.size * nvar])
        if valid.all():
            cols.append(np.ascontiguousarray(c4).reshape(-1))
            vals.append(v4.reshape(-1))
        else:
            m4 = np.broadcast_to(valid[:, None, :, None], v4.shape)
            cols.append(c4[m4])
            vals.append(v4[m4])
    col = np.concatenate(cols) if len(cols) != 1 else cols[0]
    val = np.concatenate(vals) if len(vals) != 1 else vals[0]
    return N, rowptr, col, val


def stencil_tables(d: int, nvar: int = 5, complex_coupling: bool = True, seed: int = 14, diag: float = -40.0):
    """What lns_stencil's entries are made of: blocks [3^d][nvar][nvar], the spatial modulation by (point % 997) and the
    diagonal by (point % 499) -- the inputs of the device-side generator (rsvdt_set_operator_stencil)."""
    offsets = list(itertools.product((-1, 0, 1), repeat=d))
    nb = len(offsets)
    rng = np.random.default_rng(seed)
    blocks = rng.standard_normal((nb, nvar, nvar)) + (1j * rng.standard_normal((nb, nvar, nvar))
                                                      if complex_coupling else 0.0)
    centre = offsets.index((0,) * d)
    absrow = np.abs(blocks).sum(axis=(0, 2)) - np.abs(blocks[centre][np.arange(nvar), np.arange(nvar)])
    blocks = (blocks * (0.45 * abs(diag) / 1.25 / absrow.max())).astype(np.complex128)
    i997 = np.arange(997, dtype=np.int64)
    i499 = np.arange(499, dtype=np.int64)
    mod_tab = 1.0 + 0.25 * np.sin(2 * np.pi * i997 / 997.0)
    diag_tab = (diag + 0.5j * np.sin(2 * np.pi * i499 / 499.0) * (1.0 if complex_coupling else 0.0)).astype(np.complex128)
    return blocks, mod_tab, diag_tab


def stencil_ghost_points(shape, pt0: int, pt1: int):
    """sorted grid points outside [pt0, pt1) that the stencil rows of the points in [pt0, pt1) reference (layout 0 ghosts);
    only the points within one stencil reach of the two ends of the range can reference any"""
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    strides = [int(np.prod(shape[i + 1:])) for i in range(d)]
    reach = sum(strides)
    cand = np.unique(np.concatenate([np.arange(pt0, min(pt1, pt0 + reach), dtype=np.int64),
                                     np.arange(max(pt0, pt1 - reach), pt1, dtype=np.int64)]))
    coords, rem = [], cand.copy()
    for s in reversed(shape):
        coords.append(rem % s)
        rem //= s
    coords = coords[::-1]
    out = []
    for off in itertools.product((-1, 0, 1), repeat=d):
        ok = np.ones(cand.size, bool)
        idx = cand.copy()
        for i in range(d):
            c = coords[i] + off[i]
            ok &= (c >= 0) & (c < shape[i])
            idx = idx + off[i] * strides[i]
        idx = idx[ok]
        out.append(idx[(idx < pt0) | (idx >= pt1)])
    return np.unique(np.concatenate(out)) if out else np.zeros(0, np.int64)


def _lns_stencil_loops(shape, nvar: int = 5, complex_coupling: bool = True, seed: int = 14, diag: float = -40.0,
                row_range=None):
    """(first implementation, one scatter per stencil block entry: kept as the brute-force check of lns_stencil)
    Block-stencil operator on a structured grid (x fastest), nvar unknowns per point (fastest index).

    Interior rows have 3^d * nvar entries.  Off-diagonal entries are fixed pseudo-random complex numbers
    (one nvar x nvar block per stencil offset, modulated smoothly in space so that rows are not all equal),
    scaled so that sum_j |a_ij| (j != i) <= 0.45 |diag|; the diagonal is `diag` + a small imaginary part.
    Hence every Gershgorin disc lies in Re < 0 and |lambda| <= 1.45 |diag| (dt = 0.003 -> dt*rho < 0.18).

    row_range=(r0, r1) builds only rows [r0, r1) (for per-GPU generation); columns stay global.
    Returns (N, rowptr, col, val).
    """
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    npts = int(np.prod(shape))
    N = npts * nvar
    r0, r1 = (0, N) if row_range is None else row_range
    assert r0 % nvar == 0 and r1 % nvar == 0
    p0, p1 = r0 // nvar, r1 // nvar
    offsets = list(itertools.product((-1, 0, 1), repeat=d))  # slowest dim first
    nb = len(offsets)
    rng = np.random.default_rng(seed)
    blocks = rng.standard_normal((nb, nvar, nvar)) + (1j * rng.standard_normal((nb, nvar, nvar))
                                                      if complex_coupling else 0.0)
    # normalise: total off-diagonal magnitude per row <= 0.45 |diag| / 1.25 (modulation <= 1.25)
    centre = offsets.index((0,) * d)
    absrow = np.abs(blocks).sum(axis=(0, 2)) - np.abs(blocks[centre][np.arange(nvar), np.arange(nvar)])
    blocks = blocks * (0.45 * abs(diag) / 1.25 / absrow.max())

    pts = np.arange(p0, p1, dtype=np.int64)
    # coordinates, x fastest: pts = x + nx*(y + ny*z); shape given as (nz, ny, nx) slowest first
    coords = []
    rem = pts.copy()
    for s in reversed(shape):
        coords.append(rem % s)
        rem //= s
    coords = coords[::-1]  # slowest first, aligned with `shape` and `offsets`
    strides = [int(np.prod(shape[i + 1:])) for i in range(d)]
    mod = 1.0 + 0.25 * np.sin(2 * np.pi * (pts % 997) / 997.0)  # smooth spatial modulation in [0.75, 1.25]

    npl = p1 - p0
    valid = np.ones((npl, nb), bool)
    nbr = np.zeros((npl, nb), np.int64)
    for b, off in enumerate(offsets):
        idx = pts.copy()
        for i in range(d):
            c = coords[i] + off[i]
            valid[:, b] &= (c >= 0) & (c < shape[i])
            idx = idx + off[i] * strides[i]
        nbr[:, b] = idx
    cnt_pt = valid.sum(axis=1) * nvar            # entries per row for each point
    rowlen = np.repeat(cnt_pt, nvar)
    rowptr = np.zeros(npl * nvar + 1, np.int64)
    np.cumsum(rowlen, out=rowptr[1:])
    nnz = int(rowptr[-1])
    col = np.empty(nnz, np.int64)
    val = np.empty(nnz, np.complex128)
    # fill by (point, variable) rows; vectorised over points with identical validity masks
    slot = np.cumsum(valid, axis=1) - 1          # position of block b among the valid blocks of the point
    for b in range(nb):
        ok = valid[:, b]
        if not ok.any():
            continue
        pidx = np.nonzero(ok)[0]
        for v in range(nvar):
            rows = pidx * nvar + v
            base = rowptr[rows] + slot[pidx, b] * nvar
            for u in range(nvar):
                col[base + u] = nbr[pidx, b] * nvar + u
                if b == centre and u == v:
                    val[base + u] = diag + 0.5j * np.sin(2 * np.pi * (pts[pidx] % 499) / 499.0) \
                        * (1.0 if complex_coupling else 0.0)
                else:
                    val[base + u] = blocks[b, v, u] * mod[pidx]
    return N, rowptr, col, val


def random_forcing(N: int, k: int, Nw_eff: int, seed: int = 14) -> np.ndarray:
    """F_hat: N x (k*Nw_eff), entries uniform [0,1) + i[0,1) like PETSc's default complex MatSetRandom
    (CreateRandomMat.c:39); frequency-major columns (col = iw*k + ik), Fortran order."""
    rng = np.random.default_rng(seed)
    F = np.empty((N, k * Nw_eff), np.complex128, order="F")
    for c in range(k * Nw_eff):
        F[:, c] = rng.random(N) + 1j * rng.random(N)
    return F


class SyntheticRows:
    """Row source with the interface of petsc_io.MatRowReader (M, N, rows(r0, r1)) that GENERATES the rows of a BASELINE.json
    operator instead of reading them: `gl:<N>`, `lns2d:<ny>x<nx>`, `lns3d:<nz>x<ny>x<nx>` (5 variables per point).  The
    multi-GPU executable accepts `OperatorDir: synthetic:<spec>` so that the 2e7-DOF configurations (a 65 GB file) can be run
    without writing and re-reading the operator; tools/make_case.py writes the very same rows to a PETSc-binary file."""

    def __init__(self, spec: str, nvar: int = 5):
        kind, _, arg = spec.partition(":")
        self.kind = kind
        if kind == "gl":
            self.M = self.N = int(arg)
            self.align = 1
        elif kind in ("lns2d", "lns3d"):
            self.shape = tuple(int(s) for s in arg.lower().split("x"))
            if len(self.shape) != (2 if kind == "lns2d" else 3):
                raise ValueError(f"{spec}: wrong number of grid dimensions")
            self.nvar = nvar
            self.M = self.N = int(np.prod(self.shape)) * nvar
            self.align = nvar
        else:
            raise ValueError(f"unknown synthetic operator '{spec}' (gl:<N>, lns2d:<ny>x<nx>, lns3d:<nz>x<ny>x<nx>)")

    def hints(self):
        """(max(row - col), max(col - row), real-valued?) without generating the operator: the band of a block stencil is that
        of an interior point; grids with a dimension < 3 (no interior) are scanned whole, they are tiny."""
        if self.kind == "gl":
            return 3, 3, False      # the 4-entry windows are shifted at the ends (explicit zeros up to 3 columns away)
        if min(self.shape) >= 3:
            centre = 0
            for s_ in self.shape:
                centre = centre * s_ + s_ // 2
            a, b = centre * self.nvar, (centre + 1) * self.nvar
        else:
            a, b = 0, self.M
        rp, ci, _ = self._rows(a, b)
        d = ci - np.repeat(np.arange(a, b, dtype=np.int64), np.diff(rp))
        return int(max(-d.min(), 0)), int(max(d.max(), 0)), False

    def rows(self, r0: int, r1: int):
        r0, r1 = max(0, int(r0)), min(self.M, int(r1))
        if getattr(self, "_memo_key", None) == (r0, r1):       # the executable asks for its own range twice in a row
            return self._memo
        out = self._rows(r0, r1)
        self._memo_key, self._memo = (r0, r1), out
        return out

    def _rows(self, r0: int, r1: int):
        if r1 <= r0:
            return np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.complex128)
        if self.kind == "gl":
            _, rp, ci, va = ginzburg_landau(self.M, row_range=(r0, r1))
            return rp, ci, va
        a = r0 - r0 % self.align
        b = r1 + (-r1) % self.align
        _, rp, ci, va = lns_stencil(self.shape, nvar=self.nvar, complex_coupling=True, row_range=(a, b))
        s, e = rp[r0 - a], rp[r1 - a]
        return rp[r0 - a: r1 - a + 1] - s, ci[s:e], va[s:e]
