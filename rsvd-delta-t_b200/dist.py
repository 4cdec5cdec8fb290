"""Row partition of the hot path over several GPUs (one process per GPU, torch.distributed for the plumbing).

The reference distributes A as MATMPIAIJ and every vector as VECMPI with PETSC_DECIDE local sizes
(LNSStructure.c:24-26, CreateTSMats.c:15-17): rank r owns floor(N/P) + (r < N mod P) consecutive rows, and PETSc's
VecScatter gathers the off-process ("ghost") entries before every mat-vec.  This module is the host-side planner of
that layout -- pure index arithmetic on numpy arrays, unit-tested on CPU with the gloo backend -- plus the bootstrap
that hands the plan to the CUDA library (include/rsvdt.h, "row partition across GPUs"), which then moves the halo rows
itself with peer stores over NVLink.

Two local layouts
  layout 0 (general sparsity)   local rows = owned rows; ghost columns renumbered after the owned ones (what MPIAIJ's
                                off-diagonal block does); one halo exchange per RK4 stage.
  layout 1 (narrow band)        local rows = owned rows + 4*lo / 4*hi neighbour rows (lo/hi = band half widths): the fused
                                RK4-step kernel recomputes the stages on the halo rows, so only y is exchanged, once per step.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


def split_ownership(N: int, nranks: int, rank: int):
    """PetscSplitOwnership with PETSC_DECIDE -> [row_begin, row_end)"""
    q, r = divmod(N, nranks)
    b = rank * q + min(rank, r)
    return b, b + q + (1 if rank < r else 0)


def all_ranges(N: int, nranks: int):
    return [split_ownership(N, nranks, p) for p in range(nranks)]


def owner_of(rows: np.ndarray, ranges) -> np.ndarray:
    starts = np.array([b for b, _ in ranges], np.int64)
    return np.searchsorted(starts, rows, side="right") - 1


@dataclass
class Plan:
    layout: int
    rank: int
    nranks: int
    N: int
    r0: int
    r1: int
    ext0: int            # global row of local row 0
    n_ext: int           # local rows of every block vector (owned + ghosts)
    own0: int            # local index of the first owned row
    # requests[p] = (global rows I need from rank p [sorted], first local ghost row they fill)
    requests: dict = field(default_factory=dict)
    ghost_global: np.ndarray = None   # layout 0: global index of ghost column j (local column n_own + j)

    @property
    def n_own(self):
        return self.r1 - self.r0

    def to_local_rows(self, global_rows: np.ndarray) -> np.ndarray:
        """local index of OWNED global rows (used to build the send lists)"""
        return np.asarray(global_rows, np.int64) - self.r0 + self.own0


def plan_general(N, nranks, rank, col_global: np.ndarray) -> Plan:
    """layout 0 from the global column indices of the owned rows"""
    ranges = all_ranges(N, nranks)
    r0, r1 = ranges[rank]
    col_global = np.asarray(col_global, np.int64)
    off = (col_global < r0) | (col_global >= r1)
    ghosts = np.unique(col_global[off])
    p = Plan(0, rank, nranks, N, r0, r1, r0, (r1 - r0) + ghosts.size, 0, {}, ghosts)
    if ghosts.size:
        own = owner_of(ghosts, ranges)
        for q in np.unique(own):
            sel = np.nonzero(own == q)[0]          # contiguous because ghosts are sorted and ranges are ordered
            p.requests[int(q)] = (ghosts[sel], (r1 - r0) + int(sel[0]))
    return p


def localize_general(plan: Plan, col_global: np.ndarray) -> np.ndarray:
    col_global = np.asarray(col_global, np.int64)
    loc = col_global - plan.r0
    off = (col_global < plan.r0) | (col_global >= plan.r1)
    if off.any():
        loc[off] = plan.n_own + np.searchsorted(plan.ghost_global, col_global[off])
    return loc


def plan_banded(N, nranks, rank, lo: int, hi: int) -> Plan:
    """layout 1: halo of 4*lo rows on the left and 4*hi on the right (the dependency depth of four RK4 stages)"""
    ranges = all_ranges(N, nranks)
    r0, r1 = ranges[rank]
    ext0, ext1 = max(0, r0 - 4 * lo), min(N, r1 + 4 * hi)
    p = Plan(1, rank, nranks, N, r0, r1, ext0, ext1 - ext0, r0 - ext0, {}, None)
    ghosts = np.concatenate([np.arange(ext0, r0, dtype=np.int64), np.arange(r1, ext1, dtype=np.int64)])
    if ghosts.size:
        own = owner_of(ghosts, ranges)
        for q in np.unique(own):
            rows = ghosts[own == q]
            p.requests[int(q)] = (rows, int(rows[0] - ext0))
    return p


def localize_banded(plan: Plan, rowptr, col_global, val):
    """rows [ext0, ext0+n_ext) of the global matrix -> local column indices; entries outside the local range are dropped
    (they only feed outer-halo rows whose values never reach an owned row)"""
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col_global, np.int64) - plan.ext0
    keep = (col >= 0) & (col < plan.n_ext)
    rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
    cnt = np.bincount(rows[keep], minlength=rowptr.size - 1)
    rp = np.zeros(rowptr.size, np.int64)
    np.cumsum(cnt, out=rp[1:])
    return rp, col[keep], np.asarray(val)[keep]


def transpose_slab(rowptr, col, val, row_offset: int, c_lo: int, c_hi: int):
    """rows [c_lo, c_hi) of A^H from a slab of rows of A (global row of slab row i is row_offset + i) that contains
    every entry with a column in [c_lo, c_hi).  The entries arrive row by row, so one STABLE counting sort by column leaves
    every row of A^H with ascending columns: scipy's CSR -> CSC conversion where scipy is installed (O(nnz); 3 s instead of
    17 s for the 1.3e8 entries a rank of the 2e6-DOF 3-D case holds), else a stable argsort."""
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    val = np.asarray(val)
    nrows = rowptr.size - 1
    try:
        from scipy import sparse
    except ImportError:
        sparse = None
    if sparse is not None and col.size:
        ncols = max(int(col.max()) + 1, c_hi)
        T = sparse.csr_matrix((val, col, rowptr), shape=(nrows, ncols)).tocsc()
        ip = np.asarray(T.indptr, np.int64)
        a, b = ip[c_lo], ip[c_hi]
        return ip[c_lo:c_hi + 1] - a, np.asarray(T.indices[a:b], np.int64) + row_offset, np.conj(T.data[a:b])
    rows = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(rowptr)) + row_offset
    sel = (col >= c_lo) & (col < c_hi)
    r, c, v = rows[sel], col[sel], np.conj(val[sel])
    order = np.argsort(c, kind="stable")
    r, c, v = r[order], c[order], v[order]
    rp = np.zeros(c_hi - c_lo + 1, np.int64)
    np.cumsum(np.bincount(c - c_lo, minlength=c_hi - c_lo), out=rp[1:])
    return rp, r, v


def slice_rows(rowptr, col, val, row_offset: int, r_lo: int, r_hi: int):
    """rows [r_lo, r_hi) (global) out of a slab whose first row is global row `row_offset`"""
    rowptr = np.asarray(rowptr, np.int64)
    a, b = r_lo - row_offset, r_hi - row_offset
    s, e = rowptr[a], rowptr[b]
    return rowptr[a:b + 1] - s, np.asarray(col)[s:e], np.asarray(val)[s:e]


def local_weight(nrows, ncols, rowptr, col, val, nranks, rank):
    """Rows of a weight / input / output matrix M (nrows x ncols, global CSR) that `rank` owns, and the rows of M^H it owns:
    PETSC_DECIDE splits of M's row space and column space (what MatLoad gives a MATMPIAIJ, ReadWeightMats.c:28-31, and what
    the in-place MatHermitianTranspose of ApplyWeightMats.c:44-74 produces).  Column indices stay global.
    -> ((rowptr, col, val) of M[r0:r1, :], (rowptr, col, val) of M^H[c0:c1, :])"""
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    val = np.asarray(val)
    r0, r1 = split_ownership(nrows, nranks, rank)
    c0, c1 = split_ownership(ncols, nranks, rank)
    p0, p1 = rowptr[r0], rowptr[r1]
    M_rows = (rowptr[r0:r1 + 1] - p0, col[p0:p1], val[p0:p1])
    MH_rows = transpose_slab(rowptr, col, val, 0, c0, c1)
    return M_rows, MH_rows


def build_messages(plan: Plan, all_requests):
    """all_requests[p][q] = (rows p needs from q, first ghost row on p).  -> my outgoing / incoming message lists"""
    out = []
    for p in range(plan.nranks):
        if p == plan.rank:
            continue
        req = all_requests[p].get(plan.rank)
        if req is None:
            continue
        rows, dest0 = req
        out.append((p, plan.to_local_rows(rows), int(dest0)))
    inc = [(int(q), int(len(rows)), int(dest0)) for q, (rows, dest0) in sorted(plan.requests.items())]
    return out, inc


# ----------------------------------------------------------------------------------------------------------------------
# bootstrap of the CUDA library (needs torch.distributed initialised with one rank per GPU)
# ----------------------------------------------------------------------------------------------------------------------
_KEEPALIVE = []


def connect(r, plan: Plan, group=None):
    """r: capi.RSVDt handle with the local operator and the grid already set.  Exchanges the halo plans, the mailbox
    offsets and the CUDA IPC handles, and registers a torch.distributed all-reduce for the Gram matrices / norms."""
    import torch
    import torch.distributed as dist

    from . import capi

    L = capi.lib()
    world = dist.get_world_size(group)
    allreq = [None] * world
    dist.all_gather_object(allreq, {q: (rows, d0) for q, (rows, d0) in plan.requests.items()}, group=group)
    out, inc = build_messages(plan, allreq)

    out_peer = np.array([p for p, _, _ in out], np.int32)
    out_count = np.array([len(rows) for _, rows, _ in out], np.int32)
    out_rows = np.concatenate([rows for _, rows, _ in out]).astype(np.int64) if out else np.zeros(0, np.int64)
    out_dest = np.array([d for _, _, d in out], np.int64)
    in_peer = np.array([q for q, _, _ in inc], np.int32)
    in_count = np.array([c for _, c, _ in inc], np.int32)
    in_dest = np.array([d for _, _, d in inc], np.int64)
    p_ = capi._p
    r._ck(L.rsvdt_dist_set_messages(r._h, C.c_int(len(out)), p_(out_peer), p_(out_count), p_(out_rows), p_(out_dest),
                                    C.c_int(len(inc)), p_(in_peer), p_(in_count), p_(in_dest)))
    in_off = np.zeros(max(len(inc), 1), np.int64)
    r._ck(L.rsvdt_dist_get_in_offsets(r._h, p_(in_off)))
    alloff = [None] * world
    dist.all_gather_object(alloff, {int(q): int(in_off[i]) for i, (q, _, _) in enumerate(inc)}, group=group)
    out_off = np.array([alloff[p][plan.rank] for p, _, _ in out], np.int64) if out else np.zeros(1, np.int64)

    mine = np.zeros(128, np.uint8)
    r._ck(L.rsvdt_dist_ipc_export(r._h, p_(mine)))
    allh = [None] * world
    dist.all_gather_object(allh, mine.tobytes(), group=group)
    blob = np.frombuffer(b"".join(allh), np.uint8).copy()
    r._ck(L.rsvdt_dist_ipc_import(r._h, p_(blob), p_(out_off)))

    class _Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def _allreduce(ptr, n, stream, user):
        try:
            t = torch.as_tensor(_Dev(ptr, n), device="cuda")
            ext = torch.cuda.ExternalStream(stream)
            cur = torch.cuda.current_stream()
            cur.wait_stream(ext)                       # the Gram kernels ran on the library's stream
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            ext.wait_stream(cur)
            return 0
        except Exception as e:  # pragma: no cover
            print("all-reduce callback failed:", e)
            return 1

    cb = capi.ALLREDUCE_FN(_allreduce)
    _KEEPALIVE.append(cb)
    r._ck(L.rsvdt_set_allreduce(r._h, cb, None))

    def _allgather(send, recv, n, stream, user):
        try:
            ts = torch.as_tensor(_Dev(send, n), device="cuda")
            tr = torch.as_tensor(_Dev(recv, n * world), device="cuda")
            ext = torch.cuda.ExternalStream(stream)
            cur = torch.cuda.current_stream()
            cur.wait_stream(ext)                       # the packing copy ran on the library's stream
            dist.all_gather_into_tensor(tr, ts, group=group)
            ext.wait_stream(cur)
            return 0
        except Exception as e:  # pragma: no cover
            print("all-gather callback failed:", e)
            return 1

    cbg = capi.ALLGATHER_FN(_allgather)
    _KEEPALIVE.append(cbg)
    r._ck(L.rsvdt_set_allgather(r._h, cbg, None))
    dist.barrier(group=group)
    return out, inc


# ----------------------------------------------------------------------------------------------------------------------
# convenience: build the local A and A^H of one rank from a row generator
# ----------------------------------------------------------------------------------------------------------------------
def local_operator(N, nranks, rank, gen_rows, max_bw, layout, band=None, with_adjoint=True):
    """gen_rows(r_lo, r_hi) -> (rowptr, col_global, val) of global rows [r_lo, r_hi).  max_bw bounds |col - row|.
    band = (lo, hi) half widths of the dominant diagonal pattern (layout 1 only).
    -> (plan, (rowptr, col_local, val) of A, same of A^H, local sum of Im(A)^2 over the owned rows)"""
    r0, r1 = split_ownership(N, nranks, rank)
    if layout == 0:
        if with_adjoint:     # one pass over the source: the slab that holds every entry of columns [r0, r1) contains the owned rows
            s0, s1 = max(0, r0 - max_bw), min(N, r1 + max_bw)
            srp, scg, sva = gen_rows(s0, s1)
            rp, cg, va = slice_rows(srp, scg, sva, s0, r0, r1)
            hrp, hcg, hva = transpose_slab(srp, scg, sva, s0, r0, r1)
            del srp, scg, sva
        else:   # measurement runs that only use the direct operator: A stands in for A^H (never applied)
            rp, cg, va = gen_rows(r0, r1)
            hrp, hcg, hva = rp, cg, va
        plan = plan_general(N, nranks, rank, np.concatenate([cg, hcg]))
        A = (rp, localize_general(plan, cg), va)
        AH = (hrp, localize_general(plan, hcg), hva)
        im2 = float(np.sum(np.asarray(va).imag ** 2))
        return plan, A, AH, im2
    lo, hi = band
    # A^H has the mirrored band (hi, lo); one halo must serve both operators
    h = max(lo, hi)
    plan = plan_banded(N, nranks, rank, h, h)
    e0, e1 = plan.ext0, plan.ext0 + plan.n_ext
    s0, s1 = max(0, e0 - max_bw), min(N, e1 + max_bw)
    srp, scg, sva = gen_rows(s0, s1)
    rp, cg, va = slice_rows(srp, scg, sva, s0, e0, e1)
    hrp, hcg, hva = transpose_slab(srp, scg, sva, s0, e0, e1)
    A = localize_banded(plan, rp, cg, va)
    AH = localize_banded(plan, hrp, hcg, hva)
    own = slice(rp[plan.own0], rp[plan.own0 + plan.n_own])
    im2 = float(np.sum(np.asarray(va)[own].imag ** 2))
    return plan, A, AH, im2


def plan_stencil(shape, nvar, nranks, rank) -> Plan:
    """layout 0 plan of a block-stencil operator from its geometry alone (synthetic.stencil_ghost_points): the PETSC_DECIDE row
    split must fall on grid points (N/P divisible by nvar ... or the rows are split inside a point, which the block kernels
    cannot take)"""
    from . import synthetic
    npts = int(np.prod(shape))
    N = npts * nvar
    ranges = all_ranges(N, nranks)
    r0, r1 = ranges[rank]
    if r0 % nvar or r1 % nvar:
        raise ValueError(f"row split [{r0}, {r1}) of rank {rank} cuts a grid point ({nvar} variables per point)")
    gp = synthetic.stencil_ghost_points(shape, r0 // nvar, r1 // nvar)
    ghosts = (gp[:, None] * nvar + np.arange(nvar, dtype=np.int64)[None, :]).reshape(-1)
    p = Plan(0, rank, nranks, N, r0, r1, r0, (r1 - r0) + ghosts.size, 0, {}, ghosts)
    if ghosts.size:
        own = owner_of(ghosts, ranges)
        for q in np.unique(own):
            sel = np.nonzero(own == q)[0]
            p.requests[int(q)] = (ghosts[sel], (r1 - r0) + int(sel[0]))
    p.ghost_points = gp
    return p


def setup_rank_stencil(pkg, device, shape, nvar, grid, group=None, with_adjoint=True, complex_coupling=True, DiscFlg=False,
                       beta=0.0):
    """row-partitioned handle on a device-generated block-stencil operator (capi.RSVDt.set_operator_stencil)"""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan = plan_stencil(shape, nvar, world, rank)
    r = pkg.capi.RSVDt(device)
    r.dist_init(rank, world, 0)
    r.set_operator_stencil(shape, nvar, (plan.r0 // nvar, plan.r1 // nvar), plan.ghost_points, with_adjoint, complex_coupling,
                           DiscFlg, beta)
    r.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"],
               grid.get("TransientRemoval", False))
    connect(r, plan, group)
    return r, plan


def setup_rank(pkg, device, N, gen_rows, max_bw, layout, band, grid, group=None, DiscFlg=False, beta=0.0,
               with_adjoint=True):
    """one-stop construction of a row-partitioned handle; grid = dict(k, Nw, Ns, Nt, ResRatio, dt, w, TransientRemoval)"""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan, A, AH, im2 = local_operator(N, world, rank, gen_rows, max_bw, layout, band, with_adjoint)
    t = torch.tensor([im2], dtype=torch.float64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, group=group)
    real = float(t.item()) ** 0.5 < 1e-12                      # LNSStructure.c:37-40 on the whole matrix
    r = pkg.capi.RSVDt(device)
    r.dist_init(rank, world, layout)
    r.LNSStructure_local(N, plan.n_own, plan.own0, plan.n_ext, A, AH, real, DiscFlg, beta)
    r.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"],
               grid.get("TransientRemoval", False))
    connect(r, plan, group)
    return r, plan
