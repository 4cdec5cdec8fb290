// stage_tiled.cu -- K1f: fused RK4 stage on the tiled block-CSR operator (stage_tiled.cuh).
//
//   TSRK4.c:18-47 / TSTransRK4.c:15-36: ksum = Op*xin + f ; xout = ybase + cx*ksum ; acc_out = (ybase|acc_in) + ca*ksum.
//
// Why this shape (B200): the block-CSR stage kernel K1e streams 10.8 KB of operator entries per block row (27 blocks of
// 5 x 5 complex) through per-thread 256-bit loads; ncu shows it latency-bound (12 warps/SM, 130 registers, issue 16 %,
// DRAM 38-46 %).  Here the operator never passes through registers on its way in: ONE producer lane per CTA issues
// cp.async.bulk copies of whole 12.9 KB chunks (the p-th block of 32 consecutive block rows + their column indices)
// into a ring of shared-memory stages, completion counted by mbarriers, so up to 100 KB per CTA are in flight whatever
// the consumer warps are doing.  Consumer thread = (block row, two test vectors): it reads the operator entries from
// shared memory (one conflict-free 128-byte wavefront per warp instruction, broadcast over the vector pairs), gathers
// the stage-vector rows with 256-bit loads one block ahead of their use, and keeps B x 2 complex accumulators in
// registers.  Persistent CTAs walk the tiles in an order that puts the halo tiles first.
//
// Bound: HBM.  Algorithmic bytes per block row and stage: nblocks*(B*B*16 + 4) (operator) + B*nv*16*(4..6) (vectors).
#include "stage_tiled.cuh"

#define DEVFN __device__ __forceinline__

namespace {

struct __align__(32) cplx2 { cplx a, b; };

DEVFN void cfma(cplx &s, const cplx a, const cplx b)
{
    s.x = fma(a.x, b.x, s.x);
    s.x = fma(-a.y, b.y, s.x);
    s.y = fma(a.x, b.y, s.y);
    s.y = fma(a.y, b.x, s.y);
}
DEVFN cplx axpy1(double c, const cplx s, const cplx y) { return make_double2(fma(c, s.x, y.x), fma(c, s.y, y.y)); }

// gathered stage-vector rows: coherent loads (ghost rows may have been written by a peer GPU while this kernel runs, after
// the acquire of its flag), cached in L1 -- a row is re-used by up to 27 block rows of the neighbourhood
DEVFN cplx2 ld_x(const cplx *p)
{
    cplx2 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(p) : "memory");
    return v;
}
// read-once block vectors (forcing, y, accumulator): keep them out of L1
DEVFN cplx2 ld_stream(const cplx *p)
{
    cplx2 v;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(p) : "memory");
    return v;
}
DEVFN void st2(cplx *p, const cplx v0, const cplx v1)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v0.x), "d"(v0.y), "d"(v1.x), "d"(v1.y) : "memory");
}
// CV test vectors handled by one thread: CV = 2 moves them with one 256-bit access, CV = 1 with a 128-bit one
template <int CV> struct VecC { cplx v[CV]; };
template <int CV> DEVFN VecC<CV> ldx(const cplx *p)      // gathered rows (coherent, L1-cached)
{
    VecC<CV> o;
    if (CV == 2) { const cplx2 t = ld_x(p); o.v[0] = t.a; o.v[CV - 1] = t.b; }
    else asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(o.v[0].x), "=d"(o.v[0].y) : "l"(p) : "memory");
    return o;
}
template <int CV> DEVFN VecC<CV> lds(const cplx *p)      // read-once block vectors
{
    VecC<CV> o;
    if (CV == 2) { const cplx2 t = ld_stream(p); o.v[0] = t.a; o.v[CV - 1] = t.b; }
    else asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(o.v[0].x), "=d"(o.v[0].y) : "l"(p) : "memory");
    return o;
}
template <int CV> DEVFN void stv(cplx *p, const VecC<CV> &o)
{
    if (CV == 2) st2(p, o.v[0], o.v[CV - 1]);
    else asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(o.v[0].x), "d"(o.v[0].y) : "memory");
}
template <int CV> DEVFN VecC<CV> axpyv(double c, const VecC<CV> &s, const VecC<CV> &y)
{
    VecC<CV> o;
#pragma unroll
    for (int i = 0; i < CV; i++) o.v[i] = axpy1(c, s.v[i], y.v[i]);
    return o;
}

DEVFN unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared-memory load that keeps its place among the barrier operations (volatile, no memory clobber: FMAs may move around it)
DEVFN cplx lds_c(unsigned addr)
{
    cplx v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
DEVFN void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVFN void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEVFN void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
DEVFN void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
DEVFN void tma_load_1d(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction policy: the operator is read once per stage (evict first), the rows of the stage vector are
// re-read by neighbouring tiles (evict last)
DEVFN void tma_load_1d_hint(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar, unsigned long long policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
DEVFN unsigned long long policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
DEVFN unsigned long long policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
DEVFN void l2_prefetch(const void *gsrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}
DEVFN void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }

DEVFN unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DEVFN void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

constexpr long long HALO_SPIN_LIMIT = 600000000ll;   // minutes of polling (a rank may be busy on its host for a while), but a
                                                      // dead peer must not hang the node for good

// ---------------------------------------------------------------------------------------------------------------------
// NCW consumer warps (TB * nv/2 threads, thread = (block row of the tile, pair of test vectors)) + 1 producer warp.
//
// Shared memory: [mbarriers][operator ring: NSA chunks][x ring: NSX windows of WMAX padded block rows].
// Chunk = [TB] block columns, header {xoff, flags, -, -}, [B*B][TB] entries:
//   xoff != TBCSR_GATHER: the block columns of the chunk are affine in the block row (col = base + tb) and their rows of the
//     stage vector sit in the current x window at block (xoff + tb); flags bit0: first chunk of its window (wait for it),
//     bit1: last chunk of its window (release it).  The window [start, start + n) -- one contiguous range of the stage
//     vector, e.g. the 34 blocks the three x-neighbours of 32 consecutive grid points share -- is brought in by the
//     producer warp with n bulk copies (one per block row, padded by 16 B so that the 8 block rows a warp reads at once
//     fall into distinct banks), so each gathered row crosses the L2 -> SM crossbar once per window instead of once per
//     block: 2.8x fewer bytes than per-thread gathers for the 27-point stencil, and no load latency in the FMA loop.
//   xoff == TBCSR_GATHER: general sparsity -- per-thread 256-bit gathers straight from global memory.
// ---------------------------------------------------------------------------------------------------------------------
// VAR (compile-time variant bits; 0 = the kernel as profiled in profiles/r2_lns3d_stage_windows_ncu.txt):
//   1  epilogue: all loads of the tile's block vectors first, then the arithmetic and the stores
//   2  tile bookkeeping and the forcing rows fetched one tile ahead
//   4  setmaxnreg: 8 warps (3 leave at once), the producer warpgroup hands registers to the consumer warpgroup (NCW == 4)
//   8  the operator ring is released and refilled piece by piece (block column j of the TB blocks of a chunk)
//  16  operator entries / x values of block column j + 1 fetched from shared memory before the FMAs of column j
//  32  an x window is ONE bulk copy (unpadded rows; the two vectors of a thread are read in an order that depends on the parity
//      of its block row, which keeps the 128-bit shared-memory reads conflict-free without the 16-byte row padding): the copy
//      engine handles ~one request per 50 cycles whatever its size, and 34 row copies per window kept it busy half of the time
constexpr int VAR_EPI = 1, VAR_NEXT = 2, VAR_SMR = 4, VAR_PIECE = 8, VAR_LDS = 16, VAR_XBULK = 32;

template <int B, int NCW, int MINB, int CV, int VAR>
__global__ void __launch_bounds__((VAR & VAR_SMR) ? 256 : (NCW + 1) * 32, MINB)
rk4_stage_tbcsr_kernel(const TbcsrDev op, const int nv, const int nvp, const int mode, const int nsa, const int nsx,
                       const int wmax, const StageArgs a, const HaloFuse hf)
{
    constexpr int TB = TBCSR_TB;
    constexpr int BB = B * B;
    constexpr unsigned CHUNK = BB * TB * sizeof(cplx) + TB * sizeof(int32_t) + 4 * sizeof(int32_t);
    constexpr int NP = (VAR & VAR_PIECE) ? B : 1;                        // release granularity: pieces per chunk
    constexpr unsigned PIECE = B * TB * sizeof(cplx);                    // block column j of the TB blocks of a chunk
    constexpr unsigned HEAD = TB * sizeof(int32_t) + 4 * sizeof(int32_t);  // block columns + header ride with the first piece
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);   // [nsa][NP] piece landed
    unsigned long long *empty = full + nsa * NP;                                 // [nsa][NP] piece released by all consumer warps
    unsigned long long *xfull = empty + nsa * NP;
    unsigned long long *xempty = xfull + nsx;
    unsigned char *ring = smem + 128 * (((2 * nsa * NP + 2 * nsx) * 8 + 127) / 128);
    const unsigned xstride = (unsigned)(B * nv) * sizeof(cplx) + ((VAR & VAR_XBULK) ? 0u : 16u);   // (padded) block row of the stage vector
    const unsigned xslot = xstride * (unsigned)wmax;
    unsigned char *xring = ring + (size_t)nsa * CHUNK;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < nsa; s++) {
            for (int j = 0; j < NP; j++) {
                mbar_init(full + s * NP + j, 1);
                mbar_init(empty + s * NP + j, NCW);
            }
        }
        for (int s = 0; s < nsx; s++) {
            mbar_init(xfull + s, 1);
            mbar_init(xempty + s, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int64_t ntiles = op.ntiles;
    const int lane = tid & 31;
    if (VAR & VAR_SMR) {
        if (tid >= NCW * 32) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
        if (tid >= (NCW + 1) * 32) return;
    }
    if (tid >= NCW * 32) {
        // ---------------- producer warp: keeps both rings full, running ahead over tile boundaries -------------------------
        // ONE stream of copies in the order the consumers need them (an x window right before the first chunk that reads it):
        // the copy engine of an SM serves its queue in order, so a window requested behind queued operator chunks would arrive
        // after the chunks that need it (measured with a separate window producer: the consumers then wait for windows).
        unsigned s = 0, ph = 0, xs = 0, xph = 0;
        const unsigned long long pol_first = policy_evict_first(), pol_last = policy_evict_last();
        for (int64_t i = blockIdx.x; i < ntiles; i += gridDim.x) {
            const int64_t t = op.tile_order ? op.tile_order[i] : i;
            const int64_t c0 = op.tile_ptr[t], c1 = op.tile_ptr[t + 1];
            int64_t w = op.win_ptr ? op.win_ptr[t] : 0;
            const int64_t w1 = op.win_ptr ? op.win_ptr[t + 1] : 0;
            if (hf.enabled && hf.epoch > 1 && w < w1 && (hf.tile_flags[t] & 1)) {
                // ghost windows are read by the copy engine: the peers' rows of the previous stage must have landed
                if (lane == 0) {
                    for (int q = 0; q < hf.n_in; q++) {
                        long long spins = 0;
                        while (ld_acquire_sys(hf.in_flag[q]) < hf.epoch - 1) {
                            if (++spins > HALO_SPIN_LIMIT) { *hf.timeout_flag = 1ull; break; }
                        }
                    }
                    asm volatile("fence.proxy.async.global;\n" ::: "memory");
                }
                __syncwarp();
            }
            if (lane == 0 && c1 > c0 && (op.flags & 32)) {
                // the tile's rows of f, y and the accumulator are read by the consumers long after this point: have L2 fetch them now
                const int64_t rows = (t * TB + TB <= op.nbrows ? TB : op.nbrows - t * TB) * (int64_t)B;
                const unsigned bytes = (unsigned)(rows * nv * (int64_t)sizeof(cplx));
                const int64_t e0 = t * TB * (int64_t)B * nv;
                if (a.f) l2_prefetch(a.f + e0, bytes);
                if (mode != 2) l2_prefetch(a.ybase + e0, bytes);
                if (mode != 0) l2_prefetch(a.acc_in + e0, bytes);
            }
            for (int64_t c = c0; c < c1; c++) {
                while (w < w1 && op.win[w].first_chunk == (int32_t)(c - c0)) {
                    const TbcsrWindow wd = op.win[w];
                    if (lane == 0) {
                        mbar_wait(xempty + xs, xph ^ 1u);
                        mbar_expect_tx(xfull + xs, (unsigned)wd.nblocks * (unsigned)(B * nv) * (unsigned)sizeof(cplx));
                    }
                    __syncwarp();
                    const cplx *src = (wd.start < op.nbcols_own) ? a.xin + (int64_t)wd.start * B * nv
                                                                  : hf.ghost_src + ((int64_t)wd.start - op.nbcols_own) * B * nv;
                    if (VAR & VAR_XBULK) {
                        if (lane == 0) {
                            const unsigned bytes = (unsigned)wd.nblocks * (unsigned)(B * nv) * (unsigned)sizeof(cplx);
                            if (op.flags & 2) tma_load_1d_hint(xring + (size_t)xs * xslot, src, bytes, xfull + xs, pol_last);
                            else tma_load_1d(xring + (size_t)xs * xslot, src, bytes, xfull + xs);
                        }
                    } else
                    for (int b = lane; b < wd.nblocks; b += 32) {
                        if (op.flags & 2)
                            tma_load_1d_hint(xring + (size_t)xs * xslot + (size_t)b * xstride, src + (int64_t)b * B * nv,
                                             (unsigned)(B * nv) * (unsigned)sizeof(cplx), xfull + xs, pol_last);
                        else
                            tma_load_1d(xring + (size_t)xs * xslot + (size_t)b * xstride, src + (int64_t)b * B * nv,
                                        (unsigned)(B * nv) * (unsigned)sizeof(cplx), xfull + xs);
                    }
                    if (++xs == (unsigned)nsx) { xs = 0; xph ^= 1u; }
                    w++;
                }
                if (lane == 0) {
                    const unsigned char *src = op.chunks + (size_t)c * CHUNK;
                    unsigned char *dst = ring + (size_t)s * CHUNK;
                    if (VAR & VAR_PIECE) {
#pragma unroll 1
                        for (int j = 0; j < B; j++) {
                            // piece 0 = columns + header + block column 0, piece j = block column j
                            const unsigned off = j ? HEAD + j * PIECE : 0u, bytes = j ? PIECE : HEAD + PIECE;
                            mbar_wait(empty + s * NP + j, ph ^ 1u);
                            mbar_expect_tx(full + s * NP + j, bytes);
                            if (op.flags & 1) tma_load_1d_hint(dst + off, src + off, bytes, full + s * NP + j, pol_first);
                            else tma_load_1d(dst + off, src + off, bytes, full + s * NP + j);
                        }
                    } else {
                        mbar_wait(empty + s, ph ^ 1u);
                        mbar_expect_tx(full + s, CHUNK);
                        if (op.flags & 1) tma_load_1d_hint(dst, src, CHUNK, full + s, pol_first);
                        else tma_load_1d(dst, src, CHUNK, full + s);
                    }
                }
                if (VAR & VAR_PIECE) __syncwarp();
                if (++s == (unsigned)nsa) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------------------------------------------------------------------------------
    const int tb = tid / nvp;
    const int iv0 = (tid - tb * nvp) * CV;
    const bool lane_on = tb < TB;
    const int tbs = lane_on ? tb : 0;
    const bool has_f = a.f != nullptr && !(op.flags & 16);
    const unsigned sw = (VAR & VAR_XBULK) ? (unsigned)(tbs & 1) : 0u;   // odd block rows read their second vector first
    unsigned s = 0, ph = 0, xs = 0, xph = 0;

    auto tile_of = [&](int64_t i) -> int64_t { return op.tile_order ? (int64_t)op.tile_order[i] : i; };
    // VAR_NEXT: tile bookkeeping (order, chunk count, halo flags) and the forcing rows are fetched one tile ahead of their use
    int64_t t_n = 0;
    int nchunk_n = 0;
    unsigned char tflags_n = 0;
    VecC<CV> fnext[B];
    if ((VAR & VAR_NEXT) && (int64_t)blockIdx.x < ntiles) {
        t_n = tile_of(blockIdx.x);
        nchunk_n = (int)(op.tile_ptr[t_n + 1] - op.tile_ptr[t_n]);
        if (hf.enabled) tflags_n = hf.tile_flags[t_n];
        const int64_t brn = t_n * TB + tbs;
        if (has_f && lane_on && brn < op.nbrows) {
#pragma unroll
            for (int r = 0; r < B; r++) fnext[r] = lds<CV>(a.f + brn * B * nv + iv0 + r * nv);
        }
    }

    for (int64_t i = blockIdx.x; i < ntiles; i += gridDim.x) {
        int64_t t;
        int nchunk;
        unsigned char tflags = 0;
        const int64_t inext = i + gridDim.x;
        if (VAR & VAR_NEXT) {
            t = t_n;
            nchunk = nchunk_n;
            tflags = tflags_n;
            if (inext < ntiles) {
                t_n = tile_of(inext);
                nchunk_n = (int)(op.tile_ptr[t_n + 1] - op.tile_ptr[t_n]);
                if (hf.enabled) tflags_n = hf.tile_flags[t_n];
            }
        } else {
            t = tile_of(i);
            nchunk = (int)(op.tile_ptr[t + 1] - op.tile_ptr[t]);
            if (hf.enabled) tflags = hf.tile_flags[t];
        }
        const int64_t br = t * TB + tbs;
        const bool row_on = lane_on && br < op.nbrows;
        const int64_t u0 = br * B * nv + iv0;

        if (hf.enabled && tflags && hf.epoch > 1) {
            // wait until every incoming peer has published epoch E-1 (their rows of the previous stage are in my mailbox,
            // and they are done reading the slot this launch's senders overwrite)
            if (lane == 0) {
                for (int q = 0; q < hf.n_in; q++) {
                    long long spins = 0;
                    while (ld_acquire_sys(hf.in_flag[q]) < hf.epoch - 1) {
                        if (++spins > HALO_SPIN_LIMIT) { *hf.timeout_flag = 1ull; break; }
                    }
                }
            }
            __syncwarp();
            asm volatile("fence.acq_rel.sys;\n" ::: "memory");
        }

        VecC<CV> acc[B];
#pragma unroll
        for (int r = 0; r < B; r++) {
#pragma unroll
            for (int q = 0; q < CV; q++) acc[r].v[q] = make_double2(0.0, 0.0);
            if (has_f && row_on) {
                if (VAR & VAR_NEXT) acc[r] = fnext[r];
                else acc[r] = lds<CV>(a.f + u0 + r * nv);
            }
        }

        for (int c = 0; c < nchunk; c++) {
            mbar_wait(full + s * NP, ph);
            unsigned long long *ful = full + s * NP;    // (pieces 1.. are awaited right before their use)
            const unsigned char *chunk = ring + (size_t)s * CHUNK;
            const int32_t *cols = reinterpret_cast<const int32_t *>(chunk);
            const int xoff = nsx ? cols[TB] : TBCSR_GATHER, cflags = cols[TB + 1];   // (no x ring: every chunk gathers)
            const int col = (row_on && !(op.flags & 8)) ? cols[tbs] : -1;   // (flag 8: measurement of the memory pipeline alone)
            const cplx *av = reinterpret_cast<const cplx *>(chunk + HEAD) + tbs;
            const unsigned avs = smem_u32(chunk) + HEAD + (unsigned)tbs * 16u;      // entry e of my block: avs + e * TB * 16
            unsigned long long *rel = empty + s * NP;
            const bool on = col >= 0;
            if (xoff != TBCSR_GATHER) {
                if (cflags & 1) mbar_wait(xfull + xs, xph);
                const unsigned char *xw = xring + (size_t)xs * xslot + (size_t)(xoff + tbs) * xstride + (size_t)iv0 * sizeof(cplx);
                if (VAR & VAR_LDS) {
                    // ordered shared-memory loads, one block column ahead of the FMAs
                    const unsigned xws = smem_u32(xw);
                    cplx e[2][B], xv[2][CV];
                    if (on) {
#pragma unroll
                        for (int r = 0; r < B; r++) e[0][r] = lds_c(avs + (unsigned)(r * TB) * 16u);
#pragma unroll
                        for (int q = 0; q < CV; q++) xv[0][q] = lds_c(xws + ((unsigned)q ^ sw) * 16u);
                    }
#pragma unroll
                    for (int j = 0; j < B; j++) {
                        if ((VAR & VAR_PIECE) && j + 1 < B) mbar_wait(ful + j + 1, ph);
                        if (on) {
                            if (j + 1 < B) {
#pragma unroll
                                for (int r = 0; r < B; r++) e[(j + 1) & 1][r] = lds_c(avs + (unsigned)(((j + 1) * B + r) * TB) * 16u);
#pragma unroll
                                for (int q = 0; q < CV; q++) xv[(j + 1) & 1][q] = lds_c(xws + ((unsigned)((j + 1) * nv + q) ^ sw) * 16u);
                            }
                            cplx xq[CV];
#pragma unroll
                            for (int q = 0; q < CV; q++) xq[q] = xv[j & 1][q];
                            if ((VAR & VAR_XBULK) && CV == 2 && sw) { const cplx t0 = xq[0]; xq[0] = xq[CV - 1]; xq[CV - 1] = t0; }
#pragma unroll
                            for (int r = 0; r < B; r++) {
#pragma unroll
                                for (int q = 0; q < CV; q++) cfma(acc[r].v[q], e[j & 1][r], xq[q]);
                            }
                        }
                        if (VAR & VAR_PIECE) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(rel + j);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < B; j++) {
                        if ((VAR & VAR_PIECE) && j > 0) mbar_wait(ful + j, ph);
                        if (on) {
                            const cplx *xp = reinterpret_cast<const cplx *>(xw + (size_t)j * nv * sizeof(cplx));
                            cplx xv[CV];
#pragma unroll
                            for (int q = 0; q < CV; q++) xv[q] = xp[(unsigned)q ^ sw];
                            if ((VAR & VAR_XBULK) && CV == 2 && sw) { const cplx t0 = xv[0]; xv[0] = xv[CV - 1]; xv[CV - 1] = t0; }
#pragma unroll
                            for (int r = 0; r < B; r++) {
                                const cplx e = av[(j * B + r) * TB];
#pragma unroll
                                for (int q = 0; q < CV; q++) cfma(acc[r].v[q], e, xv[q]);
                            }
                        }
                        if (VAR & VAR_PIECE) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(rel + j);
                        }
                    }
                }
            } else {
                VecC<CV> xr[B];
                if (on) {
                    const cplx *xp = (col < op.nbcols_own) ? a.xin + (int64_t)col * B * nv + iv0
                                                           : hf.ghost_src + ((int64_t)col - op.nbcols_own) * B * nv + iv0;
#pragma unroll
                    for (int j = 0; j < B; j++) xr[j] = ldx<CV>(xp + j * nv);
                }
#pragma unroll
                for (int j = 0; j < B; j++) {
                    if ((VAR & VAR_PIECE) && j > 0) mbar_wait(ful + j, ph);
                    if (on) {
#pragma unroll
                        for (int r = 0; r < B; r++) {
                            const cplx e = av[(j * B + r) * TB];
#pragma unroll
                            for (int q = 0; q < CV; q++) cfma(acc[r].v[q], e, xr[j].v[q]);
                        }
                    }
                    if (VAR & VAR_PIECE) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(rel + j);
                    }
                }
            }
            if (!(VAR & VAR_PIECE)) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            if (xoff != TBCSR_GATHER && (cflags & 2)) {
                if (lane == 0) mbar_arrive(xempty + xs);   // (after a __syncwarp: the whole warp is done with the window)
                if (++xs == (unsigned)nsx) { xs = 0; xph ^= 1u; }
            }
            if (++s == (unsigned)nsa) { s = 0; ph ^= 1u; }
        }

        // ---------------- epilogue: stage AXPYs, local stores, peer stores of the halo rows -----------------------------
        int64_t dest[2] = {-1, -1};
        if (hf.enabled && (tflags & 2) && row_on) {
            const int64_t si = hf.send_index[t];
            dest[0] = hf.send_dest[(si * 2 + 0) * TB + tbs];
            dest[1] = hf.send_dest[(si * 2 + 1) * TB + tbs];
        }
        if (row_on && !(op.flags & 16)) {
            VecC<CV> yv[B], av2[B];
            if (VAR & VAR_EPI) {
                // all loads of the tile's block vectors first (one exposed latency instead of B), then the arithmetic and the stores
                if (mode != 2) {
#pragma unroll
                    for (int r = 0; r < B; r++) yv[r] = lds<CV>(a.ybase + u0 + r * nv);
                }
                if (mode != 0) {
#pragma unroll
                    for (int r = 0; r < B; r++) av2[r] = lds<CV>(a.acc_in + u0 + r * nv);
                }
            }
#pragma unroll
            for (int r = 0; r < B; r++) {
                const int64_t u = u0 + r * nv;
                if (!(VAR & VAR_EPI)) {
                    if (mode != 2) yv[r] = lds<CV>(a.ybase + u);
                    if (mode != 0) av2[r] = lds<CV>(a.acc_in + u);
                }
                VecC<CV> o;   // the vector the next stage gathers (xout, or y_{n+1} in the last stage)
                if (mode == 0) {
                    o = axpyv<CV>(a.cx, acc[r], yv[r]);
                    stv<CV>(a.xout + u, o);
                    stv<CV>(a.acc_out + u, axpyv<CV>(a.ca, acc[r], yv[r]));
                } else if (mode == 1) {
                    o = axpyv<CV>(a.cx, acc[r], yv[r]);
                    stv<CV>(a.xout + u, o);
                    stv<CV>(a.acc_out + u, axpyv<CV>(a.ca, acc[r], av2[r]));
                } else {
                    o = axpyv<CV>(a.ca, acc[r], av2[r]);
                    stv<CV>(a.acc_out + u, o);
                    if (a.snap) stv<CV>(a.snap + u, o);
                }
#pragma unroll
                for (int d = 0; d < 2; d++)
                    if (dest[d] >= 0) {
                        const int slot = (int)(dest[d] >> 48);
                        const int64_t grow = dest[d] & 0xFFFFFFFFFFFFll;   // first ghost row (on the peer) of this block row
                        stv<CV>(hf.out_mail[slot] + (grow + r) * nv + iv0, o);
                    }
            }
        }
        if ((VAR & VAR_NEXT) && has_f && inext < ntiles) {
            // the forcing rows of the next tile: in flight while this tile's sends are counted and the next chunk is awaited
            const int64_t brn = t_n * TB + tbs;
            if (lane_on && brn < op.nbrows) {
#pragma unroll
                for (int r = 0; r < B; r++) fnext[r] = lds<CV>(a.f + brn * B * nv + iv0 + r * nv);
            }
        }
        if (hf.enabled && (tflags & 2)) {
            // all rows of this tile are on their way: count the tile; the last sending tile of the launch publishes the epoch
            __threadfence_system();
            named_bar_sync(1, NCW * 32);
            if (tid == 0) {
                const unsigned mask = (unsigned)hf.send_mask[hf.send_index[t]];
                for (int p = 0; p < hf.n_out; p++)
                    if (mask & (1u << p)) {
                        const unsigned long long done = atomicAdd(hf.out_done + p, 1ull) + 1ull;
                        if (done == hf.epoch * (unsigned long long)hf.out_tiles[p]) {
                            __threadfence_system();
                            st_release_sys(hf.out_flag[p], hf.epoch);
                        }
                    }
            }
        }
    }
}

// variants built for the B = 5, k = 8 shape (the 3-D LNS configuration of BASELINE.json); every other shape runs variant 0
#define TBCSR_VARIANTS(X) X(1) X(4) X(8) X(16) X(17) X(32) X(33) X(37) X(40) X(41) X(44) X(45) X(49) X(53) X(57) X(61)

template <int B, int NCW, int CV>
cudaError_t launch_t(const TbcsrDev &Op, int nv, int mode, const StageArgs &a, const HaloFuse &hf, int ctas_per_sm, int nsa_req,
                     int variant, cudaStream_t s)
{
    static int n_sm = 0;
    static size_t smem_max = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        smem_max = (size_t)v;
    }
    constexpr bool has_variants = B == 5 && NCW == 4 && CV == 2;
    if (!has_variants) variant = 0;
    if (ctas_per_sm < 1 || ctas_per_sm > 3) ctas_per_sm = 2;
    if ((variant & VAR_SMR) && ctas_per_sm != 2) variant &= ~VAR_SMR;
    const size_t chunk = tbcsr_chunk_bytes(B, TBCSR_TB);
    if (chunk != Op.chunk_bytes) return cudaErrorInvalidValue;
    // shared memory per SM: 228 KB, 1 KB reserved per CTA; the last KB of the CTA's share: mbarriers
    const size_t budget = std::min(smem_max, (size_t)(228 * 1024) / ctas_per_sm - 1024) - 1024;
    const size_t xslot = ((size_t)B * nv * sizeof(cplx) + ((variant & VAR_XBULK) ? 0 : 16)) * (size_t)Op.wmax;
    int nsx = 0, nsa = 0;
    if (Op.win_ptr && Op.wmax > 0 && 2 * xslot + 2 * chunk > budget && ctas_per_sm == 3) return launch_t<B, NCW, CV>(Op, nv, mode, a, hf, 2, nsa_req, variant, s);
    if (Op.win_ptr && Op.wmax > 0 && 2 * xslot + 2 * chunk <= budget) nsx = 2;
    TbcsrDev op = Op;
    if (!nsx) { op.win_ptr = nullptr; op.win = nullptr; }   // the x windows do not fit next to the operator ring: gather mode
    nsa = (int)((budget - (size_t)nsx * xslot) / chunk);
    if (nsa > 12) nsa = 12;
    if (nsa_req >= 2 && nsa_req < nsa) nsa = nsa_req;
    if (nsa < 2) return cudaErrorInvalidValue;
    const int np = (variant & VAR_PIECE) ? B : 1;
    const size_t smem = 128 * (((size_t)(2 * nsa * np + 2 * nsx) * 8 + 127) / 128) + (size_t)nsa * chunk + (size_t)nsx * xslot;
    const int threads = (variant & VAR_SMR) ? 256 : (NCW + 1) * 32;
    const unsigned grid = (unsigned)std::min<int64_t>(Op.ntiles, (int64_t)n_sm * ctas_per_sm);
    if (grid == 0) return cudaSuccess;
    const int nvp = nv / CV;
    cudaError_t e;
#define GO(MINB, VAR)                                                                                                         \
    do {                                                                                                                      \
        e = cudaFuncSetAttribute(rk4_stage_tbcsr_kernel<B, NCW, MINB, CV, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                       \
        rk4_stage_tbcsr_kernel<B, NCW, MINB, CV, VAR><<<grid, threads, smem, s>>>(op, nv, nvp, mode, nsa, nsx, Op.wmax, a, hf);  \
    } while (0)
    if (has_variants && variant != 0 && ctas_per_sm <= 2) {
        switch (variant) {
#define CASE(V) case V: GO(2, (has_variants ? V : 0)); break;
            TBCSR_VARIANTS(CASE)
#undef CASE
            default: return cudaErrorInvalidValue;
        }
    } else if (ctas_per_sm <= 2) GO(2, 0);   // (1 CTA/SM runs the 2-CTA build: same registers, deeper ring)
    else GO(3, 0);
#undef GO
    return cudaGetLastError();
}

#ifndef TBCSR_DEV_ONE
template <int B>
cudaError_t launch_b(const TbcsrDev &Op, int nv, int mode, const StageArgs &a, const HaloFuse &hf, int cps, int nsa, int cv, int var, cudaStream_t s)
{
    // consumer warps: TB block rows x (nv / CV) threads each.  CV = 1 doubles the warps per CTA at twice the shared-memory reads
    // per FMA (measured slower: 1.67 vs 1.58 ms on the 3-D operator); it serves odd vector counts and is there on request.
    if (cv == 1 || nv % 2) {
        switch (nv) {
            case 1: return launch_t<B, 1, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 2: return launch_t<B, 2, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 3: return launch_t<B, 3, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 4: return launch_t<B, 4, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 5: return launch_t<B, 5, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 6: return launch_t<B, 6, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 7: return launch_t<B, 7, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            case 8: return launch_t<B, 8, 1>(Op, nv, mode, a, hf, cps, nsa, var, s);
            default: break;
        }
        if (nv % 2) return cudaErrorInvalidValue;
    }
    switch (nv / 2) {
        case 1: return launch_t<B, 1, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 2: return launch_t<B, 2, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 3: return launch_t<B, 3, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 4: return launch_t<B, 4, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 5: return launch_t<B, 5, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 6: return launch_t<B, 6, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 7: return launch_t<B, 7, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        case 8: return launch_t<B, 8, 2>(Op, nv, mode, a, hf, cps, nsa, var, s);
        default: return cudaErrorInvalidValue;
    }
}
#endif

}  // namespace

bool tbcsr_stage_supported(int B, int nv)
{
    if (!(B == 3 || B == 4 || B == 5) || nv < 1) return false;
    return nv <= 8 || (nv % 2 == 0 && nv <= 16);   // one vector per thread up to 8, two per thread up to 16
}

cudaError_t launch_rk4_stage_tbcsr(const TbcsrDev &Op, int nv, int mode, const StageArgs &a, const HaloFuse &hf, int ctas_per_sm,
                                   int ring_stages, int vectors_per_thread, int variant, cudaStream_t s)
{
    if (!tbcsr_stage_supported(Op.B, nv)) return cudaErrorInvalidValue;
#ifdef TBCSR_DEV_ONE   // development aid: build the B = 5, k = 8 instances alone (ptxas -v / SASS in seconds)
    return launch_t<5, 4, 2>(Op, nv, mode, a, hf, ctas_per_sm, ring_stages, variant, s);
#else
    switch (Op.B) {
        case 3: return launch_b<3>(Op, nv, mode, a, hf, ctas_per_sm, ring_stages, vectors_per_thread, variant, s);
        case 4: return launch_b<4>(Op, nv, mode, a, hf, ctas_per_sm, ring_stages, vectors_per_thread, variant, s);
        case 5: return launch_b<5>(Op, nv, mode, a, hf, ctas_per_sm, ring_stages, vectors_per_thread, variant, s);
        default: return cudaErrorInvalidValue;
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// x windows: which chunks read their rows of the stage vector from a shared-memory window, and which windows a tile needs
// ---------------------------------------------------------------------------------------------------------------------
void tbcsr_plan_windows(int64_t ntiles, const int64_t *tile_ptr, const int32_t *cols, int64_t nb_own, int64_t nb_total, int TB,
                        int wmax, TbcsrPlan &out)
{
    (void)nb_total;
    const int64_t nchunks = tile_ptr[ntiles];
    out.hdr.assign((size_t)nchunks * 4, 0);
    out.win_ptr.assign((size_t)ntiles + 1, 0);
    out.win.clear();
    out.window_chunks = out.gather_chunks = 0;
    for (int64_t t = 0; t < ntiles; t++) {
        int64_t cur = -1, cur_last = -1;   // open window (index into out.win) and the last chunk that used it
        bool cur_ghost = false;
        auto close = [&]() {
            if (cur >= 0 && cur_last >= 0) out.hdr[(size_t)cur_last * 4 + 1] |= 2;
            cur = -1;
            cur_last = -1;
        };
        for (int64_t p = tile_ptr[t]; p < tile_ptr[t + 1]; p++) {
            const int32_t *c = cols + p * TB;
            bool any = false, affine = true;
            int64_t base = 0, lo = 0, hi = 0;
            for (int tb = 0; tb < TB; tb++) {
                if (c[tb] < 0) continue;
                const int64_t b = (int64_t)c[tb] - tb;
                if (!any) { any = true; base = b; lo = c[tb]; }
                else if (b != base) affine = false;
                hi = (int64_t)c[tb] + 1;
            }
            bool ok = any && affine && hi - lo <= wmax;
            const bool ghost = ok && lo >= nb_own;
            if (ok && !ghost && hi > nb_own) ok = false;   // straddles owned rows and ghost rows: two sources
            if (!ok) {
                out.hdr[(size_t)p * 4] = TBCSR_GATHER;
                if (any) out.gather_chunks++;
                continue;
            }
            if (cur >= 0 && cur_ghost == ghost && lo >= out.win[(size_t)cur].start && hi <= (int64_t)out.win[(size_t)cur].start + wmax) {
                out.win[(size_t)cur].nblocks = (int32_t)std::max<int64_t>(out.win[(size_t)cur].nblocks, hi - out.win[(size_t)cur].start);
            } else {
                close();
                TbcsrWindow w;
                w.start = (int32_t)lo;
                w.nblocks = (int32_t)(hi - lo);
                w.first_chunk = (int32_t)(p - tile_ptr[t]);
                w.pad = 0;
                out.win.push_back(w);
                cur = (int64_t)out.win.size() - 1;
                cur_ghost = ghost;
                out.hdr[(size_t)p * 4 + 1] |= 1;
            }
            out.hdr[(size_t)p * 4] = (int32_t)(base - out.win[(size_t)cur].start);
            cur_last = p;
            out.window_chunks++;
        }
        close();
        out.win_ptr[(size_t)t + 1] = (int64_t)out.win.size();
    }
}
