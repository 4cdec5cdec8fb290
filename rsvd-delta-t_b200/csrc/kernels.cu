// kernels.cu -- hand-written sm_100a kernels of the RSVD-Delta-t time-stepping path.
//
// Everything here is FP64 complex arithmetic on HBM-resident data; the roofline that bounds each kernel is
// stated next to it (DESIGN.md section 4 has the byte counts).  No tensor-core reshaping: the SpMM stage is a
// gather/stream kernel, the Fourier contractions are register-tiled FP64 FMA loops.
#include "kernels.cuh"

#include <math.h>

#define DEVFN __device__ __forceinline__

// s += a*b (complex), 4 FMAs; the order (re*re first, then -im*im) is fixed so results are reproducible
DEVFN void cfma(cplx &s, const cplx a, const cplx b)
{
    s.x = fma(a.x, b.x, s.x);
    s.x = fma(-a.y, b.y, s.x);
    s.y = fma(a.x, b.y, s.y);
    s.y = fma(a.y, b.x, s.y);
}
// s += conj(a)*b
DEVFN void cfma_conj(cplx &s, const cplx a, const cplx b)
{
    s.x = fma(a.x, b.x, s.x);
    s.x = fma(a.y, b.y, s.x);
    s.y = fma(a.x, b.y, s.y);
    s.y = fma(-a.y, b.x, s.y);
}
DEVFN cplx cmul(const cplx a, const cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DEVFN cplx ldg(const cplx *p) { return __ldg(p); }
// read-once data (operator entries, forcing): do not allocate in L1, the gathered stage-vector rows live there
DEVFN cplx ldg_stream(const cplx *p)
{
    cplx v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
DEVFN int32_t ldg_stream_i32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// =====================================================================================================
// K1: one RK4 stage = CSR SpMM over nv columns + forcing add + stage AXPYs   (TSRK4.c:18-47, TSTransRK4.c:15-36)
//   thread <-> unit u = row*nv + iv.  The nv threads of a row read the same A entry (one broadcast sector) and
//   nv*16 contiguous bytes of the gathered row of xin, so every global access is coalesced.
//   Bound: HBM.  Algorithmic bytes per unit and stage: nnz_row*20/nv + 8/nv (A) + 16*(reads+writes of block vectors).
// =====================================================================================================
template <int MODE, bool HAS_F>
__global__ void __launch_bounds__(256)
rk4_stage_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                     int nv, int64_t u_begin, int64_t n_units, StageArgs a)
{
    const int64_t u = u_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // rows [row_begin, row_begin + count)
    if (u >= u_begin + n_units) return;
    const int64_t row = u / nv;
    const int iv = (int)(u - row * nv);
    cplx s = make_double2(0.0, 0.0);
    if (HAS_F) s = ldg(a.f + u);
    int64_t p = __ldg(rowptr + row);
    const int64_t p1 = __ldg(rowptr + row + 1);
    const cplx *__restrict__ xin = a.xin + iv;
    // 4-way unrolled: issue the loads of four entries before the dependent FMA chain (memory-level parallelism)
    for (; p + 4 <= p1; p += 4) {
        const int32_t c0 = __ldg(col + p), c1 = __ldg(col + p + 1), c2 = __ldg(col + p + 2), c3 = __ldg(col + p + 3);
        const cplx a0 = ldg(val + p), a1 = ldg(val + p + 1), a2 = ldg(val + p + 2), a3 = ldg(val + p + 3);
        const cplx x0 = ldg(xin + (int64_t)c0 * nv), x1 = ldg(xin + (int64_t)c1 * nv);
        const cplx x2 = ldg(xin + (int64_t)c2 * nv), x3 = ldg(xin + (int64_t)c3 * nv);
        cfma(s, a0, x0);
        cfma(s, a1, x1);
        cfma(s, a2, x2);
        cfma(s, a3, x3);
    }
    for (; p < p1; ++p) cfma(s, ldg(val + p), ldg(xin + (int64_t)__ldg(col + p) * nv));

    if (MODE == 0) {
        const cplx y = ldg(a.ybase + u);
        a.xout[u] = make_double2(fma(a.cx, s.x, y.x), fma(a.cx, s.y, y.y));
        a.acc_out[u] = make_double2(fma(a.ca, s.x, y.x), fma(a.ca, s.y, y.y));
    } else if (MODE == 1) {
        const cplx y = ldg(a.ybase + u);
        const cplx ac = a.acc_in[u];
        a.xout[u] = make_double2(fma(a.cx, s.x, y.x), fma(a.cx, s.y, y.y));
        a.acc_out[u] = make_double2(fma(a.ca, s.x, ac.x), fma(a.ca, s.y, ac.y));
    } else {
        const cplx ac = a.acc_in[u];
        const cplx yn = make_double2(fma(a.ca, s.x, ac.x), fma(a.ca, s.y, ac.y));
        a.acc_out[u] = yn;
        if (a.snap) a.snap[u] = yn;
    }
}

// =====================================================================================================
// K1e: the same stage on a block-CSR (BCSR) operator with dense B x B blocks (LNS operators: B = number of flow variables
//   per grid point).  thread <-> (block row, CV consecutive test vectors): one column index per block instead of per entry
//   (16.16 instead of 20 bytes per nonzero), each gathered x value feeds B rows and each operator entry feeds CV test
//   vectors from registers -- the CSR kernel is bound by the L1 -> register-file fill rate (ncu: L1 wavefronts 72 %, DRAM
//   30 % on the 3-D operator), this one needs (16/CV + 16/B) instead of 36 bytes of fill per complex FMA.
//   Summation order per row = ascending columns, identical to the CSR kernel (bit-identical results).
// =====================================================================================================
// two complex numbers moved by one 256-bit access (LDG.E.256 / STG.E.256 on sm_100a)
struct __align__(32) cplx2 { cplx a, b; };
DEVFN cplx2 ldg2(const cplx *p)
{
    cplx2 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(p));
    return v;
}
// streaming variant for the operator entries (read exactly once): keep them out of L1 so the gathered x rows stay there
DEVFN cplx2 ldg2_stream(const cplx *p)
{
    cplx2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y) : "l"(p));
    return v;
}
DEVFN void st2(cplx *p, const cplx v0, const cplx v1)
{
    cplx2 v;
    v.a = v0;
    v.b = v1;
    *reinterpret_cast<cplx2 *>(p) = v;
}
DEVFN cplx axpy1(double c, const cplx s, const cplx y) { return make_double2(fma(c, s.x, y.x), fma(c, s.y, y.y)); }

// BS = padded block size in complex entries (B*B rounded up to even so that every block starts 32-byte aligned)
template <int B, int CV, int MODE, bool HAS_F>
__global__ void __launch_bounds__(128, 4)
rk4_stage_bcsr_kernel(const int64_t *__restrict__ browptr, const int32_t *__restrict__ bcol, const cplx *__restrict__ bval,
                      int nv, int nvp, int64_t br_begin, int64_t n_threads, StageArgs a)
{
    constexpr int BS = (B * B + 1) & ~1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const int64_t br = br_begin + t / nvp;
    const int iv0 = (int)(t % nvp) * CV;
    const int64_t u0 = br * B * nv + iv0;   // unit of (first row of the block row, iv0); next row: + nv
    cplx s[B][CV];
#pragma unroll
    for (int r = 0; r < B; r++) {
        if (CV == 2) {
            cplx2 v;
            v.a = v.b = make_double2(0.0, 0.0);
            if (HAS_F) v = ldg2_stream(a.f + u0 + r * nv);
            s[r][0] = v.a;
            s[r][CV - 1] = v.b;
        } else {
            s[r][0] = HAS_F ? ldg(a.f + u0 + r * nv) : make_double2(0.0, 0.0);
        }
    }
    const int64_t p1 = __ldg(browptr + br + 1);
    int64_t p = __ldg(browptr + br);
    int32_t cnext = (p < p1) ? __ldg(bcol + p) : 0;
    for (; p < p1; ++p) {
        const cplx *__restrict__ xb = a.xin + (int64_t)cnext * B * nv + iv0;
        if (p + 1 < p1) cnext = __ldg(bcol + p + 1);   // the next block's column index is in flight during this block's FMAs
        const cplx *__restrict__ av = bval + p * BS;
        // Block entries are stored column by column ([j][r]) and fetched two at a time; only one gathered x row (plus the
        // prefetched next one) is live, which keeps the register count low enough for 4-5 CTAs per SM.
        cplx xc[CV], xn[CV];
        if (CV == 2) { const cplx2 v = ldg2(xb); xn[0] = v.a; xn[CV - 1] = v.b; } else { xn[0] = ldg(xb); }
#pragma unroll
        for (int e = 0; e < B * B; e += 2) {
            const cplx2 aa = ldg2_stream(av + e);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ee = e + h;
                if (ee < B * B) {
                    const int j = ee / B, r = ee % B;
                    if (r == 0) {
#pragma unroll
                        for (int c = 0; c < CV; c++) xc[c] = xn[c];
                        if (j + 1 < B) {
                            if (CV == 2) { const cplx2 v = ldg2(xb + (j + 1) * nv); xn[0] = v.a; xn[CV - 1] = v.b; }
                            else { xn[0] = ldg(xb + (j + 1) * nv); }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CV; c++) cfma(s[r][c], h ? aa.b : aa.a, xc[c]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < B; r++) {
        const int64_t u = u0 + r * nv;
        if (CV == 2) {
            const cplx s0 = s[r][0], s1 = s[r][CV - 1];
            if (MODE == 0) {
                const cplx2 y = ldg2_stream(a.ybase + u);
                st2(a.xout + u, axpy1(a.cx, s0, y.a), axpy1(a.cx, s1, y.b));
                st2(a.acc_out + u, axpy1(a.ca, s0, y.a), axpy1(a.ca, s1, y.b));
            } else if (MODE == 1) {
                const cplx2 y = ldg2_stream(a.ybase + u);
                const cplx2 ac = *reinterpret_cast<const cplx2 *>(a.acc_in + u);
                st2(a.xout + u, axpy1(a.cx, s0, y.a), axpy1(a.cx, s1, y.b));
                st2(a.acc_out + u, axpy1(a.ca, s0, ac.a), axpy1(a.ca, s1, ac.b));
            } else {
                const cplx2 ac = *reinterpret_cast<const cplx2 *>(a.acc_in + u);
                const cplx y0 = axpy1(a.ca, s0, ac.a), y1 = axpy1(a.ca, s1, ac.b);
                st2(a.acc_out + u, y0, y1);
                if (a.snap) st2(a.snap + u, y0, y1);
            }
        } else {
            const cplx sv = s[r][0];
            if (MODE == 0) {
                const cplx y = ldg(a.ybase + u);
                a.xout[u] = axpy1(a.cx, sv, y);
                a.acc_out[u] = axpy1(a.ca, sv, y);
            } else if (MODE == 1) {
                const cplx y = ldg(a.ybase + u);
                const cplx ac = a.acc_in[u];
                a.xout[u] = axpy1(a.cx, sv, y);
                a.acc_out[u] = axpy1(a.ca, sv, ac);
            } else {
                const cplx yn = axpy1(a.ca, sv, a.acc_in[u]);
                a.acc_out[u] = yn;
                if (a.snap) a.snap[u] = yn;
            }
        }
    }
}

// CSR stage with two test vectors per thread (nv even): one operator entry and one 256-bit gather feed two complex FMAs,
// halving the load instructions per unit for operators with long rows and no block structure.
template <int MODE, bool HAS_F>
__global__ void __launch_bounds__(256)
rk4_stage_csr2_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                      int nv, int nvp, int64_t t_begin, int64_t n_threads, StageArgs a)
{
    const int64_t t = t_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_begin + n_threads) return;
    const int64_t row = t / nvp;
    const int iv0 = (int)(t - row * nvp) * 2;
    const int64_t u = row * nv + iv0;
    cplx s0 = make_double2(0.0, 0.0), s1 = s0;
    if (HAS_F) { const cplx2 f = ldg2(a.f + u); s0 = f.a; s1 = f.b; }
    int64_t p = __ldg(rowptr + row);
    const int64_t p1 = __ldg(rowptr + row + 1);
    const cplx *__restrict__ xin = a.xin + iv0;
    for (; p + 2 <= p1; p += 2) {
        const int32_t c0 = __ldg(col + p), c1 = __ldg(col + p + 1);
        const cplx a0 = ldg(val + p), a1 = ldg(val + p + 1);
        const cplx2 x0 = ldg2(xin + (int64_t)c0 * nv), x1 = ldg2(xin + (int64_t)c1 * nv);
        cfma(s0, a0, x0.a); cfma(s1, a0, x0.b);
        cfma(s0, a1, x1.a); cfma(s1, a1, x1.b);
    }
    if (p < p1) {
        const cplx a0 = ldg(val + p);
        const cplx2 x0 = ldg2(xin + (int64_t)__ldg(col + p) * nv);
        cfma(s0, a0, x0.a); cfma(s1, a0, x0.b);
    }
    if (MODE == 0) {
        const cplx2 y = ldg2(a.ybase + u);
        st2(a.xout + u, axpy1(a.cx, s0, y.a), axpy1(a.cx, s1, y.b));
        st2(a.acc_out + u, axpy1(a.ca, s0, y.a), axpy1(a.ca, s1, y.b));
    } else if (MODE == 1) {
        const cplx2 y = ldg2(a.ybase + u);
        const cplx2 ac = *reinterpret_cast<const cplx2 *>(a.acc_in + u);
        st2(a.xout + u, axpy1(a.cx, s0, y.a), axpy1(a.cx, s1, y.b));
        st2(a.acc_out + u, axpy1(a.ca, s0, ac.a), axpy1(a.ca, s1, ac.b));
    } else {
        const cplx2 ac = *reinterpret_cast<const cplx2 *>(a.acc_in + u);
        const cplx y0 = axpy1(a.ca, s0, ac.a), y1 = axpy1(a.ca, s1, ac.b);
        st2(a.acc_out + u, y0, y1);
        if (a.snap) st2(a.snap + u, y0, y1);
    }
}

template <int B, int CV>
static cudaError_t bcsr_launch(const BcsrDev &Op, int nv, int mode, const StageArgs &a, cudaStream_t s, int64_t br_begin,
                               int64_t br_count)
{
    const int nvp = nv / CV;
    const int64_t n_threads = br_count * nvp;
    if (n_threads <= 0) return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_threads + threads - 1) / threads);
    const bool hf = a.f != nullptr;
#define LAUNCH(M, F) rk4_stage_bcsr_kernel<B, CV, M, F><<<blocks, threads, 0, s>>>(Op.browptr, Op.bcol, Op.bval, nv, nvp, br_begin, n_threads, a)
    if (mode == 0) { if (hf) LAUNCH(0, true); else LAUNCH(0, false); }
    else if (mode == 1) { if (hf) LAUNCH(1, true); else LAUNCH(1, false); }
    else { if (hf) LAUNCH(2, true); else LAUNCH(2, false); }
#undef LAUNCH
    return cudaGetLastError();
}

bool bcsr_stage_supported(int B, int nv) { return (B == 3 || B == 4 || B == 5) && nv >= 1; }

cudaError_t launch_rk4_stage_bcsr(const BcsrDev &Op, int nv, int mode, const StageArgs &a, int variant, cudaStream_t s,
                                  int64_t row_begin, int64_t row_count)
{
    const bool two = (nv % 2 == 0) && variant != 1;
    if (row_count < 0) row_count = Op.nbrows * Op.B - row_begin;
    if (row_begin % Op.B || row_count % Op.B) return cudaErrorInvalidValue;   // ranges must be whole block rows
    const int64_t b0 = row_begin / Op.B, bn = row_count / Op.B;
    switch (Op.B) {
        case 3: return two ? bcsr_launch<3, 2>(Op, nv, mode, a, s, b0, bn) : bcsr_launch<3, 1>(Op, nv, mode, a, s, b0, bn);
        case 4: return two ? bcsr_launch<4, 2>(Op, nv, mode, a, s, b0, bn) : bcsr_launch<4, 1>(Op, nv, mode, a, s, b0, bn);
        case 5: return two ? bcsr_launch<5, 2>(Op, nv, mode, a, s, b0, bn) : bcsr_launch<5, 1>(Op, nv, mode, a, s, b0, bn);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_rk4_stage(const CsrDev &Op, int nv, int mode, const StageArgs &a, int variant, cudaStream_t s,
                             int64_t row_begin, int64_t row_count)
{
    (void)variant;
    if (row_count < 0) row_count = Op.nrows - row_begin;
    const int64_t n_units = row_count * (int64_t)nv, u_begin = row_begin * (int64_t)nv;
    if (n_units <= 0) return cudaSuccess;
    const int threads = 256;
    const bool hf = a.f != nullptr;
    if (variant != 1 && nv % 2 == 0 && Op.max_row_nnz >= 8) {   // long rows: two test vectors per thread, 256-bit gathers
        const int nvp = nv / 2;
        const int64_t n_thr = row_count * nvp, t_begin = row_begin * nvp;
        const int64_t blk2 = (n_thr + threads - 1) / threads;
#define LAUNCH2(M, F) rk4_stage_csr2_kernel<M, F><<<(unsigned)blk2, threads, 0, s>>>(Op.rowptr, Op.col, Op.val, nv, nvp, t_begin, n_thr, a)
        if (mode == 0) { if (hf) LAUNCH2(0, true); else LAUNCH2(0, false); }
        else if (mode == 1) { if (hf) LAUNCH2(1, true); else LAUNCH2(1, false); }
        else { if (hf) LAUNCH2(2, true); else LAUNCH2(2, false); }
#undef LAUNCH2
        return cudaGetLastError();
    }
    const int64_t blocks = (n_units + threads - 1) / threads;
#define LAUNCH(M, F) rk4_stage_csr_kernel<M, F><<<(unsigned)blocks, threads, 0, s>>>(Op.rowptr, Op.col, Op.val, nv, u_begin, n_units, a)
    if (mode == 0) { if (hf) LAUNCH(0, true); else LAUNCH(0, false); }
    else if (mode == 1) { if (hf) LAUNCH(1, true); else LAUNCH(1, false); }
    else { if (hf) LAUNCH(2, true); else LAUNCH(2, false); }
#undef LAUNCH
    return cudaGetLastError();
}

// =====================================================================================================
// K1b: a whole RK4 step in ONE kernel for narrow-band operators (temporal fusion of the four stages).
//   A CTA owns TR consecutive rows and recomputes a halo of 4*b rows on each side (b = matrix half bandwidth), so the
//   three intermediate stage vectors never leave shared memory.  Per unit and step the HBM traffic drops from the
//   20 block-vector passes of the stage-by-stage scheme (320 B) to: read y, f(t_n), f(t_n+dt/2), f(t_n+dt), write y
//   (80 B) times the halo factor (TR+8b)/TR.  Arithmetic and summation order per row are identical to K1.
// =====================================================================================================
template <int BT, int UPT, int MINB, bool FORCED>
__global__ void __launch_bounds__(BT, MINB)
rk4_banded_step_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                       int64_t N, int nv, int b, int TR, int E0, const int64_t *__restrict__ ranges, BandedStepArgs a)
{
    extern __shared__ cplx sm[];
    a.y_in += (int64_t)blockIdx.y * a.batch_stride;
    a.y_out += (int64_t)blockIdx.y * a.batch_stride;
    const int n_e = E0 * nv;
    cplx *Y = sm, *B1 = sm + n_e, *B2 = sm + 2 * n_e;
    // owned rows: [row0, row0 + nown) -- either the regular tiling or an explicit list of row ranges
    const int64_t row0 = ranges ? ranges[2 * blockIdx.x] : (int64_t)blockIdx.x * TR;
    const int nown = ranges ? (int)(ranges[2 * blockIdx.x + 1] - row0) : TR;
    const int64_t base = row0 - 4 * b;  // global row of local row 0 (may be negative)
    const double dt = a.dt;
    const double c_half = dt / 2, c_full = dt, w1 = dt / 6, w2 = dt / 6 * 2;  // TSRK4.c:30-34,44-47

    cplx acc[UPT];
    int lrow[UPT], liv[UPT];
    bool valid[UPT];
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int e = threadIdx.x + i * BT;
        const int lr = e / nv;
        lrow[i] = lr;
        liv[i] = e - lr * nv;
        const int64_t gr = base + lr;
        valid[i] = (e < n_e) && gr >= 0 && gr < N;
        cplx y = make_double2(0.0, 0.0);
        if (valid[i]) y = ldg(a.y_in + gr * nv + liv[i]);
        if (e < n_e) Y[e] = y;
        acc[i] = y;
    }
    __syncthreads();

    // ---- stage 1: k1 = Op y + F3(t_n) on local rows [b, E0-b) ; x2 = y + dt/2 k1 ; acc = y + dt/6 k1
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int lr = lrow[i];
        if (valid[i] && lr >= b && lr < E0 - b) {
            const int64_t gr = base + lr;
            const int64_t gu = gr * nv + liv[i];
            cplx s = make_double2(0.0, 0.0);
            if (FORCED) s = ldg(a.f_prev + gu);
            const int64_t p1 = __ldg(rowptr + gr + 1);
            for (int64_t p = __ldg(rowptr + gr); p < p1; ++p)
                cfma(s, ldg(val + p), Y[(int)(__ldg(col + p) - base) * nv + liv[i]]);
            const cplx y = acc[i];
            B1[threadIdx.x + i * BT] = make_double2(fma(c_half, s.x, y.x), fma(c_half, s.y, y.y));
            acc[i] = make_double2(fma(w1, s.x, y.x), fma(w1, s.y, y.y));
        }
    }
    __syncthreads();
    // ---- stage 2: k2 = Op x2 + F2 on [2b, E0-2b) ; x3 = y + dt/2 k2 ; acc += dt/3 k2
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int lr = lrow[i];
        if (valid[i] && lr >= 2 * b && lr < E0 - 2 * b) {
            const int64_t gr = base + lr;
            cplx s = make_double2(0.0, 0.0);
            if (FORCED) s = ldg(a.f_half + gr * nv + liv[i]);
            const int64_t p1 = __ldg(rowptr + gr + 1);
            for (int64_t p = __ldg(rowptr + gr); p < p1; ++p)
                cfma(s, ldg(val + p), B1[(int)(__ldg(col + p) - base) * nv + liv[i]]);
            const cplx y = Y[threadIdx.x + i * BT];
            B2[threadIdx.x + i * BT] = make_double2(fma(c_half, s.x, y.x), fma(c_half, s.y, y.y));
            acc[i] = make_double2(fma(w2, s.x, acc[i].x), fma(w2, s.y, acc[i].y));
        }
    }
    __syncthreads();
    // ---- stage 3: k3 = Op x3 + F2 on [3b, E0-3b) ; x4 = y + dt k3 ; acc += dt/3 k3
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int lr = lrow[i];
        if (valid[i] && lr >= 3 * b && lr < E0 - 3 * b) {
            const int64_t gr = base + lr;
            cplx s = make_double2(0.0, 0.0);
            if (FORCED) s = ldg(a.f_half + gr * nv + liv[i]);   // second read of the slice: L2 hit
            const int64_t p1 = __ldg(rowptr + gr + 1);
            for (int64_t p = __ldg(rowptr + gr); p < p1; ++p)
                cfma(s, ldg(val + p), B2[(int)(__ldg(col + p) - base) * nv + liv[i]]);
            const cplx y = Y[threadIdx.x + i * BT];
            B1[threadIdx.x + i * BT] = make_double2(fma(c_full, s.x, y.x), fma(c_full, s.y, y.y));
            acc[i] = make_double2(fma(w2, s.x, acc[i].x), fma(w2, s.y, acc[i].y));
        }
    }
    __syncthreads();
    // ---- stage 4: k4 = Op x4 + F3(t_n+dt) on the owned rows [4b, E0-4b) ; y_{n+1} = acc + dt/6 k4
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int lr = lrow[i];
        if (valid[i] && lr >= 4 * b && lr < 4 * b + nown) {
            const int64_t gr = base + lr;
            const int64_t gu = gr * nv + liv[i];
            cplx s = make_double2(0.0, 0.0);
            if (FORCED) s = ldg(a.f_next + gu);
            const int64_t p1 = __ldg(rowptr + gr + 1);
            for (int64_t p = __ldg(rowptr + gr); p < p1; ++p)
                cfma(s, ldg(val + p), B1[(int)(__ldg(col + p) - base) * nv + liv[i]]);
            const cplx yn = make_double2(fma(w1, s.x, acc[i].x), fma(w1, s.y, acc[i].y));
            a.y_out[gu] = yn;
            if (a.snap) a.snap[gu] = yn;
        }
    }
}

namespace {
struct BandedCfg { int bt, upt; };
inline BandedCfg banded_cfg(int variant)
{
    switch (variant) {
        case 1: return {512, 4};   // 1 CTA/SM, no spills
        case 2: return {512, 3};   // smaller tile, 2 CTAs/SM
        case 3: return {256, 6};   // 3 CTAs/SM
        default: return {512, 4};  // 2 CTAs/SM (64 registers)
    }
}
inline int banded_E0(int variant, int nv) { const BandedCfg c = banded_cfg(variant); return (c.bt * c.upt) / nv; }

template <int BT, int UPT, int MINB>
cudaError_t banded_launch(const CsrDev &Op, int b, int nv, int E0, const BandedStepArgs &a, const int64_t *ranges,
                          int nranges, cudaStream_t s)
{
    const int TR = E0 - 8 * b;
    const size_t smem = (size_t)3 * E0 * nv * sizeof(cplx);
    const unsigned grid = ranges ? (unsigned)nranges : (unsigned)((Op.nrows + TR - 1) / TR);
    if (grid == 0) return cudaSuccess;
    cudaError_t e;
    if (a.f_prev) {
        auto kfn = rk4_banded_step_kernel<BT, UPT, MINB, true>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<dim3(grid, (unsigned)a.nbatch), BT, smem, s>>>(Op.rowptr, Op.col, Op.val, Op.nrows, nv, b, TR, E0, ranges, a);
    } else {
        auto kfn = rk4_banded_step_kernel<BT, UPT, MINB, false>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<dim3(grid, (unsigned)a.nbatch), BT, smem, s>>>(Op.rowptr, Op.col, Op.val, Op.nrows, nv, b, TR, E0, ranges, a);
    }
    return cudaGetLastError();
}
}  // namespace

bool banded_step_supported(const CsrDev &Op, int64_t lo_bw, int64_t hi_bw, int nv, int variant)
{
    const int64_t b = lo_bw > hi_bw ? lo_bw : hi_bw;
    if (nv <= 0 || Op.nrows <= 0) return false;
    const int E0 = banded_E0(variant, nv);
    // the halo (8b rows) must stay a modest fraction of the tile
    return b >= 1 && 8 * b * 4 <= E0;
}

int banded_tile_rows(int64_t lo_bw, int64_t hi_bw, int nv, int variant)
{
    const int b = (int)(lo_bw > hi_bw ? lo_bw : hi_bw);
    return banded_E0(variant, nv) - 8 * b;
}

cudaError_t launch_rk4_banded_step(const CsrDev &Op, int64_t lo_bw, int64_t hi_bw, int nv, const BandedStepArgs &a,
                                   int variant, const int64_t *ranges, int nranges, cudaStream_t s)
{
    const int b = (int)(lo_bw > hi_bw ? lo_bw : hi_bw);
    const int E0 = banded_E0(variant, nv);
    if (E0 - 8 * b <= 0) return cudaErrorInvalidValue;
    switch (variant) {
        case 1: return banded_launch<512, 4, 1>(Op, b, nv, E0, a, ranges, nranges, s);
        case 2: return banded_launch<512, 3, 2>(Op, b, nv, E0, a, ranges, nranges, s);
        case 3: return banded_launch<256, 6, 3>(Op, b, nv, E0, a, ranges, nranges, s);
        default: return banded_launch<512, 4, 2>(Op, b, nv, E0, a, ranges, nranges, s);
    }
}

// =====================================================================================================
// K1c: the fused RK4 step for operators whose rows (almost all) follow one diagonal pattern (offsets -LO..+HI):
//   the operator is stored by diagonals (DIA: dia[row][LO+HI+1], no column indices), a thread owns R consecutive rows of
//   one test vector and keeps the R+LO+HI gathered entries of the stage vector in registers (sliding window), so the
//   shared-memory traffic per row drops from (LO+HI+1) to (R+LO+HI)/R loads and the A entries come from a shared tile.
//   Tiles that contain a row outside the pattern (e.g. the boundary rows of the Ginzburg-Landau operator) are skipped
//   here (tile_ok == 0) and handled by the CSR kernel above on an explicit row-range list.
//   Halo: 4*LO rows on the left, 4*HI on the right.
// =====================================================================================================
template <int LO, int HI, int R, bool FORCED>
__global__ void __launch_bounds__(448, 2)
rk4_dia_step_kernel(const cplx *__restrict__ dia, const unsigned char *__restrict__ tile_ok, int64_t N, int nv, int nq,
                    int TR, BandedStepArgs a)
{
    constexpr int ND = LO + HI + 1, WN = R + LO + HI;
    if (!tile_ok[blockIdx.x]) return;
    extern __shared__ cplx sm[];
    // Shared layout, bank-conflict free for every access of the stage loops: a tile row lr = q*R + rr of a block vector
    // lives at slot  rr*TP + q*nv + iv  ("rr-major"): the threads (q, iv) of a warp are consecutive in q*nv + iv, so both
    // their own rows and the neighbouring row blocks (q-1, q+1) are contiguous runs of 16-byte slots.  The operator tile
    // is stored as [rr][d][q] for the same reason (lanes of one row block broadcast, different blocks hit different banks).
    const int T = nq * nv;
    const int TP = T + ((nv - T) & 7);          // TP == nv (mod 8): consecutive rows continue the bank sequence
    const int E0 = nq * R, n_e = E0 * nv;
    cplx *BA = sm, *BB = sm + R * TP, *Asm = sm + 2 * R * TP;   // BA: y then x3 ; BB: x2 then x4
    const int64_t base = (int64_t)blockIdx.x * TR - 4 * LO;
    const double dt = a.dt;
    const double c_half = dt / 2, c_full = dt, w1 = dt / 6, w2 = dt / 6 * 2;

    {   // cooperative, coalesced tile loads: the tile of a block vector / of dia is contiguous in global memory
        const int64_t g0 = base * nv, gN = N * nv;
        for (int e = threadIdx.x; e < n_e; e += blockDim.x) {
            const int64_t g = g0 + e;
            const int lr = e / nv, iv = e - lr * nv;
            const int qq = lr / R, rr = lr - qq * R;
            BA[rr * TP + qq * nv + iv] = (g >= 0 && g < gN) ? ldg(a.y_in + g) : make_double2(0.0, 0.0);
        }
        const int64_t d0 = base * ND, dN = N * ND;
        for (int e = threadIdx.x; e < E0 * ND; e += blockDim.x) {
            const int64_t g = d0 + e;
            const int lr = e / ND, d = e - lr * ND;
            const int qq = lr / R, rr = lr - qq * R;
            Asm[(rr * ND + d) * nq + qq] = (g >= 0 && g < dN) ? ldg(dia + g) : make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    const bool active = threadIdx.x < T;
    const int q = active ? threadIdx.x / nv : 0;
    const int iv = active ? threadIdx.x - q * nv : 0;
    const int r_first = q * R;
    const int t = q * nv + iv;   // == threadIdx.x for active threads
    cplx yo[R], acc[R];
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
        yo[rr] = BA[rr * TP + t];
        acc[rr] = yo[rr];
    }

    // window slot of tile row (r_first - LO + j): previous / own / next row block
    // stages 1..3 (k_s on every tile row; rows of the outer halo hold don't-care values that never reach owned rows)
#pragma unroll
    for (int st = 1; st <= 3; st++) {
        const cplx *Xin = (st == 2) ? BB : BA;
        cplx *Xout = (st == 2) ? BA : BB;
        const cplx *fsl = (st == 1) ? a.f_prev : a.f_half;
        const double cx = (st == 3) ? c_full : c_half;
        const double ca = (st == 1) ? w1 : w2;
        if (active) {
            cplx xw[WN];
#pragma unroll
            for (int j = 0; j < WN; j++) {
                // j - LO in [-LO, R+HI): row block offset and row-in-block are compile-time
                const int off = j - LO;
                const int dq = (off < 0) ? -1 : (off >= R ? 1 : 0);
                const int rr2 = off - dq * R;
                const int q2 = q + dq;
                xw[j] = (q2 >= 0 && q2 < nq) ? Xin[rr2 * TP + t + dq * nv] : make_double2(0.0, 0.0);
            }
            cplx xo[R];
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                const int lr = r_first + rr;
                const int64_t gr = base + lr;
                cplx s = make_double2(0.0, 0.0);
                if (FORCED && lr >= st * LO && lr < E0 - st * HI && gr >= 0 && gr < N) s = ldg(fsl + gr * nv + iv);
#pragma unroll
                for (int d = 0; d < ND; d++) cfma(s, Asm[(rr * ND + d) * nq + q], xw[rr + d]);
                xo[rr] = make_double2(fma(cx, s.x, yo[rr].x), fma(cx, s.y, yo[rr].y));
                acc[rr] = make_double2(fma(ca, s.x, acc[rr].x), fma(ca, s.y, acc[rr].y));
            }
#pragma unroll
            for (int rr = 0; rr < R; rr++) Xout[rr * TP + t] = xo[rr];
        }
        __syncthreads();
    }
    // stage 4 on the owned rows
    if (active) {
        cplx xw[WN];
#pragma unroll
        for (int j = 0; j < WN; j++) {
            const int off = j - LO;
            const int dq = (off < 0) ? -1 : (off >= R ? 1 : 0);
            const int rr2 = off - dq * R;
            const int q2 = q + dq;
            xw[j] = (q2 >= 0 && q2 < nq) ? BB[rr2 * TP + t + dq * nv] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            const int lr = r_first + rr;
            const int64_t gr = base + lr;
            if (lr >= 4 * LO && lr < 4 * LO + TR && gr < N) {
                const int64_t gu = gr * nv + iv;
                cplx s = make_double2(0.0, 0.0);
                if (FORCED) s = ldg(a.f_next + gu);
#pragma unroll
                for (int d = 0; d < ND; d++) cfma(s, Asm[(rr * ND + d) * nq + q], xw[rr + d]);
                const cplx yn = make_double2(fma(w1, s.x, acc[rr].x), fma(w1, s.y, acc[rr].y));
                a.y_out[gu] = yn;
                if (a.snap) a.snap[gu] = yn;
            }
        }
    }
}

// ---- TMA (cp.async.bulk) + mbarrier helpers -------------------------------------------------------------------
DEVFN unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
DEVFN void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVFN void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
DEVFN void tma_store_1d(void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
DEVFN void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
DEVFN void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// =====================================================================================================
// K1d: the DIA fused step with TMA-staged tiles.  All inputs of a tile -- y, the operator diagonals and the three
//   forcing slices -- are contiguous ranges of global memory, so one elected thread brings them into shared memory with
//   five bulk copies (cp.async.bulk -> mbarrier) issued up front; with 2-3 CTAs resident per SM the whole input of the
//   next tile is in flight while the current one computes, which is what it takes to keep HBM busy (the CTA would
//   otherwise expose one DRAM round trip per stage).  y_{n+1} (and the snapshot) leave through bulk stores.
//   Natural [row][iv] tile layout; R is chosen so that (R-1)*nv is a multiple of 8, which makes the row-block stride
//   R*nv congruent to nv modulo the 8 16-byte slots of a shared-memory wavefront: conflict-free sliding-window accesses.
// =====================================================================================================
template <int LO, int HI, int R, int MINB, bool FORCED>
__global__ void __launch_bounds__(MINB >= 4 ? 128 : (MINB == 3 ? 160 : 320)) __maxnreg__(MINB >= 4 ? 120 : (MINB == 3 ? 112 : 96))
rk4_dia_tma_kernel(const cplx *__restrict__ dia, const unsigned char *__restrict__ tile_ok, int64_t N, int nv, int nq,
                   int TR, BandedStepArgs a)
{
    constexpr int ND = LO + HI + 1, WN = R + LO + HI;
    int tile = (int)blockIdx.x;
    if (a.bnd_flag && gridDim.x > 1) tile = blockIdx.x == 0 ? 0 : (blockIdx.x == 1 ? (int)gridDim.x - 1 : (int)blockIdx.x - 1);
    if (!tile_ok[tile]) return;
    a.y_in += (int64_t)blockIdx.y * a.batch_stride;
    a.y_out += (int64_t)blockIdx.y * a.batch_stride;
    extern __shared__ __align__(128) cplx sm[];
    const int T = nq * nv, E0 = nq * R, n_e = E0 * nv;
    cplx *BA = sm;                      // y, later x3, finally y_{n+1}
    cplx *BB = BA + n_e;                // x2, later x4
    cplx *At = BB + n_e;                // operator diagonals [row][d]
    cplx *Fp = At + E0 * ND, *Fh = Fp + n_e, *Fn = Fh + n_e;   // forcing tiles (FORCED only)
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(FORCED ? (Fn + n_e) : Fp);
    const int64_t own_end = a.own_end < 0 ? N : a.own_end;
    const int64_t base = a.own_begin + (int64_t)tile * TR - 4 * LO;
    const int lr_lo = (int)max((int64_t)0, -base), lr_hi = (int)min((int64_t)E0, N - base);   // tile rows inside [0, N)
    const double dt = a.dt;
    const double c_half = dt / 2, c_full = dt, w1 = dt / 6, w2 = dt / 6 * 2;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_proxy_async();
    }
    // rows of the tile outside the matrix: zero (disjoint from the bulk-copy destinations)
    if (lr_lo > 0 || lr_hi < E0) {
        for (int e = threadIdx.x; e < n_e; e += blockDim.x) {
            const int lr = e / nv;
            if (lr < lr_lo || lr >= lr_hi) {
                BA[e] = make_double2(0.0, 0.0);
                if (FORCED) { Fp[e] = BA[e]; Fh[e] = BA[e]; Fn[e] = BA[e]; }
            }
        }
        for (int e = threadIdx.x; e < E0 * ND; e += blockDim.x) {
            const int lr = e / ND;
            if (lr < lr_lo || lr >= lr_hi) At[e] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned vb = (unsigned)(lr_hi - lr_lo) * nv * sizeof(cplx), ab = (unsigned)(lr_hi - lr_lo) * ND * sizeof(cplx);
        const int64_t g = (base + lr_lo) * nv;
        mbar_expect_tx(bar, vb * (FORCED ? 4u : 1u) + ab);
        tma_load_1d(BA + lr_lo * nv, a.y_in + g, vb, bar);
        tma_load_1d(At + lr_lo * ND, dia + (base + lr_lo) * ND, ab, bar);
        if (FORCED) {
            tma_load_1d(Fp + lr_lo * nv, a.f_prev + g, vb, bar);
            tma_load_1d(Fh + lr_lo * nv, a.f_half + g, vb, bar);
            tma_load_1d(Fn + lr_lo * nv, a.f_next + g, vb, bar);
        }
    }
    mbar_wait(bar, 0);

    const bool active = threadIdx.x < T;
    const int q = active ? threadIdx.x / nv : 0;
    const int iv = active ? threadIdx.x - q * nv : 0;
    const int r_first = q * R;
    const int own = r_first * nv + iv;   // slot of the thread's first row; consecutive rows are nv slots apart
    // The thread's own R rows of the current stage vector stay in registers from one stage to the next (xs); only the rows
    // a neighbouring row block needs (the first HI and the last LO of the block) go through shared memory, and only the
    // LO + HI rows of the neighbouring blocks are read back: 3 + 3 instead of 8 + 5 shared accesses per stage for the
    // Ginzburg-Landau band.
    cplx yo[R], acc[R], xs[R];
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
        yo[rr] = BA[own + rr * nv];
        acc[rr] = yo[rr];
        xs[rr] = yo[rr];
    }
    constexpr bool ALLROWS = (R < LO + HI);   // tiny blocks: every row is needed by a neighbour
#pragma unroll
    for (int st = 1; st <= 3; st++) {
        const cplx *Xin = (st == 2) ? BB : BA;
        cplx *Xout = (st == 2) ? BA : BB;
        const cplx *fsl = (st == 1) ? Fp : Fh;
        const double cx = (st == 3) ? c_full : c_half;
        const double ca = (st == 1) ? w1 : w2;
        cplx xo[R];
        if (active) {
            cplx xw[WN];
#pragma unroll
            for (int j = 0; j < WN; j++) {
                if (j >= LO && j < LO + R) {
                    xw[j] = xs[j - LO];
                } else {
                    const int lr = r_first - LO + j;
                    xw[j] = (lr >= 0 && lr < E0) ? Xin[own + (j - LO) * nv] : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                cplx s = make_double2(0.0, 0.0);
                if (FORCED) s = fsl[own + rr * nv];
                const cplx *ar = At + (r_first + rr) * ND;
#pragma unroll
                for (int d = 0; d < ND; d++) cfma(s, ar[d], xw[rr + d]);
                xo[rr] = make_double2(fma(cx, s.x, yo[rr].x), fma(cx, s.y, yo[rr].y));
                acc[rr] = make_double2(fma(ca, s.x, acc[rr].x), fma(ca, s.y, acc[rr].y));
            }
        }
        if (st == 1) __syncthreads();   // every thread has read its y halo before BA is overwritten (stage 2 writes BA)
        if (active) {
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                if (ALLROWS || rr < HI || rr >= R - LO) Xout[own + rr * nv] = xo[rr];
                xs[rr] = xo[rr];
            }
        }
        __syncthreads();
    }
    // stage 4 on the owned rows; y_{n+1} goes to BA (x3 is dead) and leaves through a bulk store
    if (active) {
        cplx xw[WN];
#pragma unroll
        for (int j = 0; j < WN; j++) {
            if (j >= LO && j < LO + R) {
                xw[j] = xs[j - LO];
            } else {
                const int lr = r_first - LO + j;
                xw[j] = (lr >= 0 && lr < E0) ? BB[own + (j - LO) * nv] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            const int lr = r_first + rr;
            if (lr >= 4 * LO && lr < 4 * LO + TR) {
                cplx s = make_double2(0.0, 0.0);
                if (FORCED) s = Fn[own + rr * nv];
                const cplx *ar = At + lr * ND;
#pragma unroll
                for (int d = 0; d < ND; d++) cfma(s, ar[d], xw[rr + d]);
                BA[own + rr * nv] = make_double2(fma(w1, s.x, acc[rr].x), fma(w1, s.y, acc[rr].y));
            }
        }
    }
    fence_proxy_async();   // make the generic-proxy writes to BA visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
        const int o_lo = 4 * LO, o_hi = (int)min((int64_t)(4 * LO + TR), own_end - base);
        if (o_hi > o_lo) {
            const unsigned ob = (unsigned)(o_hi - o_lo) * nv * sizeof(cplx);
            const int64_t g = (base + o_lo) * nv;
            tma_store_1d(a.y_out + g, BA + o_lo * nv, ob);
            if (a.snap) tma_store_1d(a.snap + g, BA + o_lo * nv, ob);
            tma_store_commit_wait();
        }
        if (a.bnd_flag && (tile == 0 || tile == (int)gridDim.x - 1)) {
            asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // the rows are written, not only read out of shared memory
            __threadfence();
            atomicAdd(a.bnd_flag, 1ull);
        }
    }
}

namespace {
// R with (R-1)*nv a multiple of 8 (see above); 0 if none of the instantiated values fits
inline int tma_R(int nv)
{
    const int cand[3] = {4, 5, 3};
    for (int r : cand)
        if (((r - 1) * nv) % 8 == 0) return r;
    return 0;
}
inline int tma_nq(int nv, int R, int nd, bool forced, int minb)
{
    // shared memory per CTA = ((nbuf*nv + nd) * E0) * 16 + 16, E0 = R*nq; threads = nq*nv <= 160 (3 CTAs/SM) or 320 (2)
    const int budget = (minb >= 4 ? 55 : (minb == 3 ? 74 : 112)) * 1024;
    const int per_row = ((forced ? 5 : 2) * nv + nd) * 16;
    int nq = (budget - 64) / (per_row * R);
    const int tmax = (minb >= 4 ? 128 : (minb == 3 ? 160 : 320)) / nv;
    return nq < tmax ? nq : tmax;
}

template <int LO, int HI, int R, int MINB>
cudaError_t dia_tma_launch(const cplx *dia, const unsigned char *tile_ok, int ntiles, int64_t N, int nv, int nq,
                           const BandedStepArgs &a, cudaStream_t s)
{
    const int E0 = nq * R;
    const int TR = E0 - 4 * (LO + HI);
    const bool forced = a.f_prev != nullptr;
    const size_t smem = ((size_t)(forced ? 5 : 2) * E0 * nv + (size_t)E0 * (LO + HI + 1)) * sizeof(cplx) + 16;
    const int threads = ((nq * nv + 31) / 32) * 32;
    cudaError_t e;
    if (forced) {
        auto kfn = rk4_dia_tma_kernel<LO, HI, R, MINB, true>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        kfn<<<dim3((unsigned)ntiles, (unsigned)a.nbatch), threads, smem, s>>>(dia, tile_ok, N, nv, nq, TR, a);
    } else {
        auto kfn = rk4_dia_tma_kernel<LO, HI, R, MINB, false>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        kfn<<<dim3((unsigned)ntiles, (unsigned)a.nbatch), threads, smem, s>>>(dia, tile_ok, N, nv, nq, TR, a);
    }
    return cudaGetLastError();
}

template <int LO, int HI, int MINB>
cudaError_t dia_tma_dispatch_R(int R, const cplx *dia, const unsigned char *tile_ok, int ntiles, int64_t N, int nv, int nq,
                               const BandedStepArgs &a, cudaStream_t s)
{
    if (R == 3) return dia_tma_launch<LO, HI, 3, MINB>(dia, tile_ok, ntiles, N, nv, nq, a, s);
    if (R == 4) return dia_tma_launch<LO, HI, 4, MINB>(dia, tile_ok, ntiles, N, nv, nq, a, s);
    if (R == 5) return dia_tma_launch<LO, HI, 5, MINB>(dia, tile_ok, ntiles, N, nv, nq, a, s);
    return cudaErrorInvalidValue;
}
}  // namespace

// variant: 0 -> 3 CTAs/SM tiles, 1 -> 2 CTAs/SM tiles.  The tile height must not depend on `forced` (the plan is shared),
// so it is always sized for the forced case.
bool dia_tma_supported(int lo, int hi, int nv, int variant)
{
    const bool combo = (lo == 1 && hi == 1) || (lo == 2 && hi == 1) || (lo == 1 && hi == 2) || (lo == 2 && hi == 2);
    const int R = tma_R(nv);
    if (!combo || R == 0 || nv > 64) return false;
    const int nq = tma_nq(nv, R, lo + hi + 1, true, variant == 1 ? 2 : (variant == 2 ? 4 : 3));
    return nq >= 2 && 4 * (lo + hi) * 4 <= nq * R;
}

int dia_tma_tile_rows(int lo, int hi, int nv, int variant)
{
    const int R = tma_R(nv);
    return tma_nq(nv, R, lo + hi + 1, true, variant == 1 ? 2 : (variant == 2 ? 4 : 3)) * R - 4 * (lo + hi);
}

cudaError_t launch_rk4_dia_tma_step(const cplx *dia, int lo, int hi, const unsigned char *tile_ok, int ntiles, int64_t N,
                                    int nv, int variant, const BandedStepArgs &a, cudaStream_t s)
{
    const int R = tma_R(nv);
    const int minb = variant == 1 ? 2 : (variant == 2 ? 4 : 3);
    const int nq = tma_nq(nv, R, lo + hi + 1, true, minb);
#define DISPATCH(L, H)                                                                                    \
    if (lo == L && hi == H)                                                                               \
        return minb == 2 ? dia_tma_dispatch_R<L, H, 2>(R, dia, tile_ok, ntiles, N, nv, nq, a, s)          \
             : minb == 4 ? dia_tma_dispatch_R<L, H, 4>(R, dia, tile_ok, ntiles, N, nv, nq, a, s)          \
                         : dia_tma_dispatch_R<L, H, 3>(R, dia, tile_ok, ntiles, N, nv, nq, a, s);
    DISPATCH(1, 1)
    DISPATCH(2, 1)
    DISPATCH(1, 2)
    DISPATCH(2, 2)
#undef DISPATCH
    return cudaErrorInvalidValue;
}

namespace {
constexpr int DIA_R = 4;
inline int dia_nq(int nv) { return 448 / nv; }

template <int LO, int HI>
cudaError_t dia_launch(const cplx *dia, const unsigned char *tile_ok, int ntiles, int64_t N, int nv,
                       const BandedStepArgs &a, cudaStream_t s)
{
    const int nq = dia_nq(nv);
    const int E0 = nq * DIA_R;
    const int TR = E0 - 4 * (LO + HI);
    const int T = nq * nv, TP = T + ((nv - T) & 7);
    const size_t smem = ((size_t)2 * DIA_R * TP + (size_t)E0 * (LO + HI + 1)) * sizeof(cplx);
    const int threads = ((T + 31) / 32) * 32;
    cudaError_t e;
    if (a.f_prev) {
        auto kfn = rk4_dia_step_kernel<LO, HI, DIA_R, true>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<(unsigned)ntiles, threads, smem, s>>>(dia, tile_ok, N, nv, nq, TR, a);
    } else {
        auto kfn = rk4_dia_step_kernel<LO, HI, DIA_R, false>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<(unsigned)ntiles, threads, smem, s>>>(dia, tile_ok, N, nv, nq, TR, a);
    }
    return cudaGetLastError();
}
}  // namespace

bool dia_step_supported(int lo, int hi, int nv)
{
    if (nv <= 0 || nv > 60) return false;
    const bool combo = (lo == 1 && hi == 1) || (lo == 2 && hi == 1) || (lo == 1 && hi == 2) || (lo == 2 && hi == 2);
    if (!combo) return false;
    const int E0 = dia_nq(nv) * DIA_R;
    return 4 * (lo + hi) * 4 <= E0;   // halo at most a quarter of the tile
}

int dia_tile_rows(int lo, int hi, int nv) { return dia_nq(nv) * DIA_R - 4 * (lo + hi); }

cudaError_t launch_rk4_dia_step(const cplx *dia, int lo, int hi, const unsigned char *tile_ok, int ntiles, int64_t N, int nv,
                                const BandedStepArgs &a, cudaStream_t s)
{
    if (lo == 1 && hi == 1) return dia_launch<1, 1>(dia, tile_ok, ntiles, N, nv, a, s);
    if (lo == 2 && hi == 1) return dia_launch<2, 1>(dia, tile_ok, ntiles, N, nv, a, s);
    if (lo == 1 && hi == 2) return dia_launch<1, 2>(dia, tile_ok, ntiles, N, nv, a, s);
    if (lo == 2 && hi == 2) return dia_launch<2, 2>(dia, tile_ok, ntiles, N, nv, a, s);
    return cudaErrorInvalidValue;
}

// =====================================================================================================
// Twiddle tables (CreateDFTiDFTMats.c:42-58,76-84).  Phases are reduced in integer arithmetic before sincospi,
// so the tables are accurate to ~1 ulp (the reference's own tables carry ~1e-11 from exp() of large arguments).
// =====================================================================================================
__device__ __forceinline__ int64_t freq_harmonic(int w, int Nw_eff, int real_op)
{
    // CreateDFTiDFTMats.c:44-50 : slot -> integer harmonic m (FFT ordering for complex A)
    if (real_op) return w;
    if ((Nw_eff & 1) == 0) return (w <= Nw_eff / 2 - 1) ? w : (int64_t)w - Nw_eff;
    return (w <= (Nw_eff - 1) / 2) ? w : (int64_t)w - Nw_eff;
}

__global__ void idft_coef_kernel(cplx *coef, int S, int Nw_eff, int64_t j_start, int dir, int64_t Ns, int real_op)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * Nw_eff) return;
    const int sidx = t / Nw_eff, w = t - sidx * Nw_eff;
    const int64_t P = 2 * Ns;
    int64_t j = (j_start + (int64_t)dir * sidx) % P;
    if (j < 0) j += P;
    const int64_t m = freq_harmonic(w, Nw_eff, real_op);
    int64_t r = (m * j) % P;
    if (r < 0) r += P;
    double sn, cs;
    sincospi((double)r / (double)Ns, &sn, &cs);   // exp(2*pi*i*r/(2Ns)) = exp(i*pi*r/Ns)
    const double sc = 2.0 / (double)P;            // CreateDFTiDFTMats.c:55
    coef[t] = make_double2(cs * sc, sn * sc);
}

__global__ void dft_coef_kernel(cplx *coef, int Nw_eff, int Nw, double scale)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Nw_eff * Nw) return;
    const int wq = t / Nw, j = t - wq * Nw;
    const int64_t r = ((int64_t)wq * j) % Nw;
    double sn, cs;
    sincospi(-2.0 * (double)r / (double)Nw, &sn, &cs);   // CreateDFTiDFTMats.c:74-84
    coef[t] = make_double2(cs * scale, sn * scale);
}

cudaError_t launch_idft_coef(cplx *coef, int S, int Nw_eff, int64_t j_start, int dir, int64_t Ns, int real_op,
                             cudaStream_t s)
{
    const int n = S * Nw_eff;
    idft_coef_kernel<<<(n + 127) / 128, 128, 0, s>>>(coef, S, Nw_eff, j_start, dir, Ns, real_op);
    return cudaGetLastError();
}

cudaError_t launch_dft_coef(cplx *coef, int Nw_eff, int Nw, int64_t ResRatio, cudaStream_t s)
{
    const int n = Nw_eff * Nw;
    dft_coef_kernel<<<(n + 127) / 128, 128, 0, s>>>(coef, Nw_eff, Nw, (double)ResRatio);
    return cudaGetLastError();
}

// =====================================================================================================
// K2/K3: out[s][u] = sum_w coef[s][w] * in[w][u]   (complex double, K = #frequencies or #snapshots, small)
//   thread <-> UT consecutive-strided units, ST outputs in registers; coef tile in shared memory, read as
//   warp-wide broadcasts.  `in` is read once per s-tile (unit stride), `out` written once.
//   Bound: FP64 pipe (8*K flops per output) unless K is tiny; HBM traffic (K + S)*16 B per unit and s-tile.
// =====================================================================================================
template <int ST, int UT>
__global__ void __launch_bounds__(128)
small_cgemm_kernel(cplx *__restrict__ out, int64_t out_stride, const cplx *__restrict__ coef,
                   const cplx *__restrict__ in, int64_t in_stride, int S, int K, int64_t U)
{
    extern __shared__ cplx sm_coef[];  // [K][ST]
    const int s0 = blockIdx.y * ST;
    for (int t = threadIdx.x; t < K * ST; t += blockDim.x) {
        const int w = t / ST, ss = t - w * ST;
        sm_coef[t] = (s0 + ss < S) ? coef[(int64_t)(s0 + ss) * K + w] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int64_t ub = ((int64_t)blockIdx.x * blockDim.x) * UT + threadIdx.x;
    cplx acc[UT][ST];
#pragma unroll
    for (int i = 0; i < UT; i++)
#pragma unroll
        for (int j = 0; j < ST; j++) acc[i][j] = make_double2(0.0, 0.0);
    bool ok[UT];
#pragma unroll
    for (int i = 0; i < UT; i++) ok[i] = (ub + (int64_t)i * blockDim.x) < U;

    for (int w = 0; w < K; w++) {
        cplx v[UT];
#pragma unroll
        for (int i = 0; i < UT; i++)
            v[i] = ok[i] ? ldg(in + (int64_t)w * in_stride + ub + (int64_t)i * blockDim.x) : make_double2(0.0, 0.0);
        const cplx *cw = sm_coef + w * ST;
#pragma unroll
        for (int j = 0; j < ST; j++) {
            const cplx c = cw[j];
#pragma unroll
            for (int i = 0; i < UT; i++) cfma(acc[i][j], c, v[i]);
        }
    }
#pragma unroll
    for (int j = 0; j < ST; j++) {
        if (s0 + j < S) {
#pragma unroll
            for (int i = 0; i < UT; i++)
                if (ok[i]) out[(int64_t)(s0 + j) * out_stride + ub + (int64_t)i * blockDim.x] = acc[i][j];
        }
    }
}

template <int ST, int UT>
static cudaError_t small_cgemm_launch(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in,
                                      int64_t in_stride, int S, int K, int64_t U, cudaStream_t s)
{
    const int threads = 128;
    const size_t smem = (size_t)K * ST * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(small_cgemm_kernel<ST, UT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((U + (int64_t)threads * UT - 1) / ((int64_t)threads * UT)), (unsigned)((S + ST - 1) / ST));
    small_cgemm_kernel<ST, UT><<<grid, threads, smem, s>>>(out, out_stride, coef, in, in_stride, S, K, U);
    return cudaGetLastError();
}

cudaError_t launch_small_cgemm(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride,
                               int S, int K, int64_t U, int variant, cudaStream_t s)
{
    if (U == 0 || S == 0) return cudaSuccess;
    switch (variant) {
        case 1: return small_cgemm_launch<16, 2>(out, out_stride, coef, in, in_stride, S, K, U, s);
        case 2: return small_cgemm_launch<8, 2>(out, out_stride, coef, in, in_stride, S, K, U, s);
        case 3: return small_cgemm_launch<16, 1>(out, out_stride, coef, in, in_stride, S, K, U, s);
        default: return small_cgemm_launch<32, 1>(out, out_stride, coef, in, in_stride, S, K, U, s);
    }
}

// =====================================================================================================
// K2b: forcing synthesis through Taylor moments.  Within a chunk of S consecutive half steps centred at j_c,
//   f_s = sum_w Fhat_w g_w exp(i m_w phi (j_s - j_c)),  phi = pi/Ns,  g_w = exp(i m_w phi j_c)/Ns,
// and |m_w phi (j_s - j_c)| <= x = m_max*phi*S/2 is tiny (0.08 for the bench grid), so with tau = (j_s-j_c)/(S/2)
//   f_s = sum_{p<P} C_p tau^p,   C_p = sum_w Fhat_w g_w (i m_w phi dir S/2)^p / p!
// to a truncation error x^P/P! chosen below 2^-60 relative to sum_w |Fhat_w|/Ns, i.e. below the rounding error of the
// direct sum.  The contraction over frequencies shrinks from S to P outputs (P ~ 11 for S = 64) and the S slices are
// produced by a Horner evaluation from registers: the kernel turns from FP64-bound into HBM-bound
// (reads Fhat once per chunk, writes the S slices).
// =====================================================================================================
int idft_taylor_terms(double x)
{
    // smallest P with sum_{p>=P} x^p/p! < 2^-60; -1 if more than 16 terms would be needed
    double term = 1.0;
    for (int p = 1; p <= 16; p++) {
        term *= x / p;
        const double tail = term / (1.0 - x / (p + 1));
        if (x < p + 1 && tail < 8.7e-19) return p;
    }
    return -1;
}

__global__ void idft_taylor_coef_kernel(cplx *coef, int P, int K, int64_t J2, int dir, int64_t Ns, double h, int real_op)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= K) return;
    const int64_t m = freq_harmonic(w, K, real_op);
    const int64_t P4 = 4 * Ns;
    int64_t r = (m * (J2 % P4)) % P4;      // phase of the chunk centre j_c = J2/2 :  m*j_c*pi/Ns = pi*r/(2Ns)
    if (r < 0) r += P4;
    double sn, cs;
    sincospi((double)r / (double)(2 * Ns), &sn, &cs);
    const double sc = 1.0 / (double)Ns;    // CreateDFTiDFTMats.c:55 (2/Nt, Nt = 2Ns)
    cplx term = make_double2(cs * sc, sn * sc);
    const double q = (double)m * (3.14159265358979323846 / (double)Ns) * (double)dir * h;  // z = i*q
    coef[w] = term;
    for (int p = 1; p < P; p++) {
        const double f = q / (double)p;
        term = make_double2(-term.y * f, term.x * f);   // term *= i*q/p
        coef[(int64_t)p * K + w] = term;
    }
}

cudaError_t launch_idft_taylor_coef(cplx *coef, int P, int K, int64_t J2, int dir, int64_t Ns, double h, int real_op,
                                    cudaStream_t s)
{
    idft_taylor_coef_kernel<<<(K + 63) / 64, 64, 0, s>>>(coef, P, K, J2, dir, Ns, h, real_op);
    return cudaGetLastError();
}

// cp.async (LDGSTS) helpers: 16-byte global -> shared copies that bypass L1 and the register file
DEVFN void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
DEVFN void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPEND> DEVFN void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND) : "memory"); }

// Each thread streams its own UT units of Fhat through a private STAGES-deep shared-memory ring (no block-level
// synchronisation in the loop): enough bytes in flight to cover HBM latency at 3-4 resident CTAs per SM.
template <int PT, int UT, int STAGES>
__global__ void __launch_bounds__(128)
idft_taylor_kernel(cplx *__restrict__ out, int64_t out_stride, const cplx *__restrict__ coef,
                   const cplx *__restrict__ in, int64_t in_stride, int S, int P, int K, int64_t U, cplx *__restrict__ out_prev)
{
    extern __shared__ cplx sm_dyn[];
    cplx *sm_coef = sm_dyn;                 // [K][PT]
    cplx *ring = sm_dyn + (size_t)K * PT;   // [STAGES][UT][128]
    for (int t = threadIdx.x; t < K * PT; t += blockDim.x) {
        const int w = t / PT, p = t - w * PT;
        sm_coef[t] = (p < P) ? coef[(int64_t)p * K + w] : make_double2(0.0, 0.0);
    }
    const int64_t ub = ((int64_t)blockIdx.x * blockDim.x) * UT + threadIdx.x;
    bool ok[UT];
#pragma unroll
    for (int i = 0; i < UT; i++) ok[i] = (ub + (int64_t)i * blockDim.x) < U;
    // out-of-range lanes read zeros from their (never written) ring slots
#pragma unroll
    for (int i = 0; i < UT; i++)
        if (!ok[i])
            for (int st = 0; st < STAGES; st++) ring[(st * UT + i) * 128 + threadIdx.x] = make_double2(0.0, 0.0);
#pragma unroll
    for (int st = 0; st < STAGES - 1; st++) {
        if (st < K) {
#pragma unroll
            for (int i = 0; i < UT; i++)
                if (ok[i]) cp_async16(ring + (st * UT + i) * 128 + threadIdx.x, in + (int64_t)st * in_stride + ub + (int64_t)i * blockDim.x);
        }
        cp_async_commit();
    }
    __syncthreads();   // coefficient table ready

    cplx acc[UT][PT];
#pragma unroll
    for (int i = 0; i < UT; i++)
#pragma unroll
        for (int j = 0; j < PT; j++) acc[i][j] = make_double2(0.0, 0.0);

    for (int w = 0; w < K; w++) {
        const int wn = w + STAGES - 1;
        if (wn < K) {
            const int st = wn % STAGES;
#pragma unroll
            for (int i = 0; i < UT; i++)
                if (ok[i]) cp_async16(ring + (st * UT + i) * 128 + threadIdx.x, in + (int64_t)wn * in_stride + ub + (int64_t)i * blockDim.x);
        }
        cp_async_commit();
        cp_async_wait<STAGES - 1>();   // the group that carried slice w has landed (own data: no barrier needed)
        const int sc = w % STAGES;
        cplx v[UT];
#pragma unroll
        for (int i = 0; i < UT; i++) v[i] = ring[(sc * UT + i) * 128 + threadIdx.x];
        const cplx *cw = sm_coef + w * PT;
#pragma unroll
        for (int j = 0; j < PT; j++) {
            const cplx c = cw[j];
#pragma unroll
            for (int i = 0; i < UT; i++) cfma(acc[i][j], c, v[i]);
        }
    }
    const double inv_h = 2.0 / (double)S;
    for (int sidx = out_prev ? -1 : 0; sidx < S; sidx++) {
        const double tau = ((double)(2 * sidx - S + 1) * 0.5) * inv_h;   // (j_s - j_c) / (S/2)
#pragma unroll
        for (int i = 0; i < UT; i++) {
            cplx f = acc[i][PT - 1];
#pragma unroll
            for (int j = PT - 2; j >= 0; j--) {
                f.x = fma(f.x, tau, acc[i][j].x);
                f.y = fma(f.y, tau, acc[i][j].y);
            }
            if (ok[i]) {
                if (sidx < 0) out_prev[ub + (int64_t)i * blockDim.x] = f;
                else out[(int64_t)sidx * out_stride + ub + (int64_t)i * blockDim.x] = f;
            }
        }
    }
}

template <int PT, int UT, int STAGES>
static cudaError_t idft_taylor_launch(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride,
                                      int S, int P, int K, int64_t U, cplx *out_prev, cudaStream_t s)
{
    const int threads = 128;
    const size_t smem = ((size_t)K * PT + (size_t)STAGES * UT * 128) * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(idft_taylor_kernel<PT, UT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((U + (int64_t)threads * UT - 1) / ((int64_t)threads * UT));
    idft_taylor_kernel<PT, UT, STAGES><<<grid, threads, smem, s>>>(out, out_stride, coef, in, in_stride, S, P, K, U, out_prev);
    return cudaGetLastError();
}

cudaError_t launch_idft_taylor(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride, int S,
                               int P, int K, int64_t U, cplx *out_prev, cudaStream_t s)
{
    if (U == 0 || S == 0) return cudaSuccess;
    if (P <= 8) return idft_taylor_launch<8, 2, 8>(out, out_stride, coef, in, in_stride, S, P, K, U, out_prev, s);
    if (P <= 12) return idft_taylor_launch<12, 2, 8>(out, out_stride, coef, in, in_stride, S, P, K, U, out_prev, s);
    if (P <= 16) return idft_taylor_launch<16, 2, 8>(out, out_stride, coef, in, in_stride, S, P, K, U, out_prev, s);
    return cudaErrorInvalidValue;
}

// =====================================================================================================
// Layout conversion at the C-ABI boundary: reference column-major (column c = a*ca + b*cb) <-> stack T[a][row][b]
// =====================================================================================================
// stack side: slot stride `srows` rows, the nrows host rows map to stack rows [roff, roff + nrows)
template <bool TO_STACK>
__global__ void relayout_kernel(cplx *__restrict__ dst, const cplx *__restrict__ src, int64_t nrows, int na, int nb,
                                int64_t ca, int64_t cb, int64_t srows, int64_t roff)
{
    // tile: 32 rows x 32 columns of one slot a; the shared tile keeps both global sides coalesced
    __shared__ cplx tile[32][33];
    const int a = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int b0 = blockIdx.z * 32;
    const int nr = (int)min((int64_t)32, nrows - r0);
    const int nbt = min(32, nb - b0);
    if (TO_STACK) {
        for (int t = threadIdx.x; t < 32 * nbt; t += blockDim.x) {
            const int b = t / 32, r = t - b * 32;
            if (r < nr) tile[r][b] = src[((int64_t)a * ca + (int64_t)(b0 + b) * cb) * nrows + r0 + r];
        }
        __syncthreads();
        cplx *d = dst + ((int64_t)a * srows + roff + r0) * nb + b0;
        for (int t = threadIdx.x; t < nr * nbt; t += blockDim.x) {
            const int r = t / nbt, b = t - r * nbt;
            d[(int64_t)r * nb + b] = tile[r][b];
        }
    } else {
        const cplx *sp = src + ((int64_t)a * srows + roff + r0) * nb + b0;
        for (int t = threadIdx.x; t < nr * nbt; t += blockDim.x) {
            const int r = t / nbt, b = t - r * nbt;
            tile[r][b] = sp[(int64_t)r * nb + b];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 32 * nbt; t += blockDim.x) {
            const int b = t / 32, r = t - b * 32;
            if (r < nr) dst[((int64_t)a * ca + (int64_t)(b0 + b) * cb) * nrows + r0 + r] = tile[r][b];
        }
    }
}

cudaError_t launch_host_to_stack(cplx *dst, const cplx *src, int64_t nrows, int na, int nb, int64_t ca, int64_t cb,
                                 int64_t srows, int64_t roff, cudaStream_t s)
{
    if (nrows == 0 || na == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)((nrows + 31) / 32), (unsigned)na, (unsigned)((nb + 31) / 32));
    relayout_kernel<true><<<grid, 256, 0, s>>>(dst, src, nrows, na, nb, ca, cb, srows, roff);
    return cudaGetLastError();
}

cudaError_t launch_stack_to_host(cplx *dst, const cplx *src, int64_t nrows, int na, int nb, int64_t ca, int64_t cb,
                                 int64_t srows, int64_t roff, cudaStream_t s)
{
    if (nrows == 0 || na == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)((nrows + 31) / 32), (unsigned)na, (unsigned)((nb + 31) / 32));
    relayout_kernel<false><<<grid, 256, 0, s>>>(dst, src, nrows, na, nb, ca, cb, srows, roff);
    return cudaGetLastError();
}

// =====================================================================================================
// Per-frequency Gram matrices G_w = Y_w^H Y_w (k x k), deterministic: per-block partials then a fixed-order sum.
// =====================================================================================================
static const int GRAM_ROWS_PER_BLOCK = 2048;
int gram_num_blocks(int64_t nrows) { return (int)((nrows + GRAM_ROWS_PER_BLOCK - 1) / GRAM_ROWS_PER_BLOCK); }

__global__ void __launch_bounds__(128)
gram_partial_kernel(const cplx *__restrict__ Y, int64_t nrows, int k, int64_t slot_stride, cplx *__restrict__ partial)
{
    extern __shared__ cplx tile[];  // [64][k]
    const int w = blockIdx.y, nw = gridDim.y;
    const int64_t r_begin = (int64_t)blockIdx.x * GRAM_ROWS_PER_BLOCK;
    const int64_t r_end = min(nrows, r_begin + GRAM_ROWS_PER_BLOCK);
    const cplx *Yw = Y + (int64_t)w * slot_stride;
    const int npairs = k * k;
    // each thread owns up to 8 (i,j) pairs (k <= 32 -> 1024 pairs / 128 threads)
    cplx acc[8];
#pragma unroll
    for (int q = 0; q < 8; q++) acc[q] = make_double2(0.0, 0.0);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 64) {
        const int nr = (int)min((int64_t)64, r_end - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < nr * k; t += blockDim.x) tile[t] = Yw[r0 * k + t];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int pr = threadIdx.x + q * 128;
            if (pr < npairs) {
                const int i = pr / k, j = pr - i * k;
                cplx sacc = acc[q];
                for (int r = 0; r < nr; r++) cfma_conj(sacc, tile[r * k + i], tile[r * k + j]);
                acc[q] = sacc;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int pr = threadIdx.x + q * 128;
        if (pr < npairs) partial[((int64_t)blockIdx.x * nw + w) * npairs + pr] = acc[q];
    }
}

__global__ void gram_reduce_kernel(const cplx *__restrict__ partial, int nblk, int64_t n_out, cplx *__restrict__ G)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    cplx s = make_double2(0.0, 0.0);
    for (int b = 0; b < nblk; b++) {
        const cplx p = partial[(int64_t)b * n_out + t];
        s.x += p.x;
        s.y += p.y;
    }
    G[t] = s;
}

cudaError_t launch_gram(const cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, cplx *partial, cplx *G,
                        cudaStream_t s)
{
    if (k > 32) return cudaErrorInvalidValue;
    const int nblk = gram_num_blocks(nrows);
    dim3 grid((unsigned)nblk, (unsigned)nw);
    gram_partial_kernel<<<grid, 128, (size_t)64 * k * sizeof(cplx), s>>>(Y, nrows, k, slot_stride, partial);
    const int64_t n_out = (int64_t)nw * k * k;
    gram_reduce_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(partial, nblk, n_out, G);
    return cudaGetLastError();
}

// Cross-product SVD of a tall block through its Gram matrix (SLEPc's default SVD type, SVDAllFreqs.c:36-40):
// cyclic complex Jacobi on G_w; sigma = sqrt(lambda) sorted descending; M = V diag(1/sigma) so that U = Y M.
// One warp per frequency, G and V in shared memory (k <= 32: 2 x 16 KB): lane r owns row r in the column updates and
// column r in the row update, so a rotation is three conflict-light passes of one element pair per lane instead of 3k
// serial ones; the rotation angles and the order (p, q) are those of the serial sweep, so the result does not depend on
// the lane count.  (Round 1 ran one THREAD per frequency with 32 KB of local stack: 0.3-3.3 ms per call at Nw = 32.)
__global__ void __launch_bounds__(32)
cross_svd_small_kernel(const cplx *__restrict__ Gin, int k, int nw, cplx *__restrict__ M, double *__restrict__ sigma)
{
    extern __shared__ cplx jsm[];
    cplx *G = jsm, *V = jsm + k * k;   // row-major
    __shared__ int ord[32];
    const int w = blockIdx.x, lane = threadIdx.x;
    for (int t = lane; t < k * k; t += 32) { G[t] = Gin[(int64_t)w * k * k + t]; V[t] = make_double2(0.0, 0.0); }
    __syncwarp();
    if (lane < k) V[lane * k + lane] = make_double2(1.0, 0.0);
    __syncwarp();
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dg = 0.0;   // (every lane sums all entries in the same order: one decision for the warp)
        for (int r = 0; r < k; r++)
            for (int c = 0; c < k; c++) {
                const double a2 = G[r * k + c].x * G[r * k + c].x + G[r * k + c].y * G[r * k + c].y;
                if (r == c) dg += a2; else off += a2;
            }
        if (off <= 1e-32 * dg) break;
        for (int p = 0; p < k - 1; p++)
            for (int q = p + 1; q < k; q++) {
                const cplx gpq = G[p * k + q];
                const double apq = sqrt(gpq.x * gpq.x + gpq.y * gpq.y);
                if (apq == 0.0) continue;   // (uniform over the warp)
                const double app = G[p * k + p].x, aqq = G[q * k + q].x;
                const cplx ph = make_double2(gpq.x / apq, gpq.y / apq);  // e^{i phi}
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
                const cplx sph = make_double2(sn * ph.x, sn * ph.y);    // s e^{i phi}
                const cplx sphc = make_double2(sn * ph.x, -sn * ph.y);  // s e^{-i phi}
                __syncwarp();
                if (lane < k) {  // G <- G J : col p = c*gp - s e^{-i phi} gq ; col q = s e^{i phi} gp + c gq
                    const int r = lane;
                    const cplx gp = G[r * k + p], gq = G[r * k + q];
                    const cplx a = cmul(sphc, gq), b = cmul(sph, gp);
                    G[r * k + p] = make_double2(c * gp.x - a.x, c * gp.y - a.y);
                    G[r * k + q] = make_double2(b.x + c * gq.x, b.y + c * gq.y);
                }
                __syncwarp();
                if (lane < k) {  // G <- J^H G : row p = c*gp - s e^{i phi} gq ; row q = s e^{-i phi} gp + c gq
                    const int cc = lane;
                    const cplx gp = G[p * k + cc], gq = G[q * k + cc];
                    const cplx a = cmul(sph, gq), b = cmul(sphc, gp);
                    G[p * k + cc] = make_double2(c * gp.x - a.x, c * gp.y - a.y);
                    G[q * k + cc] = make_double2(b.x + c * gq.x, b.y + c * gq.y);
                    const int r = lane;   // V <- V J
                    const cplx vp = V[r * k + p], vq = V[r * k + q];
                    const cplx a2 = cmul(sphc, vq), b2 = cmul(sph, vp);
                    V[r * k + p] = make_double2(c * vp.x - a2.x, c * vp.y - a2.y);
                    V[r * k + q] = make_double2(b2.x + c * vq.x, b2.y + c * vq.y);
                }
                __syncwarp();
            }
    }
    // sort descending (selection sort on indices)
    if (lane == 0) {
        for (int i = 0; i < k; i++) ord[i] = i;
        for (int i = 0; i < k; i++)
            for (int j = i + 1; j < k; j++)
                if (G[ord[j] * k + ord[j]].x > G[ord[i] * k + ord[i]].x) { int tt = ord[i]; ord[i] = ord[j]; ord[j] = tt; }
    }
    __syncwarp();
    for (int mI = 0; mI < k; mI++) {
        const double lam = G[ord[mI] * k + ord[mI]].x;
        const double sg = sqrt(lam > 0.0 ? lam : 0.0);
        if (lane == 0) sigma[(int64_t)w * k + mI] = sg;
        const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
        if (lane < k) {
            const cplx v = V[lane * k + ord[mI]];
            M[(int64_t)w * k * k + lane * k + mI] = make_double2(v.x * inv, v.y * inv);
        }
    }
}

cudaError_t launch_cross_svd_small(const cplx *G, int k, int nw, cplx *M, double *sigma, cudaStream_t s)
{
    if (k > 32) return cudaErrorInvalidValue;
    if (nw <= 0) return cudaSuccess;
    cross_svd_small_kernel<<<nw, 32, 2 * (size_t)k * k * sizeof(cplx), s>>>(G, k, nw, M, sigma);
    return cudaGetLastError();
}

// Y_w <- Y_w * M_w, in place.  Block = 64-row tile of one frequency; tile and M staged in shared memory.
__global__ void __launch_bounds__(256)
apply_right_kernel(cplx *__restrict__ Y, int64_t nrows, int k, const cplx *__restrict__ M)
{
    extern __shared__ cplx sm[];  // M [k*k] then tile [64*k]
    cplx *sM = sm, *tile = sm + k * k;
    const int w = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int nr = (int)min((int64_t)64, nrows - r0);
    cplx *Yw = Y + ((int64_t)w * nrows + r0) * k;
    for (int t = threadIdx.x; t < k * k; t += blockDim.x) sM[t] = M[(int64_t)w * k * k + t];
    for (int t = threadIdx.x; t < nr * k; t += blockDim.x) tile[t] = Yw[t];
    __syncthreads();
    for (int t = threadIdx.x; t < nr * k; t += blockDim.x) {
        const int r = t / k, j = t - r * k;
        cplx sacc = make_double2(0.0, 0.0);
        for (int i = 0; i < k; i++) cfma(sacc, tile[r * k + i], sM[i * k + j]);
        Yw[t] = sacc;
    }
}

cudaError_t launch_apply_right(cplx *Y, int64_t nrows, int k, int nw, const cplx *M, cudaStream_t s)
{
    if (nrows == 0) return cudaSuccess;
    dim3 grid((unsigned)((nrows + 63) / 64), (unsigned)nw);
    const size_t smem = ((size_t)k * k + (size_t)64 * k) * sizeof(cplx);
    apply_right_kernel<<<grid, 256, smem, s>>>(Y, nrows, k, M);
    return cudaGetLastError();
}

// =====================================================================================================
// misc
// =====================================================================================================
__global__ void __launch_bounds__(256) norm2_partial_kernel(const cplx *__restrict__ q, int64_t n, double *partial)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const cplx v = q[i];
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void norm2_final_kernel(const double *partial, int nblk, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < nblk; b++) s += partial[b];
        *out = s;
    }
}

cudaError_t launch_norm2(const cplx *q, int64_t n, double *partial, double *out, cudaStream_t s)
{
    const int nblk = (int)min((int64_t)296, (n + 255) / 256);
    norm2_partial_kernel<<<nblk, 256, 0, s>>>(q, n, partial);
    norm2_final_kernel<<<1, 32, 0, s>>>(partial, nblk, out);
    return cudaGetLastError();
}

cudaError_t launch_fill_zero(cplx *p, int64_t n, cudaStream_t s)
{
    return cudaMemsetAsync(p, 0, (size_t)n * sizeof(cplx), s);
}

// out[a][r][b] = sum_c M[r][c] * in[a][c][b]   (ApplyWeightMats.c:18-76, sparse x dense on the stack layout)
__global__ void __launch_bounds__(256)
spmm_stack_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                  int64_t nrows, int64_t ncols, const cplx *__restrict__ in, cplx *__restrict__ out, int nb)
{
    const int a = blockIdx.y;
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nrows * nb) return;
    const int64_t r = u / nb;
    const int b = (int)(u - r * nb);
    const cplx *ina = in + (int64_t)a * ncols * nb + b;
    cplx s = make_double2(0.0, 0.0);
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) cfma(s, ldg(val + p), ldg(ina + (int64_t)col[p] * nb));
    out[(int64_t)a * nrows * nb + u] = s;
}

cudaError_t launch_spmm_stack(const CsrDev &M, const cplx *in, cplx *out, int na, int nb, cudaStream_t s)
{
    if (M.nrows == 0 || na == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)((M.nrows * nb + 255) / 256), (unsigned)na);
    spmm_stack_kernel<<<grid, 256, 0, s>>>(M.rowptr, M.col, M.val, M.nrows, M.ncols, in, out, nb);
    return cudaGetLastError();
}

// =====================================================================================================
// Halo exchange over NVLink peer memory (row-partitioned runs, one process per GPU; PETSc's VecScatter of MPIAIJ).
//   push : for every outgoing message, gather the rows the peer needs and store them straight into the peer's
//          mailbox (peer pointer obtained through CUDA IPC) -- remote stores ride NVLink, no staging copy;
//   signal: one thread per peer publishes the epoch with a system-scope fence (release) in the peer's flag array;
//   wait+unpack: every CTA spins (acquire) until all expected peers have published the epoch, then copies the
//          mailbox rows into the ghost rows of the destination block vector.
// =====================================================================================================
struct HaloMsgDev {
    cplx *peer_mailbox;        // base of my region in the peer's mailbox slot (already offset for this epoch parity)
    const int32_t *rows;       // local rows to send
    int count;
};

__global__ void __launch_bounds__(256)
halo_push_kernel(const cplx *__restrict__ src, int nvec, const HaloMsgDev *__restrict__ msgs, int parity, int nvec_max)
{
    const HaloMsgDev m = msgs[blockIdx.y];
    const int64_t n = (int64_t)m.count * nvec;
    cplx *dst = m.peer_mailbox + (int64_t)parity * m.count * nvec_max;   // two slots per message, chosen by epoch parity
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / nvec;
        const int iv = (int)(e - r * nvec);
        dst[e] = src[(int64_t)m.rows[r] * nvec + iv];
    }
}

__global__ void halo_signal_kernel(unsigned long long *const *__restrict__ peer_flags, int npeers, unsigned long long epoch)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npeers) return;
    __threadfence_system();   // the push kernel's stores (previous kernel in this stream) are ordered before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(peer_flags[p]), "l"(epoch) : "memory");
}

struct HaloInDev {
    const cplx *mailbox;       // my mailbox region holding this peer's rows (slot 0)
    const unsigned long long *flag;
    int64_t dest_row0;
    int count;
};

__global__ void __launch_bounds__(256)
halo_wait_unpack_kernel(cplx *__restrict__ dst, int nvec, const HaloInDev *__restrict__ ins, int parity, int nvec_max,
                        unsigned long long epoch)
{
    const HaloInDev m = ins[blockIdx.y];
    if (threadIdx.x == 0) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(m.flag) : "memory");
        } while (v < epoch);
    }
    __syncthreads();
    const int64_t n = (int64_t)m.count * nvec;
    const cplx *src = m.mailbox + (int64_t)parity * m.count * nvec_max;
    cplx *d = dst + m.dest_row0 * nvec;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        d[e] = __ldcv(src + e);   // bypass caches: the data were written by another GPU
}

// The whole exchange as ONE launch of ONE CTA, for small messages (narrow-band operators: a few rows per neighbour): push every
// outgoing message, publish the epoch, wait for the peers' epochs, unpack.  Three launches per exchange cost more than the
// exchange itself there (8 us of a 220 us RK4 step at 4 GPUs).  The wait is bounded like the fused stage kernel's.
__global__ void __launch_bounds__(512)
halo_exchange_small_kernel(cplx *__restrict__ buf, int nvec, const HaloMsgDev *__restrict__ msgs, int nout,
                           unsigned long long *const *__restrict__ peer_flags, const HaloInDev *__restrict__ ins, int nin, int parity,
                           int nvec_max, unsigned long long epoch, unsigned long long *timeout_flag,
                           const unsigned long long *ready, unsigned long long ready_target)
{
    if (ready) {
        // launched next to the step kernel that produces `buf`: wait until its boundary tiles have counted themselves
        if (threadIdx.x == 0) {
            unsigned long long v;
            long long spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(ready) : "memory");
                if (++spins > 600000000ll) { if (timeout_flag) *timeout_flag = 1ull; break; }
            } while (v < ready_target);
        }
        __syncthreads();
    }
    for (int i = 0; i < nout; i++) {
        const HaloMsgDev m = msgs[i];
        const int64_t n = (int64_t)m.count * nvec;
        cplx *dst = m.peer_mailbox + (int64_t)parity * m.count * nvec_max;
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
            const int64_t r = e / nvec;
            dst[e] = __ldcg(buf + (int64_t)m.rows[r] * nvec + (int)(e - r * nvec));   // (written by a kernel that may still be running)
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < nout)
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(peer_flags[threadIdx.x]), "l"(epoch) : "memory");
    if ((int)threadIdx.x < nin) {
        unsigned long long v;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(ins[threadIdx.x].flag) : "memory");
            if (++spins > 600000000ll) { if (timeout_flag) *timeout_flag = 1ull; break; }
        } while (v < epoch);
    }
    __syncthreads();
    for (int i = 0; i < nin; i++) {
        const HaloInDev m = ins[i];
        const int64_t n = (int64_t)m.count * nvec;
        const cplx *src = m.mailbox + (int64_t)parity * m.count * nvec_max;
        cplx *d = buf + m.dest_row0 * nvec;
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) d[e] = __ldcv(src + e);   // bypass caches: written by another GPU
    }
}

cudaError_t launch_halo_exchange_small(cplx *buf, int nvec, const void *msgs_dev, int nout, void *peer_flags_dev, const void *ins_dev,
                                       int nin, int parity, int nvec_max, unsigned long long epoch, unsigned long long *timeout_flag,
                                       const unsigned long long *ready, unsigned long long ready_target, cudaStream_t s)
{
    if (nout == 0 && nin == 0) return cudaSuccess;
    if (nout > 512 || nin > 512) return cudaErrorInvalidValue;
    halo_exchange_small_kernel<<<1, 512, 0, s>>>(buf, nvec, (const HaloMsgDev *)msgs_dev, nout, (unsigned long long *const *)peer_flags_dev,
                                                 (const HaloInDev *)ins_dev, nin, parity, nvec_max, epoch, timeout_flag, ready,
                                                 ready_target);
    return cudaGetLastError();
}

cudaError_t launch_halo_push(const cplx *src, int nvec, const void *msgs_dev, int nmsgs, int max_count, int parity,
                             int nvec_max, cudaStream_t s)
{
    if (nmsgs == 0) return cudaSuccess;
    const int64_t n = (int64_t)max_count * nvec;
    const unsigned gx = (unsigned)min((int64_t)64, (n + 255) / 256);
    dim3 grid(gx ? gx : 1, (unsigned)nmsgs);
    halo_push_kernel<<<grid, 256, 0, s>>>(src, nvec, (const HaloMsgDev *)msgs_dev, parity, nvec_max);
    return cudaGetLastError();
}

cudaError_t launch_halo_signal(void *peer_flags_dev, int npeers, unsigned long long epoch, cudaStream_t s)
{
    if (npeers == 0) return cudaSuccess;
    halo_signal_kernel<<<1, 32, 0, s>>>((unsigned long long *const *)peer_flags_dev, npeers, epoch);
    return cudaGetLastError();
}

cudaError_t launch_halo_wait_unpack(cplx *dst, int nvec, const void *ins_dev, int nin, int max_count, int parity,
                                    int nvec_max, unsigned long long epoch, cudaStream_t s)
{
    if (nin == 0) return cudaSuccess;
    const int64_t n = (int64_t)max_count * nvec;
    const unsigned gx = (unsigned)min((int64_t)64, (n + 255) / 256);
    dim3 grid(gx ? gx : 1, (unsigned)nin);
    halo_wait_unpack_kernel<<<grid, 256, 0, s>>>(dst, nvec, (const HaloInDev *)ins_dev, parity, nvec_max, epoch);
    return cudaGetLastError();
}

// =====================================================================================================
// Transient removal (TransientRemovalStrategy.c:30-40): tall-skinny products on snapshot stacks, batched over test vectors.
//   A stack element (slot a, row, vector z) lives at  base[a*slot_stride + row*nv + z].
//   These run once per action (the RK4 loop runs ~45,000 steps), so they are written for clarity and determinism.
// =====================================================================================================
// partial[chunk][z][a][b] = sum_{rows of chunk} conj(X[a][row][z]) * Y[b][row][z]
template <int PB>
__global__ void __launch_bounds__(256)
stack_gram2_kernel(const cplx *__restrict__ X, int64_t xs, int ma, const cplx *__restrict__ Y, int64_t ys, int mb, int nv,
                   int64_t row0, int64_t nrows, int rows_per_chunk, int RT, cplx *__restrict__ partial)
{
    extern __shared__ cplx sm[];   // Xt[RT][ma] then Yt[RT][mb]  (RT rows per tile, sized to the shared memory)
    cplx *Xt = sm, *Yt = sm + RT * ma;
    const int z = blockIdx.y;
    const int batch = blockIdx.z;
    const int npairs = ma * mb;
    const int64_t rb = row0 + (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t re = min(row0 + nrows, rb + rows_per_chunk);
    cplx acc[PB];
    int pa[PB], pbq[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) {
        acc[q] = make_double2(0.0, 0.0);
        const int pr = batch * 256 * PB + q * 256 + threadIdx.x;
        pa[q] = pr < npairs ? pr / mb : -1;
        pbq[q] = pr < npairs ? pr - (pr / mb) * mb : 0;
    }
    for (int64_t r0 = rb; r0 < re; r0 += RT) {
        const int nr = (int)min((int64_t)RT, re - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < RT * ma; t += blockDim.x) {
            const int a = t / RT, r = t - a * RT;
            Xt[r * ma + a] = r < nr ? X[(int64_t)a * xs + (r0 + r) * nv + z] : make_double2(0.0, 0.0);
        }
        for (int t = threadIdx.x; t < RT * mb; t += blockDim.x) {
            const int b = t / RT, r = t - b * RT;
            Yt[r * mb + b] = r < nr ? Y[(int64_t)b * ys + (r0 + r) * nv + z] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PB; q++) {
            if (pa[q] >= 0) {
                cplx s = acc[q];
                for (int r = 0; r < RT; r++) cfma_conj(s, Xt[r * ma + pa[q]], Yt[r * mb + pbq[q]]);
                acc[q] = s;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PB; q++) {
        const int pr = batch * 256 * PB + q * 256 + threadIdx.x;
        if (pr < npairs) partial[((int64_t)blockIdx.x * nv + z) * npairs + pr] = acc[q];
    }
}

int stack_gram2_chunks(int64_t nrows) { return (int)((nrows + 8191) / 8192); }

// G: [nv][ma][mb]; partial: chunks*nv*ma*mb
cudaError_t launch_stack_gram2(const cplx *X, int64_t xs, int ma, const cplx *Y, int64_t ys, int mb, int nv, int64_t row0,
                               int64_t nrows, cplx *partial, cplx *G, cudaStream_t s)
{
    constexpr int PB = 8;
    const int chunks = stack_gram2_chunks(nrows);
    const int npairs = ma * mb;
    int RT = (int)(((size_t)200 * 1024) / ((size_t)(ma + mb) * sizeof(cplx)));   // rows per shared-memory tile
    RT = RT > 32 ? 32 : RT;
    if (RT < 1) return cudaErrorInvalidValue;
    const size_t smem = (size_t)RT * (ma + mb) * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(stack_gram2_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)chunks, (unsigned)nv, (unsigned)((npairs + 256 * PB - 1) / (256 * PB)));
    stack_gram2_kernel<PB><<<grid, 256, smem, s>>>(X, xs, ma, Y, ys, mb, nv, row0, nrows, 8192, RT, partial);
    const int64_t n_out = (int64_t)nv * npairs;
    gram_reduce_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(partial, chunks, n_out, G);
    return cudaGetLastError();
}

// Out[b][row][z] = beta * Z[b][row][z] + alpha * sum_a X[a][row][z] * M[z][a][b]     (M: [nv][ma][mb] row-major)
template <int BT>
__global__ void __launch_bounds__(128)
stack_rmul_kernel(cplx *__restrict__ Out, int64_t os, const cplx *__restrict__ Z, int64_t zs, double beta,
                  const cplx *__restrict__ X, int64_t xs, int ma, const cplx *__restrict__ M, int mb, double alpha, int nv,
                  int64_t nrows)
{
    extern __shared__ cplx sM[];   // [ma][BT]
    const int z = blockIdx.z, b0 = blockIdx.y * BT;
    for (int t = threadIdx.x; t < ma * BT; t += blockDim.x) {
        const int a = t / BT, j = t - a * BT;
        sM[t] = (b0 + j < mb) ? M[((int64_t)z * ma + a) * mb + b0 + j] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int64_t u = row * nv + z;
    cplx acc[BT];
#pragma unroll
    for (int j = 0; j < BT; j++) acc[j] = make_double2(0.0, 0.0);
    for (int a = 0; a < ma; a++) {
        const cplx x = ldg(X + (int64_t)a * xs + u);
#pragma unroll
        for (int j = 0; j < BT; j++) cfma(acc[j], x, sM[a * BT + j]);
    }
#pragma unroll
    for (int j = 0; j < BT; j++) {
        if (b0 + j < mb) {
            cplx o = make_double2(alpha * acc[j].x, alpha * acc[j].y);
            if (Z) {
                const cplx zv = Z[(int64_t)(b0 + j) * zs + u];
                o.x = fma(beta, zv.x, o.x);
                o.y = fma(beta, zv.y, o.y);
            }
            Out[(int64_t)(b0 + j) * os + u] = o;
        }
    }
}

cudaError_t launch_stack_rmul(cplx *Out, int64_t os, const cplx *Z, int64_t zs, double beta, const cplx *X, int64_t xs,
                              int ma, const cplx *M, int mb, double alpha, int nv, int64_t nrows, cudaStream_t s)
{
    constexpr int BT = 16;
    const size_t smem = (size_t)ma * BT * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(stack_rmul_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((nrows + 127) / 128), (unsigned)((mb + BT - 1) / BT), (unsigned)nv);
    stack_rmul_kernel<BT><<<grid, 128, smem, s>>>(Out, os, Z, zs, beta, X, xs, ma, M, mb, alpha, nv, nrows);
    return cudaGetLastError();
}

// Cholesky G_z + shift_z*I = R^H R (upper R, positive diagonal) and Minv_z = R^{-1}; one CTA per z, matrices in global
// memory (m up to a few hundred).  G is overwritten by R.  status[z] = failing column + 1, or 0.
__global__ void __launch_bounds__(128)
chol_inverse_big_kernel(cplx *__restrict__ G, int m, double shift_scale, cplx *__restrict__ Minv, int *__restrict__ status,
                        int drop_dependent)
{
    // drop_dependent: a column whose pivot is not positive (numerically inside the span of the columns before it) is left out
    // -- its row of R and its column of R^{-1} are zero, so the basis simply has one vector fewer -- instead of being reported
    // (status[z] = -(columns left out)).  The transient-removal basis uses it: once the transient has decayed to rounding
    // level snapshot Nw + 1 repeats snapshot 1, and the projection only depends on the span (SLEPc's Gram-Schmidt carries
    // on there with a vector of rounding noise).
    __shared__ unsigned char s_drop[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_drop[i] = 0;
    const int z = blockIdx.x;
    cplx *g = G + (int64_t)z * m * m;
    cplx *mi = Minv + (int64_t)z * m * m;
    __shared__ double s_piv;
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        s_bad = 0;
        double tr = 0.0;
        for (int i = 0; i < m; i++) tr += g[i * m + i].x;
        s_piv = shift_scale * tr;
    }
    __syncthreads();
    const double shift = s_piv;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) g[i * m + i].x += shift;
    __syncthreads();
    for (int j = 0; j < m; j++) {
        if (threadIdx.x == 0) {
            double d = g[j * m + j].x;
            if (!(d > 0.0)) {
                if (drop_dependent && m <= 256) { s_drop[j] = 1; s_bad -= 1; }
                else if (!s_bad) s_bad = j + 1;
                d = 1.0;
            }
            s_piv = sqrt(d);
        }
        __syncthreads();
        const double rjj = s_piv;
        const bool dropped = s_drop[j < 256 ? j : 0] && j < 256;
        for (int c = j + threadIdx.x; c < m; c += blockDim.x) {   // row j of R
            const cplx v = g[j * m + c];
            g[j * m + c] = (c == j) ? make_double2(rjj, 0.0) : (dropped ? make_double2(0.0, 0.0) : make_double2(v.x / rjj, v.y / rjj));
        }
        __syncthreads();
        // trailing update: G[i][c] -= conj(R[j][i]) * R[j][c], i > j, c >= i
        const int nt = m - j - 1;
        for (int t = threadIdx.x; t < nt * nt; t += blockDim.x) {
            const int i = j + 1 + t / nt, c = j + 1 + t % nt;
            if (c >= i) {
                const cplx a = g[j * m + i], b = g[j * m + c];
                g[i * m + c].x -= a.x * b.x + a.y * b.y;
                g[i * m + c].y -= a.x * b.y - a.y * b.x;
            }
        }
        __syncthreads();
    }
    // Minv = R^{-1}: column c by back substitution, one thread per column
    for (int c = threadIdx.x; c < m; c += blockDim.x) {
        for (int r = m - 1; r >= 0; r--) {
            if (r > c) { mi[r * m + c] = make_double2(0.0, 0.0); continue; }
            cplx sacc = make_double2(r == c ? 1.0 : 0.0, 0.0);
            for (int p = r + 1; p <= c; p++) {
                const cplx a = g[r * m + p], b = mi[p * m + c];
                sacc.x -= a.x * b.x - a.y * b.y;
                sacc.y -= a.x * b.y + a.y * b.x;
            }
            const double rr = g[r * m + r].x;
            mi[r * m + c] = make_double2(sacc.x / rr, sacc.y / rr);
        }
        if (c < 256 && s_drop[c])
            for (int r = 0; r < m; r++) mi[r * m + c] = make_double2(0.0, 0.0);   // a column left out contributes nothing
    }
    if (threadIdx.x == 0) status[z] = s_bad;
}

cudaError_t launch_chol_inverse_big(cplx *G, int m, int nz, double shift_scale, cplx *Minv, int *status, cudaStream_t s,
                                    int drop_dependent)
{
    chol_inverse_big_kernel<<<nz, 128, 0, s>>>(G, m, shift_scale, Minv, status, drop_dependent);
    return cudaGetLastError();
}

// Galerkin systems (GalerkinProj.c:48-62): for every (z, iw) solve (alpha_iw I - Mt_z) c = alpha_iw B1_z[:,iw] - B2_z[:,iw]
// by LU with partial pivoting; C_z[a][iw] = c[a].  One CTA per system, the matrix in a global scratch slab.
__global__ void __launch_bounds__(128)
galerkin_solve_kernel(const cplx *__restrict__ Mt, const cplx *__restrict__ B1, const cplx *__restrict__ B2,
                      const cplx *__restrict__ alpha, int m, int nw, cplx *__restrict__ scratch, cplx *__restrict__ C,
                      int *__restrict__ status)
{
    const int iw = blockIdx.x, z = blockIdx.y;
    cplx *T = scratch + ((int64_t)z * nw + iw) * ((int64_t)m * (m + 1));
    cplx *rhs = T + (int64_t)m * m;
    const cplx al = alpha[iw];
    const cplx *mt = Mt + (int64_t)z * m * m;
    for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
        const int r = t / m, c = t - r * m;
        cplx v = make_double2(-mt[t].x, -mt[t].y);
        if (r == c) { v.x += al.x; v.y += al.y; }
        T[t] = v;
    }
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const cplx b1 = B1[((int64_t)z * m + r) * nw + iw], b2 = B2[((int64_t)z * m + r) * nw + iw];
        rhs[r] = make_double2(al.x * b1.x - al.y * b1.y - b2.x, al.x * b1.y + al.y * b1.x - b2.y);
    }
    __shared__ int s_p;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int c = 0; c < m; c++) {
        if (threadIdx.x == 0) {
            int p = c;
            double best = T[c * m + c].x * T[c * m + c].x + T[c * m + c].y * T[c * m + c].y;
            for (int r = c + 1; r < m; r++) {
                const double v = T[r * m + c].x * T[r * m + c].x + T[r * m + c].y * T[r * m + c].y;
                if (v > best) { best = v; p = r; }
            }
            if (best == 0.0) s_bad = c + 1;
            s_p = p;
        }
        __syncthreads();
        const int p = s_p;
        if (p != c) {
            for (int cc = threadIdx.x; cc <= m; cc += blockDim.x) {
                cplx *x = (cc < m) ? &T[c * m + cc] : &rhs[c];
                cplx *y = (cc < m) ? &T[p * m + cc] : &rhs[p];
                const cplx tmp = *x;
                *x = *y;
                *y = tmp;
            }
        }
        __syncthreads();
        const cplx piv = T[c * m + c];
        const double pn = piv.x * piv.x + piv.y * piv.y;
        const cplx pinv = pn > 0.0 ? make_double2(piv.x / pn, -piv.y / pn) : make_double2(0.0, 0.0);
        for (int r = c + 1 + threadIdx.x; r < m; r += blockDim.x) {
            const cplx l = cmul(T[r * m + c], pinv);
            for (int cc = c + 1; cc < m; cc++) {
                const cplx u = T[c * m + cc];
                T[r * m + cc].x -= l.x * u.x - l.y * u.y;
                T[r * m + cc].y -= l.x * u.y + l.y * u.x;
            }
            const cplx bc = rhs[c];
            rhs[r].x -= l.x * bc.x - l.y * bc.y;
            rhs[r].y -= l.x * bc.y + l.y * bc.x;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int r = m - 1; r >= 0; r--) {
            cplx sacc = rhs[r];
            for (int cc = r + 1; cc < m; cc++) {
                const cplx u = T[r * m + cc], x = rhs[cc];
                sacc.x -= u.x * x.x - u.y * x.y;
                sacc.y -= u.x * x.y + u.y * x.x;
            }
            const cplx piv = T[r * m + r];
            const double pn = piv.x * piv.x + piv.y * piv.y;
            rhs[r] = pn > 0.0 ? make_double2((sacc.x * piv.x + sacc.y * piv.y) / pn, (sacc.y * piv.x - sacc.x * piv.y) / pn)
                              : make_double2(0.0, 0.0);
        }
        for (int r = 0; r < m; r++) C[((int64_t)z * m + r) * nw + iw] = rhs[r];
        if (s_bad) status[z * nw + iw] = s_bad;
    }
}

cudaError_t launch_galerkin_solve(const cplx *Mt, const cplx *B1, const cplx *B2, const cplx *alpha, int m, int nw, int nz,
                                  cplx *scratch, cplx *C, int *status, cudaStream_t s)
{
    dim3 grid((unsigned)nw, (unsigned)nz);
    galerkin_solve_kernel<<<grid, 128, 0, s>>>(Mt, B1, B2, alpha, m, nw, scratch, C, status);
    return cudaGetLastError();
}

// =====================================================================================================
// K1s: latency path for small problems (the tutorial operator: N = 500, k = 10).  The whole block vector and the operator
//   fit in the shared memory of ONE CTA, so a single persistent CTA advances many RK4 steps per launch: the four stages are
//   separated by __syncthreads() instead of kernel launches (the ~760,000 dependent stages of the shipped configuration
//   cannot each pay a launch).  y and the RK4 accumulator live in registers, the two stage vectors ping-pong in shared
//   memory, the forcing slices come from L2.  blockIdx.x selects an independent block vector (snapshot slots of the
//   transient-removal step).  Arithmetic and summation order per row are those of K1.
// =====================================================================================================
template <int BT, int UPT, bool FORCED>
__global__ void __launch_bounds__(BT)
rk4_small_persistent_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                            int n, int nnz, int nv, SmallStepArgs a)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int nu = n * nv;
    cplx *sval = reinterpret_cast<cplx *>(smraw);
    cplx *XA = sval + nnz, *XB = XA + nu;
    int *scol = reinterpret_cast<int *>(XB + nu);
    int *srp = scol + nnz;
    cplx *yg = a.y + (int64_t)blockIdx.x * a.batch_stride;
    for (int t = threadIdx.x; t < nnz; t += BT) { sval[t] = val[t]; scol[t] = col[t]; }
    for (int t = threadIdx.x; t <= n; t += BT) srp[t] = (int)rowptr[t];
    cplx yv[UPT], acc[UPT];
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int e = threadIdx.x + i * BT;
        yv[i] = (e < nu) ? yg[e] : make_double2(0.0, 0.0);
        if (e < nu) XA[e] = yv[i];
    }
    __syncthreads();
    const double dt = a.dt;
    const double c_half = dt / 2, c_full = dt, w1 = dt / 6, w2 = dt / 6 * 2;
    for (int c = 0; c < a.nsteps; c++) {
        const cplx *f_prev = nullptr, *f_half = nullptr, *f_next = nullptr;
        if (FORCED) {
            f_prev = (c == 0) ? a.f_carry : a.slices + (int64_t)(2 * c - 1) * nu;
            f_half = a.slices + (int64_t)(2 * c) * nu;
            f_next = a.slices + (int64_t)(2 * c + 1) * nu;
        }
#pragma unroll
        for (int st = 1; st <= 4; st++) {
            const cplx *Xin = (st & 1) ? XA : XB;
            cplx *Xout = (st & 1) ? XB : XA;
            const cplx *fsl = (st == 1) ? f_prev : (st == 4 ? f_next : f_half);
            // the units of a thread are processed in groups of FG: the forcing values of a group are requested together, so
            // one L2 round trip is exposed per group instead of one per unit
            constexpr int FG = 5;
#pragma unroll
            for (int i0 = 0; i0 < UPT; i0 += FG) {
                cplx fv[FG];
#pragma unroll
                for (int j = 0; j < FG; j++) {
                    const int e = threadIdx.x + (i0 + j) * BT;
                    fv[j] = (FORCED && i0 + j < UPT && e < nu) ? ldg(fsl + e) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int j = 0; j < FG; j++) {
                    const int i = i0 + j;
                    const int e = threadIdx.x + i * BT;
                    if (i < UPT && e < nu) {
                        const int row = e / nv, iv = e - row * nv;
                        cplx s = fv[j];
                        const int p1 = srp[row + 1];
                        for (int p = srp[row]; p < p1; ++p) cfma(s, sval[p], Xin[scol[p] * nv + iv]);
                        if (st == 1) {
                            Xout[e] = make_double2(fma(c_half, s.x, yv[i].x), fma(c_half, s.y, yv[i].y));
                            acc[i] = make_double2(fma(w1, s.x, yv[i].x), fma(w1, s.y, yv[i].y));
                        } else if (st == 2) {
                            Xout[e] = make_double2(fma(c_half, s.x, yv[i].x), fma(c_half, s.y, yv[i].y));
                            acc[i] = make_double2(fma(w2, s.x, acc[i].x), fma(w2, s.y, acc[i].y));
                        } else if (st == 3) {
                            Xout[e] = make_double2(fma(c_full, s.x, yv[i].x), fma(c_full, s.y, yv[i].y));
                            acc[i] = make_double2(fma(w2, s.x, acc[i].x), fma(w2, s.y, acc[i].y));
                        } else {
                            yv[i] = make_double2(fma(w1, s.x, acc[i].x), fma(w1, s.y, acc[i].y));
                            Xout[e] = yv[i];   // XA: stage-1 input of the next step
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (a.snaps) {   // SaveSnapshots.c:19-28
            const int64_t i_step = a.i0 + c;
            if (i_step > a.Nt) {
                int64_t pos = i_step - a.Nt;
                if (pos % a.ResRatio == 0) {
                    pos = (pos / a.ResRatio - 1) % a.Nstore;
                    if (!a.direct) pos = a.Nstore - 1 - pos;
                    cplx *sn = a.snaps + pos * (int64_t)nu;
#pragma unroll
                    for (int i = 0; i < UPT; i++) {
                        const int e = threadIdx.x + i * BT;
                        if (e < nu) sn[e] = yv[i];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < UPT; i++) {
        const int e = threadIdx.x + i * BT;
        if (e < nu) yg[e] = yv[i];
    }
}

namespace {
constexpr int SMALL_BT = 512, SMALL_UPT = 10;
inline size_t small_smem(int64_t n, int64_t nnz, int nv)
{
    return (size_t)(nnz + 2 * n * nv) * sizeof(cplx) + (size_t)(nnz + n + 1) * sizeof(int) + 16;
}
}  // namespace

bool small_persistent_supported(const CsrDev &Op, int nv)
{
    if (Op.nrows != Op.ncols || Op.nrows <= 0) return false;
    if (Op.nrows * (int64_t)nv > (int64_t)SMALL_BT * SMALL_UPT) return false;
    return small_smem(Op.nrows, Op.nnz, nv) <= (size_t)200 * 1024;
}

cudaError_t launch_rk4_small_persistent(const CsrDev &Op, int nv, const SmallStepArgs &a, int nbatch, cudaStream_t s)
{
    const size_t smem = small_smem(Op.nrows, Op.nnz, nv);
    cudaError_t e;
    if (a.slices) {
        auto kfn = rk4_small_persistent_kernel<SMALL_BT, SMALL_UPT, true>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<(unsigned)nbatch, SMALL_BT, smem, s>>>(Op.rowptr, Op.col, Op.val, (int)Op.nrows, (int)Op.nnz, nv, a);
    } else {
        auto kfn = rk4_small_persistent_kernel<SMALL_BT, SMALL_UPT, false>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<(unsigned)nbatch, SMALL_BT, smem, s>>>(Op.rowptr, Op.col, Op.val, (int)Op.nrows, (int)Op.nnz, nv, a);
    }
    return cudaGetLastError();
}
