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

// =====================================================================================================
// K1: one RK4 stage = CSR SpMM over nv columns + forcing add + stage AXPYs   (TSRK4.c:18-47, TSTransRK4.c:15-36)
//   thread <-> unit u = row*nv + iv.  The nv threads of a row read the same A entry (one broadcast sector) and
//   nv*16 contiguous bytes of the gathered row of xin, so every global access is coalesced.
//   Bound: HBM.  Algorithmic bytes per unit and stage: nnz_row*20/nv + 8/nv (A) + 16*(reads+writes of block vectors).
// =====================================================================================================
template <int MODE, bool HAS_F>
__global__ void __launch_bounds__(256)
rk4_stage_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const cplx *__restrict__ val,
                     int nv, int64_t n_units, StageArgs a)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
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

cudaError_t launch_rk4_stage(const CsrDev &Op, int nv, int mode, const StageArgs &a, int variant, cudaStream_t s)
{
    (void)variant;
    const int64_t n_units = Op.nrows * (int64_t)nv;
    if (n_units == 0) return cudaSuccess;
    const int threads = 256;
    const int64_t blocks = (n_units + threads - 1) / threads;
    const bool hf = a.f != nullptr;
#define LAUNCH(M, F) rk4_stage_csr_kernel<M, F><<<(unsigned)blocks, threads, 0, s>>>(Op.rowptr, Op.col, Op.val, nv, n_units, a)
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
                       int64_t N, int nv, int b, int TR, int E0, BandedStepArgs a)
{
    extern __shared__ cplx sm[];
    const int n_e = E0 * nv;
    cplx *Y = sm, *B1 = sm + n_e, *B2 = sm + 2 * n_e;
    const int64_t base = (int64_t)blockIdx.x * TR - 4 * b;  // global row of local row 0 (may be negative)
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
        if (valid[i] && lr >= 4 * b && lr < E0 - 4 * b) {
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
cudaError_t banded_launch(const CsrDev &Op, int b, int nv, int E0, const BandedStepArgs &a, cudaStream_t s)
{
    const int TR = E0 - 8 * b;
    const size_t smem = (size_t)3 * E0 * nv * sizeof(cplx);
    const unsigned grid = (unsigned)((Op.nrows + TR - 1) / TR);
    cudaError_t e;
    if (a.f_prev) {
        auto kfn = rk4_banded_step_kernel<BT, UPT, MINB, true>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<grid, BT, smem, s>>>(Op.rowptr, Op.col, Op.val, Op.nrows, nv, b, TR, E0, a);
    } else {
        auto kfn = rk4_banded_step_kernel<BT, UPT, MINB, false>;
        e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<grid, BT, smem, s>>>(Op.rowptr, Op.col, Op.val, Op.nrows, nv, b, TR, E0, a);
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

cudaError_t launch_rk4_banded_step(const CsrDev &Op, int64_t lo_bw, int64_t hi_bw, int nv, const BandedStepArgs &a,
                                   int variant, cudaStream_t s)
{
    const int b = (int)(lo_bw > hi_bw ? lo_bw : hi_bw);
    const int E0 = banded_E0(variant, nv);
    if (E0 - 8 * b <= 0) return cudaErrorInvalidValue;
    switch (variant) {
        case 1: return banded_launch<512, 4, 1>(Op, b, nv, E0, a, s);
        case 2: return banded_launch<512, 3, 2>(Op, b, nv, E0, a, s);
        case 3: return banded_launch<256, 6, 3>(Op, b, nv, E0, a, s);
        default: return banded_launch<512, 4, 2>(Op, b, nv, E0, a, s);
    }
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

template <int PT, int UT>
__global__ void __launch_bounds__(128)
idft_taylor_kernel(cplx *__restrict__ out, int64_t out_stride, const cplx *__restrict__ coef,
                   const cplx *__restrict__ in, int64_t in_stride, int S, int P, int K, int64_t U)
{
    extern __shared__ cplx sm_coef[];  // [K][PT]
    for (int t = threadIdx.x; t < K * PT; t += blockDim.x) {
        const int w = t / PT, p = t - w * PT;
        sm_coef[t] = (p < P) ? coef[(int64_t)p * K + w] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int64_t ub = ((int64_t)blockIdx.x * blockDim.x) * UT + threadIdx.x;
    cplx acc[UT][PT];
#pragma unroll
    for (int i = 0; i < UT; i++)
#pragma unroll
        for (int j = 0; j < PT; j++) acc[i][j] = make_double2(0.0, 0.0);
    bool ok[UT];
#pragma unroll
    for (int i = 0; i < UT; i++) ok[i] = (ub + (int64_t)i * blockDim.x) < U;

#pragma unroll 2
    for (int w = 0; w < K; w++) {
        cplx v[UT];
#pragma unroll
        for (int i = 0; i < UT; i++)
            v[i] = ok[i] ? ldg(in + (int64_t)w * in_stride + ub + (int64_t)i * blockDim.x) : make_double2(0.0, 0.0);
        const cplx *cw = sm_coef + w * PT;
#pragma unroll
        for (int j = 0; j < PT; j++) {
            const cplx c = cw[j];
#pragma unroll
            for (int i = 0; i < UT; i++) cfma(acc[i][j], c, v[i]);
        }
    }
    const double inv_h = 2.0 / (double)S;
    for (int sidx = 0; sidx < S; sidx++) {
        const double tau = ((double)(2 * sidx - S + 1) * 0.5) * inv_h;   // (j_s - j_c) / (S/2)
#pragma unroll
        for (int i = 0; i < UT; i++) {
            cplx f = acc[i][PT - 1];
#pragma unroll
            for (int j = PT - 2; j >= 0; j--) {
                f.x = fma(f.x, tau, acc[i][j].x);
                f.y = fma(f.y, tau, acc[i][j].y);
            }
            if (ok[i]) out[(int64_t)sidx * out_stride + ub + (int64_t)i * blockDim.x] = f;
        }
    }
}

template <int PT, int UT>
static cudaError_t idft_taylor_launch(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride,
                                      int S, int P, int K, int64_t U, cudaStream_t s)
{
    const int threads = 128;
    const size_t smem = (size_t)K * PT * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(idft_taylor_kernel<PT, UT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((U + (int64_t)threads * UT - 1) / ((int64_t)threads * UT));
    idft_taylor_kernel<PT, UT><<<grid, threads, smem, s>>>(out, out_stride, coef, in, in_stride, S, P, K, U);
    return cudaGetLastError();
}

cudaError_t launch_idft_taylor(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride, int S,
                               int P, int K, int64_t U, cudaStream_t s)
{
    if (U == 0 || S == 0) return cudaSuccess;
    if (P <= 8) return idft_taylor_launch<8, 2>(out, out_stride, coef, in, in_stride, S, P, K, U, s);
    if (P <= 12) return idft_taylor_launch<12, 2>(out, out_stride, coef, in, in_stride, S, P, K, U, s);
    if (P <= 16) return idft_taylor_launch<16, 2>(out, out_stride, coef, in, in_stride, S, P, K, U, s);
    return cudaErrorInvalidValue;
}

// =====================================================================================================
// Layout conversion at the C-ABI boundary: reference column-major (column c = a*ca + b*cb) <-> stack T[a][row][b]
// =====================================================================================================
template <bool TO_STACK>
__global__ void relayout_kernel(cplx *__restrict__ dst, const cplx *__restrict__ src, int64_t nrows, int na, int nb,
                                int64_t ca, int64_t cb)
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
        cplx *d = dst + ((int64_t)a * nrows + r0) * nb + b0;
        for (int t = threadIdx.x; t < nr * nbt; t += blockDim.x) {
            const int r = t / nbt, b = t - r * nbt;
            d[(int64_t)r * nb + b] = tile[r][b];
        }
    } else {
        const cplx *sp = src + ((int64_t)a * nrows + r0) * nb + b0;
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
                                 cudaStream_t s)
{
    if (nrows == 0 || na == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)((nrows + 31) / 32), (unsigned)na, (unsigned)((nb + 31) / 32));
    relayout_kernel<true><<<grid, 256, 0, s>>>(dst, src, nrows, na, nb, ca, cb);
    return cudaGetLastError();
}

cudaError_t launch_stack_to_host(cplx *dst, const cplx *src, int64_t nrows, int na, int nb, int64_t ca, int64_t cb,
                                 cudaStream_t s)
{
    if (nrows == 0 || na == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)((nrows + 31) / 32), (unsigned)na, (unsigned)((nb + 31) / 32));
    relayout_kernel<false><<<grid, 256, 0, s>>>(dst, src, nrows, na, nb, ca, cb);
    return cudaGetLastError();
}

// =====================================================================================================
// Per-frequency Gram matrices G_w = Y_w^H Y_w (k x k), deterministic: per-block partials then a fixed-order sum.
// =====================================================================================================
static const int GRAM_ROWS_PER_BLOCK = 2048;
int gram_num_blocks(int64_t nrows) { return (int)((nrows + GRAM_ROWS_PER_BLOCK - 1) / GRAM_ROWS_PER_BLOCK); }

__global__ void __launch_bounds__(128)
gram_partial_kernel(const cplx *__restrict__ Y, int64_t nrows, int k, cplx *__restrict__ partial)
{
    extern __shared__ cplx tile[];  // [64][k]
    const int w = blockIdx.y, nw = gridDim.y;
    const int64_t r_begin = (int64_t)blockIdx.x * GRAM_ROWS_PER_BLOCK;
    const int64_t r_end = min(nrows, r_begin + GRAM_ROWS_PER_BLOCK);
    const cplx *Yw = Y + (int64_t)w * nrows * k;
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

cudaError_t launch_gram(const cplx *Y, int64_t nrows, int k, int nw, cplx *partial, cplx *G, cudaStream_t s)
{
    if (k > 32) return cudaErrorInvalidValue;
    const int nblk = gram_num_blocks(nrows);
    dim3 grid((unsigned)nblk, (unsigned)nw);
    gram_partial_kernel<<<grid, 128, (size_t)64 * k * sizeof(cplx), s>>>(Y, nrows, k, partial);
    const int64_t n_out = (int64_t)nw * k * k;
    gram_reduce_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(partial, nblk, n_out, G);
    return cudaGetLastError();
}

// Cholesky G = R^H R (R upper, positive real diagonal) and M = R^{-1}; one thread per frequency (k <= 32).
__global__ void chol_inverse_kernel(const cplx *__restrict__ G, int k, int nw, cplx *__restrict__ M, int *status)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    const cplx *g = G + (int64_t)w * k * k;
    cplx *m = M + (int64_t)w * k * k;
    cplx R[32 * 32];  // local memory; row-major upper triangular
    int bad = 0;
    for (int j = 0; j < k; j++) {
        // diagonal
        double d = g[j * k + j].x;
        for (int p = 0; p < j; p++) d -= R[p * k + j].x * R[p * k + j].x + R[p * k + j].y * R[p * k + j].y;
        if (!(d > 0.0)) { bad = j + 1; d = 1.0; }
        const double rjj = sqrt(d);
        R[j * k + j] = make_double2(rjj, 0.0);
        for (int c = j + 1; c < k; c++) {
            cplx sacc = g[j * k + c];  // G[j][c]
            for (int p = 0; p < j; p++) {  // s -= conj(R[p][j]) * R[p][c]
                const cplx a = R[p * k + j], b = R[p * k + c];
                sacc.x -= a.x * b.x + a.y * b.y;
                sacc.y -= a.x * b.y - a.y * b.x;
            }
            R[j * k + c] = make_double2(sacc.x / rjj, sacc.y / rjj);
        }
    }
    // M = R^{-1}, upper triangular, by back substitution column by column
    for (int c = 0; c < k; c++) {
        for (int r = k - 1; r >= 0; r--) {
            if (r > c) { m[r * k + c] = make_double2(0.0, 0.0); continue; }
            cplx sacc = make_double2(r == c ? 1.0 : 0.0, 0.0);
            for (int p = r + 1; p <= c; p++) {  // s -= R[r][p] * M[p][c]
                const cplx a = R[r * k + p], b = m[p * k + c];
                sacc.x -= a.x * b.x - a.y * b.y;
                sacc.y -= a.x * b.y + a.y * b.x;
            }
            const double rr = R[r * k + r].x;
            m[r * k + c] = make_double2(sacc.x / rr, sacc.y / rr);
        }
    }
    status[w] = bad;
}

cudaError_t launch_chol_inverse(const cplx *G, int k, int nw, cplx *M, int *status, cudaStream_t s)
{
    if (k > 32) return cudaErrorInvalidValue;
    chol_inverse_kernel<<<(nw + 31) / 32, 32, 0, s>>>(G, k, nw, M, status);
    return cudaGetLastError();
}

// Cross-product SVD of a tall block through its Gram matrix (SLEPc's default SVD type, SVDAllFreqs.c:36-40):
// cyclic complex Jacobi on G_w; sigma = sqrt(lambda) sorted descending; M = V diag(1/sigma) so that U = Y M.
__global__ void cross_svd_small_kernel(const cplx *__restrict__ Gin, int k, int nw, cplx *__restrict__ M,
                                       double *__restrict__ sigma)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    cplx G[32 * 32], V[32 * 32];  // row-major
    for (int t = 0; t < k * k; t++) { G[t] = Gin[(int64_t)w * k * k + t]; V[t] = make_double2(0.0, 0.0); }
    for (int i = 0; i < k; i++) V[i * k + i] = make_double2(1.0, 0.0);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dg = 0.0;
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
                if (apq == 0.0) continue;
                const double app = G[p * k + p].x, aqq = G[q * k + q].x;
                const cplx ph = make_double2(gpq.x / apq, gpq.y / apq);  // e^{i phi}
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
                const cplx sph = make_double2(sn * ph.x, sn * ph.y);    // s e^{i phi}
                const cplx sphc = make_double2(sn * ph.x, -sn * ph.y);  // s e^{-i phi}
                for (int r = 0; r < k; r++) {  // G <- G J : col p = c*gp - s e^{-i phi} gq ; col q = s e^{i phi} gp + c gq
                    const cplx gp = G[r * k + p], gq = G[r * k + q];
                    const cplx a = cmul(sphc, gq), b = cmul(sph, gp);
                    G[r * k + p] = make_double2(c * gp.x - a.x, c * gp.y - a.y);
                    G[r * k + q] = make_double2(b.x + c * gq.x, b.y + c * gq.y);
                }
                for (int cc = 0; cc < k; cc++) {  // G <- J^H G : row p = c*gp - s e^{i phi} gq ; row q = s e^{-i phi} gp + c gq
                    const cplx gp = G[p * k + cc], gq = G[q * k + cc];
                    const cplx a = cmul(sph, gq), b = cmul(sphc, gp);
                    G[p * k + cc] = make_double2(c * gp.x - a.x, c * gp.y - a.y);
                    G[q * k + cc] = make_double2(b.x + c * gq.x, b.y + c * gq.y);
                }
                for (int r = 0; r < k; r++) {  // V <- V J
                    const cplx vp = V[r * k + p], vq = V[r * k + q];
                    const cplx a = cmul(sphc, vq), b = cmul(sph, vp);
                    V[r * k + p] = make_double2(c * vp.x - a.x, c * vp.y - a.y);
                    V[r * k + q] = make_double2(b.x + c * vq.x, b.y + c * vq.y);
                }
            }
    }
    // sort descending (selection sort on indices)
    int ord[32];
    for (int i = 0; i < k; i++) ord[i] = i;
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++)
            if (G[ord[j] * k + ord[j]].x > G[ord[i] * k + ord[i]].x) { int tt = ord[i]; ord[i] = ord[j]; ord[j] = tt; }
    for (int mI = 0; mI < k; mI++) {
        const double lam = G[ord[mI] * k + ord[mI]].x;
        const double sg = sqrt(lam > 0.0 ? lam : 0.0);
        sigma[(int64_t)w * k + mI] = sg;
        const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
        for (int r = 0; r < k; r++) {
            const cplx v = V[r * k + ord[mI]];
            M[(int64_t)w * k * k + r * k + mI] = make_double2(v.x * inv, v.y * inv);
        }
    }
}

cudaError_t launch_cross_svd_small(const cplx *G, int k, int nw, cplx *M, double *sigma, cudaStream_t s)
{
    if (k > 32) return cudaErrorInvalidValue;
    cross_svd_small_kernel<<<(nw + 31) / 32, 32, 0, s>>>(G, k, nw, M, sigma);
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
