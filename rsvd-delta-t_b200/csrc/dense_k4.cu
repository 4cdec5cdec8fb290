// dense_k4.cu -- K4: the Gram matrices of QRAllFreqs.c:22-32 / SVDAllFreqs*.c on FP64 tensor cores (DMMA), optionally of
// the TRANSFORMED block Y_w * M_w formed on the fly -- the second pass of CholeskyQR2 without writing the first pass's Q.
//
//   G_w = Y_w^H Y_w,  Y_w: nrows x k complex, row-major (one frequency slot of the stack T[w][row][ik]).
//
// Real formulation: Z = [Re Y | Im Y] (nrows x 2KP real, KP = k rounded up to 8), S = Z^T Z, and
//   Re G[i][j] = S[i][j] + S[KP+i][KP+j],   Im G[i][j] = S[i][KP+j] - S[KP+i][j].
// S is accumulated by mma.sync.m8n8k4.f64: four rows of Z per instruction.  For both operands lane l of a warp needs
// Z[row0 + l%4][8t + l/4] -- the Re/Im parts of ONE complex entry Y[row0 + l%4][8t' + l/4] -- so a warp reads its 4 rows with
// one perfectly coalesced 128-bit load per lane and 8-column tile, straight from global memory: no shared memory, no
// transposition, 2*(2KT)^2 accumulator registers (KT = KP/8).  With TRANSFORM the lane first forms its entry of Y*M from the
// row of Y (k loads that the 8 lanes of a row share, k complex FMAs on the FP64 pipes) and feeds that to the tensor cores.
//
// Bound: HBM (16 B per entry read once).  Per 4 rows and KT = 1: 512 B, 4 DMMA (+ 256 DFMA with TRANSFORM).
// The FMA formulation this replaces (gram_partial_kernel, kernels.cu) needs 2 shared-memory reads per complex FMA and
// leaves half of its threads idle at k = 8; profiles/r2_dense_phases_ncu.txt holds the ncu comparison north_star asks for.
#include "kernels.cuh"

#include <algorithm>

namespace {

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int DG_WARPS = 8;             // warps per CTA
constexpr int DG_ROWS_PER_CTA = 8192;   // rows one CTA sums (its warps interleave groups of 4 rows)

// partial[(blk*nw + w)][2KT][2KT][64]  (tile-major, then the 2 x 32 accumulator registers of a warp, summed over the CTA's
// warps in warp order: deterministic)
template <int KT, bool TRANSFORM>
__global__ void __launch_bounds__(DG_WARPS * 32)
gram_dmma_kernel(const cplx *__restrict__ Y, int64_t nrows, int k, int64_t slot_stride, const cplx *__restrict__ M,
                 double *__restrict__ partial)
{
    constexpr int NT = 2 * KT;
    const int w = blockIdx.y, nw = gridDim.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = lane & 3, ml = lane >> 2;
    const cplx *Yw = Y + (int64_t)w * slot_stride;
    const int64_t r_begin = (int64_t)blockIdx.x * DG_ROWS_PER_CTA;
    const int64_t r_end = min(nrows, r_begin + DG_ROWS_PER_CTA);
    double acc[NT][NT][2];
#pragma unroll
    for (int a = 0; a < NT; a++)
#pragma unroll
        for (int b = 0; b < NT; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    // TRANSFORM: column (8t + ml) of M_w in registers, one column per m-tile
    cplx mcol[TRANSFORM ? KT : 1][TRANSFORM ? 8 * KT : 1];
    if (TRANSFORM) {
        const cplx *Mw = M + (int64_t)w * k * k;
#pragma unroll
        for (int t = 0; t < KT; t++)
#pragma unroll
            for (int i = 0; i < 8 * KT; i++) {
                const int c = 8 * t + ml;
                mcol[t][i] = (i < k && c < k) ? Mw[i * k + c] : make_double2(0.0, 0.0);
            }
    }
    for (int64_t r0 = r_begin + 4 * warp; r0 < r_end; r0 += 4 * DG_WARPS) {
        const int64_t row = r0 + rl;
        const bool on = row < r_end;
        double frag[NT];
        if (!TRANSFORM) {
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int c = 8 * t + ml;
                cplx v = make_double2(0.0, 0.0);
                if (on && c < k) v = __ldg(Yw + row * k + c);
                frag[t] = v.x;
                frag[KT + t] = v.y;
            }
        } else {
            cplx s[KT];
#pragma unroll
            for (int t = 0; t < KT; t++) s[t] = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 8 * KT; i++) {
                cplx y = make_double2(0.0, 0.0);
                if (on && i < k) y = __ldg(Yw + row * k + i);
#pragma unroll
                for (int t = 0; t < KT; t++) {   // s += y * M[i][c]
                    s[t].x = fma(y.x, mcol[t][i].x, s[t].x);
                    s[t].x = fma(-y.y, mcol[t][i].y, s[t].x);
                    s[t].y = fma(y.x, mcol[t][i].y, s[t].y);
                    s[t].y = fma(y.y, mcol[t][i].x, s[t].y);
                }
            }
#pragma unroll
            for (int t = 0; t < KT; t++) { frag[t] = s[t].x; frag[KT + t] = s[t].y; }
        }
#pragma unroll
        for (int a = 0; a < NT; a++)
#pragma unroll
            for (int b = 0; b < NT; b++) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], frag[a], frag[b]);
    }
    __shared__ double red[NT * NT * 64];
    for (int wp = 0; wp < DG_WARPS; wp++) {
        if (warp == wp) {
#pragma unroll
            for (int a = 0; a < NT; a++)
#pragma unroll
                for (int b = 0; b < NT; b++) {
                    double *q = red + (a * NT + b) * 64;
                    q[lane] = (wp ? q[lane] : 0.0) + acc[a][b][0];
                    q[32 + lane] = (wp ? q[32 + lane] : 0.0) + acc[a][b][1];
                }
        }
        __syncthreads();
    }
    double *out = partial + (((int64_t)blockIdx.x * nw + w) * NT * NT) * 64;
    for (int t = threadIdx.x; t < NT * NT * 64; t += blockDim.x) out[t] = red[t];
}

// fixed-order sum of the warp partials and S -> G.  One thread per (w, i, j).
template <int KT>
__global__ void gram_dmma_reduce_kernel(const double *__restrict__ partial, int nblk, int nw, int k, cplx *__restrict__ G)
{
    constexpr int NT = 2 * KT, KP = 8 * KT;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nw * k * k) return;
    const int w = (int)(t / (k * k)), i = (int)((t / k) % k), j = (int)(t % k);
    // element (r, c) of S lives in tile (r/8, c/8), accumulator row r%8 = lane/4, column c%8 = 2*(lane%4) + reg
    auto S = [&](int r, int c) {
        const int a = r >> 3, b = c >> 3, lane = (r & 7) * 4 + ((c & 7) >> 1), reg = c & 1;
        double s = 0.0;
        for (int blk = 0; blk < nblk; blk++) s += partial[((((int64_t)blk * nw + w) * NT * NT) + (a * NT + b)) * 64 + reg * 32 + lane];
        return s;
    };
    G[t] = make_double2(S(i, j) + S(KP + i, KP + j), S(i, KP + j) - S(KP + i, j));
}

// Y_w <- Y_w * M_w in place on FP64 tensor cores, and (GRAM) the Gram matrix of the ROUNDED product in the same pass --
// pass 1 of CholeskyQR2 fused with the Gram of pass 2: 5 instead of 6 passes over Y_hat, and no FP64-pipe work at all.
//   Real form: Z' = Z * Mreal, Z = [Re Y | Im Y] (rows x 2KP), Mreal = [[Mr, Mi], [-Mi, Mr]] (2KP x 2KP).
//   mma m8n8k4: 8 rows per instruction as the M dimension.  A fragment of k-step s: Z[row0 + lane/4][4s + lane%4], i.e. the
//   Re (s < KP/4) or Im part of Y[row][4j + lane%4] -- KP/4 128-bit loads per lane, the 4 lanes of a row reading 64
//   contiguous bytes.  B fragment: Mreal[4s + lane%4][8t + lane/4], kept in registers.  D tile t: Z'[row0 + lane/4][8t +
//   2(lane%4) + {0,1}] -> the lane owns Re and Im of two neighbouring entries of Y': one 256-bit store, 128 contiguous
//   bytes per row.  For the Gram the D fragments are re-laid out by warp shuffles into the operand layout of
//   gram_dmma_kernel (4 rows per k-step) and fed to the same (2KT)^2 accumulating MMAs.
template <int KT, bool GRAM>
__global__ void __launch_bounds__(DG_WARPS * 32)
apply_dmma_kernel(cplx *__restrict__ Y, int64_t nrows, int k, int64_t slot_stride, const cplx *__restrict__ M,
                  double *__restrict__ partial)
{
    constexpr int NT = 2 * KT, KP = 8 * KT, NS = KP / 2;   // NS k-steps of 4 real columns
    const int w = blockIdx.y, nw = gridDim.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, g = lane >> 2;
    cplx *Yw = Y + (int64_t)w * slot_stride;
    const cplx *Mw = M + (int64_t)w * k * k;
    const int64_t r_begin = (int64_t)blockIdx.x * DG_ROWS_PER_CTA;
    const int64_t r_end = min(nrows, r_begin + DG_ROWS_PER_CTA);
    // B fragments: Mreal[4s + q][8t + g]
    double bf[NS][NT];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int t = 0; t < NT; t++) {
            const int r = 4 * s + q, c = 8 * t + g;           // real row / column of Mreal
            const int ri = r % KP, ci = c % KP;               // complex indices
            double v = 0.0;
            if (ri < k && ci < k) {
                const cplx m = Mw[ri * k + ci];
                v = (r < KP) ? (c < KP ? m.x : m.y) : (c < KP ? -m.y : m.x);
            }
            bf[s][t] = v;
        }
    double acc[GRAM ? NT : 1][GRAM ? NT : 1][2];
    if (GRAM) {
#pragma unroll
        for (int a = 0; a < NT; a++)
#pragma unroll
            for (int b = 0; b < NT; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    }
    for (int64_t r0 = r_begin + 8 * warp; r0 < r_end; r0 += 8 * DG_WARPS) {
        const int64_t row = r0 + g;
        const bool on = row < r_end;
        double d[NT][2];
#pragma unroll
        for (int t = 0; t < NT; t++) d[t][0] = d[t][1] = 0.0;
#pragma unroll
        for (int j = 0; j < KP / 4; j++) {
            const int c = 4 * j + q;
            cplx y = make_double2(0.0, 0.0);
            if (on && c < k) y = Yw[row * k + c];
#pragma unroll
            for (int t = 0; t < NT; t++) {
                dmma_m8n8k4(d[t][0], d[t][1], y.x, bf[j][t]);            // k-step j: real parts
                dmma_m8n8k4(d[t][0], d[t][1], y.y, bf[KP / 4 + j][t]);   // k-step KP/4 + j: imaginary parts
            }
        }
#pragma unroll
        for (int t = 0; t < KT; t++) {
            const int c = 8 * t + 2 * q;
            if (on && c + 1 < k) {
                double4 o = make_double4(d[t][0], d[KT + t][0], d[t][1], d[KT + t][1]);
                *reinterpret_cast<double4 *>(Yw + row * k + c) = o;
            } else if (on && c < k) {
                Yw[row * k + c] = make_double2(d[t][0], d[KT + t][0]);
            }
        }
        if (GRAM) {
            // rows outside the range contribute zeros (their Y was read as zero, so d = 0)
#pragma unroll
            for (int u = 0; u < 2; u++) {   // the 8 rows as two k-steps of 4
                double frag[NT];
                const int src = (4 * u + q) * 4 + (g >> 1);   // lane that owns Z'[row0 + 4u + q][8t + g] ...
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const double v0 = __shfl_sync(0xffffffffu, d[t][0], src);
                    const double v1 = __shfl_sync(0xffffffffu, d[t][1], src);
                    frag[t] = (g & 1) ? v1 : v0;              // ... in register g % 2
                }
#pragma unroll
                for (int a = 0; a < NT; a++)
#pragma unroll
                    for (int b = 0; b < NT; b++) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], frag[a], frag[b]);
            }
        }
    }
    if (GRAM) {
        __shared__ double red[NT * NT * 64];
        for (int wp = 0; wp < DG_WARPS; wp++) {
            if (warp == wp) {
#pragma unroll
                for (int a = 0; a < NT; a++)
#pragma unroll
                    for (int b = 0; b < NT; b++) {
                        double *qq = red + (a * NT + b) * 64;
                        qq[lane] = (wp ? qq[lane] : 0.0) + acc[a][b][0];
                        qq[32 + lane] = (wp ? qq[32 + lane] : 0.0) + acc[a][b][1];
                    }
            }
            __syncthreads();
        }
        double *out = partial + (((int64_t)blockIdx.x * nw + w) * NT * NT) * 64;
        for (int t = threadIdx.x; t < NT * NT * 64; t += blockDim.x) out[t] = red[t];
    }
}

template <int KT>
cudaError_t go_apply(cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial, cplx *G, cudaStream_t s)
{
    const int nblk = gram_dmma_num_blocks(nrows);
    dim3 grid((unsigned)nblk, (unsigned)nw);
    if (G) {
        apply_dmma_kernel<KT, true><<<grid, DG_WARPS * 32, 0, s>>>(Y, nrows, k, slot_stride, M, partial);
        const int64_t n_out = (int64_t)nw * k * k;
        gram_dmma_reduce_kernel<KT><<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(partial, nblk, nw, k, G);
    } else {
        apply_dmma_kernel<KT, false><<<grid, DG_WARPS * 32, 0, s>>>(Y, nrows, k, slot_stride, M, nullptr);
    }
    return cudaGetLastError();
}

template <int KT>
cudaError_t go(const cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial, cplx *G, cudaStream_t s)
{
    const int nblk = gram_dmma_num_blocks(nrows);
    dim3 grid((unsigned)nblk, (unsigned)nw);
    if (M) gram_dmma_kernel<KT, true><<<grid, DG_WARPS * 32, 0, s>>>(Y, nrows, k, slot_stride, M, partial);
    else gram_dmma_kernel<KT, false><<<grid, DG_WARPS * 32, 0, s>>>(Y, nrows, k, slot_stride, nullptr, partial);
    const int64_t n_out = (int64_t)nw * k * k;
    gram_dmma_reduce_kernel<KT><<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(partial, nblk, nw, k, G);
    return cudaGetLastError();
}

// M3_w = M1_w * M2_w (k x k each): the two triangular factors of CholeskyQR2 folded into one right factor
__global__ void small_matmul_kernel(const cplx *__restrict__ A, const cplx *__restrict__ B, int k, int nw, cplx *__restrict__ C)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nw * k * k) return;
    const int w = (int)(t / (k * k)), i = (int)((t / k) % k), j = (int)(t % k);
    const cplx *a = A + (int64_t)w * k * k, *b = B + (int64_t)w * k * k;
    cplx s = make_double2(0.0, 0.0);
    for (int p = 0; p < k; p++) {
        s.x = fma(a[i * k + p].x, b[p * k + j].x, s.x);
        s.x = fma(-a[i * k + p].y, b[p * k + j].y, s.x);
        s.y = fma(a[i * k + p].x, b[p * k + j].y, s.y);
        s.y = fma(a[i * k + p].y, b[p * k + j].x, s.y);
    }
    C[t] = s;
}

}  // namespace

int gram_dmma_num_blocks(int64_t nrows) { return (int)std::max<int64_t>(1, (nrows + DG_ROWS_PER_CTA - 1) / DG_ROWS_PER_CTA); }
size_t gram_dmma_partial_doubles(int64_t nrows, int k, int nw)
{
    const int NT = 2 * ((k + 7) / 8);
    return (size_t)gram_dmma_num_blocks(nrows) * nw * NT * NT * 64;
}

cudaError_t launch_gram_dmma(const cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial,
                             cplx *G, cudaStream_t s)
{
    switch ((k + 7) / 8) {
        case 1: return go<1>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        case 2: return go<2>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        case 3: return M ? cudaErrorInvalidValue : go<3>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        case 4: return M ? cudaErrorInvalidValue : go<4>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_apply_dmma(cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial, cplx *G,
                              cudaStream_t s)
{
    if (nrows == 0) return cudaSuccess;
    if ((k * (int64_t)sizeof(cplx)) % 32) {   // the 256-bit stores need rows that start 32-byte aligned: even k
        return cudaErrorInvalidValue;
    }
    switch ((k + 7) / 8) {
        case 1: return go_apply<1>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        case 2: return go_apply<2>(Y, nrows, k, nw, slot_stride, M, partial, G, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_small_matmul(const cplx *A, const cplx *B, int k, int nw, cplx *C, cudaStream_t s)
{
    const int64_t n = (int64_t)nw * k * k;
    small_matmul_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(A, B, k, nw, C);
    return cudaGetLastError();
}
