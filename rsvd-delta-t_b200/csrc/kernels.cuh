// kernels.cuh -- launchers of the sm_100a kernels behind the C ABI (include/rsvdt.h).
//
// Device data layout.  All values are complex double (double2, 16-B aligned).
//   block vector      X[row][iv]                 iv = 0..nv-1 fastest: one RK4 state for nv test vectors
//   frequency stack   T[a][row][iv]              a = frequency slot iw (Y_hat, F_hat) or snapshot slot pos (Y_all)
// so that (i) a gather of row `c` of a block vector is nv*16 contiguous bytes, (ii) every frequency block
// Y_hat_w is a contiguous row-major N x k matrix (QR / SVD operate on it in place) and (iii) the iDFT / DFT
// contractions read and write unit-stride along (row, iv).  The reference's column permutations
// (PermuteMat.c / ReversePermuteMat.c) are therefore index arithmetic, not data movement.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

struct CsrDev {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int64_t *rowptr = nullptr;  // nrows+1
    int32_t *col = nullptr;     // nnz
    cplx *val = nullptr;        // nnz
    int64_t max_row_nnz = 0;
};

// One RK4 stage  (TSRK4.c:18-47, TSTransRK4.c:15-36)
//   ksum = Op * xin + f
//   mode 0 (stage 1)  : xout = ybase + cx*ksum ; acc_out = ybase  + ca*ksum
//   mode 1 (stage 2,3): xout = ybase + cx*ksum ; acc_out = acc_in + ca*ksum
//   mode 2 (stage 4)  : acc_out = acc_in + ca*ksum (this is y_{n+1}; acc_out may alias ybase) ; snap = same (optional)
struct StageArgs {
    const cplx *xin = nullptr;
    const cplx *ybase = nullptr;
    const cplx *f = nullptr;
    const cplx *acc_in = nullptr;
    cplx *xout = nullptr;
    cplx *acc_out = nullptr;
    cplx *snap = nullptr;
    double cx = 0, ca = 0;
};

// rows [row_begin, row_begin + row_count) only (row_count < 0: to the end): the interior / boundary split that overlaps
// the halo exchange of a row-partitioned run with the interior rows
cudaError_t launch_rk4_stage(const CsrDev &Op, int nv, int mode, const StageArgs &a, int variant, cudaStream_t s,
                             int64_t row_begin = 0, int64_t row_count = -1);

// Block-CSR form (dense B x B blocks, e.g. the 5 flow variables of an LNS operator): browptr[nbrows+1], bcol[nblocks]
// (block column), bval[nblocks][BS] with the B*B entries column-major ([j][r]) and BS = B*B rounded up to even (32-byte aligned
// blocks for 256-bit loads).
struct BcsrDev {
    int B = 0;
    int64_t nbrows = 0, nblocks = 0;
    int64_t *browptr = nullptr;
    int32_t *bcol = nullptr;
    cplx *bval = nullptr;
};
bool bcsr_stage_supported(int B, int nv);
// variant 1 forces one test vector per thread (default: two when nv is even)
cudaError_t launch_rk4_stage_bcsr(const BcsrDev &Op, int nv, int mode, const StageArgs &a, int variant, cudaStream_t s,
                                  int64_t row_begin = 0, int64_t row_count = -1);

// Whole RK4 step in one kernel for narrow-band operators (all 4 stages, halo recomputation in shared memory).
//   f_prev / f_half / f_next : forcing slices at t_n, t_n + dt/2, t_n + dt (nullptr -> unforced)
struct BandedStepArgs {
    const cplx *y_in = nullptr;
    cplx *y_out = nullptr;
    const cplx *f_prev = nullptr, *f_half = nullptr, *f_next = nullptr;
    cplx *snap = nullptr;
    double dt = 0;
    // rows [own_begin, own_end) of the local arrays are advanced by this rank; rows outside are halo copies of the
    // neighbours' rows (row-partitioned runs).  own_end < 0 means "all rows".
    int64_t own_begin = 0, own_end = -1;
    // batch of independent block vectors advanced by one launch (blockIdx.y): vector b lives at y_in/y_out + b*batch_stride
    int nbatch = 1;
    int64_t batch_stride = 0;
    // row-partitioned runs (TMA-staged kernel, nbatch == 1): the first and the last tile of the rank -- the ones whose rows the
    // neighbours need -- are given to blocks 0 and 1, and count themselves in *bnd_flag once their rows are in global memory,
    // so that the halo exchange (a kernel on another stream polling the counter) runs under the interior tiles
    unsigned long long *bnd_flag = nullptr;
};
bool banded_step_supported(const CsrDev &Op, int64_t lo_bw, int64_t hi_bw, int nv, int variant);
int banded_tile_rows(int64_t lo_bw, int64_t hi_bw, int nv, int variant);
// ranges (device, optional): explicit list of nranges owned row ranges [r0, r1), each at most banded_tile_rows long
cudaError_t launch_rk4_banded_step(const CsrDev &Op, int64_t lo_bw, int64_t hi_bw, int nv, const BandedStepArgs &a,
                                   int variant, const int64_t *ranges, int nranges, cudaStream_t s);

// Same fused step on the diagonal (DIA) form of the operator: dia[row][lo+hi+1], offsets -lo..+hi; tile_ok[t] == 0
// marks tiles (of dia_tile_rows rows) containing rows that do not follow the pattern -- they are left untouched.
bool dia_step_supported(int lo, int hi, int nv);
int dia_tile_rows(int lo, int hi, int nv);
cudaError_t launch_rk4_dia_step(const cplx *dia, int lo, int hi, const unsigned char *tile_ok, int ntiles, int64_t N, int nv,
                                const BandedStepArgs &a, cudaStream_t s);

// TMA-staged form of the DIA step (bulk copies of whole tiles, 2-3 CTAs per SM); variant 0: 3 CTAs/SM, 1: 2 CTAs/SM
bool dia_tma_supported(int lo, int hi, int nv, int variant);
int dia_tma_tile_rows(int lo, int hi, int nv, int variant);
cudaError_t launch_rk4_dia_tma_step(const cplx *dia, int lo, int hi, const unsigned char *tile_ok, int ntiles, int64_t N,
                                    int nv, int variant, const BandedStepArgs &a, cudaStream_t s);

// Small problems: one persistent CTA advances `nsteps` RK4 steps of a whole block vector per launch (see kernels.cu, K1s).
//   y: state in/out (block vector b at y + b*batch_stride); forced runs: slices[2*nsteps] + f_carry as in forced_steps();
//   snapshots written by the kernel itself following SaveSnapshots.c:19-28 (i0 = 1-based index of the first step).
struct SmallStepArgs {
    cplx *y = nullptr;
    int64_t batch_stride = 0;
    int nsteps = 0;
    double dt = 0;
    const cplx *slices = nullptr, *f_carry = nullptr;
    cplx *snaps = nullptr;
    int64_t i0 = 0, Nt = 0, ResRatio = 1, Nstore = 1;
    int direct = 1;
};
bool small_persistent_supported(const CsrDev &Op, int nv);
cudaError_t launch_rk4_small_persistent(const CsrDev &Op, int nv, const SmallStepArgs &a, int nbatch, cudaStream_t s);

// Twiddle coefficient tables (CreateDFTiDFTMats.c:42-58,76-84), generated on the device with exactly reduced phases.
//   idft: coef[s][w] = (1/Ns) * exp(+2*pi*i * m(w) * j_s / (2Ns)),  j_s = (j_start + dir*s) mod 2Ns,  s < S, w < Nw_eff
//   dft : coef[w'][j] = ResRatio * exp(-2*pi*i * w' * j / Nw),       w' < Nw_eff, j < Nw
cudaError_t launch_idft_coef(cplx *coef, int S, int Nw_eff, int64_t j_start, int dir, int64_t Ns, int real_op,
                             cudaStream_t s);
cudaError_t launch_dft_coef(cplx *coef, int Nw_eff, int Nw, int64_t ResRatio, cudaStream_t s);

// out[s][u] = sum_{w<K} coef[s*K + w] * in[w*in_stride + u],  s < S, u < U   (complex double contraction)
//   forcing synthesis (CreateForcingOnFly.c:17-19 batched over a chunk of half steps) and DFT (DFT.c:26-39).
cudaError_t launch_small_cgemm(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride,
                               int S, int K, int64_t U, int variant, cudaStream_t s);

// Forcing synthesis through Taylor moments about the chunk centre (see kernels.cu, K2b).  J2 = 2*j_centre (integer),
// h = S/2.  coef: [P][K].  idft_taylor_terms(x) returns the number of terms for |m*phi*(j-j_c)| <= x, or -1.
int idft_taylor_terms(double x);
cudaError_t launch_idft_taylor_coef(cplx *coef, int P, int K, int64_t J2, int dir, int64_t Ns, double h, int real_op,
                                    cudaStream_t s);
// out_prev (optional): additionally receives the slice one half step before slice 0 (s = -1)
cudaError_t launch_idft_taylor(cplx *out, int64_t out_stride, const cplx *coef, const cplx *in, int64_t in_stride, int S,
                               int P, int K, int64_t U, cplx *out_prev, cudaStream_t s);

// reference (column-major, column c = a*ca + b*cb) <-> device stack T[a][row][b]
// The stack has `srows` rows per slot; the nrows host rows are its rows [roff, roff + nrows) (row-partitioned runs keep
// halo rows around the owned ones).
cudaError_t launch_host_to_stack(cplx *dst, const cplx *src_colmajor, int64_t nrows, int na, int nb, int64_t ca,
                                 int64_t cb, int64_t srows, int64_t roff, cudaStream_t s);
cudaError_t launch_stack_to_host(cplx *dst_colmajor, const cplx *src, int64_t nrows, int na, int nb, int64_t ca,
                                 int64_t cb, int64_t srows, int64_t roff, cudaStream_t s);

// ---- per-frequency small dense work (QRAllFreqs.c, SVDAllFreqs*.c) -----------------------------------
// G[w] = Y_w^H Y_w for all w (k x k each, row-major), deterministic two-stage reduction. partial: nblk*nw*k*k
int gram_num_blocks(int64_t nrows);
// Y_w starts at Y + w*slot_stride (complex elements); only the nrows rows given are summed (owned rows of a partition)
cudaError_t launch_gram(const cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, cplx *partial, cplx *G,
                        cudaStream_t s);
// eigen-decomposition of G_w (Hermitian Jacobi): sigma[w][m] = sqrt(lambda_m) descending, M[w] = V * diag(1/sigma)
cudaError_t launch_cross_svd_small(const cplx *G, int k, int nw, cplx *M, double *sigma, cudaStream_t s);
// Y_w <- Y_w * M_w in place (row tiles staged through shared memory)
cudaError_t launch_apply_right(cplx *Y, int64_t nrows, int k, int nw, const cplx *M, cudaStream_t s);

// The same Gram matrices on FP64 tensor cores (dense_k4.cu, mma.sync m8n8k4.f64), optionally of the transformed block
// Y_w * M_w formed on the fly (M != null; k <= 16): the second Gram of CholeskyQR2 without storing the first pass's Q.
// partial: gram_dmma_partial_doubles() doubles.
int gram_dmma_num_blocks(int64_t nrows);
size_t gram_dmma_partial_doubles(int64_t nrows, int k, int nw);
cudaError_t launch_gram_dmma(const cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial,
                             cplx *G, cudaStream_t s);
// Y_w <- Y_w * M_w in place on FP64 tensor cores (even k <= 16); G != null: also the Gram matrices of the rounded product
// (partial as for launch_gram_dmma) -- the apply of CholeskyQR2's first pass fused with the Gram of its second
cudaError_t launch_apply_dmma(cplx *Y, int64_t nrows, int k, int nw, int64_t slot_stride, const cplx *M, double *partial, cplx *G,
                              cudaStream_t s);
// C_w = A_w * B_w, k x k each
cudaError_t launch_small_matmul(const cplx *A, const cplx *B, int k, int nw, cplx *C, cudaStream_t s);

// ||q||_2^2 partial sums -> out (deterministic two-stage)
cudaError_t launch_norm2(const cplx *q, int64_t n, double *partial, double *out, cudaStream_t s);

// y <- 0 etc.
cudaError_t launch_fill_zero(cplx *p, int64_t n, cudaStream_t s);

// sparse (CSR, nrows x ncols) times stack: out[a][r][b] = sum_c M[r][c] in[a][c][b]  (ApplyWeightMats.c:18-76)
cudaError_t launch_spmm_stack(const CsrDev &M, const cplx *in, cplx *out, int na, int nb, cudaStream_t s);

// ---- halo exchange over peer memory (row-partitioned runs) ---------------------------------------------------------
// device-side descriptors (layout must match kernels.cu)
struct HaloMsgHost { cplx *peer_mailbox; const int32_t *rows; int count; int pad; };
struct HaloInHost { const cplx *mailbox; const unsigned long long *flag; int64_t dest_row0; int count; int pad; };
// push + signal + wait + unpack in one launch of one CTA (small messages)
cudaError_t launch_halo_exchange_small(cplx *buf, int nvec, const void *msgs_dev, int nout, void *peer_flags_dev, const void *ins_dev,
                                       int nin, int parity, int nvec_max, unsigned long long epoch, unsigned long long *timeout_flag,
                                       const unsigned long long *ready, unsigned long long ready_target, cudaStream_t s);
cudaError_t launch_halo_push(const cplx *src, int nvec, const void *msgs_dev, int nmsgs, int max_count, int parity,
                             int nvec_max, cudaStream_t s);
cudaError_t launch_halo_signal(void *peer_flags_dev, int npeers, unsigned long long epoch, cudaStream_t s);
cudaError_t launch_halo_wait_unpack(cplx *dst, int nvec, const void *ins_dev, int nin, int max_count, int parity,
                                    int nvec_max, unsigned long long epoch, cudaStream_t s);

// ---- transient removal (TransientRemovalStrategy.c): batched tall-skinny algebra on snapshot stacks ---------------------
// stack element (slot a, row, vector z) at base[a*slot_stride + row*nv + z]
int stack_gram2_chunks(int64_t nrows);
// G[z][a][b] = sum_{row0 <= row < row0+nrows} conj(X[a][row][z]) * Y[b][row][z]; partial: chunks*nv*ma*mb elements
cudaError_t launch_stack_gram2(const cplx *X, int64_t xs, int ma, const cplx *Y, int64_t ys, int mb, int nv, int64_t row0,
                               int64_t nrows, cplx *partial, cplx *G, cudaStream_t s);
// Out[b][row][z] = beta*Z[b][row][z] + alpha * sum_a X[a][row][z] * M[z][a][b]   (Z may be null)
cudaError_t launch_stack_rmul(cplx *Out, int64_t os, const cplx *Z, int64_t zs, double beta, const cplx *X, int64_t xs,
                              int ma, const cplx *M, int mb, double alpha, int nv, int64_t nrows, cudaStream_t s);
// G_z + shift_scale*trace(G_z)*I = R^H R; Minv_z = R^{-1}; G overwritten; status[z] != 0 on breakdown
// drop_dependent != 0: columns with a non-positive pivot are left out of the basis (status[z] = -(their number)) instead of
// being reported as a breakdown (status[z] = column + 1)
cudaError_t launch_chol_inverse_big(cplx *G, int m, int nz, double shift_scale, cplx *Minv, int *status, cudaStream_t s,
                                    int drop_dependent = 0);
// C_z[:, iw] = (alpha_iw I - Mt_z)^{-1} (alpha_iw B1_z[:, iw] - B2_z[:, iw]); scratch: nz*nw*m*(m+1) elements
cudaError_t launch_galerkin_solve(const cplx *Mt, const cplx *B1, const cplx *B2, const cplx *alpha, int m, int nw, int nz,
                                  cplx *scratch, cplx *C, int *status, cudaStream_t s);
