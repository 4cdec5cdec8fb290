/*
 * rsvdt_oracle.h -- CPU oracle for the RSVD-Delta-t time-stepping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (rsvd-delta-t_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (AliFarghadan/RSVD-Delta-t) ships no tests, golden vectors
 * or known-answer fixtures for this path, and it cannot be built here (PETSc 3.19.4 / SLEPc 3.19 /
 * MPI are absent).  The oracle is a restatement of the reference's operation sequence with the
 * PETSc/SLEPc calls replaced by their published mathematical definitions (SpMV, AXPY, dense
 * mat-vec, GEMM, Gram-Schmidt thin QR with diag(R)>0, cross-product SVD, dense LU solve).
 * It is pinned only by (i) the analytic resolvent identity of SURVEY.md A.6, (ii) the exact
 * resolvent gains of the shipped tutorial operator A_GL and (iii)
 * uniqueness properties (thin QR, SVD).  See tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line (relative to /root/reference/SourceCode/) it follows.
 * All dense matrices are column-major complex double, exactly like a one-rank PETSc MATDENSE.
 */
#ifndef RSVDT_ORACLE_H
#define RSVDT_ORACLE_H

#include <complex.h>
#include <stdint.h>

typedef double complex orc_cplx;
typedef int64_t orc_int; /* PetscInt of the recommended --with-64-bit-indices build (README.md:123-125) */

/* CSR operator, 64-bit indices, complex values: a one-rank MATMPIAIJ (LNSStructure.c:24-26). */
typedef struct {
    orc_int nrows, ncols;
    const orc_int *rowptr; /* nrows+1 */
    const orc_int *col;    /* nnz */
    const orc_cplx *val;   /* nnz */
} orc_csr;

/* Variables.h:5-28 (RSVD_vars + TS_vars), only the members the hot path reads. */
typedef struct {
    orc_int N, k, Nw, Nw_eff, Ns, Nt, ResRatio;
    double dt, w;
    int RealOperator, TransientRemoval, DirAdj;
} orc_vars;

/* SetupTimeFreqGrid.c:16-21 */
int orc_setup_time_freq_grid(double w, orc_int Nw, double dt_in, double TransientLength,
                             double *dt, orc_int *Ns, orc_int *Nt, orc_int *ResRatio);

/* CreateDFTiDFTMats.c:42-58,76-84,93.  dft: Nw x Nw_eff, idft: Nw_eff x 2Ns (after the in-place transpose).
 * faithful=1 evaluates exp() of the un-reduced argument exactly as the reference does (error ~1e-11);
 * faithful=0 reduces the integer phase first (error ~1e-16). */
int orc_create_dft_idft_mats(orc_int Nw, orc_int Nw_eff, orc_int Ns, int RealOperator, int faithful,
                             orc_cplx *dft, orc_cplx *idft);

/* conj-transpose of a CSR matrix (so that MatMultHermitianTranspose becomes a gather SpMV). */
int orc_csr_hermitian_transpose(const orc_csr *A, orc_int *t_rowptr, orc_int *t_col, orc_cplx *t_val);

/* y = A x [+ f]  (MatMult / MatMultAdd, TSRK4.c:18,31,33,35; TSTransRK4.c:16-22) */
void orc_mat_mult_add(const orc_csr *A, const orc_cplx *x, const orc_cplx *f, orc_cplx *y);

/* CreateForcingOnFly.c:17-19 : F = F_hat_k (N x Nw_eff) * idft[:, jt_cyc] */
void orc_create_forcing_on_fly(orc_int N, orc_int Nw_eff, const orc_cplx *F_hat_k, const orc_cplx *idft,
                               orc_int jt_cyc, orc_cplx *F);

/* TSRK4.c:18-47 one forced RK4 step.  Op must be A (DirAdj=1) or A^H as CSR (DirAdj=0).
 * work: 6*N complex (F2,F3,k1..k4 share: layout F2,F3,k1,k2,k3,k4,y_temp -> 7*N). */
void orc_tsrk4(const orc_csr *Op, const orc_vars *v, const orc_cplx *F_hat_k, const orc_cplx *idft,
               orc_int i, orc_cplx *y, orc_cplx *work7N);

/* TSTransRK4.c:15-36 one unforced RK4 step; work: 5*N complex (k1..k4,y_temp). */
void orc_tstrans_rk4(const orc_csr *Op, double dt, orc_cplx *y, orc_cplx *work5N);

/* TSActionRK4.c:38-148 + DFT.c:26-39 (+ TransientRemovalStrategy.c:30-40 when v->TransientRemoval).
 * Y_hat_in: N x (k*Nw_eff) frequency-major (col = iw*k+ik); Y_hat_out: same shape.
 * Op = A (direct) or A^H (adjoint).  snapshots_out (optional, may be NULL): N x (Nstore*k) test-vector-major.
 * max_steps<=0: full rend; otherwise truncate the loop (timing windows only). */
int orc_ts_action_rk4(const orc_csr *Op, const orc_vars *v, const orc_cplx *dft, const orc_cplx *idft,
                      const orc_cplx *Y_hat_in, orc_cplx *Y_hat_out, orc_cplx *snapshots_out,
                      orc_int max_steps);

/* PermuteMat.c:24-38 / ReversePermuteMat.c:24-38 (out of place) */
void orc_permute_mat(orc_int N, orc_int k, orc_int Nw, const orc_cplx *F, orc_cplx *F_permuted);
void orc_reverse_permute_mat(orc_int N, orc_int k, orc_int Nw, const orc_cplx *F, orc_cplx *F_permuted);

/* DFT.c:26-39 : Y_all N x (Nw*k) test-vector-major -> Y_hat N x (k*Nw_eff) frequency-major */
void orc_dft(orc_int N, orc_int k, orc_int Nw, orc_int Nw_eff, orc_int ResRatio, const orc_cplx *dft,
             const orc_cplx *Y_all, orc_cplx *Y_hat);

/* QRDecomposition.c:16-20 : in-place thin QR (Gram-Schmidt, R discarded, diag(R)>0) of an N x m block */
int orc_qr_decomposition(orc_int N, orc_int m, orc_cplx *V);

/* QRAllFreqs.c:22-32 */
int orc_qr_all_freqs(orc_int N, orc_int k, orc_int Nw_eff, orc_cplx *Y_hat);

/* TSExpADeltaT.c:25-34 : ResRatio unforced steps on every column of U (N x ncols) */
void orc_ts_exp_a_deltat(const orc_csr *Op, const orc_vars *v, orc_int ncols, orc_cplx *U);

/* TransientRemovalStrategy.c:30-40 (CreateMtilde.c:41-54, GalerkinProjAllModes.c:28-39, GalerkinProj.c:26-74)
 * Y_all: N x ((Nw+1)*k) test-vector-major; Y_hat: N x (k*Nw_eff) frequency-major */
int orc_transient_removal_strategy(const orc_csr *Op, const orc_vars *v, const orc_cplx *dft,
                                   const orc_cplx *Y_all, orc_cplx *Y_hat);

/* TransientRunRK4.c:128-158 : loop of unforced steps with norm every TransSaveMod steps.
 * norms: floor(rend/TransSaveMod)+1 doubles (norms[0] = ||q0||).  Returns number of steps done in *steps_done,
 * stop reason in *reason (0 finished, 1 NaN, 2 diverged, 3 converged). */
int orc_transient_run_rk4(const orc_csr *Op, double dt, orc_int rend, orc_int TransSaveMod, double TransDivVal,
                          double TransConVal, orc_cplx *q0, double *norms, orc_int *steps_done, int *reason);

/* SVDAllFreqsBeforeAdjoint.c:31-53 / SVDAllFreqs.c:34-64 (SLEPc default = cross-product SVD):
 * per frequency block Y_w (N x k): left singular vectors -> U (N x k*Nw_eff), singular values -> S (k x Nw_eff, may be NULL) */
int orc_svd_all_freqs(orc_int N, orc_int k, orc_int Nw_eff, const orc_cplx *Y_hat, orc_cplx *U, double *S);

/* ApplyWeightMats.c:18-76 building block: Y = M X or Y = M^H X for sparse M (MatMatMult sparse x dense) */
void orc_mat_mat_mult(const orc_csr *M, int hermitian_transpose, orc_int ncols, const orc_cplx *X, orc_cplx *Y);

/* timing helper for the CPU baseline: runs `nsteps` forced RK4 steps for `ncols` test vectors one at a time
 * (exactly the loop of TSActionRK4.c:89-122, truncated) and returns wall seconds */
double orc_time_forced_steps(const orc_csr *Op, const orc_vars *v, const orc_cplx *idft, const orc_cplx *Y_hat_in,
                             orc_int ncols, orc_int nsteps);

/* the state y after `nsteps` forced RK4 steps of the same loop for test vectors 0..ncols-1: y_out is N x ncols, column-major
 * (what TSActionRK4.c:116-122 holds in `y` after i = nsteps) */
void orc_forced_state(const orc_csr *Op, const orc_vars *v, const orc_cplx *idft, const orc_cplx *Y_hat_in, orc_int ncols,
                      orc_int nsteps, orc_cplx *y_out);

int orc_num_threads(void);

#endif
