/*
 * rsvdt_oracle.c -- CPU oracle for the RSVD-Delta-t time-stepping hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rsvdt_oracle.h).  PARITY UNPINNED: the reference has no golden vectors.
 *
 * The code is deliberately structured like the reference's PETSc path: one test vector at a time,
 * CSR with 64-bit indices and complex-double values, separate SpMV / WAXPY / AXPY passes, and a dense
 * mat-vec for the forcing synthesis.  OpenMP row blocks stand in for the MPI ranks of PETSc's MPIAIJ.
 * Citations are to /root/reference/SourceCode/<file>:<lines>.
 */
#include "rsvdt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PAR_MIN 16384 /* below this, fork/join costs more than the loop */
#define ORC_PI 3.14159265358979323846264338327950288

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double orc_wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------
 * SetupTimeFreqGrid.c:16-21
 * ---------------------------------------------------------------------------------------------- */
int orc_setup_time_freq_grid(double w, orc_int Nw, double dt_in, double TransientLength,
                             double *dt, orc_int *Ns, orc_int *Nt, orc_int *ResRatio)
{
    double T_ss = 2 * ORC_PI / w;                      /* :16 */
    double dt_w = T_ss / (double)Nw;                   /* :17 */
    *dt = dt_w / ceil(dt_w / dt_in);                   /* :18 */
    *Ns = (orc_int)round(T_ss / *dt);                  /* :19 */
    *Nt = (orc_int)round(TransientLength / *dt);       /* :20 */
    *ResRatio = *Ns / Nw;                              /* :21 integer division */
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CreateDFTiDFTMats.c
 *   idft (before transpose) is 2Ns x Nw_eff; column iw = (2/Nt) * conj(exp(v*J)), v = alpha*m', alpha=-2*pi*i/Nt
 *   with m' = iw (real A, :44) or iw / (iw-Nw)+Nt (complex A, :46-50); then transposed in place (:93).
 *   dft is Nw x Nw_eff, dft[j,iw] = exp(-2*pi*i*iw*j/Nw) (:74-84; for real A the vector J has 2*Nw_eff = Nw rows).
 * ---------------------------------------------------------------------------------------------- */
static orc_cplx unit_phase_exact(long long num, long long den)
{
    /* exp(2*pi*i*num/den) with the integer phase reduced first, evaluated in long double */
    long long r = num % den;
    if (r < 0) r += den;
    long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)r / (long double)den;
    return (double)cosl(ang) + I * (double)sinl(ang);
}

int orc_create_dft_idft_mats(orc_int Nw, orc_int Nw_eff, orc_int Ns, int RealOperator, int faithful,
                             orc_cplx *dft, orc_cplx *idft)
{
    const orc_int Nt = 2 * Ns;   /* rows of idft before the transpose (:25) */
    const orc_int Nwl = Nw_eff;  /* the local "Nw" of the reference after MatGetSize (:27) */
    for (orc_int iw = 0; iw < Nwl; iw++) {
        orc_int mp; /* multiplier of alpha */
        if (RealOperator) mp = iw;                                         /* :44 */
        else if (Nwl % 2 == 0) mp = (iw <= Nwl / 2 - 1) ? iw : (iw - Nwl) + Nt; /* :46 */
        else mp = (iw <= (Nwl - 1) / 2) ? iw : (iw - Nwl) + Nt;            /* :48 */
        if (faithful) {
            orc_cplx alpha = -2 * ORC_PI * I / (double)Nt;                 /* :42 */
            orc_cplx v = alpha * (double)mp;
            for (orc_int it = 0; it < Nt; it++) {
                orc_cplx e = cexp(v * (orc_cplx)(double)it);               /* VecPointwiseMult + VecExp :52-53 */
                e = conj(e);                                               /* :54 */
                e *= 2. / (double)Nt;                                      /* :55 */
                idft[iw + it * Nwl] = e;                                   /* transposed storage (:93) */
            }
        } else {
            for (orc_int it = 0; it < Nt; it++)
                idft[iw + it * Nwl] = unit_phase_exact((long long)mp * (long long)it, Nt) * (2. / (double)Nt);
        }
    }
    for (orc_int iw = 0; iw < Nwl; iw++) {
        if (faithful) {
            orc_cplx alpha = -2 * ORC_PI * I / (double)Nw;                 /* :74 (2*Nw_eff = Nw for real A) */
            orc_cplx v = alpha * (double)iw;                               /* :78 */
            for (orc_int j = 0; j < Nw; j++) dft[j + iw * Nw] = cexp(v * (orc_cplx)(double)j);
        } else {
            for (orc_int j = 0; j < Nw; j++) dft[j + iw * Nw] = unit_phase_exact(-(long long)iw * (long long)j, Nw);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Sparse kernels (PETSc MatMult / MatMultAdd on AIJ; MatMultHermitianTranspose[Add] through an explicit
 * conjugate transpose so that it is the same gather SpMV).
 * ---------------------------------------------------------------------------------------------- */
int orc_csr_hermitian_transpose(const orc_csr *A, orc_int *t_rowptr, orc_int *t_col, orc_cplx *t_val)
{
    const orc_int nr = A->nrows, nc = A->ncols, nnz = A->rowptr[nr];
    memset(t_rowptr, 0, (size_t)(nc + 1) * sizeof(orc_int));
    for (orc_int p = 0; p < nnz; p++) t_rowptr[A->col[p] + 1]++;
    for (orc_int c = 0; c < nc; c++) t_rowptr[c + 1] += t_rowptr[c];
    orc_int *next = (orc_int *)malloc((size_t)(nc > 0 ? nc : 1) * sizeof(orc_int));
    if (!next) return 1;
    memcpy(next, t_rowptr, (size_t)nc * sizeof(orc_int));
    for (orc_int r = 0; r < nr; r++)
        for (orc_int p = A->rowptr[r]; p < A->rowptr[r + 1]; p++) {
            orc_int q = next[A->col[p]]++;
            t_col[q] = r;
            t_val[q] = conj(A->val[p]);
        }
    free(next);
    return 0;
}

void orc_mat_mult_add(const orc_csr *A, const orc_cplx *x, const orc_cplx *f, orc_cplx *y)
{
    const orc_int n = A->nrows;
#pragma omp parallel for schedule(static) if (n >= ORC_PAR_MIN)
    for (orc_int r = 0; r < n; r++) {
        orc_cplx s = f ? f[r] : 0.0;
        for (orc_int p = A->rowptr[r]; p < A->rowptr[r + 1]; p++) s += A->val[p] * x[A->col[p]];
        y[r] = s;
    }
}

void orc_mat_mat_mult(const orc_csr *M, int hermitian_transpose, orc_int ncols, const orc_cplx *X, orc_cplx *Y)
{
    /* ApplyWeightMats.c:18-76 : MatMatMult(sparse, dense).  X is ncols columns of length (ht ? nrows : ncols). */
    if (!hermitian_transpose) {
        for (orc_int c = 0; c < ncols; c++)
            orc_mat_mult_add(M, X + c * M->ncols, NULL, Y + c * M->nrows);
        return;
    }
    const orc_int nnz = M->rowptr[M->nrows];
    orc_int *tr = (orc_int *)malloc((size_t)(M->ncols + 1) * sizeof(orc_int));
    orc_int *tc = (orc_int *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(orc_int));
    orc_cplx *tv = (orc_cplx *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(orc_cplx));
    orc_csr_hermitian_transpose(M, tr, tc, tv);
    orc_csr T = {M->ncols, M->nrows, tr, tc, tv};
    for (orc_int c = 0; c < ncols; c++) orc_mat_mult_add(&T, X + c * M->nrows, NULL, Y + c * M->ncols);
    free(tr); free(tc); free(tv);
}

/* VecWAXPY(w, alpha, x, y): w = alpha*x + y  (TSRK4.c:30,32,34) */
static void vec_waxpy(orc_int n, orc_cplx *w, double alpha, const orc_cplx *x, const orc_cplx *y)
{
#pragma omp parallel for schedule(static) if (n >= ORC_PAR_MIN)
    for (orc_int i = 0; i < n; i++) w[i] = alpha * x[i] + y[i];
}

/* VecAXPY(y, alpha, x): y += alpha*x  (TSRK4.c:44-47) */
static void vec_axpy(orc_int n, orc_cplx *y, double alpha, const orc_cplx *x)
{
#pragma omp parallel for schedule(static) if (n >= ORC_PAR_MIN)
    for (orc_int i = 0; i < n; i++) y[i] += alpha * x[i];
}

/* ------------------------------------------------------------------------------------------------
 * CreateForcingOnFly.c:17-19 : dense MatMult, F = F_hat_k * idft[:, jt_cyc]
 * ---------------------------------------------------------------------------------------------- */
void orc_create_forcing_on_fly(orc_int N, orc_int Nw_eff, const orc_cplx *F_hat_k, const orc_cplx *idft,
                               orc_int jt_cyc, orc_cplx *F)
{
    const orc_cplx *W_col = idft + jt_cyc * Nw_eff; /* MatDenseGetColumnVecRead(idft, jt_cyc) :17 */
    enum { BLK = 512 };
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
    for (orc_int r0 = 0; r0 < N; r0 += BLK) {
        orc_int nb = (N - r0 < BLK) ? (N - r0) : BLK;
        orc_cplx acc[BLK];
        for (orc_int i = 0; i < nb; i++) acc[i] = 0.0;
        for (orc_int iw = 0; iw < Nw_eff; iw++) { /* gemv 'N' : column sweep */
            const orc_cplx wv = W_col[iw];
            const orc_cplx *col = F_hat_k + iw * N + r0;
            for (orc_int i = 0; i < nb; i++) acc[i] += col[i] * wv;
        }
        for (orc_int i = 0; i < nb; i++) F[r0 + i] = acc[i];
    }
}

/* forcing column indices of TSRK4.c:21-22,25-26 */
static inline orc_int pos_mod(orc_int a, orc_int m)
{
    orc_int r = a % m;
    return r < 0 ? r + m : r;
}

static void forcing_indices(const orc_vars *v, orc_int i, orc_int *j2, orc_int *j3)
{
    const orc_int P = 2 * v->Ns;
    if (v->DirAdj) {
        *j2 = pos_mod(2 * (i - 1) + 1, P);                 /* TSRK4.c:21 */
        *j3 = pos_mod(2 * (i - 1) + 2, P);                 /* TSRK4.c:25 */
    } else {
        /* 1e6*Ns is an exact multiple of 2*Ns, so fmod(1e6*Ns - 1 - 2(i-1), 2Ns) = (-1-2(i-1)) mod 2Ns */
        *j2 = pos_mod(-1 - 2 * (i - 1), P);                /* TSRK4.c:22 */
        *j3 = pos_mod(-1 - (2 * (i - 1) + 1), P);          /* TSRK4.c:26 */
    }
}

/* ------------------------------------------------------------------------------------------------
 * TSRK4.c:18-47
 * ---------------------------------------------------------------------------------------------- */
void orc_tsrk4(const orc_csr *Op, const orc_vars *v, const orc_cplx *F_hat_k, const orc_cplx *idft,
               orc_int i, orc_cplx *y, orc_cplx *work)
{
    const orc_int N = v->N;
    const double dt = v->dt;
    orc_cplx *F2 = work, *F3 = work + N, *k1 = work + 2 * N, *k2 = work + 3 * N, *k3 = work + 4 * N,
             *k4 = work + 5 * N, *y_temp = work + 6 * N;
    orc_int j2, j3;

    orc_mat_mult_add(Op, y, F3, k1);                                       /* :18-19 */
    forcing_indices(v, i, &j2, &j3);
    orc_create_forcing_on_fly(N, v->Nw_eff, F_hat_k, idft, j2, F2);         /* :23 */
    orc_create_forcing_on_fly(N, v->Nw_eff, F_hat_k, idft, j3, F3);         /* :27 */

    vec_waxpy(N, y_temp, dt / 2, k1, y);                                   /* :30 */
    orc_mat_mult_add(Op, y_temp, F2, k2);                                  /* :31 */
    vec_waxpy(N, y_temp, dt / 2, k2, y);                                   /* :32 */
    orc_mat_mult_add(Op, y_temp, F2, k3);                                  /* :33 */
    vec_waxpy(N, y_temp, dt, k3, y);                                       /* :34 */
    orc_mat_mult_add(Op, y_temp, F3, k4);                                  /* :35 */

    vec_axpy(N, y, dt / 6, k1);                                            /* :44 */
    vec_axpy(N, y, dt / 6 * 2, k2);                                        /* :45 */
    vec_axpy(N, y, dt / 6 * 2, k3);                                        /* :46 */
    vec_axpy(N, y, dt / 6, k4);                                            /* :47 */
}

/* TSTransRK4.c:15-36 */
void orc_tstrans_rk4(const orc_csr *Op, double dt, orc_cplx *y, orc_cplx *work)
{
    const orc_int N = Op->nrows;
    orc_cplx *k1 = work, *k2 = work + N, *k3 = work + 2 * N, *k4 = work + 3 * N, *y_temp = work + 4 * N;
    orc_mat_mult_add(Op, y, NULL, k1);                                     /* :16 */
    vec_waxpy(N, y_temp, dt / 2, k1, y);                                   /* :17 */
    orc_mat_mult_add(Op, y_temp, NULL, k2);                                /* :18 */
    vec_waxpy(N, y_temp, dt / 2, k2, y);                                   /* :19 */
    orc_mat_mult_add(Op, y_temp, NULL, k3);                                /* :20 */
    vec_waxpy(N, y_temp, dt, k3, y);                                       /* :21 */
    orc_mat_mult_add(Op, y_temp, NULL, k4);                                /* :22 */
    vec_axpy(N, y, dt / 6, k1);                                            /* :33 */
    vec_axpy(N, y, dt / 6 * 2, k2);                                        /* :34 */
    vec_axpy(N, y, dt / 6 * 2, k3);                                        /* :35 */
    vec_axpy(N, y, dt / 6, k4);                                            /* :36 */
}

/* ------------------------------------------------------------------------------------------------
 * PermuteMat.c:24-38, ReversePermuteMat.c:24-38
 * ---------------------------------------------------------------------------------------------- */
void orc_permute_mat(orc_int N, orc_int k, orc_int Nw, const orc_cplx *F, orc_cplx *Fp)
{
    for (orc_int ik = 0; ik < k; ik++)
        for (orc_int iw = 0; iw < Nw; iw++) {
            orc_int ii = ik + iw * k, jj = ik * Nw + iw;                   /* :26-27 */
            memcpy(Fp + jj * N, F + ii * N, (size_t)N * sizeof(orc_cplx));
        }
}

void orc_reverse_permute_mat(orc_int N, orc_int k, orc_int Nw, const orc_cplx *F, orc_cplx *Fp)
{
    for (orc_int ik = 0; ik < k; ik++)
        for (orc_int iw = 0; iw < Nw; iw++) {
            orc_int ii = ik + iw * k, jj = ik * Nw + iw;
            memcpy(Fp + ii * N, F + jj * N, (size_t)N * sizeof(orc_cplx));
        }
}

/* dense C (N x n) = A (N x m) * B (m x n), column-major; PETSc MatMatMult dense x dense (DFT.c:29) */
static void dense_gemm(orc_int N, orc_int m, orc_int n, const orc_cplx *A, const orc_cplx *B, orc_cplx *C)
{
    enum { BLK = 256 };
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
    for (orc_int r0 = 0; r0 < N; r0 += BLK) {
        orc_int nb = (N - r0 < BLK) ? (N - r0) : BLK;
        orc_cplx acc[BLK];
        for (orc_int c = 0; c < n; c++) {
            for (orc_int i = 0; i < nb; i++) acc[i] = 0.0;
            for (orc_int j = 0; j < m; j++) {
                const orc_cplx b = B[j + c * m];
                const orc_cplx *a = A + j * N + r0;
                for (orc_int i = 0; i < nb; i++) acc[i] += a[i] * b;
            }
            for (orc_int i = 0; i < nb; i++) C[r0 + i + c * N] = acc[i];
        }
    }
}

/* DFT.c:26-39 */
void orc_dft(orc_int N, orc_int k, orc_int Nw, orc_int Nw_eff, orc_int ResRatio, const orc_cplx *dft,
             const orc_cplx *Y_all, orc_cplx *Y_hat)
{
    orc_cplx *tmp = (orc_cplx *)malloc((size_t)N * (size_t)k * (size_t)Nw_eff * sizeof(orc_cplx));
    for (orc_int ik = 0; ik < k; ik++)                                     /* :26-32 */
        dense_gemm(N, Nw, Nw_eff, Y_all + ik * Nw * N, dft, tmp + ik * Nw_eff * N);
    const orc_int tot = N * k * Nw_eff;
#pragma omp parallel for schedule(static)
    for (orc_int i = 0; i < tot; i++) tmp[i] *= (double)ResRatio;          /* :38 */
    orc_reverse_permute_mat(N, k, Nw_eff, tmp, Y_hat);                     /* :39 */
    free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * QRDecomposition.c:16-20 / QRAllFreqs.c:24-28 : SLEPc BVOrthogonalize on BVVECS = column-by-column
 * Gram-Schmidt (classical, with DGKS refinement when needed, eta = 1/sqrt(2)), R discarded.
 * The thin-QR factor with positive real diag(R) is unique, so any stable variant gives the same Q to rounding.
 * ---------------------------------------------------------------------------------------------- */
static double vec_norm2(orc_int n, const orc_cplx *x)
{
    double s = 0;
#pragma omp parallel for schedule(static) if (n >= ORC_PAR_MIN) reduction(+ : s)
    for (orc_int i = 0; i < n; i++) s += creal(x[i]) * creal(x[i]) + cimag(x[i]) * cimag(x[i]);
    return sqrt(s);
}

static orc_cplx vec_dot_conj(orc_int n, const orc_cplx *q, const orc_cplx *x) /* q^H x */
{
    double sr = 0, si = 0;
#pragma omp parallel for schedule(static) if (n >= ORC_PAR_MIN) reduction(+ : sr, si)
    for (orc_int i = 0; i < n; i++) {
        orc_cplx t = conj(q[i]) * x[i];
        sr += creal(t);
        si += cimag(t);
    }
    return sr + I * si;
}

int orc_qr_decomposition(orc_int N, orc_int m, orc_cplx *V)
{
    orc_cplx *h = (orc_cplx *)malloc((size_t)(m > 0 ? m : 1) * sizeof(orc_cplx));
    for (orc_int j = 0; j < m; j++) {
        orc_cplx *v = V + j * N;
        double nrm_old = vec_norm2(N, v);
        for (int pass = 0; pass < 3; pass++) {
            for (orc_int c = 0; c < j; c++) h[c] = vec_dot_conj(N, V + c * N, v);  /* classical GS */
            for (orc_int c = 0; c < j; c++) {
                const orc_cplx hc = h[c];
                const orc_cplx *q = V + c * N;
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
                for (orc_int i = 0; i < N; i++) v[i] -= hc * q[i];
            }
            double nrm_new = vec_norm2(N, v);
            int ok = (nrm_new >= 0.7071067811865475 * nrm_old) || j == 0;
            nrm_old = nrm_new;
            if (ok) break;
        }
        if (!(nrm_old > 0)) { free(h); return 1; }
        const double inv = 1.0 / nrm_old;
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
        for (orc_int i = 0; i < N; i++) v[i] *= inv;
    }
    free(h);
    return 0;
}

int orc_qr_all_freqs(orc_int N, orc_int k, orc_int Nw_eff, orc_cplx *Y_hat)
{
    for (orc_int iw = 0; iw < Nw_eff; iw++) {                              /* QRAllFreqs.c:22 */
        int e = orc_qr_decomposition(N, k, Y_hat + iw * k * N);
        if (e) return e;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * TSExpADeltaT.c:25-34
 * ---------------------------------------------------------------------------------------------- */
void orc_ts_exp_a_deltat(const orc_csr *Op, const orc_vars *v, orc_int ncols, orc_cplx *U)
{
    const orc_int N = v->N;
    orc_cplx *work = (orc_cplx *)malloc((size_t)5 * (size_t)N * sizeof(orc_cplx));
    for (orc_int is = 0; is < ncols; is++)                                 /* :25 */
        for (orc_int it = 0; it < v->ResRatio; it++)                       /* :29 */
            orc_tstrans_rk4(Op, v->dt, U + is * N, work);                  /* :30 */
    free(work);
}

/* ------------------------------------------------------------------------------------------------
 * Small dense helpers for the Galerkin projection (replicated (Nw+1)^2 problems)
 * ---------------------------------------------------------------------------------------------- */
/* C (m x n) = A^H (A: N x m) * B (N x n)   -- MatTransposeMatMult on the conjugated V (CreateMtilde.c:49-51) */
static void dense_AhB(orc_int N, orc_int m, orc_int n, const orc_cplx *A, const orc_cplx *B, orc_cplx *C)
{
    for (orc_int c = 0; c < n; c++)
        for (orc_int r = 0; r < m; r++) C[r + c * m] = vec_dot_conj(N, A + r * N, B + c * N);
}

/* in-place LU with partial pivoting + solve, column-major n x n (stands in for KSPSolve on the dense
 * (Nw+1)^2 system, GalerkinProj.c:57-62; PGMRES preconditioned by a full LU is a direct solve) */
static int dense_lu_solve(orc_int n, orc_cplx *M, orc_cplx *b)
{
    for (orc_int c = 0; c < n; c++) {
        orc_int p = c;
        double best = cabs(M[c + c * n]);
        for (orc_int r = c + 1; r < n; r++)
            if (cabs(M[r + c * n]) > best) { best = cabs(M[r + c * n]); p = r; }
        if (best == 0.0) return 1;
        if (p != c) {
            for (orc_int cc = 0; cc < n; cc++) { orc_cplx t = M[c + cc * n]; M[c + cc * n] = M[p + cc * n]; M[p + cc * n] = t; }
            orc_cplx t = b[c]; b[c] = b[p]; b[p] = t;
        }
        const orc_cplx piv = M[c + c * n];
        for (orc_int r = c + 1; r < n; r++) {
            const orc_cplx l = M[r + c * n] / piv;
            M[r + c * n] = l;
            for (orc_int cc = c + 1; cc < n; cc++) M[r + cc * n] -= l * M[c + cc * n];
            b[r] -= l * b[c];
        }
    }
    for (orc_int r = n - 1; r >= 0; r--) {
        orc_cplx s = b[r];
        for (orc_int cc = r + 1; cc < n; cc++) s -= M[r + cc * n] * b[cc];
        b[r] = s / M[r + r * n];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * TransientRemovalStrategy.c:30-40
 * ---------------------------------------------------------------------------------------------- */
int orc_transient_removal_strategy(const orc_csr *Op, const orc_vars *v, const orc_cplx *dft,
                                   const orc_cplx *Y_all, orc_cplx *Y_hat)
{
    const orc_int N = v->N, k = v->k, Nw = v->Nw, Nw_eff = v->Nw_eff, Nstore = Nw + 1;
    const double deltaT = v->dt * (double)v->ResRatio;                      /* GalerkinProj.c:17 */
    orc_cplx *V = (orc_cplx *)malloc((size_t)N * Nstore * sizeof(orc_cplx));
    orc_cplx *MV = (orc_cplx *)malloc((size_t)N * Nstore * sizeof(orc_cplx));
    orc_cplx *Mt = (orc_cplx *)malloc((size_t)Nstore * Nstore * sizeof(orc_cplx));
    orc_cplx *Mtmp = (orc_cplx *)malloc((size_t)Nstore * Nstore * sizeof(orc_cplx));
    orc_cplx *Y_up = (orc_cplx *)malloc((size_t)N * k * Nw_eff * sizeof(orc_cplx));
    orc_cplx *y1 = (orc_cplx *)malloc((size_t)N * sizeof(orc_cplx));
    orc_cplx *y2 = (orc_cplx *)malloc((size_t)N * sizeof(orc_cplx));
    orc_cplx *yt = (orc_cplx *)malloc((size_t)N * sizeof(orc_cplx));
    orc_cplx *b = (orc_cplx *)malloc((size_t)Nstore * sizeof(orc_cplx));
    int err = 0;

    for (orc_int ik = 0; ik < k && !err; ik++) {
        const orc_cplx *Y_all_k = Y_all + ik * Nstore * N;
        /* CreateMtilde.c:44-51 : V = QR(Y_all_k); MV = exp(A dT) V; M_tilde = V^H MV */
        memcpy(V, Y_all_k, (size_t)N * Nstore * sizeof(orc_cplx));
        err = orc_qr_decomposition(N, Nstore, V);
        if (err) break;
        memcpy(MV, V, (size_t)N * Nstore * sizeof(orc_cplx));
        orc_ts_exp_a_deltat(Op, v, Nstore, MV);
        dense_AhB(N, Nstore, Nstore, V, MV, Mt);
        /* GalerkinProjAllModes.c:31-35 recomputes the same QR; V is reused here */
        for (orc_int iw = 0; iw < Nw_eff && !err; iw++) {                   /* GalerkinProj.c:26 */
            const orc_cplx *A1 = v->DirAdj ? Y_all_k : Y_all_k + N;         /* :30-31 */
            const orc_cplx *A2 = v->DirAdj ? Y_all_k + N : Y_all_k;         /* :40-41 */
            const orc_cplx *d = dft + iw * Nw;
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
            for (orc_int r = 0; r < N; r++) {
                orc_cplx s1 = 0, s2 = 0;
                for (orc_int j = 0; j < Nw; j++) {
                    s1 += A1[r + j * N] * d[j];                             /* :35 */
                    s2 += A2[r + j * N] * d[j];                             /* :43 */
                }
                y1[r] = s1 * (double)v->ResRatio;                           /* :36 */
                y2[r] = s2 * (double)v->ResRatio;                           /* :44 */
            }
            const orc_cplx alpha = v->DirAdj ? cexp(I * v->w * (double)iw * deltaT)
                                             : cexp(-I * v->w * (double)iw * deltaT);  /* :49-50 */
            for (orc_int i = 0; i < Nstore * Nstore; i++) Mtmp[i] = -Mt[i];  /* :48 */
            for (orc_int i = 0; i < Nstore; i++) Mtmp[i + i * Nstore] += alpha; /* :49-50 MatShift */
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
            for (orc_int r = 0; r < N; r++) yt[r] = alpha * y1[r] - y2[r];  /* :51-54 */
            for (orc_int c = 0; c < Nstore; c++) b[c] = vec_dot_conj(N, V + c * N, yt); /* :55 */
            err = dense_lu_solve(Nstore, Mtmp, b);                          /* :56-62 */
            if (err) break;
            orc_cplx *out = Y_up + (ik * Nw_eff + iw) * N;                  /* :67-70 */
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
            for (orc_int r = 0; r < N; r++) {
                orc_cplx s = 0;
                for (orc_int c = 0; c < Nstore; c++) s += V[r + c * N] * b[c]; /* :65 */
                out[r] = y1[r] - s;                                         /* :66 */
            }
        }
    }
    if (!err) orc_reverse_permute_mat(N, k, Nw_eff, Y_up, Y_hat);           /* TransientRemovalStrategy.c:38 */
    free(V); free(MV); free(Mt); free(Mtmp); free(Y_up); free(y1); free(y2); free(yt); free(b);
    return err;
}

/* ------------------------------------------------------------------------------------------------
 * TSActionRK4.c:38-148
 * ---------------------------------------------------------------------------------------------- */
int orc_ts_action_rk4(const orc_csr *Op, const orc_vars *v, const orc_cplx *dft, const orc_cplx *idft,
                      const orc_cplx *Y_hat_in, orc_cplx *Y_hat_out, orc_cplx *snapshots_out, orc_int max_steps)
{
    const orc_int N = v->N, k = v->k, Nw = v->Nw, Nw_eff = v->Nw_eff;
    orc_int rend = v->Nt + v->Ns;                                           /* :41 */
    rend += v->TransientRemoval ? v->ResRatio : 0;                          /* :42 */
    const orc_int Nstore = v->TransientRemoval ? Nw + 1 : Nw;               /* :43 */
    if (max_steps > 0 && max_steps < rend) rend = max_steps;

    orc_cplx *F_hat = (orc_cplx *)malloc((size_t)N * k * Nw_eff * sizeof(orc_cplx));
    orc_cplx *Y_all = (orc_cplx *)calloc((size_t)N * Nstore * k, sizeof(orc_cplx));
    orc_cplx *y = (orc_cplx *)malloc((size_t)N * sizeof(orc_cplx));
    orc_cplx *work = (orc_cplx *)malloc((size_t)7 * N * sizeof(orc_cplx));
    if (!F_hat || !Y_all || !y || !work) return 1;

    orc_permute_mat(N, k, Nw_eff, Y_hat_in, F_hat);                         /* :49,76 */

    for (orc_int ik = 0; ik < k; ik++) {                                    /* :89 */
        memset(y, 0, (size_t)N * sizeof(orc_cplx));                         /* :91 */
        const orc_cplx *F_hat_k = F_hat + ik * Nw_eff * N;                  /* :93-97 */
        orc_cplx *Y_all_k = Y_all + ik * Nstore * N;
        orc_create_forcing_on_fly(N, Nw_eff, F_hat_k, idft, 0, work + N);   /* :103 -> F3 */
        for (orc_int i = 1; i <= rend; i++) {                               /* :116 */
            orc_tsrk4(Op, v, F_hat_k, idft, i, y, work);                    /* :118 */
            if (i > v->Nt) {                                                /* SaveSnapshots.c:19 */
                orc_int pos = i - v->Nt;
                if (pos % v->ResRatio == 0) {                               /* :21 */
                    pos = pos / v->ResRatio - 1;                            /* :22 */
                    pos = pos_mod(pos, Nstore);                             /* :23 */
                    pos = v->DirAdj ? pos : Nstore - 1 - pos;               /* :24 */
                    pos = pos_mod(pos, Nstore);                             /* :25 */
                    memcpy(Y_all_k + pos * N, y, (size_t)N * sizeof(orc_cplx)); /* :26-28 */
                }
            }
        }
    }
    free(y); free(F_hat); free(work);
    if (snapshots_out) memcpy(snapshots_out, Y_all, (size_t)N * Nstore * k * sizeof(orc_cplx));

    int err = 0;
    if (v->TransientRemoval) err = orc_transient_removal_strategy(Op, v, dft, Y_all, Y_hat_out);  /* :147 */
    else orc_dft(N, k, Nw, Nw_eff, v->ResRatio, dft, Y_all, Y_hat_out);                          /* :148 */
    free(Y_all);
    return err;
}

double orc_time_forced_steps(const orc_csr *Op, const orc_vars *v, const orc_cplx *idft, const orc_cplx *Y_hat_in,
                             orc_int ncols, orc_int nsteps)
{
    const orc_int N = v->N, Nw_eff = v->Nw_eff;
    orc_cplx *F_hat_k = (orc_cplx *)malloc((size_t)N * Nw_eff * sizeof(orc_cplx));
    orc_cplx *y = (orc_cplx *)malloc((size_t)N * sizeof(orc_cplx));
    orc_cplx *work = (orc_cplx *)malloc((size_t)7 * N * sizeof(orc_cplx));
    double t = 0;
    for (orc_int ik = 0; ik < ncols; ik++) {
        for (orc_int iw = 0; iw < Nw_eff; iw++) /* the test-vector-major slice of PermuteMat */
            memcpy(F_hat_k + iw * N, Y_hat_in + (ik + iw * v->k) * N, (size_t)N * sizeof(orc_cplx));
        memset(y, 0, (size_t)N * sizeof(orc_cplx));
        orc_create_forcing_on_fly(N, Nw_eff, F_hat_k, idft, 0, work + N);
        double t0 = orc_wtime();
        for (orc_int i = 1; i <= nsteps; i++) orc_tsrk4(Op, v, F_hat_k, idft, i, y, work);
        t += orc_wtime() - t0;
    }
    free(F_hat_k); free(y); free(work);
    return t;
}

/* TSActionRK4.c:89-122 truncated after nsteps steps: the state of every test vector (checker of the bench-size run) */
void orc_forced_state(const orc_csr *Op, const orc_vars *v, const orc_cplx *idft, const orc_cplx *Y_hat_in, orc_int ncols,
                      orc_int nsteps, orc_cplx *y_out)
{
    const orc_int N = v->N, Nw_eff = v->Nw_eff;
    orc_cplx *F_hat_k = (orc_cplx *)malloc((size_t)N * Nw_eff * sizeof(orc_cplx));
    orc_cplx *work = (orc_cplx *)malloc((size_t)7 * N * sizeof(orc_cplx));
    for (orc_int ik = 0; ik < ncols; ik++) {
        orc_cplx *y = y_out + ik * N;
        for (orc_int iw = 0; iw < Nw_eff; iw++) /* the test-vector-major slice of PermuteMat */
            memcpy(F_hat_k + iw * N, Y_hat_in + (ik + iw * v->k) * N, (size_t)N * sizeof(orc_cplx));
        memset(y, 0, (size_t)N * sizeof(orc_cplx));                               /* TSActionRK4.c:91 */
        orc_create_forcing_on_fly(N, Nw_eff, F_hat_k, idft, 0, work + N);         /* :103 */
        for (orc_int i = 1; i <= nsteps; i++) orc_tsrk4(Op, v, F_hat_k, idft, i, y, work);   /* :116-118 */
    }
    free(F_hat_k); free(work);
}

/* ------------------------------------------------------------------------------------------------
 * TransientRunRK4.c:128-158
 * ---------------------------------------------------------------------------------------------- */
int orc_transient_run_rk4(const orc_csr *Op, double dt, orc_int rend, orc_int TransSaveMod, double TransDivVal,
                          double TransConVal, orc_cplx *q0, double *norms, orc_int *steps_done, int *reason)
{
    const orc_int N = Op->nrows;
    orc_cplx *work = (orc_cplx *)malloc((size_t)5 * N * sizeof(orc_cplx));
    norms[0] = vec_norm2(N, q0);                                            /* :106-108 */
    *reason = 0;
    orc_int i;
    for (i = 1; i <= rend; i++) {                                           /* :128 */
        orc_tstrans_rk4(Op, dt, q0, work);                                  /* :130 */
        if (i % TransSaveMod == 0) {                                        /* :137 */
            orc_int pos = i / TransSaveMod;
            double norm = vec_norm2(N, q0);                                 /* :141 */
            if (isnan(norm)) { *reason = 1; break; }                        /* :143-146 */
            if (norm > TransDivVal) { *reason = 2; break; }                 /* :147-151 */
            if (norm < TransConVal) { *reason = 3; break; }                 /* :152-156 */
            norms[pos] = norm;                                              /* :166 */
        }
    }
    *steps_done = (i > rend) ? rend : i;
    free(work);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SVDAllFreqsBeforeAdjoint.c:31-53, SVDAllFreqs.c:34-64.  SLEPc's default SVD solver is the cross-product
 * one: eigenpairs (lambda, v) of Y^H Y, sigma = sqrt(lambda), left vector u = Y v / sigma, sorted descending.
 * ---------------------------------------------------------------------------------------------- */
static void hermitian_jacobi(orc_int n, orc_cplx *G, orc_cplx *Vv, double *lam)
{
    /* cyclic complex Jacobi on a column-major Hermitian n x n matrix G; Vv returns eigenvectors (columns) */
    for (orc_int i = 0; i < n * n; i++) Vv[i] = 0;
    for (orc_int i = 0; i < n; i++) Vv[i + i * n] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (orc_int c = 0; c < n; c++)
            for (orc_int r = 0; r < n; r++) {
                double a = cabs(G[r + c * n]);
                if (r == c) diag += a * a; else off += a * a;
            }
        if (off <= 1e-32 * diag) break;
        for (orc_int p = 0; p < n - 1; p++)
            for (orc_int q = p + 1; q < n; q++) {
                orc_cplx gpq = G[p + q * n];
                double apq = cabs(gpq);
                if (apq == 0.0) continue;
                double app = creal(G[p + p * n]), aqq = creal(G[q + q * n]);
                orc_cplx ph = gpq / apq; /* e^{i phi} */
                double tau = (aqq - app) / (2 * apq);
                double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1 + tau * tau));
                double c = 1 / sqrt(1 + t * t), s = t * c;
                /* rotation J: columns p,q <- [c, -s*ph; s*conj(ph), c] */
                for (orc_int r = 0; r < n; r++) { /* G <- G J */
                    orc_cplx gp = G[r + p * n], gq = G[r + q * n];
                    G[r + p * n] = c * gp - s * conj(ph) * gq;
                    G[r + q * n] = s * ph * gp + c * gq;
                }
                for (orc_int cc = 0; cc < n; cc++) { /* G <- J^H G */
                    orc_cplx gp = G[p + cc * n], gq = G[q + cc * n];
                    G[p + cc * n] = c * gp - s * ph * gq;
                    G[q + cc * n] = s * conj(ph) * gp + c * gq;
                }
                for (orc_int r = 0; r < n; r++) { /* V <- V J */
                    orc_cplx vp = Vv[r + p * n], vq = Vv[r + q * n];
                    Vv[r + p * n] = c * vp - s * conj(ph) * vq;
                    Vv[r + q * n] = s * ph * vp + c * vq;
                }
            }
    }
    for (orc_int i = 0; i < n; i++) lam[i] = creal(G[i + i * n]);
}

int orc_svd_all_freqs(orc_int N, orc_int k, orc_int Nw_eff, const orc_cplx *Y_hat, orc_cplx *U, double *S)
{
    orc_cplx *G = (orc_cplx *)malloc((size_t)k * k * sizeof(orc_cplx));
    orc_cplx *Vv = (orc_cplx *)malloc((size_t)k * k * sizeof(orc_cplx));
    double *lam = (double *)malloc((size_t)k * sizeof(double));
    orc_int *ord = (orc_int *)malloc((size_t)k * sizeof(orc_int));
    for (orc_int iw = 0; iw < Nw_eff; iw++) {
        const orc_cplx *Y = Y_hat + iw * k * N;
        dense_AhB(N, k, k, Y, Y, G);
        hermitian_jacobi(k, G, Vv, lam);
        for (orc_int i = 0; i < k; i++) ord[i] = i;
        for (orc_int i = 0; i < k; i++)
            for (orc_int j = i + 1; j < k; j++)
                if (lam[ord[j]] > lam[ord[i]]) { orc_int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
        for (orc_int m = 0; m < k; m++) {
            const orc_cplx *vv = Vv + ord[m] * k;
            double sigma = sqrt(lam[ord[m]] > 0 ? lam[ord[m]] : 0.0);
            if (S) S[m + iw * k] = sigma;
            orc_cplx *u = U + (iw * k + m) * N;
            const double inv = sigma > 0 ? 1.0 / sigma : 0.0;
#pragma omp parallel for schedule(static) if (N >= ORC_PAR_MIN)
            for (orc_int r = 0; r < N; r++) {
                orc_cplx s = 0;
                for (orc_int c = 0; c < k; c++) s += Y[r + c * N] * vv[c];
                u[r] = s * inv;
            }
        }
    }
    free(G); free(Vv); free(lam); free(ord);
    return 0;
}
