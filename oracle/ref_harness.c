/* ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * ctypes-callable entry points that build the reference's own argument structures (Variables.h) from plain arrays
 * and call the reference's own functions -- compiled unmodified from /root/reference/SourceCode by Makefile.ref
 * against the PETSc/SLEPc API shim -- so that tests can compare the oracle restatement (rsvdt_oracle.c) and the CUDA
 * path against what the reference code itself computes, function by function, and so that bench.py can time the
 * reference's RK4 step (TSRK4.c) on the box's host cores.  Dense matrices cross this boundary as column-major
 * complex double (the memory image of a one-rank MATDENSE). Every function returns the PetscErrorCode (0 = ok). */
#include <petscmat.h>

#include <Variables.h>

#include <ApplyWeightMats.h>
#include <CreateDFTiDFTMats.h>
#include <CreateTSMats.h>
#include <DestroyTSMats.h>
#include <AdjointActionRK4.h>
#include <DirectActionRK4.h>
#include <CreateForcingOnFly.h>
#include <PermuteMat.h>
#include <QRAllFreqs.h>
#include <QRDecomposition.h>
#include <ReversePermuteMat.h>
#include <SVDAllFreqs.h>
#include <SVDAllFreqsBeforeAdjoint.h>
#include <SetupTimeFreqGrid.h>
#include <StoreU.h>
#include <TSActionRK4.h>
#include <TSExpADeltaT.h>
#include <TSRK4.h>
#include <TSTransRK4.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct {
    PetscInt N, k, Nw, Nw_eff, Ns, Nt, ResRatio;
    PetscReal dt, w;
    int RealOperator, TransientRemoval, DirAdj;
} ref_vars; /* same field order as the oracle's orc_vars, so the Python side can share the structure */

static void fill_vars(const ref_vars *v, RSVDt_vars *R)
{
    memset(R, 0, sizeof *R);
    R->RSVD.N = R->RSVD.Nb = R->RSVD.Nc = v->N;
    R->RSVD.k = v->k;
    R->RSVD.Nw = v->Nw;
    R->RSVD.Nw_eff = v->Nw_eff;
    R->RSVD.w = v->w;
    R->TS.Ns = v->Ns;
    R->TS.Nt = v->Nt;
    R->TS.ResRatio = v->ResRatio;
    R->TS.dt = v->dt;
    R->TS.RealOperator = v->RealOperator ? PETSC_TRUE : PETSC_FALSE;
    R->TS.TransientRemoval = v->TransientRemoval ? PETSC_TRUE : PETSC_FALSE;
    R->TS.DirAdj = v->DirAdj ? PETSC_TRUE : PETSC_FALSE;
    R->Display = 0;
    R->SaveResultsOpt = 1;
}

static void copy_out(Mat A, PetscScalar *out)
{
    PetscInt m, n, lda;
    MatGetSize(A, &m, &n);
    const PetscScalar *a = ShimMatDenseArray(A, &lda);
    for (PetscInt j = 0; j < n; j++) memcpy(out + m * j, a + lda * j, (size_t)m * sizeof(PetscScalar));
}

/* SetupTimeFreqGrid.c */
EXPORT int ref_setup_time_freq_grid(double w, PetscInt Nw, double dt_in, double TransientLength, double *dt, PetscInt *Ns, PetscInt *Nt, PetscInt *RR)
{
    RSVDt_vars R;
    TransRun_vars T;
    memset(&R, 0, sizeof R);
    memset(&T, 0, sizeof T);
    T.TransRun = PETSC_TRUE; /* only silences the banner */
    R.RSVD.w = w;
    R.RSVD.Nw = Nw;
    R.TS.dt = dt_in;
    R.TS.TransientLength = TransientLength;
    PetscErrorCode e = SetupTimeFreqGrid(&R, &T);
    *dt = R.TS.dt;
    *Ns = R.TS.Ns;
    *Nt = R.TS.Nt;
    *RR = R.TS.ResRatio;
    return e;
}

/* CreateDFTiDFTMats.c: dft is (Nw or 2*Nw_eff) x Nw_eff ... as the reference sizes it: Nw x Nw_eff; idft is Nw_eff x 2Ns */
EXPORT int ref_create_dft_idft_mats(const ref_vars *v, PetscScalar *dft_out, PetscScalar *idft_out)
{
    RSVDt_vars R;
    DFT_matrices D;
    LNS_vars L;
    memset(&L, 0, sizeof L);
    fill_vars(v, &R);
    PetscErrorCode e = CreateDFTiDFTMats(&R, &D, &L);
    if (e) return e;
    if (dft_out) copy_out(D.dft, dft_out);
    if (idft_out) copy_out(D.idft, idft_out);
    MatDestroy(&D.dft);
    MatDestroy(&D.idft);
    return 0;
}

static PetscErrorCode make_ctx(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, RSVDt_vars *R, LNS_vars *L, DFT_matrices *D)
{
    PetscErrorCode ierr;
    fill_vars(v, R);
    memset(L, 0, sizeof *L);
    ierr = ShimMatCreateAIJFrom(v->N, v->N, rowptr, col, val, &L->A);CHKERRQ(ierr);
    if (D) {
        ierr = CreateDFTiDFTMats(R, D, L);CHKERRQ(ierr);
    }
    return 0;
}

/* TSActionRK4.c (with DFT.c or TransientRemovalStrategy.c inside): Y_hat [N x k*Nw_eff] frequency-major in and out */
EXPORT int ref_ts_action_rk4(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, const PetscScalar *Y_in, PetscScalar *Y_out)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    LNS_vars L;
    DFT_matrices D;
    RSVD_matrices M;
    TS_removal_matrices TSR;
    Directories dirs;
    memset(&M, 0, sizeof M);
    memset(&TSR, 0, sizeof TSR);
    memset(&dirs, 0, sizeof dirs);
    ShimSetQuiet(1);
    ierr = make_ctx(v, rowptr, col, val, &R, &L, &D);CHKERRQ(ierr);
    ierr = ShimMatCreateDenseFrom(v->N, v->k * v->Nw_eff, Y_in, &M.Y_hat);CHKERRQ(ierr);
    ierr = TSActionRK4(&M, &D, &L, &R, &dirs, &TSR);CHKERRQ(ierr);
    copy_out(M.Y_hat, Y_out);
    MatDestroy(&M.Y_hat);
    MatDestroy(&D.dft);
    MatDestroy(&D.idft);
    MatDestroy(&L.A);
    return 0;
}

/* DirectActionRK4.c / AdjointActionRK4.c with the optional weight, input and output matrices (CSR triples, NULL = off).
 * Sizes: B N x Nb, C Nc x N, W_f Nb x Nb, W_q Nc x Nc.  Y_in rows: Nb (direct) / Nc (adjoint); Y_out rows: Nc / Nb. */
typedef struct {
    PetscInt m, n;
    const PetscInt *rowptr, *col;
    const PetscScalar *val;
} ref_csr;

EXPORT int ref_weighted_action(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, const ref_csr *B, const ref_csr *C,
                               const ref_csr *Wf_inv, const ref_csr *Wq, PetscInt rows_in, const PetscScalar *Y_in, PetscInt rows_out, PetscScalar *Y_out)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    LNS_vars L;
    DFT_matrices D;
    RSVD_matrices M;
    TS_removal_matrices TSR;
    Directories dirs;
    Weight_matrices W;
    memset(&M, 0, sizeof M);
    memset(&TSR, 0, sizeof TSR);
    memset(&dirs, 0, sizeof dirs);
    memset(&W, 0, sizeof W);
    ShimSetQuiet(1);
    ierr = make_ctx(v, rowptr, col, val, &R, &L, &D);CHKERRQ(ierr);
    if (B) {
        ierr = ShimMatCreateAIJFrom(B->m, B->n, B->rowptr, B->col, B->val, &W.B);CHKERRQ(ierr);
        W.InputMatrixFlg = PETSC_TRUE;
        R.RSVD.Nb = B->n;
    }
    if (C) {
        ierr = ShimMatCreateAIJFrom(C->m, C->n, C->rowptr, C->col, C->val, &W.C);CHKERRQ(ierr);
        W.OutputMatrixFlg = PETSC_TRUE;
        R.RSVD.Nc = C->m;
    }
    if (Wf_inv) {
        ierr = ShimMatCreateAIJFrom(Wf_inv->m, Wf_inv->n, Wf_inv->rowptr, Wf_inv->col, Wf_inv->val, &W.W_f_sqrt_inv);CHKERRQ(ierr);
        W.InvInputWeightFlg = PETSC_TRUE;
    }
    if (Wq) {
        ierr = ShimMatCreateAIJFrom(Wq->m, Wq->n, Wq->rowptr, Wq->col, Wq->val, &W.W_q_sqrt);CHKERRQ(ierr);
        W.OutputWeightFlg = PETSC_TRUE;
    }
    ierr = ShimMatCreateDenseFrom(rows_in, v->k * v->Nw_eff, Y_in, &M.Y_hat);CHKERRQ(ierr);
    ierr = v->DirAdj ? DirectActionRK4(&M, &R, &L, &D, &W, &dirs, &TSR) : AdjointActionRK4(&M, &R, &L, &D, &W, &dirs, &TSR);CHKERRQ(ierr);
    PetscInt m, n;
    MatGetSize(M.Y_hat, &m, &n);
    if (m != rows_out) return PetscShimError(__FILE__, __LINE__, __func__, PETSC_ERR_ARG_SIZ, "output has %lld rows, caller expected %lld", (long long)m, (long long)rows_out);
    copy_out(M.Y_hat, Y_out);
    MatDestroy(&M.Y_hat);
    MatDestroy(&W.B);
    MatDestroy(&W.C);
    MatDestroy(&W.W_f_sqrt_inv);
    MatDestroy(&W.W_q_sqrt);
    MatDestroy(&D.dft);
    MatDestroy(&D.idft);
    MatDestroy(&L.A);
    return 0;
}

/* QRAllFreqs.c, in place */
EXPORT int ref_qr_all_freqs(PetscInt N, PetscInt k, PetscInt Nw_eff, PetscScalar *Y)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    RSVD_matrices M;
    memset(&R, 0, sizeof R);
    memset(&M, 0, sizeof M);
    R.RSVD.N = N;
    R.RSVD.k = k;
    R.RSVD.Nw_eff = Nw_eff;
    ShimSetQuiet(1);
    ierr = ShimMatCreateDenseFrom(N, k * Nw_eff, Y, &M.Y_hat);CHKERRQ(ierr);
    ierr = QRAllFreqs(&M, &R);CHKERRQ(ierr);
    copy_out(M.Y_hat, Y);
    MatDestroy(&M.Y_hat);
    return 0;
}

/* QRDecomposition.c, in place on an N x n matrix */
EXPORT int ref_qr_decomposition(PetscInt N, PetscInt n, PetscScalar *V)
{
    PetscErrorCode ierr;
    Mat A;
    ierr = ShimMatCreateDenseFrom(N, n, V, &A);CHKERRQ(ierr);
    ierr = QRDecomposition(A);CHKERRQ(ierr);
    copy_out(A, V);
    MatDestroy(&A);
    return 0;
}

/* SVDAllFreqsBeforeAdjoint.c: Y <- left singular vectors, per frequency */
EXPORT int ref_svd_before_adjoint(PetscInt N, PetscInt k, PetscInt Nw_eff, PetscScalar *Y)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    RSVD_matrices M;
    memset(&R, 0, sizeof R);
    memset(&M, 0, sizeof M);
    R.RSVD.N = R.RSVD.Nc = R.RSVD.Nb = N;
    R.RSVD.k = k;
    R.RSVD.Nw_eff = Nw_eff;
    ShimSetQuiet(1);
    ierr = ShimMatCreateDenseFrom(N, k * Nw_eff, Y, &M.Y_hat);CHKERRQ(ierr);
    ierr = SVDAllFreqsBeforeAdjoint(&M, &R);CHKERRQ(ierr);
    copy_out(M.Y_hat, Y);
    MatDestroy(&M.Y_hat);
    return 0;
}

/* SVDAllFreqs.c (weights off): V_out [N x k*Nw_eff] left singular vectors of Y, S_out [k x Nw_eff] (real parts) */
EXPORT int ref_svd_all_freqs(PetscInt N, PetscInt k, PetscInt Nw_eff, const PetscScalar *Y, PetscScalar *V_out, double *S_out)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    RSVD_matrices M;
    Weight_matrices W;
    Resolvent_matrices Res;
    memset(&R, 0, sizeof R);
    memset(&M, 0, sizeof M);
    memset(&W, 0, sizeof W);
    memset(&Res, 0, sizeof Res);
    R.RSVD.N = R.RSVD.Nc = R.RSVD.Nb = N;
    R.RSVD.k = k;
    R.RSVD.Nw_eff = Nw_eff;
    ShimSetQuiet(1);
    ierr = ShimMatCreateDenseFrom(N, k * Nw_eff, Y, &M.Y_hat);CHKERRQ(ierr);
    ierr = SVDAllFreqs(&M, &R, &W, &Res);CHKERRQ(ierr);
    copy_out(Res.V_hat, V_out);
    PetscInt lda;
    const PetscScalar *s = ShimMatDenseArray(Res.S_hat, &lda);
    for (PetscInt iw = 0; iw < Nw_eff; iw++)
        for (PetscInt ik = 0; ik < k; ik++) S_out[ik + k * iw] = creal(s[ik + lda * iw]);
    MatDestroy(&Res.V_hat);
    MatDestroy(&Res.S_hat);
    return 0;
}

/* TSExpADeltaT.c: U [N x ncols] <- RK4^{ResRatio}(U), column by column */
EXPORT int ref_ts_exp_a_deltat(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, PetscInt ncols, PetscScalar *U)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    LNS_vars L;
    Mat Um;
    ierr = make_ctx(v, rowptr, col, val, &R, &L, NULL);CHKERRQ(ierr);
    ierr = ShimMatCreateDenseFrom(v->N, ncols, U, &Um);CHKERRQ(ierr);
    ierr = TSExpADeltaT(&L, &R, Um);CHKERRQ(ierr);
    copy_out(Um, U);
    MatDestroy(&Um);
    MatDestroy(&L.A);
    return 0;
}

/* PermuteMat.c / ReversePermuteMat.c in place on an N x (k*Nw) matrix */
EXPORT int ref_permute_mat(PetscInt N, PetscInt k, PetscInt Nw, int reverse, PetscScalar *F)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    Mat A;
    memset(&R, 0, sizeof R);
    R.RSVD.k = k;
    ierr = ShimMatCreateDenseFrom(N, k * Nw, F, &A);CHKERRQ(ierr);
    ierr = reverse ? ReversePermuteMat(A, &R) : PermuteMat(A, &R);CHKERRQ(ierr);
    copy_out(A, F);
    MatDestroy(&A);
    return 0;
}

/* Unforced stepping of one vector with TSTransRK4.c (what TransientRunRK4.c's loop does between its norm checks) */
EXPORT int ref_trans_steps(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, PetscInt nsteps, PetscScalar *q)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    LNS_vars L;
    TS_matrices TS;
    Vec y;
    ierr = make_ctx(v, rowptr, col, val, &R, &L, NULL);CHKERRQ(ierr);
    ierr = CreateTSMats(&TS, v->N);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(v->N, q, &y);CHKERRQ(ierr);
    for (PetscInt i = 1; i <= nsteps; i++) {
        ierr = TSTransRK4(&L, &TS, &R, i, y);CHKERRQ(ierr);
    }
    memcpy(q, ShimVecArray(y), (size_t)v->N * sizeof(PetscScalar));
    VecDestroy(&y);
    DestroyTSMats(&TS);
    MatDestroy(&L.A);
    return 0;
}

/* Times `nsteps` calls of the reference's forced RK4 step (TSRK4.c: 4 MatMultAdd + 2 CreateForcingOnFly + 7 vector
 * operations) on `ncols` test vectors one after the other, as TSActionRK4.c:89-122 runs them, after `warm` untimed
 * steps.  F_hat_k is N x Nw_eff (one test vector's forcing; reused for every column -- the cost does not depend on
 * the values).  Returns the wall time in seconds through *seconds; y_out (N, optional) receives the last state. */
EXPORT int ref_time_forced_steps(const ref_vars *v, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, const PetscScalar *F_hat_k,
                                 PetscInt ncols, PetscInt warm, PetscInt nsteps, double *seconds, PetscScalar *y_out)
{
    PetscErrorCode ierr;
    RSVDt_vars R;
    LNS_vars L;
    DFT_matrices D;
    TS_matrices TS;
    Mat F;
    Vec y;
    PetscLogDouble t0, t1, total = 0;
    ShimSetQuiet(1);
    ierr = make_ctx(v, rowptr, col, val, &R, &L, &D);CHKERRQ(ierr);
    ierr = CreateTSMats(&TS, v->N);CHKERRQ(ierr);
    ierr = ShimMatCreateDenseFrom(v->N, v->Nw_eff, F_hat_k, &F);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(v->N, NULL, &y);CHKERRQ(ierr);
    for (PetscInt ik = 0; ik < ncols; ik++) {
        ierr = VecZeroEntries(y);CHKERRQ(ierr);
        ierr = CreateForcingOnFly(F, &D, 0, TS.F3);CHKERRQ(ierr);
        for (PetscInt i = 1; i <= warm; i++) {
            ierr = TSRK4(F, &D, &L, &TS, &R, i, y);CHKERRQ(ierr);
        }
        PetscTime(&t0);
        for (PetscInt i = warm + 1; i <= warm + nsteps; i++) {
            ierr = TSRK4(F, &D, &L, &TS, &R, i, y);CHKERRQ(ierr);
        }
        PetscTime(&t1);
        total += t1 - t0;
    }
    *seconds = total;
    if (y_out) memcpy(y_out, ShimVecArray(y), (size_t)v->N * sizeof(PetscScalar));
    VecDestroy(&y);
    MatDestroy(&F);
    DestroyTSMats(&TS);
    MatDestroy(&D.dft);
    MatDestroy(&D.idft);
    MatDestroy(&L.A);
    return 0;
}

EXPORT int ref_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------------------
 * Self-test entry points for the shim's primitives (tests/test_petsc_shim.py checks them against numpy / LAPACK and
 * against the independent PETSc-binary reader/writer rsvd-delta-t_b200/petsc_io.py).  Not reference code.
 * ------------------------------------------------------------------------------------------------------------------ */
/* mode 0: y = A x, 1: y = add + A x, 2: y = A^H x, 3: y = add + A^H x; dense != 0 converts A to MATDENSE first */
EXPORT int shim_mat_mult(PetscInt m, PetscInt n, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, int dense, int mode,
                         const PetscScalar *x, const PetscScalar *add, PetscScalar *y)
{
    PetscErrorCode ierr;
    Mat A;
    Vec vx, va = NULL, vy;
    const PetscInt nx = (mode < 2) ? n : m, ny = (mode < 2) ? m : n;
    if (dense) {
        ierr = ShimMatCreateDenseFrom(m, n, NULL, &A);CHKERRQ(ierr);
        PetscInt lda;
        PetscScalar *a = ShimMatDenseArray(A, &lda);
        for (PetscInt r = 0; r < m; r++)
            for (PetscInt p = rowptr[r]; p < rowptr[r + 1]; p++) a[r + lda * col[p]] += val[p];
    } else {
        ierr = ShimMatCreateAIJFrom(m, n, rowptr, col, val, &A);CHKERRQ(ierr);
    }
    ierr = ShimVecCreateFrom(nx, x, &vx);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(ny, NULL, &vy);CHKERRQ(ierr);
    if (add) { ierr = ShimVecCreateFrom(ny, add, &va);CHKERRQ(ierr); }
    switch (mode) {
    case 0: ierr = MatMult(A, vx, vy); break;
    case 1: ierr = MatMultAdd(A, vx, va, vy); break;
    case 2: ierr = MatMultHermitianTranspose(A, vx, vy); break;
    default: ierr = MatMultHermitianTransposeAdd(A, vx, va, vy); break;
    }
    CHKERRQ(ierr);
    memcpy(y, ShimVecArray(vy), (size_t)ny * sizeof(PetscScalar));
    VecDestroy(&vx); VecDestroy(&vy); VecDestroy(&va); MatDestroy(&A);
    return 0;
}

/* op 0: C = A*B (A AIJ m x n, B dense n x c), 1: same with MAT_REUSE_MATRIX into a submatrix view of a wider matrix,
 * 2: C = A^T * B with A dense (MatTransposeMatMult), 3: in-place MatHermitianTranspose of the AIJ matrix then C = A^H * B */
EXPORT int shim_mat_mat(int op, PetscInt m, PetscInt n, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, PetscInt c,
                        const PetscScalar *B, PetscScalar *Cout)
{
    PetscErrorCode ierr;
    Mat A, Bm, Cm = NULL, Wide = NULL, Sub = NULL;
    ierr = ShimMatCreateAIJFrom(m, n, rowptr, col, val, &A);CHKERRQ(ierr);
    if (op == 0 || op == 1) {
        ierr = ShimMatCreateDenseFrom(n, c, B, &Bm);CHKERRQ(ierr);
        if (op == 0) {
            ierr = MatMatMult(A, Bm, MAT_INITIAL_MATRIX, PETSC_DEFAULT, &Cm);CHKERRQ(ierr);
            copy_out(Cm, Cout);
        } else {
            ierr = ShimMatCreateDenseFrom(m, c + 3, NULL, &Wide);CHKERRQ(ierr);
            ierr = MatDenseGetSubMatrix(Wide, PETSC_DECIDE, PETSC_DECIDE, 2, 2 + c, &Sub);CHKERRQ(ierr);
            ierr = MatMatMult(A, Bm, MAT_REUSE_MATRIX, PETSC_DEFAULT, &Sub);CHKERRQ(ierr);
            copy_out(Sub, Cout);
            ierr = MatDenseRestoreSubMatrix(Wide, &Sub);CHKERRQ(ierr);
        }
    } else if (op == 2) {
        Mat Ad;
        ierr = ShimMatCreateDenseFrom(m, n, NULL, &Ad);CHKERRQ(ierr);
        PetscInt lda;
        PetscScalar *a = ShimMatDenseArray(Ad, &lda);
        for (PetscInt r = 0; r < m; r++)
            for (PetscInt p = rowptr[r]; p < rowptr[r + 1]; p++) a[r + lda * col[p]] += val[p];
        ierr = ShimMatCreateDenseFrom(m, c, B, &Bm);CHKERRQ(ierr);
        ierr = MatTransposeMatMult(Ad, Bm, MAT_INITIAL_MATRIX, PETSC_DEFAULT, &Cm);CHKERRQ(ierr);
        copy_out(Cm, Cout);
        MatDestroy(&Ad);
    } else {
        ierr = ShimMatCreateDenseFrom(m, c, B, &Bm);CHKERRQ(ierr);
        ierr = MatHermitianTranspose(A, MAT_INPLACE_MATRIX, &A);CHKERRQ(ierr);
        ierr = MatMatMult(A, Bm, MAT_INITIAL_MATRIX, PETSC_DEFAULT, &Cm);CHKERRQ(ierr);
        copy_out(Cm, Cout);
    }
    MatDestroy(&A); MatDestroy(&Bm); MatDestroy(&Cm); MatDestroy(&Wide);
    return 0;
}

/* KSPSolve on a dense n x n system, in place on b like GalerkinProj.c:61 */
EXPORT int shim_ksp_solve(PetscInt n, const PetscScalar *A, PetscScalar *b)
{
    PetscErrorCode ierr;
    Mat Am;
    Vec vb;
    KSP ksp;
    ierr = ShimMatCreateDenseFrom(n, n, A, &Am);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(n, b, &vb);CHKERRQ(ierr);
    ierr = KSPCreate(PETSC_COMM_WORLD, &ksp);CHKERRQ(ierr);
    ierr = KSPSetOperators(ksp, Am, Am);CHKERRQ(ierr);
    ierr = KSPSetType(ksp, KSPPGMRES);CHKERRQ(ierr);
    ierr = KSPSetTolerances(ksp, 1.e-15, PETSC_DEFAULT, PETSC_DEFAULT, PETSC_DEFAULT);CHKERRQ(ierr);
    ierr = KSPSolve(ksp, vb, vb);CHKERRQ(ierr);
    memcpy(b, ShimVecArray(vb), (size_t)n * sizeof(PetscScalar));
    KSPDestroy(&ksp); VecDestroy(&vb); MatDestroy(&Am);
    return 0;
}

/* w = alpha x + y (VecWAXPY), then y += beta w (VecAXPY), then w = x + gamma w (VecAYPX); norm2 of w returned */
EXPORT int shim_vec_ops(PetscInt n, const PetscScalar *x, PetscScalar *y, PetscScalar *w, double ar, double ai, double br, double bi, double gr, double gi, double *nrm)
{
    PetscErrorCode ierr;
    Vec vx, vy, vw;
    ierr = ShimVecCreateFrom(n, x, &vx);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(n, y, &vy);CHKERRQ(ierr);
    ierr = ShimVecCreateFrom(n, NULL, &vw);CHKERRQ(ierr);
    ierr = VecWAXPY(vw, ar + ai * PETSC_i, vx, vy);CHKERRQ(ierr);
    ierr = VecAXPY(vy, br + bi * PETSC_i, vw);CHKERRQ(ierr);
    ierr = VecAYPX(vw, gr + gi * PETSC_i, vx);CHKERRQ(ierr);
    ierr = VecNorm(vw, NORM_2, nrm);CHKERRQ(ierr);
    memcpy(y, ShimVecArray(vy), (size_t)n * sizeof(PetscScalar));
    memcpy(w, ShimVecArray(vw), (size_t)n * sizeof(PetscScalar));
    VecDestroy(&vx); VecDestroy(&vy); VecDestroy(&vw);
    return 0;
}

/* MatLoad (as MATMPIAIJ or MATDENSE) + optional MatShift + MatView, and VecLoad + VecView: file round trips */
EXPORT int shim_mat_roundtrip(const char *in, const char *out, int as_dense, double shift)
{
    PetscErrorCode ierr;
    PetscViewer fi, fo;
    Mat A;
    ierr = PetscViewerBinaryOpen(PETSC_COMM_WORLD, in, FILE_MODE_READ, &fi);CHKERRQ(ierr);
    ierr = MatCreate(PETSC_COMM_WORLD, &A);CHKERRQ(ierr);
    ierr = MatSetType(A, as_dense ? MATDENSE : MATMPIAIJ);CHKERRQ(ierr);
    ierr = MatLoad(A, fi);CHKERRQ(ierr);
    if (shift != 0.0) { ierr = MatShift(A, shift);CHKERRQ(ierr); }
    ierr = PetscViewerBinaryOpen(PETSC_COMM_WORLD, out, FILE_MODE_WRITE, &fo);CHKERRQ(ierr);
    ierr = MatView(A, fo);CHKERRQ(ierr);
    PetscViewerDestroy(&fi); PetscViewerDestroy(&fo); MatDestroy(&A);
    return 0;
}

EXPORT int shim_vec_roundtrip(const char *in, const char *out)
{
    PetscErrorCode ierr;
    PetscViewer fi, fo;
    Vec v;
    ierr = PetscViewerBinaryOpen(PETSC_COMM_WORLD, in, FILE_MODE_READ, &fi);CHKERRQ(ierr);
    ierr = VecCreate(PETSC_COMM_WORLD, &v);CHKERRQ(ierr);
    ierr = VecLoad(v, fi);CHKERRQ(ierr);
    ierr = PetscViewerBinaryOpen(PETSC_COMM_WORLD, out, FILE_MODE_WRITE, &fo);CHKERRQ(ierr);
    ierr = VecView(v, fo);CHKERRQ(ierr);
    PetscViewerDestroy(&fi); PetscViewerDestroy(&fo); VecDestroy(&v);
    return 0;
}

/* options database: insert a YAML file, then look one key up as string / bool / int / real (type 0..3).
 * returns 0 and *set; value as text in buf */
EXPORT int shim_option_lookup(const char *yaml, const char *key, int type, int *set, char *buf, PetscInt buflen)
{
    PetscErrorCode ierr;
    PetscBool flg = PETSC_FALSE;
    ierr = PetscOptionsClear(NULL);CHKERRQ(ierr);
    ierr = PetscOptionsInsertFileYAML(PETSC_COMM_WORLD, NULL, yaml, PETSC_TRUE);CHKERRQ(ierr);
    buf[0] = 0;
    if (type == 0) {
        ierr = PetscOptionsGetString(NULL, NULL, key, buf, (size_t)buflen, &flg);CHKERRQ(ierr);
    } else if (type == 1) {
        PetscBool v = PETSC_FALSE;
        ierr = PetscOptionsGetBool(NULL, NULL, key, &v, &flg);CHKERRQ(ierr);
        snprintf(buf, (size_t)buflen, "%d", (int)v);
    } else if (type == 2) {
        PetscInt v = 0;
        ierr = PetscOptionsGetInt(NULL, NULL, key, &v, &flg);CHKERRQ(ierr);
        snprintf(buf, (size_t)buflen, "%lld", (long long)v);
    } else {
        PetscReal v = 0;
        ierr = PetscOptionsGetReal(NULL, NULL, key, &v, &flg);CHKERRQ(ierr);
        snprintf(buf, (size_t)buflen, "%.17g", v);
    }
    *set = (int)flg;
    ierr = PetscOptionsClear(NULL);CHKERRQ(ierr);
    return 0;
}
