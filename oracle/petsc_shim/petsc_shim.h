/* petsc_shim.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A one-rank, PETSc-free restatement of exactly the subset of the PETSc 3.19 / SLEPc 3.19 C API that the reference's
 * SourceCode/*.c calls (110 entry points, listed by `grep` in oracle/README.md), so that the reference's OWN source
 * files compile unmodified from /root/reference into oracle/_ref/ (the recipe is oracle/Makefile.ref) and can be run
 * here to pin the oracle and to generate golden vectors.  What is restated is the published definition of each call
 * (MatMultAdd: z = y + A x, VecWAXPY: w = a x + y, BVOrthogonalize: thin QR with diag(R) > 0, SVDSolve: thin SVD by
 * the cross-product method, KSPSolve: the solution of the small dense system, MatLoad/MatView/VecLoad/VecView: the
 * PETSc binary record formats); everything above those calls -- loop structure, index arithmetic, twiddle tables,
 * permutations, snapshot positions, the transient-removal algebra -- is the reference's code itself.
 *
 * Scalars are complex double, PetscInt is 64-bit (the reference's recommended configure, README.md:123-125),
 * one process stands for PETSC_COMM_WORLD.  Nothing in the product (rsvd-delta-t_b200/) includes or links this. */
#ifndef PETSC_SHIM_H
#define PETSC_SHIM_H

#include <complex.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int PetscErrorCode;
typedef int64_t PetscInt;
typedef double PetscReal;
typedef double _Complex PetscScalar;
typedef double PetscLogDouble;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
typedef int MPI_Comm;
typedef int PetscMPIInt;

typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef struct _p_BV *BV;
typedef struct _p_SVD *SVD;
typedef struct _p_KSP *KSP;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscRandom *PetscRandom;
typedef void *PetscOptions;

typedef const char *MatType;
typedef const char *VecType;
typedef const char *BVType;
typedef const char *KSPType;

#define MATMPIAIJ "mpiaij"
#define MATSEQAIJ "seqaij"
#define MATAIJ "aij"
#define MATDENSE "dense"
#define MATMPIDENSE "mpidense"
#define MATSEQDENSE "seqdense"
#define VECMPI "mpi"
#define VECSEQ "seq"
#define VECSTANDARD "standard"
#define BVVECS "vecs"
#define BVMAT "mat"
#define BVSVEC "svec"
#define KSPPGMRES "pgmres"
#define KSPGMRES "gmres"

#define PETSC_MAX_PATH_LEN 4096
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#define MPI_COMM_SELF 1
#define MPI_COMM_WORLD 0
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_PI 3.1415926535897932384626433832795029
#define PETSC_i _Complex_I
#define PETSC_ERR_USER 83
#define PETSC_ERR_FILE_OPEN 65
#define PETSC_ERR_ARG_SIZ 60
#define PETSC_ERR_SUP 56

typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_DO_NOT_COPY_VALUES, MAT_COPY_VALUES, MAT_SHARE_NONZERO_PATTERN } MatDuplicateOption;
typedef enum { DIFFERENT_NONZERO_PATTERN, SUBSET_NONZERO_PATTERN, SAME_NONZERO_PATTERN } MatStructure;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX, MAT_IGNORE_MATRIX, MAT_INPLACE_MATRIX } MatReuse;
typedef enum { FILE_MODE_UNDEFINED = -1, FILE_MODE_READ = 0, FILE_MODE_WRITE, FILE_MODE_APPEND } PetscFileMode;

/* error handling: the reference uses the pre-3.18 `ierr = ...; CHKERRQ(ierr);` idiom throughout */
PetscErrorCode PetscShimError(const char *file, int line, const char *func, PetscErrorCode code, const char *fmt, ...);
#define CHKERRQ(ierr) \
    do { \
        PetscErrorCode ierr__ = (ierr); \
        if (ierr__) return PetscShimError(__FILE__, __LINE__, __func__, ierr__, NULL); \
    } while (0)
#define CHKERRMPI(ierr) CHKERRQ(ierr)
#define SETERRQ(comm, code, ...) return PetscShimError(__FILE__, __LINE__, __func__, (code), __VA_ARGS__)
#define PetscFunctionBeginUser \
    do { \
    } while (0)
#define PetscFunctionBegin PetscFunctionBeginUser
#define PetscFunctionReturn(x) return (x)

#define PetscFmodReal(a, b) fmod((double)(a), (double)(b))
#define PetscCeilReal(a) ceil(a)
#define PetscFloorReal(a) floor(a)
#define PetscExpComplex(a) cexp(a)
#define PetscIsNanReal(a) isnan(a)
#define PetscRealPart(a) creal(a)
#define PetscImaginaryPart(a) cimag(a)

/* sys */
PetscErrorCode SlepcInitialize(int *argc, char ***args, const char file[], const char help[]);
PetscErrorCode SlepcFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char format[], ...);
PetscErrorCode PetscSNPrintf(char *str, size_t len, const char format[], ...);
PetscErrorCode PetscTime(PetscLogDouble *t);
PetscErrorCode PetscGetWorkingDirectory(char path[], size_t len);
PetscErrorCode PetscFOpen(MPI_Comm comm, const char name[], const char mode[], FILE **fp);
PetscErrorCode PetscFClose(MPI_Comm comm, FILE *fp);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);

/* options database */
PetscErrorCode PetscOptionsInsertFileYAML(MPI_Comm comm, PetscOptions options, const char file[], PetscBool require);
PetscErrorCode PetscOptionsSetValue(PetscOptions options, const char name[], const char value[]);
PetscErrorCode PetscOptionsGetString(PetscOptions options, const char pre[], const char name[], char string[], size_t len, PetscBool *set);
PetscErrorCode PetscOptionsGetBool(PetscOptions options, const char pre[], const char name[], PetscBool *ivalue, PetscBool *set);
PetscErrorCode PetscOptionsGetInt(PetscOptions options, const char pre[], const char name[], PetscInt *ivalue, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(PetscOptions options, const char pre[], const char name[], PetscReal *dvalue, PetscBool *set);
PetscErrorCode PetscOptionsClear(PetscOptions options);

/* random */
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r);
PetscErrorCode PetscRandomSetSeed(PetscRandom r, unsigned long seed);
PetscErrorCode PetscRandomSeed(PetscRandom r);
PetscErrorCode PetscRandomDestroy(PetscRandom *r);

/* viewers */
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char name[], PetscFileMode mode, PetscViewer *viewer);
PetscErrorCode PetscViewerDestroy(PetscViewer *viewer);

/* Vec */
PetscErrorCode VecCreate(MPI_Comm comm, Vec *v);
PetscErrorCode VecSetType(Vec v, VecType type);
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N);
PetscErrorCode VecSetUp(Vec v);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *low, PetscInt *high);
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode iora);
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar va, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecSet(Vec v, PetscScalar alpha);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecScale(Vec x, PetscScalar alpha);
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecExp(Vec v);
PetscErrorCode VecConjugate(Vec v);
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val);
PetscErrorCode VecNormalize(Vec x, PetscReal *val);
PetscErrorCode VecSetRandom(Vec x, PetscRandom r);
PetscErrorCode VecLoad(Vec v, PetscViewer viewer);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode VecGetSize(Vec v, PetscInt *n);

/* Mat */
PetscErrorCode MatCreate(MPI_Comm comm, Mat *A);
PetscErrorCode MatSetType(Mat A, MatType type);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetUp(Mat A);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType type);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type);
PetscErrorCode MatLoad(Mat A, PetscViewer viewer);
PetscErrorCode MatView(Mat A, PetscViewer viewer);
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat *B);
PetscErrorCode MatCopy(Mat A, Mat B, MatStructure str);
PetscErrorCode MatScale(Mat A, PetscScalar a);
PetscErrorCode MatShift(Mat A, PetscScalar a);
PetscErrorCode MatConjugate(Mat A);
PetscErrorCode MatImaginaryPart(Mat A);
PetscErrorCode MatNorm(Mat A, NormType type, PetscReal *nrm);
PetscErrorCode MatSetRandom(Mat A, PetscRandom r);
PetscErrorCode MatTranspose(Mat A, MatReuse reuse, Mat *B);
PetscErrorCode MatHermitianTranspose(Mat A, MatReuse reuse, Mat *B);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat A, Vec v1, Vec v2, Vec v3);
PetscErrorCode MatMultHermitianTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatMultHermitianTransposeAdd(Mat A, Vec v1, Vec v2, Vec v3);
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C);
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatDenseGetColumnVecRead(Mat A, PetscInt col, Vec *v);
PetscErrorCode MatDenseRestoreColumnVecRead(Mat A, PetscInt col, Vec *v);
PetscErrorCode MatDenseGetColumnVecWrite(Mat A, PetscInt col, Vec *v);
PetscErrorCode MatDenseRestoreColumnVecWrite(Mat A, PetscInt col, Vec *v);
PetscErrorCode MatDenseGetSubMatrix(Mat A, PetscInt rbegin, PetscInt rend, PetscInt cbegin, PetscInt cend, Mat *v);
PetscErrorCode MatDenseRestoreSubMatrix(Mat A, Mat *v);

/* KSP */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat Amat, Mat Pmat);
PetscErrorCode KSPSetType(KSP ksp, KSPType type);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPDestroy(KSP *ksp);

/* SLEPc BV */
PetscErrorCode BVCreateFromMat(Mat A, BV *bv);
PetscErrorCode BVSetType(BV bv, BVType type);
PetscErrorCode BVOrthogonalize(BV V, Mat R);
PetscErrorCode BVCreateMat(BV bv, Mat *A);
PetscErrorCode BVDestroy(BV *bv);

/* SLEPc SVD */
PetscErrorCode SVDCreate(MPI_Comm comm, SVD *svd);
PetscErrorCode SVDSetOperators(SVD svd, Mat A, Mat B);
PetscErrorCode SVDSetDimensions(SVD svd, PetscInt nsv, PetscInt ncv, PetscInt mpd);
PetscErrorCode SVDSolve(SVD svd);
PetscErrorCode SVDGetSingularTriplet(SVD svd, PetscInt i, PetscReal *sigma, Vec u, Vec v);
PetscErrorCode SVDDestroy(SVD *svd);

/* ---- shim-only accessors used by oracle/ref_harness.c (not PETSc API) ---- */
PetscErrorCode ShimMatCreateDenseFrom(PetscInt m, PetscInt n, const PetscScalar *colmajor, Mat *A);
PetscErrorCode ShimMatCreateAIJFrom(PetscInt m, PetscInt n, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, Mat *A);
PetscScalar *ShimMatDenseArray(Mat A, PetscInt *lda);
PetscErrorCode ShimVecCreateFrom(PetscInt n, const PetscScalar *a, Vec *v);
PetscScalar *ShimVecArray(Vec v);
void ShimSetQuiet(int quiet);

#endif
