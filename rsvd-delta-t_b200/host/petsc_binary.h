/* petsc_binary.h -- PETSc-free readers/writers of the PETSc binary formats the reference uses
 * (MatLoad / MatView / VecLoad / VecView through PetscViewerBinaryOpen: LNSStructure.c:22-27, ReadWeightMats.c:26-99,
 * CreateRandomMat.c:23-31, SaveResults.c:32-84, TransientRunRK4.c:82-86,201-209).  Big-endian; 32- or 64-bit indices are
 * auto-detected on input from the class id, written as 64-bit (the reference's recommended build, README.md:123-125). */
#ifndef PETSC_BINARY_H
#define PETSC_BINARY_H
#include <stdint.h>

#define MAT_FILE_CLASSID 1211216
#define VEC_FILE_CLASSID 1211214

typedef struct {
    int64_t nrows, ncols, nnz;
    int64_t *rowptr; /* nrows+1 */
    int64_t *col;    /* nnz */
    double *val;     /* 2*nnz, (re, im) */
} pb_csr;

/* all return 0 on success; on failure a message is left in pb_error() */
int pb_read_mat(const char *path, pb_csr *out);
void pb_free_csr(pb_csr *m);
/* any Mat record -> dense column-major complex (2*nrows*ncols doubles, caller frees) */
int pb_read_mat_dense(const char *path, int64_t *nrows, int64_t *ncols, double **colmajor);
/* what MatView(dense) writes in the default binary format: an AIJ record holding all entries, row by row.
 * colmajor: nrows x ncols complex, leading dimension ld (in complex elements). */
int pb_write_mat_dense(const char *path, int64_t nrows, int64_t ncols, const double *colmajor, int64_t ld);
int pb_read_vec(const char *path, int64_t *n, double **v);
int pb_write_vec(const char *path, int64_t n, const double *v);
const char *pb_error(void);
#endif
