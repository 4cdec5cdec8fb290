/* petsc_binary.c -- see petsc_binary.h */
#include "petsc_binary.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[512] = "";
const char *pb_error(void) { return g_err; }
static int pb_fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}

static uint32_t bs32(uint32_t v) { return __builtin_bswap32(v); }
static uint64_t bs64(uint64_t v) { return __builtin_bswap64(v); }

static int read_ints(FILE *f, int isz, int64_t n, int64_t *out)
{
    if (isz == 8) {
        if (fread(out, 8, (size_t)n, f) != (size_t)n) return 1;
        for (int64_t i = 0; i < n; i++) out[i] = (int64_t)bs64((uint64_t)out[i]);
    } else {
        int32_t *tmp = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * 4);
        if (!tmp) return 1;
        if (fread(tmp, 4, (size_t)n, f) != (size_t)n) { free(tmp); return 1; }
        for (int64_t i = 0; i < n; i++) out[i] = (int32_t)bs32((uint32_t)tmp[i]);
        free(tmp);
    }
    return 0;
}

static int read_doubles(FILE *f, int64_t n, double *out)
{
    if (fread(out, 8, (size_t)n, f) != (size_t)n) return 1;
    uint64_t *u = (uint64_t *)out;
    for (int64_t i = 0; i < n; i++) u[i] = bs64(u[i]);
    return 0;
}

/* detect the integer width from the class id at the start of the record */
static int detect_isz(FILE *f, int64_t classid)
{
    unsigned char b[8];
    if (fread(b, 1, 8, f) != 8) return 0;
    rewind(f);
    uint32_t v32;
    uint64_t v64;
    memcpy(&v32, b, 4);
    memcpy(&v64, b, 8);
    if ((int64_t)(int32_t)bs32(v32) == classid) return 4;
    if ((int64_t)bs64(v64) == classid) return 8;
    return 0;
}

static int64_t file_size(FILE *f)
{
    long cur = ftell(f);
    fseek(f, 0, SEEK_END);
    long e = ftell(f);
    fseek(f, cur, SEEK_SET);
    return (int64_t)e;
}

int pb_read_mat(const char *path, pb_csr *out)
{
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "rb");
    if (!f) return pb_fail("cannot open %s", path);
    const int isz = detect_isz(f, MAT_FILE_CLASSID);
    if (!isz) { fclose(f); return pb_fail("%s is not a PETSc binary Mat", path); }
    int64_t hdr[4];
    if (read_ints(f, isz, 4, hdr)) { fclose(f); return pb_fail("%s: short header", path); }
    const int64_t M = hdr[1], N = hdr[2];
    int64_t nnz = hdr[3];
    const int native_dense = (nnz == -1);
    if (native_dense) nnz = M * N;
    if (M < 0 || N < 0 || nnz < 0) { fclose(f); return pb_fail("%s: bad sizes", path); }
    out->nrows = M; out->ncols = N; out->nnz = nnz;
    out->rowptr = (int64_t *)malloc((size_t)(M + 1) * 8);
    out->col = (int64_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * 8);
    out->val = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * 16);
    if (!out->rowptr || !out->col || !out->val) { fclose(f); pb_free_csr(out); return pb_fail("out of memory"); }
    if (native_dense) {
        /* PETSc >= 3.14 (MatView_Dense_Binary / MatLoad_Dense_Binary; the reference targets 3.19): the M*N scalars of the
         * dense-native record are stored ROW by row, val[r*N + c]; convert to a full CSR */
        const int64_t rest = file_size(f) - ftell(f);
        const int cplx = rest >= M * N * 16;
        double *tmp = (double *)malloc((size_t)M * N * (cplx ? 16 : 8));
        if (!tmp || read_doubles(f, M * N * (cplx ? 2 : 1), tmp)) { free(tmp); fclose(f); pb_free_csr(out); return pb_fail("%s: short data", path); }
        for (int64_t r = 0; r <= M; r++) out->rowptr[r] = r * N;
        for (int64_t r = 0; r < M; r++)
            for (int64_t c = 0; c < N; c++) {
                out->col[r * N + c] = c;
                out->val[2 * (r * N + c)] = cplx ? tmp[2 * (r * N + c)] : tmp[r * N + c];
                out->val[2 * (r * N + c) + 1] = cplx ? tmp[2 * (r * N + c) + 1] : 0.0;
            }
        free(tmp);
        fclose(f);
        return 0;
    }
    int64_t *rl = (int64_t *)malloc((size_t)(M > 0 ? M : 1) * 8);
    if (!rl || read_ints(f, isz, M, rl) || read_ints(f, isz, nnz, out->col)) { free(rl); fclose(f); pb_free_csr(out); return pb_fail("%s: short index data", path); }
    out->rowptr[0] = 0;
    for (int64_t r = 0; r < M; r++) out->rowptr[r + 1] = out->rowptr[r] + rl[r];
    free(rl);
    if (out->rowptr[M] != nnz) { fclose(f); pb_free_csr(out); return pb_fail("%s: row lengths do not sum to nnz", path); }
    const int64_t rest = file_size(f) - ftell(f);
    if (rest >= nnz * 16) {
        if (read_doubles(f, 2 * nnz, out->val)) { fclose(f); pb_free_csr(out); return pb_fail("%s: short values", path); }
    } else if (rest >= nnz * 8) { /* real-scalar file */
        double *tmp = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * 8);
        if (!tmp || read_doubles(f, nnz, tmp)) { free(tmp); fclose(f); pb_free_csr(out); return pb_fail("%s: short values", path); }
        for (int64_t p = 0; p < nnz; p++) { out->val[2 * p] = tmp[p]; out->val[2 * p + 1] = 0.0; }
        free(tmp);
    } else { fclose(f); pb_free_csr(out); return pb_fail("%s: short values", path); }
    fclose(f); /* trailing bytes (e.g. MATLAB PetscBinaryWrite option records after A_GL) are ignored, like MatLoad */
    return 0;
}

void pb_free_csr(pb_csr *m)
{
    free(m->rowptr); free(m->col); free(m->val);
    memset(m, 0, sizeof *m);
}

int pb_read_mat_dense(const char *path, int64_t *nrows, int64_t *ncols, double **colmajor)
{
    pb_csr m;
    if (pb_read_mat(path, &m)) return 1;
    double *d = (double *)calloc((size_t)(m.nrows * m.ncols > 0 ? m.nrows * m.ncols : 1), 16);
    if (!d) { pb_free_csr(&m); return pb_fail("out of memory"); }
    for (int64_t r = 0; r < m.nrows; r++)
        for (int64_t p = m.rowptr[r]; p < m.rowptr[r + 1]; p++) {
            const int64_t c = m.col[p];
            if (c < 0 || c >= m.ncols) { free(d); pb_free_csr(&m); return pb_fail("%s: column out of range", path); }
            d[2 * (c * m.nrows + r)] = m.val[2 * p];
            d[2 * (c * m.nrows + r) + 1] = m.val[2 * p + 1];
        }
    *nrows = m.nrows; *ncols = m.ncols; *colmajor = d;
    pb_free_csr(&m);
    return 0;
}

static int write_i64(FILE *f, int64_t v)
{
    uint64_t u = bs64((uint64_t)v);
    return fwrite(&u, 8, 1, f) != 1;
}

int pb_write_mat_dense(const char *path, int64_t M, int64_t N, const double *a, int64_t ld)
{
    FILE *f = fopen(path, "wb");
    if (!f) return pb_fail("cannot create %s", path);
    int bad = write_i64(f, MAT_FILE_CLASSID) | write_i64(f, M) | write_i64(f, N) | write_i64(f, M * N);
    for (int64_t r = 0; r < M && !bad; r++) bad |= write_i64(f, N);
    uint64_t *buf = (uint64_t *)malloc((size_t)(N > 0 ? N : 1) * 16);
    if (!buf) { fclose(f); return pb_fail("out of memory"); }
    for (int64_t c = 0; c < N; c++) buf[c] = bs64((uint64_t)c);
    for (int64_t r = 0; r < M && !bad; r++) bad |= fwrite(buf, 8, (size_t)N, f) != (size_t)N;
    for (int64_t r = 0; r < M && !bad; r++) { /* values row by row */
        for (int64_t c = 0; c < N; c++) {
            uint64_t re, im;
            memcpy(&re, &a[2 * (c * ld + r)], 8);
            memcpy(&im, &a[2 * (c * ld + r) + 1], 8);
            buf[2 * c] = bs64(re);
            buf[2 * c + 1] = bs64(im);
        }
        bad |= fwrite(buf, 16, (size_t)N, f) != (size_t)N;
    }
    free(buf);
    if (fclose(f)) bad = 1;
    return bad ? pb_fail("write error on %s", path) : 0;
}

int pb_read_vec(const char *path, int64_t *n, double **v)
{
    FILE *f = fopen(path, "rb");
    if (!f) return pb_fail("cannot open %s", path);
    const int isz = detect_isz(f, VEC_FILE_CLASSID);
    if (!isz) { fclose(f); return pb_fail("%s is not a PETSc binary Vec", path); }
    int64_t hdr[2];
    if (read_ints(f, isz, 2, hdr) || hdr[1] < 0) { fclose(f); return pb_fail("%s: short header", path); }
    double *d = (double *)malloc((size_t)(hdr[1] > 0 ? hdr[1] : 1) * 16);
    if (!d) { fclose(f); return pb_fail("out of memory"); }
    const int64_t rest = file_size(f) - ftell(f);
    if (rest >= hdr[1] * 16) {
        if (read_doubles(f, 2 * hdr[1], d)) { free(d); fclose(f); return pb_fail("%s: short data", path); }
    } else {
        double *tmp = (double *)malloc((size_t)(hdr[1] > 0 ? hdr[1] : 1) * 8);
        if (!tmp || read_doubles(f, hdr[1], tmp)) { free(tmp); free(d); fclose(f); return pb_fail("%s: short data", path); }
        for (int64_t i = 0; i < hdr[1]; i++) { d[2 * i] = tmp[i]; d[2 * i + 1] = 0.0; }
        free(tmp);
    }
    fclose(f);
    *n = hdr[1];
    *v = d;
    return 0;
}

int pb_write_vec(const char *path, int64_t n, const double *v)
{
    FILE *f = fopen(path, "wb");
    if (!f) return pb_fail("cannot create %s", path);
    int bad = write_i64(f, VEC_FILE_CLASSID) | write_i64(f, n);
    for (int64_t i = 0; i < 2 * n && !bad; i++) {
        uint64_t u;
        memcpy(&u, &v[i], 8);
        u = bs64(u);
        bad |= fwrite(&u, 8, 1, f) != 1;
    }
    if (fclose(f)) bad = 1;
    return bad ? pb_fail("write error on %s", path) : 0;
}
