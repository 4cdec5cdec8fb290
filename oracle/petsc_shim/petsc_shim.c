/* petsc_shim.c -- TEST INFRASTRUCTURE ONLY.  One-rank restatement of the PETSc 3.19 / SLEPc 3.19 calls the reference
 * makes (see petsc_shim.h).  Each primitive follows the published definition of the PETSc/SLEPc call it stands for;
 * where the summation order of the sequential PETSc kernel is documented it is kept (MatMult_SeqAIJ: row sums in
 * storage order; MatMultAdd: the sum starts from the addend; MatMultHermitianTransposeAdd: a scatter-add over rows). */
#define _GNU_SOURCE
#include "petsc_shim.h"

#include <ctype.h>
#include <sys/time.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ objects */
enum { MT_UNSET = 0, MT_AIJ = 1, MT_DENSE = 2 };

struct _p_Vec {
    PetscInt n;
    PetscScalar *a;
    int owns;
};

struct _p_Mat {
    int type;
    PetscInt m, n;
    /* dense: column-major, leading dimension lda */
    PetscScalar *a;
    PetscInt lda;
    int owns;
    /* aij */
    PetscInt *ia, *ja;
    PetscScalar *va;
    /* outstanding column vector / submatrix handed out by MatDenseGet... */
    struct _p_Vec colvec;
    int colvec_out;
    Mat sub;
};

struct _p_BV {
    PetscInt n, k;
    PetscScalar *a;
};

struct _p_SVD {
    Mat A;
    PetscInt nsv, m, n;
    PetscReal *sigma;
    PetscScalar *U, *V;
};

struct _p_KSP {
    Mat A;
};

struct _p_PetscViewer {
    char path[PETSC_MAX_PATH_LEN];
    PetscFileMode mode;
    FILE *f;
};

struct _p_PetscRandom {
    uint64_t state;
    unsigned long seed;
};

static int shim_quiet = 0;
void ShimSetQuiet(int q) { shim_quiet = q; }

PetscErrorCode PetscShimError(const char *file, int line, const char *func, PetscErrorCode code, const char *fmt, ...)
{
    if (fmt) {
        va_list ap;
        va_start(ap, fmt);
        fprintf(stderr, "[0]PETSC ERROR: ");
        vfprintf(stderr, fmt, ap);
        fprintf(stderr, "\n");
        va_end(ap);
    }
    fprintf(stderr, "[0]PETSC ERROR: #%d %s() at %s:%d\n", code, func, file, line);
    return code ? code : PETSC_ERR_USER;
}
#define SHIM_FAIL(code, ...) return PetscShimError(__FILE__, __LINE__, __func__, (code), __VA_ARGS__)

static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "petsc_shim: out of memory (%zu x %zu)\n", n, sz);
        abort();
    }
    return p;
}

/* ------------------------------------------------------------------------------------------------ sys */
#define MAXOPT 256
static struct {
    char *name, *value;
} opts[MAXOPT];
static int nopts = 0;

PetscErrorCode PetscOptionsSetValue(PetscOptions o, const char name[], const char value[])
{
    (void)o;
    for (int i = 0; i < nopts; i++)
        if (!strcmp(opts[i].name, name)) {
            free(opts[i].value);
            opts[i].value = strdup(value ? value : "");
            return 0;
        }
    if (nopts == MAXOPT) SHIM_FAIL(PETSC_ERR_USER, "too many options");
    opts[nopts].name = strdup(name);
    opts[nopts].value = strdup(value ? value : "");
    nopts++;
    return 0;
}

PetscErrorCode PetscOptionsClear(PetscOptions o)
{
    (void)o;
    for (int i = 0; i < nopts; i++) {
        free(opts[i].name);
        free(opts[i].value);
    }
    nopts = 0;
    return 0;
}

static const char *opt_find(const char *name)
{
    for (int i = 0; i < nopts; i++)
        if (!strcmp(opts[i].name, name)) return opts[i].value;
    return NULL;
}

PetscErrorCode SlepcInitialize(int *argc, char ***args, const char file[], const char help[])
{
    (void)file;
    (void)help;
    /* command line: "-name value" pairs, a name followed by another option (or nothing) has an empty value */
    if (argc && args)
        for (int i = 1; i < *argc; i++) {
            const char *a = (*args)[i];
            if (a[0] == '-' && !(isdigit((unsigned char)a[1]) || a[1] == '.')) {
                const char *v = "";
                if (i + 1 < *argc) {
                    const char *b = (*args)[i + 1];
                    if (!(b[0] == '-' && !(isdigit((unsigned char)b[1]) || b[1] == '.'))) {
                        v = b;
                        i++;
                    }
                }
                PetscErrorCode e = PetscOptionsSetValue(NULL, a, v);
                if (e) return e;
            }
        }
    return 0;
}

PetscErrorCode SlepcFinalize(void) { return 0; }

PetscErrorCode PetscPrintf(MPI_Comm comm, const char format[], ...)
{
    (void)comm;
    if (shim_quiet) return 0;
    va_list ap;
    va_start(ap, format);
    vfprintf(stdout, format, ap);
    va_end(ap);
    fflush(stdout);
    return 0;
}

PetscErrorCode PetscSNPrintf(char *str, size_t len, const char format[], ...)
{
    /* the reference formats into a buffer that is also one of the arguments (dirs->IO_dir); go through a temporary */
    char *tmp = (char *)xcalloc(len + 1, 1);
    va_list ap;
    va_start(ap, format);
    vsnprintf(tmp, len, format, ap);
    va_end(ap);
    memcpy(str, tmp, len);
    str[len - 1] = 0;
    free(tmp);
    return 0;
}

PetscErrorCode PetscTime(PetscLogDouble *t)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    *t = (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
    return 0;
}

PetscErrorCode PetscGetWorkingDirectory(char path[], size_t len)
{
    if (!getcwd(path, len)) SHIM_FAIL(PETSC_ERR_USER, "getcwd failed");
    return 0;
}

PetscErrorCode PetscFOpen(MPI_Comm comm, const char name[], const char mode[], FILE **fp)
{
    (void)comm;
    *fp = fopen(name, mode);
    if (!*fp) SHIM_FAIL(PETSC_ERR_FILE_OPEN, "Unable to open file %s", name);
    return 0;
}

PetscErrorCode PetscFClose(MPI_Comm comm, FILE *fp)
{
    (void)comm;
    if (fp) fclose(fp);
    return 0;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank)
{
    (void)comm;
    *rank = 0;
    return 0;
}
int MPI_Comm_size(MPI_Comm comm, int *size)
{
    (void)comm;
    *size = 1;
    return 0;
}

/* YAML subset PETSc's options reader accepts for a flat file: "key: value  # comment" per line -> option -key value */
PetscErrorCode PetscOptionsInsertFileYAML(MPI_Comm comm, PetscOptions o, const char file[], PetscBool require)
{
    (void)comm;
    (void)o;
    FILE *f = fopen(file, "r");
    if (!f) {
        if (require) SHIM_FAIL(PETSC_ERR_FILE_OPEN, "Unable to open YAML options file %s", file);
        /* PETSc errors as well when the file named by the user cannot be read */
        SHIM_FAIL(PETSC_ERR_FILE_OPEN, "Unable to open YAML options file %s", file);
    }
    char line[8192];
    while (fgets(line, sizeof line, f)) {
        char *s = line;
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s || *s == '#' || !strncmp(s, "---", 3) || !strncmp(s, "...", 3)) continue;
        char *colon = strchr(s, ':');
        if (!colon) continue;
        *colon = 0;
        char *key = s, *ke = colon - 1;
        while (ke >= key && isspace((unsigned char)*ke)) *ke-- = 0;
        char *v = colon + 1;
        while (*v && isspace((unsigned char)*v)) v++;
        /* strip a trailing comment: '#' at the start of the value or preceded by white space */
        for (char *p = v; *p; p++)
            if (*p == '#' && (p == v || isspace((unsigned char)p[-1]))) {
                *p = 0;
                break;
            }
        char *ve = v + strlen(v) - 1;
        while (ve >= v && isspace((unsigned char)*ve)) *ve-- = 0;
        size_t lv = strlen(v);
        if (lv >= 2 && ((v[0] == '"' && v[lv - 1] == '"') || (v[0] == '\'' && v[lv - 1] == '\''))) {
            v[lv - 1] = 0;
            v++;
        }
        char name[1024];
        snprintf(name, sizeof name, "-%s", key);
        PetscErrorCode e = PetscOptionsSetValue(NULL, name, v);
        if (e) {
            fclose(f);
            return e;
        }
    }
    fclose(f);
    return 0;
}

PetscErrorCode PetscOptionsGetString(PetscOptions o, const char pre[], const char name[], char string[], size_t len, PetscBool *set)
{
    (void)o;
    (void)pre;
    const char *v = opt_find(name);
    if (v) {
        strncpy(string, v, len);
        string[len - 1] = 0;
    }
    if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
    return 0;
}

static int ieq(const char *a, const char *b)
{
    for (; *a && *b; a++, b++)
        if (tolower((unsigned char)*a) != tolower((unsigned char)*b)) return 0;
    return !*a && !*b;
}

PetscErrorCode PetscOptionsGetBool(PetscOptions o, const char pre[], const char name[], PetscBool *ivalue, PetscBool *set)
{
    (void)o;
    (void)pre;
    const char *v = opt_find(name);
    if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
    if (!v) return 0;
    if (!*v || ieq(v, "true") || ieq(v, "yes") || ieq(v, "1") || ieq(v, "on")) *ivalue = PETSC_TRUE;
    else if (ieq(v, "false") || ieq(v, "no") || ieq(v, "0") || ieq(v, "off")) *ivalue = PETSC_FALSE;
    else SHIM_FAIL(PETSC_ERR_USER, "Unknown logical value: %s", v);
    return 0;
}

PetscErrorCode PetscOptionsGetInt(PetscOptions o, const char pre[], const char name[], PetscInt *ivalue, PetscBool *set)
{
    (void)o;
    (void)pre;
    const char *v = opt_find(name);
    if (set) *set = (v && *v) ? PETSC_TRUE : PETSC_FALSE;
    if (!v || !*v) return 0;
    char *end;
    long long x = strtoll(v, &end, 10);
    if (*end) SHIM_FAIL(PETSC_ERR_USER, "Input string %s has no integer value (do not include . in it)", v);
    *ivalue = (PetscInt)x;
    return 0;
}

PetscErrorCode PetscOptionsGetReal(PetscOptions o, const char pre[], const char name[], PetscReal *dvalue, PetscBool *set)
{
    (void)o;
    (void)pre;
    const char *v = opt_find(name);
    if (set) *set = (v && *v) ? PETSC_TRUE : PETSC_FALSE;
    if (!v || !*v) return 0;
    char *end;
    double x = strtod(v, &end);
    if (*end) SHIM_FAIL(PETSC_ERR_USER, "Input string %s has no numeric value", v);
    *dvalue = x;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ random */
static struct _p_PetscRandom default_random = {0x12345678ULL, 0x12345678UL};

PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r)
{
    (void)comm;
    *r = (PetscRandom)xcalloc(1, sizeof **r);
    (*r)->seed = 0x12345678UL;
    (*r)->state = (*r)->seed;
    return 0;
}
PetscErrorCode PetscRandomSetSeed(PetscRandom r, unsigned long seed)
{
    r->seed = seed;
    return 0;
}
PetscErrorCode PetscRandomSeed(PetscRandom r)
{
    r->state = ((uint64_t)r->seed << 16) | 0x330EULL; /* srand48 convention */
    return 0;
}
PetscErrorCode PetscRandomDestroy(PetscRandom *r)
{
    if (r && *r) {
        free(*r);
        *r = NULL;
    }
    return 0;
}
/* 48-bit linear congruential generator (the drand48 recurrence, which PETSc's default "rander48" type also uses);
 * the exact stream of a PETSc build is not reproduced -- parity runs read their random data from files. */
static double rand48(PetscRandom r)
{
    r->state = (r->state * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
    return (double)r->state / 281474976710656.0;
}
static PetscScalar rand_scalar(PetscRandom r)
{
    if (!r) r = &default_random;
    double re = rand48(r), im = rand48(r);
    return re + im * _Complex_I;
}

/* ------------------------------------------------------------------------------------------------ binary IO */
#define MAT_FILE_CLASSID 1211216
#define VEC_FILE_CLASSID 1211214

PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char name[], PetscFileMode mode, PetscViewer *viewer)
{
    (void)comm;
    PetscViewer v = (PetscViewer)xcalloc(1, sizeof *v);
    strncpy(v->path, name, sizeof v->path - 1);
    v->mode = mode;
    v->f = fopen(name, mode == FILE_MODE_READ ? "rb" : (mode == FILE_MODE_APPEND ? "ab" : "wb"));
    if (!v->f) {
        free(v);
        SHIM_FAIL(PETSC_ERR_FILE_OPEN, "Cannot open file %s for %s", name, mode == FILE_MODE_READ ? "reading" : "writing");
    }
    *viewer = v;
    return 0;
}

PetscErrorCode PetscViewerDestroy(PetscViewer *viewer)
{
    if (viewer && *viewer) {
        if ((*viewer)->f) fclose((*viewer)->f);
        free(*viewer);
        *viewer = NULL;
    }
    return 0;
}

static int rd_ints(FILE *f, int isz, PetscInt n, PetscInt *out)
{
    for (PetscInt i = 0; i < n; i++) {
        if (isz == 8) {
            uint64_t u;
            if (fread(&u, 8, 1, f) != 1) return 1;
            out[i] = (PetscInt)__builtin_bswap64(u);
        } else {
            uint32_t u;
            if (fread(&u, 4, 1, f) != 1) return 1;
            out[i] = (PetscInt)(int32_t)__builtin_bswap32(u);
        }
    }
    return 0;
}
static int rd_doubles(FILE *f, PetscInt n, double *out)
{
    if ((PetscInt)fread(out, 8, (size_t)n, f) != n) return 1;
    uint64_t *u = (uint64_t *)out;
    for (PetscInt i = 0; i < n; i++) u[i] = __builtin_bswap64(u[i]);
    return 0;
}
static void wr_int(FILE *f, PetscInt v)
{
    uint64_t u = __builtin_bswap64((uint64_t)v);
    fwrite(&u, 8, 1, f);
}
static void wr_scalar(FILE *f, PetscScalar s)
{
    double d[2] = {creal(s), cimag(s)};
    uint64_t u[2];
    memcpy(u, d, 16);
    u[0] = __builtin_bswap64(u[0]);
    u[1] = __builtin_bswap64(u[1]);
    fwrite(u, 8, 2, f);
}
/* the class id tells the index width of the writer: 32-bit PETSc builds write 4-byte ints, 64-bit builds 8-byte ints */
static int detect_isz(FILE *f, PetscInt classid)
{
    unsigned char b[8];
    long pos = ftell(f);
    if (fread(b, 1, 8, f) != 8) return 0;
    fseek(f, pos, SEEK_SET);
    uint32_t v32;
    uint64_t v64;
    memcpy(&v32, b, 4);
    memcpy(&v64, b, 8);
    if ((PetscInt)(int32_t)__builtin_bswap32(v32) == classid) return 4;
    if ((PetscInt)__builtin_bswap64(v64) == classid) return 8;
    return 0;
}
static long bytes_left(FILE *f)
{
    long pos = ftell(f);
    fseek(f, 0, SEEK_END);
    long end = ftell(f);
    fseek(f, pos, SEEK_SET);
    return end - pos;
}

/* ------------------------------------------------------------------------------------------------ Vec */
PetscErrorCode VecCreate(MPI_Comm comm, Vec *v)
{
    (void)comm;
    *v = (Vec)xcalloc(1, sizeof **v);
    (*v)->n = -1;
    (*v)->owns = 1;
    return 0;
}
PetscErrorCode VecSetType(Vec v, VecType type)
{
    (void)v;
    (void)type;
    return 0;
}
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N)
{
    (void)n;
    if (N < 0) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "global size must be given");
    free(v->a);
    v->n = N;
    v->a = (PetscScalar *)xcalloc((size_t)N, sizeof(PetscScalar));
    return 0;
}
PetscErrorCode VecSetUp(Vec v)
{
    (void)v;
    return 0;
}
PetscErrorCode VecDestroy(Vec *v)
{
    if (v && *v) {
        if ((*v)->owns) free((*v)->a);
        free(*v);
        *v = NULL;
    }
    return 0;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n)
{
    *n = v->n;
    return 0;
}
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *low, PetscInt *high)
{
    if (low) *low = 0;
    if (high) *high = v->n;
    return 0;
}
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode iora)
{
    for (PetscInt i = 0; i < ni; i++) {
        if (ix[i] < 0) continue;
        if (ix[i] >= v->n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Out of range index value %lld maximum %lld", (long long)ix[i], (long long)v->n);
        if (iora == ADD_VALUES) v->a[ix[i]] += y[i];
        else v->a[ix[i]] = y[i];
    }
    return 0;
}
PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar va, InsertMode mode) { return VecSetValues(v, 1, &i, &va, mode); }
PetscErrorCode VecAssemblyBegin(Vec v)
{
    (void)v;
    return 0;
}
PetscErrorCode VecAssemblyEnd(Vec v)
{
    (void)v;
    return 0;
}
#define SHIM_PRAGMA(x) _Pragma(#x)
#define OMP_FOR_N(n) SHIM_PRAGMA(omp parallel for schedule(static) if ((n) > 50000))

PetscErrorCode VecSet(Vec v, PetscScalar alpha)
{
    OMP_FOR_N(v->n)
    for (PetscInt i = 0; i < v->n; i++) v->a[i] = alpha;
    return 0;
}
PetscErrorCode VecZeroEntries(Vec v) { return VecSet(v, 0.0); }
static int vec_same_size(Vec x, Vec y) { return x->n == y->n; }
PetscErrorCode VecCopy(Vec x, Vec y)
{
    if (!vec_same_size(x, y)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Incompatible vector sizes %lld != %lld", (long long)x->n, (long long)y->n);
    if (x->a != y->a) memcpy(y->a, x->a, (size_t)x->n * sizeof(PetscScalar));
    return 0;
}
PetscErrorCode VecScale(Vec x, PetscScalar alpha)
{
    OMP_FOR_N(x->n)
    for (PetscInt i = 0; i < x->n; i++) x->a[i] *= alpha;
    return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x)
{
    if (!vec_same_size(x, y)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Incompatible vector sizes");
    PetscScalar *restrict ya = y->a;
    const PetscScalar *restrict xa = x->a;
    OMP_FOR_N(y->n)
    for (PetscInt i = 0; i < y->n; i++) ya[i] += alpha * xa[i];
    return 0;
}
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x)
{
    if (!vec_same_size(x, y)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Incompatible vector sizes");
    OMP_FOR_N(y->n)
    for (PetscInt i = 0; i < y->n; i++) y->a[i] = x->a[i] + beta * y->a[i];
    return 0;
}
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y)
{
    if (!vec_same_size(x, y) || !vec_same_size(w, x)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Incompatible vector sizes");
    PetscScalar *restrict wa = w->a;
    const PetscScalar *xa = x->a, *ya = y->a;
    OMP_FOR_N(w->n)
    for (PetscInt i = 0; i < w->n; i++) wa[i] = alpha * xa[i] + ya[i];
    return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
    if (!vec_same_size(x, y) || !vec_same_size(w, x)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Incompatible vector sizes");
    for (PetscInt i = 0; i < w->n; i++) w->a[i] = x->a[i] * y->a[i];
    return 0;
}
PetscErrorCode VecExp(Vec v)
{
    for (PetscInt i = 0; i < v->n; i++) v->a[i] = cexp(v->a[i]);
    return 0;
}
PetscErrorCode VecConjugate(Vec v)
{
    for (PetscInt i = 0; i < v->n; i++) v->a[i] = conj(v->a[i]);
    return 0;
}
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val)
{
    if (type != NORM_2 && type != NORM_FROBENIUS) SHIM_FAIL(PETSC_ERR_SUP, "only NORM_2");
    double s = 0;
    for (PetscInt i = 0; i < x->n; i++) s += creal(x->a[i]) * creal(x->a[i]) + cimag(x->a[i]) * cimag(x->a[i]);
    *val = sqrt(s);
    return 0;
}
PetscErrorCode VecNormalize(Vec x, PetscReal *val)
{
    PetscReal nrm;
    VecNorm(x, NORM_2, &nrm);
    if (nrm > 0) VecScale(x, 1.0 / nrm);
    if (val) *val = nrm;
    return 0;
}
PetscErrorCode VecSetRandom(Vec x, PetscRandom r)
{
    for (PetscInt i = 0; i < x->n; i++) x->a[i] = rand_scalar(r);
    return 0;
}
PetscErrorCode VecLoad(Vec v, PetscViewer viewer)
{
    FILE *f = viewer->f;
    int isz = detect_isz(f, VEC_FILE_CLASSID);
    if (!isz) SHIM_FAIL(PETSC_ERR_USER, "Not a vector next in file %s", viewer->path);
    PetscInt hdr[2];
    if (rd_ints(f, isz, 2, hdr)) SHIM_FAIL(PETSC_ERR_USER, "short vector header");
    if (v->n >= 0 && v->n != hdr[1]) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Vector in file different size (%lld) than input vector (%lld)", (long long)hdr[1], (long long)v->n);
    if (v->n < 0) VecSetSizes(v, PETSC_DECIDE, hdr[1]);
    if (rd_doubles(f, 2 * v->n, (double *)v->a)) SHIM_FAIL(PETSC_ERR_USER, "short vector data");
    return 0;
}
PetscErrorCode VecView(Vec v, PetscViewer viewer)
{
    FILE *f = viewer->f;
    wr_int(f, VEC_FILE_CLASSID);
    wr_int(f, v->n);
    for (PetscInt i = 0; i < v->n; i++) wr_scalar(f, v->a[i]);
    fflush(f);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ Mat basics */
PetscErrorCode MatCreate(MPI_Comm comm, Mat *A)
{
    (void)comm;
    *A = (Mat)xcalloc(1, sizeof **A);
    (*A)->m = (*A)->n = -1;
    (*A)->owns = 1;
    return 0;
}
PetscErrorCode MatSetType(Mat A, MatType type)
{
    if (strstr(type, "dense")) A->type = MT_DENSE;
    else if (strstr(type, "aij")) A->type = MT_AIJ;
    else SHIM_FAIL(PETSC_ERR_SUP, "Unknown Mat type %s", type);
    return 0;
}
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N)
{
    (void)m;
    (void)n;
    A->m = M;
    A->n = N;
    return 0;
}
static void mat_alloc_dense(Mat A)
{
    free(A->a);
    A->lda = A->m > 0 ? A->m : 1;
    A->a = (PetscScalar *)xcalloc((size_t)A->lda * (size_t)(A->n > 0 ? A->n : 1), sizeof(PetscScalar));
    A->owns = 1;
}
PetscErrorCode MatSetUp(Mat A)
{
    if (A->m < 0 || A->n < 0) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Must call MatSetSizes() first");
    if (A->type == MT_UNSET) A->type = MT_AIJ;
    if (A->type == MT_DENSE) {
        if (!A->a) mat_alloc_dense(A);
    } else if (!A->ia) {
        A->ia = (PetscInt *)xcalloc((size_t)A->m + 1, sizeof(PetscInt));
        A->ja = (PetscInt *)xcalloc(1, sizeof(PetscInt));
        A->va = (PetscScalar *)xcalloc(1, sizeof(PetscScalar));
    }
    return 0;
}
PetscErrorCode MatDestroy(Mat *A)
{
    if (A && *A) {
        Mat M = *A;
        if (M->owns) free(M->a);
        free(M->ia);
        free(M->ja);
        free(M->va);
        if (M->sub) free(M->sub);
        free(M);
        *A = NULL;
    }
    return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n)
{
    if (m) *m = A->m;
    if (n) *n = A->n;
    return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType type)
{
    (void)A;
    (void)type;
    return 0;
}
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type)
{
    (void)A;
    (void)type;
    return 0;
}

PetscErrorCode ShimMatCreateDenseFrom(PetscInt m, PetscInt n, const PetscScalar *colmajor, Mat *A)
{
    MatCreate(0, A);
    (*A)->type = MT_DENSE;
    (*A)->m = m;
    (*A)->n = n;
    mat_alloc_dense(*A);
    if (colmajor) memcpy((*A)->a, colmajor, (size_t)m * (size_t)n * sizeof(PetscScalar));
    return 0;
}
PetscErrorCode ShimMatCreateAIJFrom(PetscInt m, PetscInt n, const PetscInt *rowptr, const PetscInt *col, const PetscScalar *val, Mat *A)
{
    MatCreate(0, A);
    Mat M = *A;
    M->type = MT_AIJ;
    M->m = m;
    M->n = n;
    PetscInt nz = rowptr[m];
    M->ia = (PetscInt *)xcalloc((size_t)m + 1, sizeof(PetscInt));
    M->ja = (PetscInt *)xcalloc((size_t)nz, sizeof(PetscInt));
    M->va = (PetscScalar *)xcalloc((size_t)nz, sizeof(PetscScalar));
    memcpy(M->ia, rowptr, ((size_t)m + 1) * sizeof(PetscInt));
    memcpy(M->ja, col, (size_t)nz * sizeof(PetscInt));
    memcpy(M->va, val, (size_t)nz * sizeof(PetscScalar));
    return 0;
}
PetscScalar *ShimMatDenseArray(Mat A, PetscInt *lda)
{
    if (lda) *lda = A->lda;
    return A->a;
}
PetscErrorCode ShimVecCreateFrom(PetscInt n, const PetscScalar *a, Vec *v)
{
    VecCreate(0, v);
    VecSetSizes(*v, PETSC_DECIDE, n);
    if (a) memcpy((*v)->a, a, (size_t)n * sizeof(PetscScalar));
    return 0;
}
PetscScalar *ShimVecArray(Vec v) { return v->a; }

/* MatLoad: binary Mat record {classid, M, N, nz, rowlens[M], cols[nz], vals[nz]}; nz = -1 marks the dense-native
 * record (M*N values stored row by row, MatView_Dense_Binary of PETSc >= 3.14).  Either record can be loaded into either an AIJ or a dense Mat, as in PETSc. */
PetscErrorCode MatLoad(Mat A, PetscViewer viewer)
{
    FILE *f = viewer->f;
    int isz = detect_isz(f, MAT_FILE_CLASSID);
    if (!isz) SHIM_FAIL(PETSC_ERR_USER, "Not a matrix object in file %s", viewer->path);
    PetscInt hdr[4];
    if (rd_ints(f, isz, 4, hdr)) SHIM_FAIL(PETSC_ERR_USER, "short matrix header");
    PetscInt M = hdr[1], N = hdr[2], nz = hdr[3];
    if (A->m >= 0 && (A->m != M || A->n != N)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Matrix in file of different size (%lld x %lld)", (long long)M, (long long)N);
    A->m = M;
    A->n = N;
    if (A->type == MT_UNSET) A->type = MT_AIJ;
    PetscInt *rl = NULL, *cols = NULL;
    PetscScalar *vals = NULL;
    if (nz == -1) {
        vals = (PetscScalar *)xcalloc((size_t)M * (size_t)N, sizeof(PetscScalar));
        if (rd_doubles(f, 2 * M * N, (double *)vals)) SHIM_FAIL(PETSC_ERR_USER, "short dense data");
    } else {
        rl = (PetscInt *)xcalloc((size_t)M, sizeof(PetscInt));
        cols = (PetscInt *)xcalloc((size_t)nz, sizeof(PetscInt));
        vals = (PetscScalar *)xcalloc((size_t)nz, sizeof(PetscScalar));
        if (rd_ints(f, isz, M, rl) || rd_ints(f, isz, nz, cols)) SHIM_FAIL(PETSC_ERR_USER, "short index data");
        if (bytes_left(f) >= (long)nz * 16) {
            if (rd_doubles(f, 2 * nz, (double *)vals)) SHIM_FAIL(PETSC_ERR_USER, "short values");
        } else { /* a file written by a real-scalar PETSc (or MATLAB PetscBinaryWrite of a real matrix) */
            double *tmp = (double *)xcalloc((size_t)nz, sizeof(double));
            if (rd_doubles(f, nz, tmp)) SHIM_FAIL(PETSC_ERR_USER, "short values");
            for (PetscInt p = 0; p < nz; p++) vals[p] = tmp[p];
            free(tmp);
        }
    }
    if (A->type == MT_DENSE) {
        mat_alloc_dense(A);
        if (nz == -1) {   /* dense-native values are stored row by row (MatLoad_Dense_Binary, PETSc >= 3.14) */
            for (PetscInt r = 0; r < M; r++)
                for (PetscInt c = 0; c < N; c++) A->a[r + A->lda * c] = vals[r * N + c];
        }
        else {
            PetscInt p = 0;
            for (PetscInt r = 0; r < M; r++)
                for (PetscInt q = 0; q < rl[r]; q++, p++) A->a[r + A->lda * cols[p]] = vals[p];
        }
    } else {
        free(A->ia);
        free(A->ja);
        free(A->va);
        A->ia = (PetscInt *)xcalloc((size_t)M + 1, sizeof(PetscInt));
        if (nz == -1) {
            A->ja = (PetscInt *)xcalloc((size_t)M * (size_t)N, sizeof(PetscInt));
            A->va = (PetscScalar *)xcalloc((size_t)M * (size_t)N, sizeof(PetscScalar));
            for (PetscInt r = 0; r < M; r++) {
                A->ia[r + 1] = (r + 1) * N;
                for (PetscInt c = 0; c < N; c++) {
                    A->ja[r * N + c] = c;
                    A->va[r * N + c] = vals[r * N + c];
                }
            }
        } else {
            for (PetscInt r = 0; r < M; r++) A->ia[r + 1] = A->ia[r] + rl[r];
            if (A->ia[M] != nz) SHIM_FAIL(PETSC_ERR_USER, "row lengths do not sum to nz");
            A->ja = cols;
            A->va = vals;
            cols = NULL;
            vals = NULL;
        }
    }
    free(rl);
    free(cols);
    free(vals);
    return 0;
}

/* MatView to a binary viewer in the default format: a dense matrix is written as an AIJ record holding every entry */
PetscErrorCode MatView(Mat A, PetscViewer viewer)
{
    FILE *f = viewer->f;
    wr_int(f, MAT_FILE_CLASSID);
    wr_int(f, A->m);
    wr_int(f, A->n);
    if (A->type == MT_DENSE) {
        wr_int(f, A->m * A->n);
        for (PetscInt r = 0; r < A->m; r++) wr_int(f, A->n);
        for (PetscInt r = 0; r < A->m; r++)
            for (PetscInt c = 0; c < A->n; c++) wr_int(f, c);
        for (PetscInt r = 0; r < A->m; r++)
            for (PetscInt c = 0; c < A->n; c++) wr_scalar(f, A->a[r + A->lda * c]);
    } else {
        PetscInt nz = A->ia[A->m];
        wr_int(f, nz);
        for (PetscInt r = 0; r < A->m; r++) wr_int(f, A->ia[r + 1] - A->ia[r]);
        for (PetscInt p = 0; p < nz; p++) wr_int(f, A->ja[p]);
        for (PetscInt p = 0; p < nz; p++) wr_scalar(f, A->va[p]);
    }
    fflush(f);
    return 0;
}

PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat *B)
{
    MatCreate(0, B);
    Mat C = *B;
    C->type = A->type;
    C->m = A->m;
    C->n = A->n;
    if (A->type == MT_DENSE) {
        mat_alloc_dense(C);
        if (op == MAT_COPY_VALUES)
            for (PetscInt j = 0; j < A->n; j++) memcpy(C->a + C->lda * j, A->a + A->lda * j, (size_t)A->m * sizeof(PetscScalar));
    } else {
        PetscInt nz = A->ia[A->m];
        C->ia = (PetscInt *)xcalloc((size_t)A->m + 1, sizeof(PetscInt));
        C->ja = (PetscInt *)xcalloc((size_t)nz, sizeof(PetscInt));
        C->va = (PetscScalar *)xcalloc((size_t)nz, sizeof(PetscScalar));
        memcpy(C->ia, A->ia, ((size_t)A->m + 1) * sizeof(PetscInt));
        memcpy(C->ja, A->ja, (size_t)nz * sizeof(PetscInt));
        if (op == MAT_COPY_VALUES) memcpy(C->va, A->va, (size_t)nz * sizeof(PetscScalar));
    }
    return 0;
}

PetscErrorCode MatCopy(Mat A, Mat B, MatStructure str)
{
    (void)str;
    if (A->m != B->m || A->n != B->n)
        SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Mat A,Mat B: global dim (%lld,%lld) (%lld,%lld)", (long long)A->m, (long long)B->m, (long long)A->n, (long long)B->n);
    if (A->type != MT_DENSE || B->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatCopy: dense only");
    if (A->a == B->a) return 0;
    for (PetscInt j = 0; j < A->n; j++) memcpy(B->a + B->lda * j, A->a + A->lda * j, (size_t)A->m * sizeof(PetscScalar));
    return 0;
}

PetscErrorCode MatScale(Mat A, PetscScalar a)
{
    if (A->type == MT_DENSE) {
        for (PetscInt j = 0; j < A->n; j++) {
            PetscScalar *c = A->a + A->lda * j;
            OMP_FOR_N(A->m)
            for (PetscInt i = 0; i < A->m; i++) c[i] *= a;
        }
    } else
        for (PetscInt p = 0; p < A->ia[A->m]; p++) A->va[p] *= a;
    return 0;
}

PetscErrorCode MatShift(Mat A, PetscScalar a)
{
    PetscInt d = A->m < A->n ? A->m : A->n;
    if (A->type == MT_DENSE) {
        for (PetscInt i = 0; i < d; i++) A->a[i + A->lda * i] += a;
        return 0;
    }
    /* AIJ: add to the diagonal entries, inserting the ones missing from the pattern (keeps columns sorted) */
    PetscInt missing = 0;
    for (PetscInt r = 0; r < d; r++) {
        int found = 0;
        for (PetscInt p = A->ia[r]; p < A->ia[r + 1]; p++)
            if (A->ja[p] == r) {
                A->va[p] += a;
                found = 1;
                break;
            }
        if (!found) missing++;
    }
    if (!missing) return 0;
    PetscInt nz = A->ia[A->m] + missing;
    PetscInt *ia = (PetscInt *)xcalloc((size_t)A->m + 1, sizeof(PetscInt));
    PetscInt *ja = (PetscInt *)xcalloc((size_t)nz, sizeof(PetscInt));
    PetscScalar *va = (PetscScalar *)xcalloc((size_t)nz, sizeof(PetscScalar));
    PetscInt q = 0;
    for (PetscInt r = 0; r < A->m; r++) {
        int has = 0;
        for (PetscInt p = A->ia[r]; p < A->ia[r + 1]; p++)
            if (A->ja[p] == r) has = 1;
        int placed = has || r >= d;
        for (PetscInt p = A->ia[r]; p < A->ia[r + 1]; p++) {
            if (!placed && A->ja[p] > r) {
                ja[q] = r;
                va[q++] = a;
                placed = 1;
            }
            ja[q] = A->ja[p];
            va[q++] = A->va[p];
        }
        if (!placed) {
            ja[q] = r;
            va[q++] = a;
        }
        ia[r + 1] = q;
    }
    free(A->ia);
    free(A->ja);
    free(A->va);
    A->ia = ia;
    A->ja = ja;
    A->va = va;
    return 0;
}

PetscErrorCode MatConjugate(Mat A)
{
    if (A->type == MT_DENSE) {
        for (PetscInt j = 0; j < A->n; j++)
            for (PetscInt i = 0; i < A->m; i++) A->a[i + A->lda * j] = conj(A->a[i + A->lda * j]);
    } else
        for (PetscInt p = 0; p < A->ia[A->m]; p++) A->va[p] = conj(A->va[p]);
    return 0;
}

PetscErrorCode MatImaginaryPart(Mat A)
{
    if (A->type == MT_DENSE) {
        for (PetscInt j = 0; j < A->n; j++)
            for (PetscInt i = 0; i < A->m; i++) A->a[i + A->lda * j] = cimag(A->a[i + A->lda * j]);
    } else
        for (PetscInt p = 0; p < A->ia[A->m]; p++) A->va[p] = cimag(A->va[p]);
    return 0;
}

PetscErrorCode MatNorm(Mat A, NormType type, PetscReal *nrm)
{
    if (type != NORM_FROBENIUS) SHIM_FAIL(PETSC_ERR_SUP, "only NORM_FROBENIUS");
    double s = 0;
    if (A->type == MT_DENSE) {
        for (PetscInt j = 0; j < A->n; j++)
            for (PetscInt i = 0; i < A->m; i++) {
                PetscScalar v = A->a[i + A->lda * j];
                s += creal(v) * creal(v) + cimag(v) * cimag(v);
            }
    } else
        for (PetscInt p = 0; p < A->ia[A->m]; p++) s += creal(A->va[p]) * creal(A->va[p]) + cimag(A->va[p]) * cimag(A->va[p]);
    *nrm = sqrt(s);
    return 0;
}

PetscErrorCode MatSetRandom(Mat A, PetscRandom r)
{
    if (A->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatSetRandom: dense only");
    for (PetscInt j = 0; j < A->n; j++)
        for (PetscInt i = 0; i < A->m; i++) A->a[i + A->lda * j] = rand_scalar(r);
    return 0;
}

static PetscErrorCode mat_transpose(Mat A, MatReuse reuse, Mat *B, int herm)
{
    if (reuse != MAT_INPLACE_MATRIX && reuse != MAT_INITIAL_MATRIX) SHIM_FAIL(PETSC_ERR_SUP, "transpose: MAT_REUSE_MATRIX not supported");
    Mat T;
    MatCreate(0, &T);
    T->type = A->type;
    T->m = A->n;
    T->n = A->m;
    if (A->type == MT_DENSE) {
        mat_alloc_dense(T);
        for (PetscInt j = 0; j < A->n; j++)
            for (PetscInt i = 0; i < A->m; i++) {
                PetscScalar v = A->a[i + A->lda * j];
                T->a[j + T->lda * i] = herm ? conj(v) : v;
            }
    } else {
        PetscInt nz = A->ia[A->m];
        T->ia = (PetscInt *)xcalloc((size_t)T->m + 1, sizeof(PetscInt));
        T->ja = (PetscInt *)xcalloc((size_t)nz, sizeof(PetscInt));
        T->va = (PetscScalar *)xcalloc((size_t)nz, sizeof(PetscScalar));
        for (PetscInt p = 0; p < nz; p++) T->ia[A->ja[p] + 1]++;
        for (PetscInt r = 0; r < T->m; r++) T->ia[r + 1] += T->ia[r];
        PetscInt *next = (PetscInt *)xcalloc((size_t)T->m, sizeof(PetscInt));
        memcpy(next, T->ia, (size_t)T->m * sizeof(PetscInt));
        for (PetscInt r = 0; r < A->m; r++)
            for (PetscInt p = A->ia[r]; p < A->ia[r + 1]; p++) {
                PetscInt q = next[A->ja[p]]++;
                T->ja[q] = r;
                T->va[q] = herm ? conj(A->va[p]) : A->va[p];
            }
        free(next);
    }
    if (reuse == MAT_INPLACE_MATRIX) {
        /* swap the contents into A (B == &A in the reference's calls) */
        if (A->colvec_out || A->sub) SHIM_FAIL(PETSC_ERR_USER, "in-place transpose of a matrix with outstanding views");
        if (A->owns) free(A->a);
        free(A->ia);
        free(A->ja);
        free(A->va);
        *A = *T;
        free(T);
        if (B) *B = A;
    } else
        *B = T;
    return 0;
}
PetscErrorCode MatTranspose(Mat A, MatReuse reuse, Mat *B) { return mat_transpose(A, reuse, B, 0); }
PetscErrorCode MatHermitianTranspose(Mat A, MatReuse reuse, Mat *B) { return mat_transpose(A, reuse, B, 1); }

/* ------------------------------------------------------------------------------------------------ products */
static PetscErrorCode mult_core(Mat A, Vec x, Vec add, Vec y)
{
    if (x->n != A->n || y->n != A->m || (add && add->n != A->m))
        SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Mat mat,Vec x: global dim %lld %lld / Vec y: %lld %lld", (long long)A->n, (long long)x->n, (long long)A->m, (long long)y->n);
    if (x->a == y->a) SHIM_FAIL(PETSC_ERR_USER, "x and y must be different vectors");
    const PetscScalar *restrict xa = x->a;
    PetscScalar *ya = y->a;
    if (A->type == MT_AIJ) {
        const PetscInt *ia = A->ia, *ja = A->ja;
        const PetscScalar *va = A->va;
        const PetscScalar *ad = add ? add->a : NULL;
        OMP_FOR_N(A->m)
        for (PetscInt r = 0; r < A->m; r++) {
            PetscScalar sum = ad ? ad[r] : 0.0;
            for (PetscInt p = ia[r]; p < ia[r + 1]; p++) sum += va[p] * xa[ja[p]];
            ya[r] = sum;
        }
    } else {
        /* dense GEMV, column sweep (the order of the reference BLAS zgemv 'N'); threads split the rows */
        const PetscInt m = A->m, n = A->n, lda = A->lda;
        const PetscScalar *a = A->a;
        const PetscScalar *ad = add ? add->a : NULL;
#pragma omp parallel if (m > 50000)
        {
            PetscInt lo = 0, hi = m;
#ifdef _OPENMP
            PetscInt nt = omp_get_num_threads(), t = omp_get_thread_num();
            lo = m * t / nt;
            hi = m * (t + 1) / nt;
#endif
            for (PetscInt i = lo; i < hi; i++) ya[i] = ad ? ad[i] : 0.0;
            for (PetscInt j = 0; j < n; j++) {
                const PetscScalar xj = xa[j];
                const PetscScalar *restrict c = a + lda * j;
                for (PetscInt i = lo; i < hi; i++) ya[i] += c[i] * xj;
            }
        }
    }
    return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return mult_core(A, x, NULL, y); }
PetscErrorCode MatMultAdd(Mat A, Vec v1, Vec v2, Vec v3)
{
    if (v2->a == v3->a) { /* z = z + A x */
        Vec t;
        ShimVecCreateFrom(v2->n, v2->a, &t);
        PetscErrorCode e = mult_core(A, v1, t, v3);
        VecDestroy(&t);
        return e;
    }
    return mult_core(A, v1, v2, v3);
}

static PetscErrorCode multh_core(Mat A, Vec x, Vec add, Vec y)
{
    if (x->n != A->m || y->n != A->n || (add && add->n != A->n)) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "MatMultHermitianTranspose: incompatible sizes");
    if (x->a == y->a) SHIM_FAIL(PETSC_ERR_USER, "x and y must be different vectors");
    PetscScalar *ya = y->a;
    if (add) {
        if (add->a != y->a) memcpy(ya, add->a, (size_t)y->n * sizeof(PetscScalar));
    } else
        memset(ya, 0, (size_t)y->n * sizeof(PetscScalar));
    if (A->type == MT_AIJ) {
        for (PetscInt r = 0; r < A->m; r++) {
            const PetscScalar alpha = x->a[r];
            for (PetscInt p = A->ia[r]; p < A->ia[r + 1]; p++) ya[A->ja[p]] += alpha * conj(A->va[p]);
        }
    } else {
        for (PetscInt j = 0; j < A->n; j++) {
            const PetscScalar *c = A->a + A->lda * j;
            PetscScalar s = 0;
            for (PetscInt i = 0; i < A->m; i++) s += conj(c[i]) * x->a[i];
            ya[j] += s;
        }
    }
    return 0;
}
PetscErrorCode MatMultHermitianTranspose(Mat A, Vec x, Vec y) { return multh_core(A, x, NULL, y); }
PetscErrorCode MatMultHermitianTransposeAdd(Mat A, Vec v1, Vec v2, Vec v3) { return multh_core(A, v1, v2, v3); }

/* C = A*B with B dense; A dense or AIJ.  MAT_REUSE_MATRIX writes into the caller's dense C (which the reference also
 * passes aliased to B for the square weight matrices: the product is formed in a temporary first). */
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C)
{
    (void)fill;
    if (B->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatMatMult: B must be dense");
    if (A->n != B->m) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Matrix dimensions are incompatible, %lld != %lld", (long long)A->n, (long long)B->m);
    Mat T;
    MatCreate(0, &T);
    T->type = MT_DENSE;
    T->m = A->m;
    T->n = B->n;
    mat_alloc_dense(T);
    struct _p_Vec xv = {B->m, NULL, 0}, yv = {T->m, NULL, 0};
    for (PetscInt j = 0; j < B->n; j++) {
        xv.a = B->a + B->lda * j;
        yv.a = T->a + T->lda * j;
        PetscErrorCode e = mult_core(A, &xv, NULL, &yv);
        if (e) return e;
    }
    if (scall == MAT_REUSE_MATRIX) {
        Mat D = *C;
        if (D->type != MT_DENSE || D->m != T->m || D->n != T->n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "MatMatMult: reused C has the wrong size/type");
        MatCopy(T, D, SAME_NONZERO_PATTERN);
        MatDestroy(&T);
    } else
        *C = T;
    return 0;
}

/* C = A^T * B (plain transpose, as PETSc defines it; the reference conjugates A beforehand), A and B dense */
PetscErrorCode MatTransposeMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat *C)
{
    (void)fill;
    if (A->type != MT_DENSE || B->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatTransposeMatMult: dense only");
    if (A->m != B->m) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Matrix dimensions are incompatible");
    Mat T;
    MatCreate(0, &T);
    T->type = MT_DENSE;
    T->m = A->n;
    T->n = B->n;
    mat_alloc_dense(T);
    for (PetscInt j = 0; j < B->n; j++)
        for (PetscInt i = 0; i < A->n; i++) {
            const PetscScalar *a = A->a + A->lda * i, *b = B->a + B->lda * j;
            PetscScalar s = 0;
            for (PetscInt r = 0; r < A->m; r++) s += a[r] * b[r];
            T->a[i + T->lda * j] = s;
        }
    if (scall == MAT_REUSE_MATRIX) {
        Mat D = *C;
        if (D->type != MT_DENSE || D->m != T->m || D->n != T->n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "MatTransposeMatMult: reused C has the wrong size/type");
        MatCopy(T, D, SAME_NONZERO_PATTERN);
        MatDestroy(&T);
    } else
        *C = T;
    return 0;
}

PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
    if (right) ShimVecCreateFrom(A->n, NULL, right);
    if (left) ShimVecCreateFrom(A->m, NULL, left);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ dense views */
static PetscErrorCode get_colvec(Mat A, PetscInt col, Vec *v)
{
    if (A->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatDenseGetColumnVec: dense only");
    if (A->colvec_out) SHIM_FAIL(PETSC_ERR_USER, "Need to call MatDenseRestoreColumnVec() first");
    if (col < 0 || col >= A->n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Column %lld out of range [0,%lld)", (long long)col, (long long)A->n);
    A->colvec.n = A->m;
    A->colvec.a = A->a + A->lda * col;
    A->colvec.owns = 0;
    A->colvec_out = 1;
    *v = &A->colvec;
    return 0;
}
static PetscErrorCode restore_colvec(Mat A, Vec *v)
{
    if (!A->colvec_out) SHIM_FAIL(PETSC_ERR_USER, "Need to call MatDenseGetColumnVec() first");
    A->colvec_out = 0;
    if (v) *v = NULL;
    return 0;
}
PetscErrorCode MatDenseGetColumnVecRead(Mat A, PetscInt col, Vec *v) { return get_colvec(A, col, v); }
PetscErrorCode MatDenseGetColumnVecWrite(Mat A, PetscInt col, Vec *v) { return get_colvec(A, col, v); }
PetscErrorCode MatDenseRestoreColumnVecRead(Mat A, PetscInt col, Vec *v)
{
    (void)col;
    return restore_colvec(A, v);
}
PetscErrorCode MatDenseRestoreColumnVecWrite(Mat A, PetscInt col, Vec *v)
{
    (void)col;
    return restore_colvec(A, v);
}

PetscErrorCode MatDenseGetSubMatrix(Mat A, PetscInt rbegin, PetscInt rend, PetscInt cbegin, PetscInt cend, Mat *v)
{
    if (A->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "MatDenseGetSubMatrix: dense only");
    if (A->sub) SHIM_FAIL(PETSC_ERR_USER, "Need to call MatDenseRestoreSubMatrix() first");
    if (rbegin == PETSC_DECIDE) rbegin = 0;
    if (rend == PETSC_DECIDE) rend = A->m;
    if (cbegin == PETSC_DECIDE) cbegin = 0;
    if (cend == PETSC_DECIDE) cend = A->n;
    if (rbegin < 0 || rend > A->m || rbegin > rend || cbegin < 0 || cend > A->n || cbegin > cend)
        SHIM_FAIL(PETSC_ERR_ARG_SIZ, "Invalid submatrix range [%lld,%lld) x [%lld,%lld) of %lld x %lld", (long long)rbegin, (long long)rend, (long long)cbegin,
                  (long long)cend, (long long)A->m, (long long)A->n);
    Mat S = (Mat)xcalloc(1, sizeof *S);
    S->type = MT_DENSE;
    S->m = rend - rbegin;
    S->n = cend - cbegin;
    S->lda = A->lda;
    S->a = A->a + rbegin + A->lda * cbegin;
    S->owns = 0;
    A->sub = S;
    *v = S;
    return 0;
}
PetscErrorCode MatDenseRestoreSubMatrix(Mat A, Mat *v)
{
    if (!A->sub) SHIM_FAIL(PETSC_ERR_USER, "Need to call MatDenseGetSubMatrix() first");
    if (v && *v != A->sub) SHIM_FAIL(PETSC_ERR_USER, "Not the submatrix obtained from MatDenseGetSubMatrix()");
    if (A->sub->colvec_out || A->sub->sub) SHIM_FAIL(PETSC_ERR_USER, "submatrix still has outstanding views");
    free(A->sub);
    A->sub = NULL;
    if (v) *v = NULL;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ KSP */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp)
{
    (void)comm;
    *ksp = (KSP)xcalloc(1, sizeof **ksp);
    return 0;
}
PetscErrorCode KSPSetOperators(KSP ksp, Mat Amat, Mat Pmat)
{
    (void)Pmat;
    ksp->A = Amat;
    return 0;
}
PetscErrorCode KSPSetType(KSP ksp, KSPType type)
{
    (void)ksp;
    (void)type;
    return 0;
}
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
    (void)ksp;
    (void)rtol;
    (void)abstol;
    (void)dtol;
    (void)maxits;
    return 0;
}
PetscErrorCode KSPDestroy(KSP *ksp)
{
    if (ksp && *ksp) {
        free(*ksp);
        *ksp = NULL;
    }
    return 0;
}
/* The reference asks for a Krylov solve to 1e-15 of a small dense system; what that converges to is the solution of
 * the system, computed here by LU with partial pivoting. */
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x)
{
    Mat A = ksp->A;
    if (!A || A->type != MT_DENSE || A->m != A->n) SHIM_FAIL(PETSC_ERR_SUP, "KSPSolve: square dense operator expected");
    const PetscInt n = A->m;
    if (b->n != n || x->n != n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "KSPSolve: incompatible sizes");
    PetscScalar *L = (PetscScalar *)xcalloc((size_t)n * (size_t)n, sizeof(PetscScalar));
    PetscScalar *r = (PetscScalar *)xcalloc((size_t)n, sizeof(PetscScalar));
    for (PetscInt j = 0; j < n; j++)
        for (PetscInt i = 0; i < n; i++) L[i + n * j] = A->a[i + A->lda * j];
    memcpy(r, b->a, (size_t)n * sizeof(PetscScalar));
    for (PetscInt c = 0; c < n; c++) {
        PetscInt piv = c;
        double best = cabs(L[c + n * c]);
        for (PetscInt i = c + 1; i < n; i++)
            if (cabs(L[i + n * c]) > best) {
                best = cabs(L[i + n * c]);
                piv = i;
            }
        if (best == 0.0) {
            free(L);
            free(r);
            SHIM_FAIL(PETSC_ERR_USER, "KSPSolve: singular matrix");
        }
        if (piv != c) {
            for (PetscInt j = 0; j < n; j++) {
                PetscScalar t = L[c + n * j];
                L[c + n * j] = L[piv + n * j];
                L[piv + n * j] = t;
            }
            PetscScalar t = r[c];
            r[c] = r[piv];
            r[piv] = t;
        }
        for (PetscInt i = c + 1; i < n; i++) {
            PetscScalar f = L[i + n * c] / L[c + n * c];
            L[i + n * c] = f;
            for (PetscInt j = c + 1; j < n; j++) L[i + n * j] -= f * L[c + n * j];
            r[i] -= f * r[c];
        }
    }
    for (PetscInt i = n - 1; i >= 0; i--) {
        PetscScalar s = r[i];
        for (PetscInt j = i + 1; j < n; j++) s -= L[i + n * j] * r[j];
        r[i] = s / L[i + n * i];
    }
    memcpy(x->a, r, (size_t)n * sizeof(PetscScalar));
    free(L);
    free(r);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ BV */
PetscErrorCode BVCreateFromMat(Mat A, BV *bv)
{
    if (A->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "BVCreateFromMat: dense only");
    BV b = (BV)xcalloc(1, sizeof *b);
    b->n = A->m;
    b->k = A->n;
    b->a = (PetscScalar *)xcalloc((size_t)b->n * (size_t)b->k, sizeof(PetscScalar));
    for (PetscInt j = 0; j < b->k; j++) memcpy(b->a + b->n * j, A->a + A->lda * j, (size_t)b->n * sizeof(PetscScalar));
    *bv = b;
    return 0;
}
PetscErrorCode BVSetType(BV bv, BVType type)
{
    (void)bv;
    (void)type;
    return 0;
}
PetscErrorCode BVDestroy(BV *bv)
{
    if (bv && *bv) {
        free((*bv)->a);
        free(*bv);
        *bv = NULL;
    }
    return 0;
}
PetscErrorCode BVCreateMat(BV bv, Mat *A) { return ShimMatCreateDenseFrom(bv->n, bv->k, bv->a, A); }

/* thin QR by Gram-Schmidt, column by column: classical Gram-Schmidt with one conditional re-orthogonalisation
 * (refinement "if needed", eta = 1/sqrt(2)) -- SLEPc's default BV orthogonalisation.  R has a positive real diagonal
 * (the column norms), so Q is the unique thin-QR factor. */
static PetscErrorCode gs_orthogonalize(PetscInt n, PetscInt k, PetscScalar *Q, PetscScalar *R /* k x k or NULL */)
{
    PetscScalar *h = (PetscScalar *)xcalloc((size_t)k, sizeof(PetscScalar));
    for (PetscInt j = 0; j < k; j++) {
        PetscScalar *v = Q + n * j;
        double onrm = 0;
        for (PetscInt i = 0; i < n; i++) onrm += creal(v[i]) * creal(v[i]) + cimag(v[i]) * cimag(v[i]);
        onrm = sqrt(onrm);
        double nrm = onrm;
        if (R)
            for (PetscInt c = 0; c < k; c++) R[c + k * j] = 0;
        for (int pass = 0; pass < 3; pass++) {
            for (PetscInt c = 0; c < j; c++) {
                const PetscScalar *q = Q + n * c;
                PetscScalar s = 0;
                for (PetscInt i = 0; i < n; i++) s += conj(q[i]) * v[i];
                h[c] = s;
            }
            for (PetscInt c = 0; c < j; c++) {
                const PetscScalar *q = Q + n * c;
                const PetscScalar hc = h[c];
                for (PetscInt i = 0; i < n; i++) v[i] -= hc * q[i];
                if (R) R[c + k * j] += hc;
            }
            double s2 = 0;
            for (PetscInt i = 0; i < n; i++) s2 += creal(v[i]) * creal(v[i]) + cimag(v[i]) * cimag(v[i]);
            nrm = sqrt(s2);
            if (j == 0 || nrm >= 0.7071 * onrm) break;
            onrm = nrm;
        }
        if (nrm == 0.0) {
            free(h);
            SHIM_FAIL(PETSC_ERR_USER, "BVOrthogonalize: linearly dependent column %lld", (long long)j);
        }
        for (PetscInt i = 0; i < n; i++) v[i] /= nrm;
        if (R) R[j + k * j] = nrm;
    }
    free(h);
    return 0;
}

PetscErrorCode BVOrthogonalize(BV V, Mat R)
{
    PetscScalar *r = NULL;
    if (R) {
        if (R->type != MT_DENSE || R->m != V->k || R->n != V->k || R->lda != V->k) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "BVOrthogonalize: R must be k x k dense");
        r = R->a;
    }
    return gs_orthogonalize(V->n, V->k, V->a, r);
}

/* ------------------------------------------------------------------------------------------------ SVD */
PetscErrorCode SVDCreate(MPI_Comm comm, SVD *svd)
{
    (void)comm;
    *svd = (SVD)xcalloc(1, sizeof **svd);
    return 0;
}
PetscErrorCode SVDSetOperators(SVD svd, Mat A, Mat B)
{
    if (B) SHIM_FAIL(PETSC_ERR_SUP, "generalized SVD not supported");
    svd->A = A;
    return 0;
}
PetscErrorCode SVDSetDimensions(SVD svd, PetscInt nsv, PetscInt ncv, PetscInt mpd)
{
    (void)ncv;
    (void)mpd;
    svd->nsv = nsv;
    return 0;
}
PetscErrorCode SVDDestroy(SVD *svd)
{
    if (svd && *svd) {
        free((*svd)->sigma);
        free((*svd)->U);
        free((*svd)->V);
        free(*svd);
        *svd = NULL;
    }
    return 0;
}

/* Hermitian eigen-decomposition G = W diag(lam) W^H by cyclic complex Jacobi rotations (G: n x n col-major, destroyed) */
static void herm_jacobi(PetscInt n, PetscScalar *G, PetscScalar *W, double *lam)
{
    for (PetscInt j = 0; j < n; j++)
        for (PetscInt i = 0; i < n; i++) W[i + n * j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (PetscInt j = 0; j < n; j++)
            for (PetscInt i = 0; i < n; i++) {
                double a2 = creal(G[i + n * j]) * creal(G[i + n * j]) + cimag(G[i + n * j]) * cimag(G[i + n * j]);
                if (i == j) diag += a2;
                else off += a2;
            }
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (PetscInt p = 0; p < n - 1; p++)
            for (PetscInt q = p + 1; q < n; q++) {
                PetscScalar gpq = G[p + n * q];
                double ab = cabs(gpq);
                if (ab == 0.0) continue;
                double app = creal(G[p + n * p]), aqq = creal(G[q + n * q]);
                PetscScalar ph = gpq / ab; /* e^{i phi} */
                double tau = (aqq - app) / (2.0 * ab);
                double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                /* rotation J: columns p,q <- [c*p - s*conj(ph)*q , s*ph*p + c*q] */
                for (PetscInt i = 0; i < n; i++) {
                    PetscScalar gip = G[i + n * p], giq = G[i + n * q];
                    G[i + n * p] = c * gip - s * conj(ph) * giq;
                    G[i + n * q] = s * ph * gip + c * giq;
                }
                for (PetscInt i = 0; i < n; i++) {
                    PetscScalar gpi = G[p + n * i], gqi = G[q + n * i];
                    G[p + n * i] = c * gpi - s * ph * gqi;
                    G[q + n * i] = s * conj(ph) * gpi + c * gqi;
                }
                for (PetscInt i = 0; i < n; i++) {
                    PetscScalar wip = W[i + n * p], wiq = W[i + n * q];
                    W[i + n * p] = c * wip - s * conj(ph) * wiq;
                    W[i + n * q] = s * ph * wip + c * wiq;
                }
                G[p + n * q] = 0;
                G[q + n * p] = 0;
                G[p + n * p] = creal(G[p + n * p]);
                G[q + n * q] = creal(G[q + n * q]);
            }
    }
    for (PetscInt i = 0; i < n; i++) lam[i] = creal(G[i + n * i]);
}

/* SLEPc's default SVD solver is the cross-product method: eigenpairs of A^H A give sigma^2 and the right vectors;
 * the left vectors are A v, orthonormalised.  Triplets are sorted by decreasing sigma. */
PetscErrorCode SVDSolve(SVD svd)
{
    Mat A = svd->A;
    if (!A || A->type != MT_DENSE) SHIM_FAIL(PETSC_ERR_SUP, "SVDSolve: dense operator expected");
    const PetscInt m = A->m, n = A->n;
    if (svd->nsv <= 0 || svd->nsv > n) svd->nsv = n;
    svd->m = m;
    svd->n = n;
    PetscScalar *G = (PetscScalar *)xcalloc((size_t)n * (size_t)n, sizeof(PetscScalar));
    PetscScalar *W = (PetscScalar *)xcalloc((size_t)n * (size_t)n, sizeof(PetscScalar));
    double *lam = (double *)xcalloc((size_t)n, sizeof(double));
    for (PetscInt j = 0; j < n; j++)
        for (PetscInt i = 0; i <= j; i++) {
            const PetscScalar *a = A->a + A->lda * i, *b = A->a + A->lda * j;
            PetscScalar s = 0;
            for (PetscInt r = 0; r < m; r++) s += conj(a[r]) * b[r];
            G[i + n * j] = s;
            G[j + n * i] = conj(s);
        }
    for (PetscInt i = 0; i < n; i++) G[i + n * i] = creal(G[i + n * i]);
    herm_jacobi(n, G, W, lam);
    PetscInt *ord = (PetscInt *)xcalloc((size_t)n, sizeof(PetscInt));
    for (PetscInt i = 0; i < n; i++) ord[i] = i;
    for (PetscInt i = 1; i < n; i++) {
        PetscInt o = ord[i], j = i - 1;
        while (j >= 0 && lam[ord[j]] < lam[o]) {
            ord[j + 1] = ord[j];
            j--;
        }
        ord[j + 1] = o;
    }
    free(svd->sigma);
    free(svd->U);
    free(svd->V);
    svd->sigma = (PetscReal *)xcalloc((size_t)n, sizeof(PetscReal));
    svd->V = (PetscScalar *)xcalloc((size_t)n * (size_t)n, sizeof(PetscScalar));
    svd->U = (PetscScalar *)xcalloc((size_t)m * (size_t)n, sizeof(PetscScalar));
    for (PetscInt c = 0; c < n; c++) {
        svd->sigma[c] = sqrt(lam[ord[c]] > 0 ? lam[ord[c]] : 0.0);
        memcpy(svd->V + n * c, W + n * ord[c], (size_t)n * sizeof(PetscScalar));
    }
    /* U = A V, then orthonormalise (R = diag(sigma) up to rounding) */
    for (PetscInt c = 0; c < n; c++) {
        PetscScalar *u = svd->U + m * c;
        for (PetscInt j = 0; j < n; j++) {
            const PetscScalar vj = svd->V[j + n * c];
            const PetscScalar *a = A->a + A->lda * j;
            for (PetscInt r = 0; r < m; r++) u[r] += a[r] * vj;
        }
    }
    PetscErrorCode e = gs_orthogonalize(m, n, svd->U, NULL);
    free(G);
    free(W);
    free(lam);
    free(ord);
    return e;
}

PetscErrorCode SVDGetSingularTriplet(SVD svd, PetscInt i, PetscReal *sigma, Vec u, Vec v)
{
    if (!svd->sigma) SHIM_FAIL(PETSC_ERR_USER, "Must call SVDSolve() first");
    if (i < 0 || i >= svd->nsv) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "The index can be nconv-1 at most");
    if (sigma) *sigma = svd->sigma[i];
    if (u) {
        if (u->n != svd->m) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "left vector has the wrong size");
        memcpy(u->a, svd->U + svd->m * i, (size_t)svd->m * sizeof(PetscScalar));
    }
    if (v) {
        if (v->n != svd->n) SHIM_FAIL(PETSC_ERR_ARG_SIZ, "right vector has the wrong size");
        memcpy(v->a, svd->V + svd->n * i, (size_t)svd->n * sizeof(PetscScalar));
    }
    return 0;
}
