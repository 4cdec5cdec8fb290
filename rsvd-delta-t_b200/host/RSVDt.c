/*
 * RSVDt.c -- PETSc-free C host driver, drop-in for the reference executable:
 *
 *     RSVDt -inputs variables.yaml            (reference: README.md:221, SourceCode/ReadUserInput.c:20-22)
 *
 * Same YAML keys (variables.yaml:1-129), same PETSc-binary inputs (A, optional B/C/weights, F_hat, IC vector), same
 * results layout (ResolventModes_<i>/{S_hat,U_hat_*,V_hat_*}, TransientSnapshots_<i>/q_transient_*), same sequence as
 * SourceCode/main.c:107-150.  Each function below re-states the reference function of the same name and hands the
 * numerical work to the CUDA library through the C ABI of include/rsvdt.h -- the host never touches the RK4 loop.
 * Exit code: first non-zero error code, else 0 (main.c:107-150 with CHKERRQ).
 */
#define _POSIX_C_SOURCE 200809L
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/rsvdt.h"
#include "petsc_binary.h"

#define MAX_PATH_LEN 4096 /* PETSC_MAX_PATH_LEN */
#define ERR_USER RSVDT_ERR_USER

/* ---- Variables.h:5-110 ------------------------------------------------------------------------------------- */
typedef struct {
    int64_t N, Nb, Nc, k, q, Nw, Nw_eff;
    double w;
    int TwoPI, InputForcingFlg;
    int64_t RandSeed;
} RSVD_vars;
typedef struct {
    int64_t Ns, Nt, ResRatio;
    double TransientLength, dt;
    int RealOperator, TransientRemoval, DirAdj;
} TS_vars;
typedef struct { int DiscFlg; double beta; } Discounting;
typedef struct {
    RSVD_vars RSVD;
    TS_vars TS;
    Discounting Disc;
    int64_t Display, SaveResultsOpt;
} RSVDt_vars;
typedef struct {
    int TransRun, TransSave, TransRemovalEst, TransICFlg;
    int64_t TransSaveMod, TransPeriods;
    double TransDivVal, TransConVal;
} TransRun_vars;
typedef struct {
    int InvInputWeightFlg, OutputWeightFlg, InvOutputWeightFlg, InputMatrixFlg, OutputMatrixFlg;
} Weight_matrices;
typedef struct {
    char RootDir[MAX_PATH_LEN], ResultsDir[MAX_PATH_LEN], OperatorDir[MAX_PATH_LEN], IO_dir[2 * MAX_PATH_LEN + 64],
        FolderDir[2 * MAX_PATH_LEN + 64], InputForcingDir[MAX_PATH_LEN], InvInputWeightDir[MAX_PATH_LEN],
        OutputWeightDir[MAX_PATH_LEN], InvOutputWeightDir[MAX_PATH_LEN], InputMatrixDir[MAX_PATH_LEN],
        OutputMatrixDir[MAX_PATH_LEN], TransICDir[MAX_PATH_LEN], inputs[MAX_PATH_LEN];
} Directories;

static rsvdt_handle H = NULL;

static int seterr(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[RSVDt] ERROR: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    return code;
}
#define CHK(call)                                                                        \
    do {                                                                                 \
        int e__ = (call);                                                                \
        if (e__) {                                                                       \
            if (H && rsvdt_last_error(H)[0]) fprintf(stderr, "[RSVDt] %s\n", rsvdt_last_error(H)); \
            return e__;                                                                  \
        }                                                                                \
    } while (0)

static double wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static void print_hms(const char *prefix, double sec, const char *suffix)
{
    int hh = (int)(sec / 3600), mm = (int)((sec - 3600 * hh) / 60), ss = (int)(sec - 3600 * hh - 60 * mm);
    printf("%s%02d:%02d:%02d%s", prefix, hh, mm, ss, suffix);
}

/* ---- options database: "-key value" pairs from the YAML file (PetscOptionsInsertFileYAML, ReadUserInput.c:22) -- */
typedef struct { char key[128]; char val[MAX_PATH_LEN]; } opt_t;
static opt_t g_opts[256];
static int g_nopts = 0;

static void trim(char *s)
{
    char *p = s;
    while (*p && isspace((unsigned char)*p)) p++;
    if (p != s) memmove(s, p, strlen(p) + 1);
    size_t n = strlen(s);
    while (n > 0 && isspace((unsigned char)s[n - 1])) s[--n] = 0;
}

static int options_insert_file_yaml(const char *file)
{
    FILE *f = fopen(file, "r");
    if (!f) return seterr(ERR_USER, "Unable to open YAML options file %s", file);
    char line[2 * MAX_PATH_LEN];
    while (fgets(line, sizeof line, f)) {
        /* a comment starts at a '#' that opens the line or follows white space (the same rule as rsvdt_mpi.py: a '#' inside a
           value, e.g. a directory name, is kept) */
        for (char *c = line; *c; c++)
            if (*c == '#' && (c == line || isspace((unsigned char)c[-1]))) { *c = 0; break; }
        char *colon = strchr(line, ':');
        if (!colon) continue;
        *colon = 0;
        char key[128], val[MAX_PATH_LEN];
        snprintf(key, sizeof key, "%s", line);
        snprintf(val, sizeof val, "%s", colon + 1);
        trim(key);
        trim(val);
        if (!key[0] || g_nopts >= 256) continue;
        size_t n = strlen(val);
        if (n >= 2 && ((val[0] == '"' && val[n - 1] == '"') || (val[0] == '\'' && val[n - 1] == '\''))) {
            val[n - 1] = 0;
            memmove(val, val + 1, n - 1);
        }
        int found = -1;
        for (int i = 0; i < g_nopts; i++)
            if (!strcmp(g_opts[i].key, key)) found = i;
        if (found < 0) found = g_nopts++;
        snprintf(g_opts[found].key, sizeof g_opts[found].key, "%s", key);
        snprintf(g_opts[found].val, sizeof g_opts[found].val, "%s", val);
    }
    fclose(f);
    return 0;
}

static const char *opt_find(const char *key)
{
    for (int i = 0; i < g_nopts; i++)
        if (!strcmp(g_opts[i].key, key)) return g_opts[i].val;
    return NULL;
}
static int opt_real(const char *key, double *v, int *set)
{
    const char *s = opt_find(key);
    *set = 0;
    if (!s || !s[0]) return 0;
    char *end;
    errno = 0;
    double d = strtod(s, &end);
    if (end == s || *end || errno) return seterr(ERR_USER, "Input string '%s' for '%s' has no numeric value", s, key);
    *v = d;
    *set = 1;
    return 0;
}
static int opt_int(const char *key, int64_t *v, int *set)
{
    const char *s = opt_find(key);
    *set = 0;
    if (!s || !s[0]) return 0;
    char *end;
    errno = 0;
    long long d = strtoll(s, &end, 10);
    if (end == s || *end || errno) return seterr(ERR_USER, "Input string '%s' for '%s' has no integer value (do not include . in it)", s, key);
    *v = d;
    *set = 1;
    return 0;
}
static int opt_bool(const char *key, int *v, int *set)
{
    const char *s = opt_find(key);
    *set = 0;
    if (!s) return 0;
    if (!s[0]) { *v = 1; *set = 1; return 0; }   /* `Flag:` without a value is true, PETSc's rule (PetscOptionsGetBool) */
    if (!strcasecmp(s, "true") || !strcmp(s, "1") || !strcasecmp(s, "yes") || !strcasecmp(s, "on")) *v = 1;
    else if (!strcasecmp(s, "false") || !strcmp(s, "0") || !strcasecmp(s, "no") || !strcasecmp(s, "off")) *v = 0;
    else return seterr(ERR_USER, "Unknown logical value '%s' for '%s'", s, key);
    *set = 1;
    return 0;
}
static int opt_string(const char *key, char *dst, int *set)
{
    const char *s = opt_find(key);
    *set = 0;
    if (!s) return 0;
    snprintf(dst, MAX_PATH_LEN, "%s", s);
    *set = 1;
    return 0;
}

/* ---- ReadUserInput.c:20-173 ------------------------------------------------------------------------------------ */
static int flag_with_dir(const char *flag, const char *dirkey, const char *warn_name, int *val, char *dir)
{
    int set;
    CHK(opt_bool(flag, val, &set));
    if (!set) {
        *val = 0;
        printf("Warning: '%s' variable not found. Setting '%s' to default value: %d\n", flag, warn_name, *val);
    } else if (*val) {
        CHK(opt_string(dirkey, dir, &set));
        if (!set) return seterr(ERR_USER, "Must specify '%s'", dirkey);
    }
    return 0;
}

static int ReadUserInput(int argc, char **argv, RSVDt_vars *RSVDt, Weight_matrices *Weight_mat, TransRun_vars *TR_vars,
                         Directories *dirs)
{
    int set;
    printf("\n*** Reading user inputs ***\n");
    dirs->inputs[0] = 0;
    for (int i = 1; i + 1 < argc; i++)
        if (!strcmp(argv[i], "-inputs")) snprintf(dirs->inputs, MAX_PATH_LEN, "%s", argv[i + 1]);
    if (!dirs->inputs[0]) return seterr(ERR_USER, "Must specify inputs variables via -inputs");           /* :21 */
    CHK(options_insert_file_yaml(dirs->inputs));                                                           /* :22 */
    CHK(opt_real("TransientLength", &RSVDt->TS.TransientLength, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'TransientLength'");                                    /* :24 */
    CHK(opt_real("dt", &RSVDt->TS.dt, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'dt'");                                                 /* :26 */
    CHK(opt_real("w", &RSVDt->RSVD.w, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'dt'");                                                 /* :28 (sic: the reference says 'dt') */
    CHK(flag_with_dir("InputForcingFlg", "InputForcingDir", "InputForcingFlg", &RSVDt->RSVD.InputForcingFlg, dirs->InputForcingDir));
    CHK(flag_with_dir("InvInputWeightFlg", "InvInputWeightDir", "InputWeightFlg", &Weight_mat->InvInputWeightFlg, dirs->InvInputWeightDir));
    CHK(flag_with_dir("OutputWeightFlg", "OutputWeightDir", "OutputWeightFlg", &Weight_mat->OutputWeightFlg, dirs->OutputWeightDir));
    CHK(flag_with_dir("InvOutputWeightFlg", "InvOutputWeightDir", "InvOutputWeightFlg", &Weight_mat->InvOutputWeightFlg, dirs->InvOutputWeightDir));
    CHK(flag_with_dir("InputMatrixFlg", "InputMatrixDir", "InputMatrixFlg", &Weight_mat->InputMatrixFlg, dirs->InputMatrixDir));
    CHK(flag_with_dir("OutputMatrixFlg", "OutputMatrixDir", "OutputMatrixFlg", &Weight_mat->OutputMatrixFlg, dirs->OutputMatrixDir));
    CHK(opt_bool("TwoPI", &RSVDt->RSVD.TwoPI, &set));
    if (!set) { RSVDt->RSVD.TwoPI = 0; printf("Warning: 'TwoPI' variable not found. Setting 'TwoPI' to default value: %d\n", 0); }
    RSVDt->RSVD.w *= RSVDt->RSVD.TwoPI ? 2 * 3.14159265358979323846264338327950288 : 1;                     /* :82 */
    CHK(opt_bool("TransRun", &TR_vars->TransRun, &set));
    if (!set) { TR_vars->TransRun = 0; printf("Warning: 'TransRun' variable not found. Setting 'TransRun' to default value: %d\n", 0); }
    CHK(opt_bool("TransientRemoval", &RSVDt->TS.TransientRemoval, &set));
    if (!set) { RSVDt->TS.TransientRemoval = 0; printf("Warning: 'TransientRemoval' variable not found. Setting 'TransientRemoval' to default value: %d\n", 0); }
    CHK(opt_bool("DiscFlg", &RSVDt->Disc.DiscFlg, &set));
    if (!set) { RSVDt->Disc.DiscFlg = 0; printf("Warning: 'DiscFlg' variable not found. Setting 'DiscFlg' to default value: %d\n", 0); }
    CHK(opt_int("SaveResultsOpt", &RSVDt->SaveResultsOpt, &set));
    if (!set) { RSVDt->SaveResultsOpt = 1; printf("Warning: 'SaveResultsOpt' variable not found. Setting 'SaveResultsOpt' to default value: %d\n", 1); }
    CHK(opt_string("RootDir", dirs->RootDir, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'RootDir'");
    CHK(opt_string("ResultsDir", dirs->ResultsDir, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'ResultsDir'");
    CHK(opt_string("OperatorDir", dirs->OperatorDir, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'OperatorDir'");
    CHK(opt_int("k", &RSVDt->RSVD.k, &set));
    if (!set) { RSVDt->RSVD.k = 3; printf("Warning: 'k' variable not found. Setting 'k' to default value: %d\n", 3); }
    if (RSVDt->RSVD.k > RSVDT_MAX_K)   /* (not in the reference: a limit of this build, include/rsvdt.h; said before anything is loaded) */
        return seterr(ERR_USER, "'k' = %d: this build supports at most %d test vectors", RSVDt->RSVD.k, RSVDT_MAX_K);
    CHK(opt_int("q", &RSVDt->RSVD.q, &set));
    if (!set) { RSVDt->RSVD.q = 0; printf("Warning: 'q' variable not found. Setting 'q' to default value: %d\n", 0); }
    CHK(opt_int("Nw", &RSVDt->RSVD.Nw, &set));
    if (!set) return seterr(ERR_USER, "Must specify 'Nw'");
    CHK(opt_int("RandSeed", &RSVDt->RSVD.RandSeed, &set));
    if (!set) { RSVDt->RSVD.RandSeed = 1373; printf("Warning: 'RandSeed' variable not found. Setting 'RandSeed' to default value: %d\n", 1373); }
    CHK(opt_int("Display", &RSVDt->Display, &set));
    if (!set) { RSVDt->Display = 2; printf("Warning: 'Display' variable not found. Setting 'Display' to default value: %d\n", 2); }
    CHK(opt_real("beta", &RSVDt->Disc.beta, &set));
    if (!set && RSVDt->Disc.DiscFlg) return seterr(ERR_USER, "Discounting flag is on! Either set Discounting to zero or specify 'beta'");
    if (RSVDt->Disc.beta < 0 && RSVDt->Disc.DiscFlg) return seterr(ERR_USER, "'beta' must be positive, current value: %g", RSVDt->Disc.beta);
    CHK(opt_bool("TransICFlg", &TR_vars->TransICFlg, &set));
    if (!set && TR_vars->TransRun) {
        TR_vars->TransICFlg = 0;
        printf("Warning: 'TransICFlg' variable not found. Setting 'TransICFlg' to default value: %d\n", 0);
    } else {
        /* the reference requires TransICDir whenever this else-branch is reached (ReadUserInput.c:134-141), even when
         * TransRun is false; it is only *used* when TransRun && TransICFlg.  We keep the lookup but only fail when the
         * value is actually needed -- a conscious fix of that quirk. */
        CHK(opt_string("TransICDir", dirs->TransICDir, &set));
        if (!set && TR_vars->TransRun && TR_vars->TransICFlg) return seterr(ERR_USER, "Must specify 'TransICDir'");
    }
    CHK(opt_real("TransDivVal", &TR_vars->TransDivVal, &set));
    if (!set) { TR_vars->TransDivVal = 1e3; if (TR_vars->TransRun) printf("Warning: 'TransDivVal' variable not found. Setting 'TransDivVal' to default value: %g\n", 1e3); }
    CHK(opt_real("TransConVal", &TR_vars->TransConVal, &set));
    if (!set) { TR_vars->TransConVal = 1e-8; if (TR_vars->TransRun) printf("Warning: 'TransConVal' variable not found. Setting 'TransConVal' to default value: %g\n", 1e-8); }
    CHK(opt_bool("TransSave", &TR_vars->TransSave, &set));
    if (!set) { TR_vars->TransSave = 0; if (TR_vars->TransRun) printf("Warning: 'TransSave' variable not found. Setting 'TransSave' to default value: %d\n", 0); }
    CHK(opt_bool("TransRemovalEst", &TR_vars->TransRemovalEst, &set));
    if (!set) { TR_vars->TransRemovalEst = 0; if (TR_vars->TransRun) printf("Warning: 'TransRemovalEst' variable not found. Setting 'TransRemovalEst' to default value: %d\n", 0); }
    CHK(opt_int("TransSaveMod", &TR_vars->TransSaveMod, &set));
    if (!set) { TR_vars->TransSaveMod = 100; if (TR_vars->TransRun) printf("Warning: 'TransSaveMod' variable not found. Setting 'TransSaveMod' to default value: %d\n", 100); }
    CHK(opt_int("TransPeriods", &TR_vars->TransPeriods, &set));
    if (!set) { TR_vars->TransPeriods = 1; if (TR_vars->TransRun) printf("Warning: 'TransPeriods' variable not found. Setting 'TransPeriods' to default value: %d\n", 1); }
    printf("\n");
    return 0;
}

/* ---- SetupTimeFreqGrid.c:16-40 ----------------------------------------------------------------------------------- */
static int SetupTimeFreqGrid(RSVDt_vars *RSVDt, TransRun_vars *TR_vars)
{
    double dt;
    int64_t Ns, Nt, RR;
    if (rsvdt_setup_time_freq_grid(RSVDt->RSVD.w, RSVDt->RSVD.Nw, RSVDt->TS.dt, RSVDt->TS.TransientLength, &dt, &Ns, &Nt, &RR))
        return seterr(RSVDT_ERR_ARG, "invalid w / Nw / dt");
    RSVDt->TS.dt = dt; RSVDt->TS.Ns = Ns; RSVDt->TS.Nt = Nt; RSVDt->TS.ResRatio = RR;
    if (!TR_vars->TransRun) {
        printf("RSVD parameters:\n@ k        = %d\n@ q        = %d\n@ w        = %g\n\n", (int)RSVDt->RSVD.k, (int)RSVDt->RSVD.q, RSVDt->RSVD.w);
        printf("Time-stepping parameters:\n@ TransientLength = %g\n@ dt              = %g\n@ Ns              = %d\n@ Nt              = %d\n",
               RSVDt->TS.TransientLength, RSVDt->TS.dt, (int)RSVDt->TS.Ns, (int)RSVDt->TS.Nt);
        printf(RSVDt->TS.TransientRemoval ? "Transient removal strategy is requested\n\n" : "Transient removal strategy is NOT requested\n\n");
    }
    return 0;
}

/* ---- LNSStructure.c:19-58 ----------------------------------------------------------------------------------------- */
static int LNSStructure(RSVDt_vars *RSVDt, TransRun_vars *TR_vars, Directories *dirs)
{
    double t1 = wtime();
    printf("Loading the operator: %s%s \n", dirs->RootDir, dirs->OperatorDir);
    snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->RootDir, dirs->OperatorDir);
    pb_csr A;
    if (pb_read_mat(dirs->IO_dir, &A)) return seterr(ERR_USER, "%s", pb_error());
    if (A.nrows != A.ncols) { pb_free_csr(&A); return seterr(RSVDT_ERR_ARG, "the operator must be square (%lld x %lld)", (long long)A.nrows, (long long)A.ncols); }
    RSVDt->RSVD.N = A.nrows;
    int real_op = 0;
    int rc = rsvdt_set_operator_csr(H, A.nrows, A.rowptr, A.col, A.val, RSVDt->Disc.DiscFlg, RSVDt->Disc.beta, &real_op);
    pb_free_csr(&A);
    CHK(rc);
    printf("The operator is loaded in %f seconds (N = %d) \n", wtime() - t1, (int)RSVDt->RSVD.N);
    RSVDt->TS.RealOperator = real_op;
    if (!TR_vars->TransRun) printf(real_op ? "\nThe LNS operator is real-valued!\n" : "\nThe LNS operator is complex-valued!\n");
    RSVDt->RSVD.Nw_eff = real_op ? RSVDt->RSVD.Nw / 2 : RSVDt->RSVD.Nw;                                       /* :44 */
    if (!TR_vars->TransRun) printf("The effective number of frequencies, Nw = %d\n", (int)RSVDt->RSVD.Nw_eff);
    if (RSVDt->Disc.DiscFlg) printf("\n---- Discounting with beta = %g ----\n\n", RSVDt->Disc.beta);         /* :53-56 */
    return 0;
}

/* ---- ReadWeightMats.c:26-99 ------------------------------------------------------------------------------------------ */
static int load_weight(int which, const char *label, const char *rel, Directories *dirs, int64_t *rows, int64_t *cols)
{
    snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->RootDir, rel);
    printf("Reading the %s: %s\n", label, dirs->IO_dir);
    pb_csr M;
    if (pb_read_mat(dirs->IO_dir, &M)) return seterr(ERR_USER, "%s", pb_error());
    *rows = M.nrows;
    *cols = M.ncols;
    int rc = rsvdt_set_weight_csr(H, which, M.nrows, M.ncols, M.rowptr, M.col, M.val);
    pb_free_csr(&M);
    return rc;
}

static int ReadWeightMats(RSVDt_vars *RSVDt, Weight_matrices *W, Directories *dirs)
{
    const int64_t N = RSVDt->RSVD.N;
    int64_t row1 = N, col1 = N, row2, col2, rq = -1, cq;
    if (W->InvInputWeightFlg) CHK(load_weight(0, "inverse input weight matrix  ", dirs->InvInputWeightDir, dirs, &row1, &col1));
    if (W->InputMatrixFlg) {
        CHK(load_weight(1, "input matrix                 ", dirs->InputMatrixDir, dirs, &row2, &col2));
        if (col2 != col1) return seterr(ERR_USER, "Size mismatch between input matrix (B) and input weight matrix (W_f_sqrt_inv), %d != %d", (int)col2, (int)col1);
        RSVDt->RSVD.Nb = col2;
    } else {
        RSVDt->RSVD.Nb = N;
        if (W->InvInputWeightFlg && RSVDt->RSVD.Nb != col1) return seterr(ERR_USER, "Size mismatch between input matrix (B) and input weight matrix (W_f_sqrt_inv)");
    }
    if (W->InvOutputWeightFlg) CHK(load_weight(2, "inverse output weight matrix ", dirs->InvOutputWeightDir, dirs, &rq, &cq));
    row1 = N;
    if (W->OutputWeightFlg) CHK(load_weight(3, "output weight matrix         ", dirs->OutputWeightDir, dirs, &row1, &col1));
    if (W->OutputMatrixFlg) {
        CHK(load_weight(4, "output matrix                ", dirs->OutputMatrixDir, dirs, &row2, &col2));
        RSVDt->RSVD.Nc = row2;
        if (row2 != row1) return seterr(ERR_USER, "Size mismatch between output matrix (C) and output weight matrix (W_q_sqrt), %d != %d", (int)row2, (int)row1);
        if (W->InvOutputWeightFlg && row2 != rq) return seterr(ERR_USER, "Size mismatch between output matrix (C) and inverse output weight matrix (W_q_sqrt_inv), %d != %d", (int)row2, (int)rq);
    } else {
        RSVDt->RSVD.Nc = N;
        if (W->OutputWeightFlg && RSVDt->RSVD.Nc != row1) return seterr(ERR_USER, "Size mismatch between output matrix (C) and output weight matrix (W_q_sqrt)");
        if (W->InvOutputWeightFlg && RSVDt->RSVD.Nc != rq) return seterr(ERR_USER, "Size mismatch between output matrix (C) and inverse output weight matrix (W_q_sqrt_inv)");
    }
    return 0;
}

/* ---- CreateRandomMat.c:23-40 ---------------------------------------------------------------------------------------- */
static uint64_t g_rng;
static double rng_next(void) /* splitmix64 -> uniform [0,1) */
{
    uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static int push_grid(RSVDt_vars *RSVDt)
{
    rsvdt_grid g;
    g.k = RSVDt->RSVD.k; g.Nw = RSVDt->RSVD.Nw; g.Nw_eff = RSVDt->RSVD.Nw_eff; g.Ns = RSVDt->TS.Ns; g.Nt = RSVDt->TS.Nt;
    g.ResRatio = RSVDt->TS.ResRatio; g.dt = RSVDt->TS.dt; g.w = RSVDt->RSVD.w; g.RealOperator = RSVDt->TS.RealOperator;
    g.TransientRemoval = RSVDt->TS.TransientRemoval;
    return rsvdt_set_grid(H, &g);
}

static int CreateRandomMat(RSVDt_vars *RSVDt, Directories *dirs)
{
    const int64_t Nb = RSVDt->RSVD.Nb;
    double *Y = NULL;
    if (RSVDt->RSVD.InputForcingFlg) {
        snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->RootDir, dirs->InputForcingDir);
        printf("Reading the input forcing matrix         : %s\n", dirs->IO_dir);
        int64_t row, col;
        if (pb_read_mat_dense(dirs->IO_dir, &row, &col, &Y)) return seterr(ERR_USER, "%s", pb_error());
        if (row != Nb) { free(Y); return seterr(ERR_USER, "Input forcing must be have %d rows, current size = %d x %d", (int)Nb, (int)row, (int)col); }
        RSVDt->RSVD.k = col / RSVDt->RSVD.Nw_eff;                                                            /* :30 */
        if (RSVDt->RSVD.k > RSVDT_MAX_K) { free(Y); return seterr(ERR_USER, "the input forcing holds k = %d test vectors: this build supports at most %d", RSVDt->RSVD.k, RSVDT_MAX_K); }
        printf("Based on the input forcing: k = %d\n", (int)RSVDt->RSVD.k);
        if (RSVDt->RSVD.k * RSVDt->RSVD.Nw_eff != col) { free(Y); return seterr(ERR_USER, "Input forcing columns (%d) must be a multiple of Nw_eff (%d)", (int)col, (int)RSVDt->RSVD.Nw_eff); }
    } else {
        printf("Generating a random forcing matrix\n");
        /* The reference seeds a PetscRandom but then passes NULL to MatSetRandom (CreateRandomMat.c:36-39), so RandSeed
         * does not control its forcing.  Here RandSeed does: entries uniform [0,1) + i[0,1) like PETSc's complex default. */
        const int64_t cols = RSVDt->RSVD.Nw_eff * RSVDt->RSVD.k;
        Y = (double *)malloc((size_t)Nb * cols * 16);
        if (!Y) return seterr(RSVDT_ERR_MEM, "out of memory");
        g_rng = 0x853c49e6748fea9bull ^ (uint64_t)RSVDt->RSVD.RandSeed;
        for (int64_t i = 0; i < 2 * Nb * cols; i++) Y[i] = rng_next();
    }
    int rc = push_grid(RSVDt);
    if (!rc) rc = rsvdt_set_yhat(H, Y, Nb);
    free(Y);
    return rc;
}

/* ---- CreateResultsDir.c:18-29 ----------------------------------------------------------------------------------------- */
static int CreateResultsDir(Directories *dirs, const char *FileName)
{
    int FolderInd = 0;
    snprintf(dirs->FolderDir, sizeof dirs->FolderDir, "%s%s%s%d/", dirs->RootDir, dirs->ResultsDir, FileName, FolderInd);
    while (access(dirs->FolderDir, F_OK) == 0) {
        FolderInd++;
        snprintf(dirs->FolderDir, sizeof dirs->FolderDir, "%s%s%s%d/", dirs->RootDir, dirs->ResultsDir, FileName, FolderInd);
    }
    if (mkdir(dirs->FolderDir, 0777)) return seterr(ERR_USER, "cannot create %s: %s", dirs->FolderDir, strerror(errno));
    return 0;
}

/* ---- SaveInputVarsCopy.c:27-37 ------------------------------------------------------------------------------------------ */
static int SaveInputVarsCopy(Directories *dirs)
{
    char dst[3 * MAX_PATH_LEN];
    const char *base = strrchr(dirs->inputs, '/');
    base = base ? base + 1 : dirs->inputs;
    snprintf(dst, sizeof dst, "%s%s", dirs->FolderDir, base);
    FILE *s = fopen(dirs->inputs, "r"), *d = fopen(dst, "w");
    if (!s || !d) { if (s) fclose(s); if (d) fclose(d); return seterr(ERR_USER, "cannot copy %s to %s", dirs->inputs, dst); }
    char buffer[1024];
    size_t n;
    while ((n = fread(buffer, 1, sizeof buffer, s)) > 0) fwrite(buffer, 1, n, d);
    fclose(s);
    fclose(d);
    return 0;
}

/* ---- TransientRunRK4.c:11-228 ----------------------------------------------------------------------------------------- */
typedef struct { Directories *dirs; int save; } snap_ctx;
static int snapshot_cb(int64_t pos, const double *q, int64_t N, void *user)
{
    snap_ctx *c = (snap_ctx *)user;
    if (!c->save) return 0;
    snprintf(c->dirs->IO_dir, sizeof c->dirs->IO_dir, "%s%s%d", c->dirs->FolderDir, "q_transient_", (int)pos);   /* :161 */
    return pb_write_vec(c->dirs->IO_dir, N, q);
}

/* ---- TransientRemovalEst.c: how well could the transient be removed, period by period (diagnostic, `TransRemovalEst`) ---------
 * The transient run keeps the last Nw snapshots, one every Nt_period / Nw steps, in Q_all (TransientRunRK4.c:174-184; here
 * through the second callback cadence of rsvdt_transient_run_rk4_est) and at the end of every period computes
 *   "initial" norms  || Q_all dft[:, i] / Nw ||                        (TransientRemovalEst.c:58-71)
 *   "updated" norms  || q_ss - U U^H q_ss || / || q_ss ||,  U = orth(q_ss e^{i w_i s dT} + Q_all[:, s])     (:75-119)
 * for the non-negative frequencies w_i = i w, q_ss = -q0.  Host arithmetic on Nw vectors (complex as (re, im) pairs). */
typedef struct {
    int64_t N, Nw, Nw_eff, Nt_period, est_mod;
    double w;
    double *Q_all, *qss;      /* N x Nw column-major, N */
    Directories *dirs;
    char *text[64];           /* what the estimate of a period prints, shown in step order after the run */
    int64_t text_step[64];
    int ntext;
} est_ctx;

static void est_printf(char **buf, size_t *len, size_t *cap, const char *fmt, ...)
{
    va_list ap;
    char line[256];
    va_start(ap, fmt);
    int n = vsnprintf(line, sizeof line, fmt, ap);
    va_end(ap);
    if (n < 0) return;
    if (*len + (size_t)n + 1 > *cap) {
        *cap = 2 * (*cap + (size_t)n + 1);
        *buf = (char *)realloc(*buf, *cap);
    }
    memcpy(*buf + *len, line, (size_t)n + 1);
    *len += (size_t)n;
}

static double cnorm2(const double *x, int64_t n)
{
    double s = 0;
    for (int64_t i = 0; i < 2 * n; i++) s += x[i] * x[i];
    return s;
}

static int TransientRemovalEst(est_ctx *c, int64_t period_index, int64_t step)
{
    const int64_t N = c->N, Nw = c->Nw, nhalf = Nw / 2;
    const double PI = 3.14159265358979323846264338327950288;
    const double deltaT = 2 * PI / c->w / (double)Nw;                                                         /* :26 */
    const int64_t skip = period_index == 1 ? 1 : 0, Ns = Nw - skip;                                           /* :28 */
    char *buf = NULL;
    size_t len = 0, cap = 0;
    est_printf(&buf, &len, &cap, "**** deltaT = %g ****\n", deltaT);
    est_printf(&buf, &len, &cap, "**** Subspace dim = %d ****\n", (int)Ns);
    est_printf(&buf, &len, &cap, "**** period index = %d ****\n", (int)period_index);
    est_printf(&buf, &len, &cap, "**** Displaying transient norms for (positive) frequencies ****\n");
    double *init = (double *)calloc((size_t)nhalf * 2, sizeof(double)), *upd = (double *)calloc((size_t)nhalf * 2, sizeof(double));
    double *Y = (double *)malloc((size_t)N * (size_t)Ns * 16), *d = (double *)malloc((size_t)N * 16);
    if (!init || !upd || !Y || !d) { free(init); free(upd); free(Y); free(d); free(buf); return seterr(RSVDT_ERR_MEM, "out of memory"); }
    for (int64_t i = 0; i < nhalf && i < c->Nw_eff; i++) {                                                    /* :58-71 */
        memset(d, 0, (size_t)N * 16);
        for (int64_t s = 0; s < Nw; s++) {
            const double ph = -2 * PI * (double)s * (double)i / (double)Nw, cr = cos(ph) / (double)Nw, ci = sin(ph) / (double)Nw;
            const double *q = c->Q_all + (size_t)s * (size_t)N * 2;
            for (int64_t r = 0; r < N; r++) {
                d[2 * r] += q[2 * r] * cr - q[2 * r + 1] * ci;
                d[2 * r + 1] += q[2 * r] * ci + q[2 * r + 1] * cr;
            }
        }
        init[2 * i] = sqrt(cnorm2(d, N));
        est_printf(&buf, &len, &cap, "**** Norm of the ratio norm(qt)/norm(qs): %g @ w = %g ****\n", init[2 * i], c->w * (double)i);
    }
    est_printf(&buf, &len, &cap, "**** Displaying updated transient norms for (positive) frequencies ****\n");
    const double qn = sqrt(cnorm2(c->qss, N));
    for (int64_t i = 0; i < nhalf; i++) {
        const double wi = c->w * (double)i;
        for (int64_t s = 0; s < Ns; s++) {                                                                    /* :86-98 */
            const double ph = wi * (double)s * deltaT, cr = cos(ph), ci = sin(ph);
            const double *q = c->Q_all + (size_t)(s + skip) * (size_t)N * 2;
            double *y = Y + (size_t)s * (size_t)N * 2;
            for (int64_t r = 0; r < N; r++) {
                y[2 * r] = q[2 * r] + c->qss[2 * r] * cr - c->qss[2 * r + 1] * ci;
                y[2 * r + 1] = q[2 * r + 1] + c->qss[2 * r] * ci + c->qss[2 * r + 1] * cr;
            }
        }
        /* orthonormal basis of the columns (QRDecomposition.c: BVOrthogonalize): Gram-Schmidt, every column swept twice */
        for (int64_t s = 0; s < Ns; s++) {
            double *y = Y + (size_t)s * (size_t)N * 2;
            for (int pass = 0; pass < 2; pass++)
                for (int64_t t = 0; t < s; t++) {
                    const double *u = Y + (size_t)t * (size_t)N * 2;
                    double pr = 0, pi = 0;   /* u^H y */
                    for (int64_t r = 0; r < N; r++) { pr += u[2 * r] * y[2 * r] + u[2 * r + 1] * y[2 * r + 1]; pi += u[2 * r] * y[2 * r + 1] - u[2 * r + 1] * y[2 * r]; }
                    for (int64_t r = 0; r < N; r++) { y[2 * r] -= u[2 * r] * pr - u[2 * r + 1] * pi; y[2 * r + 1] -= u[2 * r] * pi + u[2 * r + 1] * pr; }
                }
            const double nr = sqrt(cnorm2(y, N));
            const double inv = nr > 0 ? 1.0 / nr : 0.0;
            for (int64_t r = 0; r < 2 * N; r++) y[r] *= inv;
        }
        memcpy(d, c->qss, (size_t)N * 16);                                                                    /* :106-111: q_ss - U U^H q_ss */
        for (int pass = 0; pass < 2; pass++)
            for (int64_t t = 0; t < Ns; t++) {
                const double *u = Y + (size_t)t * (size_t)N * 2;
                double pr = 0, pi = 0;
                for (int64_t r = 0; r < N; r++) { pr += u[2 * r] * d[2 * r] + u[2 * r + 1] * d[2 * r + 1]; pi += u[2 * r] * d[2 * r + 1] - u[2 * r + 1] * d[2 * r]; }
                for (int64_t r = 0; r < N; r++) { d[2 * r] -= u[2 * r] * pr - u[2 * r + 1] * pi; d[2 * r + 1] -= u[2 * r] * pi + u[2 * r + 1] * pr; }
            }
        upd[2 * i] = sqrt(cnorm2(d, N)) / qn;
        est_printf(&buf, &len, &cap, "**** Norm of the ratio norm(qt_up)/norm(qs): %g @ w = %g ****\n", upd[2 * i], wi);
    }
    int e = 0;
    snprintf(c->dirs->IO_dir, sizeof c->dirs->IO_dir, "%s%s%d", c->dirs->FolderDir, "Initial_transient_norm_period_", (int)period_index);   /* :123 */
    e = pb_write_vec(c->dirs->IO_dir, nhalf, init);
    snprintf(c->dirs->IO_dir, sizeof c->dirs->IO_dir, "%s%s%d", c->dirs->FolderDir, "Updated_transient_norm_period_", (int)period_index);   /* :127 */
    if (!e) e = pb_write_vec(c->dirs->IO_dir, nhalf, upd);
    free(init); free(upd); free(Y); free(d);
    if (c->ntext < 64) { c->text[c->ntext] = buf; c->text_step[c->ntext] = step; c->ntext++; } else free(buf);
    return e;
}

static int est_cb(int64_t pos_e, const double *q, int64_t N, void *user)
{
    est_ctx *c = (est_ctx *)user;
    const int64_t step = pos_e * c->est_mod, pos = pos_e % c->Nw;                                             /* TransientRunRK4.c:176-183 */
    if (pos == 0 && TransientRemovalEst(c, step / c->Nt_period, step)) return 1;
    memcpy(c->Q_all + (size_t)pos * (size_t)N * 2, q, (size_t)N * 16);
    return 0;
}

static int TransientRunRK4(TransRun_vars *TR_vars, RSVDt_vars *RSVDt, Directories *dirs)
{
    printf("\n********************************************\n************** Transient run ***************\n********************************************\n\n");
    const int64_t N = RSVDt->RSVD.N, Ns = RSVDt->RSVD.Nw;
    const double TSS = 2 * 3.14159265358979323846264338327950288 / RSVDt->RSVD.w;                             /* :40 */
    const double deltaT = TSS / (double)Ns;
    RSVDt->TS.dt = deltaT / ceil(deltaT / RSVDt->TS.dt);                                                      /* :42 */
    const int64_t Nt_period = (int64_t)round(TSS / RSVDt->TS.dt);
    const int64_t rend = Nt_period * TR_vars->TransPeriods;                                                   /* :45 */
    const int64_t Nt_saved = rend / TR_vars->TransSaveMod + 1;                                                /* :46 */
    printf("Total integration time = %g = %d x %g with dt = %g\n", RSVDt->TS.dt * rend, (int)TR_vars->TransPeriods, TSS, RSVDt->TS.dt);
    if (TR_vars->TransSave) printf("Total snapshots to be saved = %d\n", (int)Nt_saved);
    else printf("Saving snapshots is not requested\n");
    /* the grid is only needed for dt here */
    RSVDt->RSVD.Nw_eff = RSVDt->TS.RealOperator ? RSVDt->RSVD.Nw / 2 : RSVDt->RSVD.Nw;
    RSVDt->TS.Ns = Nt_period; RSVDt->TS.Nt = 0; RSVDt->TS.ResRatio = Nt_period / Ns > 0 ? Nt_period / Ns : 1;
    if (RSVDt->RSVD.k < 1) RSVDt->RSVD.k = 1;
    CHK(push_grid(RSVDt));
    double *q0 = NULL;
    if (TR_vars->TransICFlg) {
        printf("Reading the initial vector\n");
        snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s%s", dirs->RootDir, dirs->ResultsDir, dirs->TransICDir);   /* :84 */
        int64_t n;
        if (pb_read_vec(dirs->IO_dir, &n, &q0)) return seterr(ERR_USER, "%s", pb_error());
        if (n != N) { free(q0); return seterr(ERR_USER, "initial vector has %d entries, expected %d", (int)n, (int)N); }
    } else {
        printf("Random vector is generated\n\n");
        q0 = (double *)malloc((size_t)N * 16);
        if (!q0) return seterr(RSVDT_ERR_MEM, "out of memory");
        g_rng = 0x853c49e6748fea9bull ^ (uint64_t)RSVDt->RSVD.RandSeed;                                       /* :89-92 */
        double nrm = 0;
        for (int64_t i = 0; i < 2 * N; i++) { q0[i] = rng_next(); nrm += q0[i] * q0[i]; }
        nrm = sqrt(nrm);
        for (int64_t i = 0; i < 2 * N; i++) q0[i] /= nrm;                                                     /* :93 VecNormalize */
    }
    double *norms = (double *)calloc((size_t)Nt_saved, sizeof(double));
    if (!norms) { free(q0); return seterr(RSVDT_ERR_MEM, "out of memory"); }
    snap_ctx ctx = {dirs, TR_vars->TransSave};
    if (TR_vars->TransSave) { int e = snapshot_cb(0, q0, N, &ctx); if (e) { free(q0); free(norms); return seterr(ERR_USER, "%s", pb_error()); } }   /* :116-120 */
    int64_t steps_done = 0;
    int reason = 0;
    est_ctx ec;
    memset(&ec, 0, sizeof ec);
    if (TR_vars->TransRemovalEst) {                                                                           /* :56-65, :96-99, :110-114 */
        ec.N = N; ec.Nw = Ns; ec.Nw_eff = RSVDt->RSVD.Nw_eff; ec.Nt_period = Nt_period; ec.est_mod = Nt_period / Ns;
        ec.w = RSVDt->RSVD.w; ec.dirs = dirs;
        ec.Q_all = (double *)calloc((size_t)N * (size_t)Ns * 2, sizeof(double));
        ec.qss = (double *)malloc((size_t)N * 16);
        if (!ec.Q_all || !ec.qss || ec.est_mod <= 0) { free(ec.Q_all); free(ec.qss); free(q0); free(norms); return seterr(RSVDT_ERR_MEM, "out of memory"); }
        for (int64_t i = 0; i < 2 * N; i++) ec.qss[i] = -q0[i];
        memcpy(ec.Q_all, q0, (size_t)N * 16);
    }
    double t1 = wtime();
    int rc = rsvdt_transient_run_rk4_est(H, 1, q0, rend, TR_vars->TransSaveMod, TR_vars->TransDivVal, TR_vars->TransConVal, norms,
                                         &steps_done, &reason, snapshot_cb, &ctx, ec.est_mod, TR_vars->TransRemovalEst ? est_cb : NULL, &ec);
    free(ec.Q_all); free(ec.qss);
    if (rc) { for (int i = 0; i < ec.ntext; i++) free(ec.text[i]); free(q0); free(norms); CHK(rc); }
    printf("Norm of the initial snapshot: %f @ t = %d\n", norms[0], 0);
    const int64_t nchecks = steps_done / TR_vars->TransSaveMod;
    int it = 0;   /* the estimate of a step follows the norm line of that step (TransientRunRK4.c:137-184) */
    for (int64_t p = 1; p <= nchecks; p++) {
        for (; it < ec.ntext && ec.text_step[it] < p * TR_vars->TransSaveMod; it++) { fputs(ec.text[it], stdout); free(ec.text[it]); }
        if (!(reason && p == nchecks)) printf("Norm of the snapshot: %f @ t = %f\n", norms[p], (double)(p * TR_vars->TransSaveMod) * RSVDt->TS.dt);
    }
    for (; it < ec.ntext; it++) { fputs(ec.text[it], stdout); free(ec.text[it]); }
    if (reason == 1) printf("Stopped! The norm of the last snapshot is NAN!\n");
    if (reason == 2) printf("Diverged! The norm of the last snapshot exceeds the divergence threshold = %g\n", TR_vars->TransDivVal);
    if (reason == 3) printf("Converged! The norm of the last snapshot is smaller than the convergence threshold = %g\n", TR_vars->TransConVal);
    print_hms("*** Transient simulation elapsed time = ", wtime() - t1, " ***\n");
    printf("\nThe results directory is: %s\n\n", dirs->FolderDir);
    /* q_transient_norms: a Vec whose entries are the norms (real parts), :201-205 */
    double *nv = (double *)calloc((size_t)Nt_saved * 2, sizeof(double));
    for (int64_t i = 0; i < Nt_saved; i++) nv[2 * i] = norms[i];
    snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->FolderDir, "q_transient_norms");
    int e = pb_write_vec(dirs->IO_dir, Nt_saved, nv);
    snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->FolderDir, "q_transient_last_snapshot");         /* :207-209 */
    if (!e) e = pb_write_vec(dirs->IO_dir, N, q0);
    free(nv); free(norms); free(q0);
    if (e) return seterr(ERR_USER, "%s", pb_error());
    CHK(SaveInputVarsCopy(dirs));                                                                             /* :222 */
    printf("DONE :))\n\n");
    return 0;
}

/* ---- the actions: banners of TSActionRK4.c:31-32,150-160; numerics behind the C ABI ------------------------------------- */
static int DirectActionRK4(RSVDt_vars *RSVDt)
{
    RSVDt->TS.DirAdj = 1;                                                                                     /* DirectActionRK4.c:24 */
    printf("\n*** Direct action begins! ***\n");
    double t0 = wtime();
    CHK(rsvdt_direct_action_rk4(H));
    print_hms("*** Direct action elapsed time = ", wtime() - t0, " ***\n");
    return 0;
}
static int AdjointActionRK4(RSVDt_vars *RSVDt)
{
    RSVDt->TS.DirAdj = 0;                                                                                     /* AdjointActionRK4.c:24 */
    printf("\n*** Adjoint action begins! ***\n");
    double t0 = wtime();
    CHK(rsvdt_adjoint_action_rk4(H));
    print_hms("*** Adjoint action elapsed time = ", wtime() - t0, " ***\n");
    return 0;
}
static int QRAllFreqs(void)
{
    printf("\n*** QR decomposition begins! ***\n");
    double t1 = wtime();
    CHK(rsvdt_qr_all_freqs(H));
    print_hms("*** QR decomposition elapsed time = ", wtime() - t1, " ***\n");
    return 0;
}
/* PowerIterationRK4.c:20-29 */
static int PowerIterationRK4(RSVDt_vars *RSVDt)
{
    for (int64_t iq = 0; iq < RSVDt->RSVD.q; iq++) {
        printf("\n******** Inside power iteration, %d/%d *******\n", (int)iq + 1, (int)RSVDt->RSVD.q);
        CHK(QRAllFreqs());
        CHK(AdjointActionRK4(RSVDt));
        CHK(QRAllFreqs());
        CHK(DirectActionRK4(RSVDt));
    }
    if (RSVDt->RSVD.q > 0) printf("\n******** Power iteration DONE! *************\n");
    return 0;
}

/* ---- SaveResults.c:32-84 ------------------------------------------------------------------------------------------------ */
static int save_modes(const char *prefix, const double *M, int64_t rows, RSVDt_vars *RSVDt, Directories *dirs)
{
    const int64_t k = RSVDt->RSVD.k, Nw = RSVDt->RSVD.Nw_eff;
    if (RSVDt->SaveResultsOpt == 1) { /* k matrices rows x Nw_eff: column iw of matrix ik = frequency-major column iw*k+ik */
        double *tmp = (double *)malloc((size_t)rows * Nw * 16);
        if (!tmp) return seterr(RSVDT_ERR_MEM, "out of memory");
        for (int64_t ik = 0; ik < k; ik++) {
            for (int64_t iw = 0; iw < Nw; iw++)
                memcpy(tmp + 2 * iw * rows, M + 2 * (iw * k + ik) * rows, (size_t)rows * 16);                 /* PermuteMat.c:26-27 */
            snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s%d%s", dirs->FolderDir, prefix, (int)ik + 1, "_allFreqs");
            if (pb_write_mat_dense(dirs->IO_dir, rows, Nw, tmp, rows)) { free(tmp); return seterr(ERR_USER, "%s", pb_error()); }
        }
        free(tmp);
    } else {
        for (int64_t iw = 0; iw < Nw; iw++) {
            /* "U_hat_k" -> "U_hat_Freq<iw>_allK" (SaveResults.c:73,79) */
            snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%.6sFreq%d%s", dirs->FolderDir, prefix, (int)iw, "_allK");
            if (pb_write_mat_dense(dirs->IO_dir, rows, k, M + 2 * iw * k * rows, rows)) return seterr(ERR_USER, "%s", pb_error());
        }
    }
    return 0;
}

static int SaveResults(RSVDt_vars *RSVDt, Directories *dirs)
{
    const int64_t k = RSVDt->RSVD.k, Nw = RSVDt->RSVD.Nw_eff, Nc = RSVDt->RSVD.Nc, Nb = RSVDt->RSVD.Nb;
    printf("\n********************************************\n************ Saving the output *************\n********************************************\n");
    printf("\nThe results directory: %s\n\n", dirs->FolderDir);
    printf("One matrix of size %d (modes) x %d (frequencies) for gains\n", (int)k, (int)Nw);
    double *S = (double *)malloc((size_t)k * Nw * 8), *Sc = (double *)calloc((size_t)k * Nw * 2, 8);
    double *U = (double *)malloc((size_t)Nc * k * Nw * 16), *V = (double *)malloc((size_t)Nb * k * Nw * 16);
    if (!S || !Sc || !U || !V) return seterr(RSVDT_ERR_MEM, "out of memory");
    CHK(rsvdt_get_s_hat(H, S));
    CHK(rsvdt_get_u_hat(H, U));
    CHK(rsvdt_get_v_hat(H, V));
    for (int64_t i = 0; i < k * Nw; i++) Sc[2 * i] = S[i];   /* real gains stored as complex scalars */
    snprintf(dirs->IO_dir, sizeof dirs->IO_dir, "%s%s", dirs->FolderDir, "S_hat");
    if (pb_write_mat_dense(dirs->IO_dir, k, Nw, Sc, k)) return seterr(ERR_USER, "%s", pb_error());
    if (RSVDt->SaveResultsOpt == 1)
        printf("%d matrices of size %d x %d for response modes\n%d matrices of size %d x %d for forcing modes\n\n", (int)k, (int)Nc, (int)Nw, (int)k, (int)Nb, (int)Nw);
    else
        printf("%d matrices of size %d x %d for response modes\n%d matrices of size %d x %d for forcing modes\n\n", (int)Nw, (int)Nc, (int)k, (int)Nw, (int)Nb, (int)k);
    CHK(save_modes("U_hat_k", U, Nc, RSVDt, dirs));
    CHK(save_modes("V_hat_k", V, Nb, RSVDt, dirs));
    free(S); free(Sc); free(U); free(V);
    CHK(SaveInputVarsCopy(dirs));
    printf("DONE :))\n\n");
    return 0;
}

/* ---- main.c:85-152 ------------------------------------------------------------------------------------------------------- */
int main(int argc, char **argv)
{
    RSVDt_vars RSVDt;
    TransRun_vars TR_vars;
    Weight_matrices Weight_mat;
    Directories dirs;
    memset(&RSVDt, 0, sizeof RSVDt);
    memset(&TR_vars, 0, sizeof TR_vars);
    memset(&Weight_mat, 0, sizeof Weight_mat);
    memset(&dirs, 0, sizeof dirs);
    setvbuf(stdout, NULL, _IOLBF, 0);

    /* `-ngpus P` stands for `mpiexec -n P` of the reference (README.md:221): one process, the rows of every matrix split
     * over P GPUs like PETSc's MATMPIAIJ / VECMPI (rsvdt_create_group); `-device d` picks the GPU of a single-GPU run. */
    int device = 0, ngpus = 1;
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "-device")) device = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "-ngpus")) ngpus = atoi(argv[i + 1]);
    }
    if (ngpus < 1) return seterr(ERR_USER, "-ngpus must be at least 1");

    /* PreProcessing.c:27-66.  The input file is read before the CUDA handle is created, so that a wrong variables.yaml is
     * reported the way the reference reports it on any machine (tests/test_host_driver.py compares the messages with the
     * reference executable's on CPU); everything after it needs the device. */
    CHK(ReadUserInput(argc, argv, &RSVDt, &Weight_mat, &TR_vars, &dirs));
    int rc = ngpus > 1 ? rsvdt_create_group(ngpus, NULL, &H) : rsvdt_create(device, &H);                      /* SlepcInitialize, main.c:107 */
    if (rc) return seterr(rc, ngpus > 1 ? "fewer usable CUDA devices than -ngpus asks for (this build has no CPU fallback)"
                                        : "no usable CUDA device (this build has no CPU fallback)");
    if (ngpus > 1) printf("Rows partitioned over %d GPUs like MATMPIAIJ (PETSC_DECIDE), one process\n", ngpus);
    if (!TR_vars.TransRun) printf("********************************************\n*************** Problem info ***************\n********************************************\n\n");
    CHK(SetupTimeFreqGrid(&RSVDt, &TR_vars));
    CHK(LNSStructure(&RSVDt, &TR_vars, &dirs));
    if (!TR_vars.TransRun) CHK(ReadWeightMats(&RSVDt, &Weight_mat, &dirs));
    if (!TR_vars.TransRun) CHK(CreateRandomMat(&RSVDt, &dirs));
    CHK(CreateResultsDir(&dirs, TR_vars.TransRun ? "TransientSnapshots_" : "ResolventModes_"));

    if (TR_vars.TransRun) {                                                                                   /* main.c:119-124 */
        CHK(TransientRunRK4(&TR_vars, &RSVDt, &dirs));
        rsvdt_destroy(H);
        return 0;
    }
    printf("\n********************************************\n*************** RSVD-\\Delta t **************\n********************************************\n");
    CHK(DirectActionRK4(&RSVDt));                                                                             /* main.c:134 */
    CHK(PowerIterationRK4(&RSVDt));                                                                           /* :136 */
    printf("\n*** Reduced SVD begins! ***\n");
    double t1 = wtime();
    CHK(rsvdt_svd_all_freqs_before_adjoint(H));                                                               /* :138 */
    print_hms("*** Reduced SVD elapsed time = ", wtime() - t1, " ***\n");
    CHK(rsvdt_store_u(H));                                                                                    /* :140 */
    CHK(AdjointActionRK4(&RSVDt));                                                                            /* :142 */
    printf("\n*** Reduced SVD begins! ***\n");
    t1 = wtime();
    CHK(rsvdt_svd_all_freqs(H));                                                                              /* :144 */
    print_hms("*** Reduced SVD elapsed time = ", wtime() - t1, " ***\n");
    CHK(SaveResults(&RSVDt, &dirs));                                                                          /* :146 */
    rsvdt_destroy(H);
    return 0;
}
