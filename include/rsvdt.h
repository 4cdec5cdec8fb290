/*
 * rsvdt.h -- C ABI of the B200-native RSVD-Delta-t time-stepping path (librsvdt_b200.so).
 *
 * The reference (AliFarghadan/RSVD-Delta-t) has no plugin/FFI interface: its seam is the set of C
 * functions that main.c calls (SourceCode/main.c:113-146).  Each entry point below replaces one of
 * those functions and keeps its argument meaning and error behaviour (int return, 0 = success, first
 * error unwinds, like PetscErrorCode + CHKERRQ).  The opaque handle plays the role of the reference's
 * structs of SourceCode/Variables.h (LNS_vars, RSVDt_vars, DFT_matrices, RSVD_matrices,
 * Resolvent_matrices, TS_matrices): it owns the device copies of A, A^H, Y_hat, U_hat, V_hat, S_hat and
 * the RK4 work vectors.
 *
 * Host buffers use the reference's own memory image: a one-rank PETSc MATDENSE, i.e. column-major
 * complex double (re, im interleaved), N x (k*Nw_eff), frequency-major columns (col = iw*k + ik;
 * SourceCode/PermuteMat.c:26-27).  Sparse inputs are CSR with 64-bit indices and complex-double values,
 * i.e. the content of a PETSc-binary AIJ record (SourceCode/LNSStructure.c:22-27).
 *
 * Threading: one host thread per handle; all calls are synchronous on return unless stated.
 * There is no CPU fallback: every compute entry point runs CUDA kernels for sm_100a or fails.
 */
#ifndef RSVDT_H
#define RSVDT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsvdt_context *rsvdt_handle;

enum {
    RSVDT_OK = 0,
    RSVDT_ERR_ARG = 63,      /* PETSC_ERR_ARG_OUTOFRANGE-like: bad argument / size mismatch */
    RSVDT_ERR_USER = 83,     /* PETSC_ERR_USER: missing or inconsistent user input */
    RSVDT_ERR_STATE = 73,    /* PETSC_ERR_ARG_WRONGSTATE: call sequence violated (e.g. action before operator) */
    RSVDT_ERR_CUDA = 97,     /* PETSC_ERR_GPU: CUDA runtime failure */
    RSVDT_ERR_MEM = 55,      /* PETSC_ERR_MEM */
    RSVDT_ERR_NUMERIC = 71   /* PETSC_ERR_MAT_CH_ZRPVT-like: breakdown (rank-deficient block in QR) */
};

/* Variables.h:5-28 : the members of RSVD_vars / TS_vars that the hot path reads. */
typedef struct {
    int64_t k;               /* number of test vectors                        (Variables.h:8)  */
    int64_t Nw;              /* number of frequencies to resolve              (Variables.h:10) */
    int64_t Nw_eff;          /* Nw (complex A) or Nw/2 (real A)               (Variables.h:11, LNSStructure.c:44) */
    int64_t Ns;              /* time steps per steady-state period            (Variables.h:19) */
    int64_t Nt;              /* transient time steps                          (Variables.h:20) */
    int64_t ResRatio;        /* Ns / Nw                                       (Variables.h:21) */
    double dt;               /* time step                                     (Variables.h:23) */
    double w;                /* base frequency                                (Variables.h:12) */
    int32_t RealOperator;    /* (Variables.h:24) */
    int32_t TransientRemoval;/* (Variables.h:25) */
} rsvdt_grid;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* SlepcInitialize / SlepcFinalize (main.c:107,148): binds a CUDA device and creates the stream. */
int rsvdt_create(int device, rsvdt_handle *h);
int rsvdt_destroy(rsvdt_handle h);
/* `mpiexec -n P RSVDt` in one process: a group handle that stands for ngpus member handles, one per GPU (devices[p], or GPU p
 * when devices is NULL).  It takes the same whole host buffers as a single-GPU handle -- the operator, the weights, Y_hat,
 * U_hat ... -- partitions their rows like PETSC_DECIDE (LNSStructure.c:24-26, CreateTSMats.c:15-17), runs every entry point
 * below on one host thread per GPU, maps peer memory directly for the halo exchange and calls NCCL itself (libnccl.so.2,
 * resolved at run time) for the all-reduce of the Gram matrices and norms and the all-gather of the weight products.
 * Entry points a group accepts: set_operator_csr, set_grid, set_weight_csr, set_yhat / get_yhat, direct / adjoint action,
 * qr_all_freqs, power_iteration, the two SVDs, store_u, get_u_hat / get_v_hat / get_s_hat, transient_run_rk4, destroy. */
int rsvdt_create_group(int ngpus, const int *devices, rsvdt_handle *h);
const char *rsvdt_last_error(rsvdt_handle h);   /* message of the last failing call (never NULL) */
const char *rsvdt_version(void);

/* ---- pre-processing ------------------------------------------------------------------------------ */
/* SetupTimeFreqGrid.c:16-21 (pure host arithmetic, bit-identical formulae). */
int rsvdt_setup_time_freq_grid(double w, int64_t Nw, double dt_in, double TransientLength,
                               double *dt, int64_t *Ns, int64_t *Nt, int64_t *ResRatio);

/* LNSStructure.c:23-56 : takes the operator (content of MatLoad), decides real vs complex
 * (||Im A||_F < 1e-12, :37-40), applies discounting A <- A - beta*I when disc_flg (:53-56), builds the
 * device CSR of A and of A^H (the adjoint is a gather SpMM on A^H instead of MatMultHermitianTranspose). */
int rsvdt_set_operator_csr(rsvdt_handle h, int64_t N, const int64_t *rowptr, const int64_t *col,
                           const double *val_re_im, int disc_flg, double beta, int *real_operator);

/* CreateDFTiDFTMats.c:5-96 + Variables.h: fixes k, Nw, Ns, ... ; twiddles are generated on the device.
 * LIMIT (not in the reference): k <= RSVDT_MAX_K -- the per-frequency k x k factorisations (Cholesky, Jacobi) keep a k x k
 * matrix per warp in shared memory; a larger k returns RSVDT_ERR_ARG here.  The host drivers check k (from variables.yaml
 * or from the column count of the forcing file) before they load the operator. */
#define RSVDT_MAX_K 32
int rsvdt_set_grid(rsvdt_handle h, const rsvdt_grid *g);

/* ReadWeightMats.c:26-99 : optional sparse B (N x Nb), C (Nc x N), W_f_sqrt_inv (Nb x Nb), W_q_sqrt (Nc x Nc),
 * W_q_sqrt_inv (Nc x Nc).  which: 0=W_f_sqrt_inv 1=B 2=W_q_sqrt_inv 3=W_q_sqrt 4=C.  Dimension checks as :34-98. */
int rsvdt_set_weight_csr(rsvdt_handle h, int which, int64_t nrows, int64_t ncols, const int64_t *rowptr,
                         const int64_t *col, const double *val_re_im);

/* The same for row-partitioned runs (after rsvdt_dist_init): every rank passes the rows of M it owns in M's row space and
 * the rows of M^H it owns in M's column space -- both PETSC_DECIDE splits, as MatLoad gives a MATMPIAIJ
 * (ReadWeightMats.c:28-31) -- with GLOBAL column indices.  Applying a weight all-gathers the owned rows of Y_hat one
 * frequency block at a time through the callback of rsvdt_set_allgather and forms the local rows of the product
 * (ApplyWeightMats.c:16-80 on MPI matrices: MatMatMult of an MPIAIJ with an MPIDENSE matrix). */
int rsvdt_set_weight_csr_local(rsvdt_handle h, int which, int64_t nrows_global, int64_t ncols_global, int64_t n_local,
                               const int64_t *rowptr, const int64_t *col_global, const double *val_re_im, int64_t nH_local,
                               const int64_t *rowptr_H, const int64_t *col_global_H, const double *val_H_re_im);

/* ---- RSVD_mat->Y_hat ----------------------------------------------------------------------------- */
/* CreateRandomMat.c:23-31 (the InputForcingFlg path): uploads Y_hat, nrows x (k*Nw_eff), reference layout. */
int rsvdt_set_yhat(rsvdt_handle h, const double *Y_host, int64_t nrows);
int rsvdt_get_yhat(rsvdt_handle h, double *Y_host, int64_t nrows);
int64_t rsvdt_yhat_rows(rsvdt_handle h);

/* ---- the hot path -------------------------------------------------------------------------------- */
/* DirectActionRK4.c:24-28 / AdjointActionRK4.c:24-28 : ApplyWeightMats(before) -> TSActionRK4 -> ApplyWeightMats(after),
 * Y_hat consumed and replaced (TSActionRK4.c:49-50, DFT.c:21-39 / TransientRemovalStrategy.c:30-40). */
int rsvdt_direct_action_rk4(rsvdt_handle h);
int rsvdt_adjoint_action_rk4(rsvdt_handle h);

/* One call with host buffers (upload F_hat, action, download Y_hat): the end-to-end form of TSActionRK4
 * (TSActionRK4.c:15-164) used for the e2e measurement.  direct != 0 -> A, else A^H. */
int rsvdt_ts_action_rk4(rsvdt_handle h, int direct, const double *F_hat_host, double *Y_hat_host);

/* QRAllFreqs.c:22-32 : per-frequency thin QR of Y_hat (CholeskyQR2), R discarded, diag(R) > 0. */
int rsvdt_qr_all_freqs(rsvdt_handle h);

/* PowerIterationRK4.c:20-29 : q x [QR, adjoint, QR, direct]. */
int rsvdt_power_iteration_rk4(rsvdt_handle h, int64_t q);

/* SVDAllFreqsBeforeAdjoint.c:31-53 : Y_hat_w <- left singular vectors of Y_hat_w. */
int rsvdt_svd_all_freqs_before_adjoint(rsvdt_handle h);
/* StoreU.c:15 : U_hat <- copy of Y_hat. */
int rsvdt_store_u(rsvdt_handle h);
/* SVDAllFreqs.c:34-64 : V_hat <- left singular vectors of Y_hat_w, S_hat <- singular values; Y_hat destroyed;
 * optional W^{-1/2} scalings of U_hat / V_hat (:63-64). */
int rsvdt_svd_all_freqs(rsvdt_handle h);
/* Resolvent_matrices readers (SaveResults.c reads these): U: Nc x (k*Nw_eff), V: Nb x (k*Nw_eff), S: k x Nw_eff (real). */
int rsvdt_get_u_hat(rsvdt_handle h, double *U_host);
int rsvdt_get_v_hat(rsvdt_handle h, double *V_host);
int rsvdt_get_s_hat(rsvdt_handle h, double *S_host);

/* TSExpADeltaT.c:25-34 : U (N x ncols, column-major, host) <- RK4^ResRatio(U), all columns batched. */
int rsvdt_ts_exp_a_deltat(rsvdt_handle h, int direct, double *U_host, int64_t ncols);

/* TransientRunRK4.c:128-186 : unforced IVP on one vector.  q0 in/out (host, N complex); norms_out has
 * floor(rend/TransSaveMod)+1 doubles (norms_out[0] = ||q0||, :106-108).  reason: 0 finished, 1 NaN, 2 diverged,
 * 3 converged (:143-158).  snapshot_cb (may be NULL) is called with (pos, q_host, N, user) every TransSaveMod
 * steps when TransSave is requested (:160-164). */
typedef int (*rsvdt_snapshot_cb)(int64_t pos, const double *q_host, int64_t N, void *user);
int rsvdt_transient_run_rk4(rsvdt_handle h, int direct, double *q0_host, int64_t rend, int64_t TransSaveMod,
                            double TransDivVal, double TransConVal, double *norms_out, int64_t *steps_done,
                            int *reason, rsvdt_snapshot_cb snapshot_cb, void *user);
/* The same with a second, independent cadence for TransientRemovalEst.c (TransientRunRK4.c:174-184 fills the N x Nw matrix
 * Q_all every Nt_period/Nw steps): est_cb(pos, q_host, N, est_user) is called after every est_mod-th step (pos = step /
 * est_mod), after the stop tests of that step if it is also a multiple of TransSaveMod.  est_mod <= 0 or est_cb == NULL:
 * exactly rsvdt_transient_run_rk4.  The estimate itself is host arithmetic on Nw snapshots per period and lives in the host
 * drivers (host/RSVDt.c TransientRemovalEst, transient_est.py). */
int rsvdt_transient_run_rk4_est(rsvdt_handle h, int direct, double *q0_host, int64_t rend, int64_t TransSaveMod,
                                double TransDivVal, double TransConVal, double *norms_out, int64_t *steps_done,
                                int *reason, rsvdt_snapshot_cb snapshot_cb, void *user, int64_t est_mod,
                                rsvdt_snapshot_cb est_cb, void *est_user);

/* ---- row partition across GPUs (one process per GPU) ----------------------------------------------------
 * Replaces PETSc's MPIAIJ distribution (LNSStructure.c:24-26: MatSetType MATMPIAIJ; CreateTSMats.c:15-17: VECMPI with
 * PETSC_DECIDE): rank r owns floor(N/P) + (r < N mod P) consecutive rows.  The planning (which rows are ghosts, who
 * sends what) is host arithmetic done by the caller (rsvd-delta-t_b200/dist.py); the library receives the LOCAL matrix
 * -- rows it stores, column indices already in local numbering -- and the halo message lists, exchanges CUDA IPC
 * handles through the caller, and from then on moves halo rows itself with peer stores over NVLink.
 *   layout 0: local rows = owned rows; ghost columns appended after them (general sparsity, one exchange per RK4 stage)
 *   layout 1: local rows = owned rows plus `halo` neighbour rows on each side (narrow-band operators advanced by the
 *             fused step kernel, one exchange per RK4 step). */
int rsvdt_plan_partition(int64_t N, int nranks, int rank, int64_t *row_begin, int64_t *row_end);   /* PETSC_DECIDE split */
int rsvdt_dist_init(rsvdt_handle h, int rank, int nranks, int layout);
/* LNSStructure.c:23-56 for one rank: n_rows stored rows with column indices in [0, n_ext); the owned rows are the local
 * rows [own0, own0 + n_own).  A and A^H are both given (the caller partitions A^H the same way, like the pre-transposed
 * operator of SURVEY.md 8e).  real_operator is the GLOBAL decision (||Im A||_F < 1e-12 over all ranks). */
int rsvdt_set_operator_csr_local(rsvdt_handle h, int64_t N_global, int64_t n_own, int64_t own0, int64_t n_ext,
                                 int64_t n_rows, const int64_t *rowptr, const int64_t *col, const double *val_re_im,
                                 const int64_t *rowptr_H, const int64_t *col_H, const double *val_H_re_im,
                                 int disc_flg, double beta, int real_operator);
/* halo messages: nout outgoing (peer, count, rows to send = local row indices, first ghost row they fill on the peer),
 * nin incoming (peer, count, first local ghost row).  Call after rsvdt_set_grid (the mailbox is sized for k). */
int rsvdt_dist_set_messages(rsvdt_handle h, int nout, const int *out_peer, const int *out_count, const int64_t *out_rows,
                            const int64_t *out_dest_row0, int nin, const int *in_peer, const int *in_count,
                            const int64_t *in_dest_row0);
/* CUDA IPC bootstrap: export 2 handles (mailbox, flags) = 128 bytes; import the nranks*128 bytes gathered from all ranks
 * plus, for every outgoing message i, the row offset of my region inside that peer's mailbox slot. */
int rsvdt_dist_ipc_export(rsvdt_handle h, unsigned char *bytes128);
int rsvdt_dist_ipc_import(rsvdt_handle h, const unsigned char *all_bytes, const int64_t *out_mailbox_row_offset);
/* offsets (in rows) of each incoming message inside my mailbox slot, to be sent to the corresponding senders */
int rsvdt_dist_get_in_offsets(rsvdt_handle h, int64_t *in_mailbox_row_offset);
/* sum-reduction across ranks of n doubles resident on the device (Gram matrices of QRAllFreqs / SVDAllFreqs, norms):
 * the caller supplies it (torch.distributed all_reduce over NCCL in the Python host); stream-ordered on `stream`. */
typedef int (*rsvdt_allreduce_fn)(void *device_ptr, int64_t n_doubles, void *cuda_stream, void *user);
int rsvdt_set_allreduce(rsvdt_handle h, rsvdt_allreduce_fn fn, void *user);
/* local layout 1 only: *ok = 1 if this rank's rows of A and A^H can be advanced by the fused band kernel for `nvec` vectors
 * (host-side check, no communication); the ranks all-reduce it and fall back to layout 0 together if any says 0. */
int rsvdt_dist_layout_supported(rsvdt_handle h, int nvec, int *ok);
/* all-gather across ranks of n doubles per rank resident on the device (send: n doubles, recv: nranks*n doubles in rank
 * order), used by the row-partitioned weight / input / output matrices; stream-ordered on `stream`. */
typedef int (*rsvdt_allgather_fn)(const void *send_device_ptr, void *recv_device_ptr, int64_t n_doubles_per_rank,
                                  void *cuda_stream, void *user);
int rsvdt_set_allgather(rsvdt_handle h, rsvdt_allgather_fn fn, void *user);
void *rsvdt_cuda_stream(rsvdt_handle h);

/* CreateRandomMat.c:36-39 (the InputForcingFlg = false path) on the device: Y_hat <- nrows x (k*Nw_eff) entries uniform in
 * [0,1) + i[0,1) from a counter-based generator keyed by (seed, global row, column): a row-partitioned run (global_row0 =
 * first owned row) fills the same matrix as one GPU.  (The reference's RandSeed is ineffective) */
int rsvdt_create_random_mat(rsvdt_handle h, int64_t nrows, int64_t global_row0, uint64_t seed);

/* ---- synthetic operators of BASELINE.json, generated on the device (measurement aid, no reference counterpart) ------------
 * The block-stencil operators of rsvd-delta-t_b200/synthetic.py (lns_stencil: nvar unknowns per grid point, 3^ndim-point
 * stencil, dense nvar x nvar blocks) written by a kernel straight into the tiled block-CSR layout of the stage kernel,
 * bit-identical to the host generator (tested), so that the 2e7-DOF configurations need no 65 GB file and no host-side
 * conversion.  blocks: [3^ndim][nvar][nvar] complex, mod_tab: [997], diag_tab: [499] complex (synthetic.stencil_tables).
 * This rank owns grid points [pt0, pt1) (x fastest); ghost_pts = sorted grid points outside that range its rows reference
 * (row-partitioned runs: after rsvdt_dist_init(layout 0); single GPU: pt0 = 0, pt1 = all points, no ghosts).
 * Only even numbers of vectors can be advanced on such an operator (no CSR copy exists for the other kernels). */
int rsvdt_set_operator_stencil(rsvdt_handle h, int ndim, const int64_t *shape, int nvar, const double *blocks_re_im,
                               const double *mod_tab, const double *diag_tab_re_im, int64_t pt0, int64_t pt1,
                               const int64_t *ghost_pts, int64_t n_ghost, int with_adjoint, int disc_flg, double beta,
                               int real_operator);

/* ---- test / measurement hooks (no reference counterpart) ------------------------------------------ */
/* the RK4 state y left on the device by the last stepping call, nrows x k column-major (host) */
int rsvdt_get_state(rsvdt_handle h, double *y_host, int64_t nrows);
/* snapshots Y_all of the last action: N x (Nstore*k), test-vector-major columns (TSActionRK4.c:62-70,126-130);
 * only kept when rsvdt_keep_snapshots(h, 1) was called before the action. */
int rsvdt_keep_snapshots(rsvdt_handle h, int keep);
int rsvdt_get_snapshots(rsvdt_handle h, double *Yall_host);
/* Runs `nsteps` forced RK4 steps of TSActionRK4's hot loop (TSActionRK4.c:116-122) on the resident F_hat
 * (= current Y_hat, not consumed), timed with CUDA events on the launching stream.  Reports the total time, the
 * time inside the RK4-stage kernels alone and inside the forcing-synthesis kernels alone (ms). */
int rsvdt_time_forced_steps(rsvdt_handle h, int direct, int64_t nsteps, float *ms_total);
/* how the stage kernel reads the operator for `nvec` vectors: out4 = {format (0 CSR, 1 block-CSR, 2 tiled block-CSR),
 * bytes of operator data one RK4 stage streams, stored entries (padding included), owned rows} */
int rsvdt_operator_info(rsvdt_handle h, int direct, int nvec, int64_t *out4);
/* number of kernels launched by this handle since creation / since the last reset */
int64_t rsvdt_kernel_launches(rsvdt_handle h, int reset);
/* tuning knobs: name in {"chunk_steps","stage_variant","idft_variant","use_graph","use_banded","use_bcsr","use_tiled",
 * "tiled_ctas","use_fused_halo",...}; returns RSVDT_ERR_ARG if unknown */
int rsvdt_set_option(rsvdt_handle h, const char *name, int64_t value);
/* device-event timing of one kernel family over the last timed window: which in {"stage","idft","dft","all"} */
int rsvdt_last_kernel_ms(rsvdt_handle h, const char *which, float *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* RSVDT_H */
