// capi.cu -- the C ABI of include/rsvdt.h: host orchestration of the CUDA kernels (one stream per handle).
//
// Mirrors the reference's call structure (SourceCode/main.c:113-146): every extern "C" function below
// replaces the reference function named in its comment.  No CPU fallback exists: if a kernel cannot be
// launched the call fails with RSVDT_ERR_CUDA.
#include "../../include/rsvdt.h"
#include "kernels.cuh"
#include "stage_tiled.cuh"

#include <algorithm>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }   // every early return through CU()/RS() frees its temporaries
    cudaError_t alloc(size_t n)
    {
        release();
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    void swap(DevBuf &o) { std::swap(p, o.p); std::swap(bytes, o.bytes); }
};

// host-side loops over block rows / tiles at operator upload (1e8-1e9 entries): plain threads, the runs are independent
template <class F>
void parallel_for(int64_t n, F fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt ? std::min(nt, 16u) : 4u;
    if (n < 4096 || nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nt - 1) / nt;
    for (unsigned i = 0; i < nt; i++) {
        const int64_t a = (int64_t)i * per, b = std::min(n, a + per);
        if (a < b) th.emplace_back([=]() { fn(a, b); });
    }
    for (auto &t : th) t.join();
}

struct CsrOwned {
    CsrDev d;
    DevBuf rowptr, col, val;
    int64_t lo_bw = 0, hi_bw = 0;  // max (row - col), max (col - row)
    bool set = false;
    // row-partitioned weight matrices: global sizes; the stored columns index the all-gathered input block
    // [rank][gather_rows][k] (gather_rows = rows of rank 0 = the largest PETSC_DECIDE share of the column space)
    int64_t g_nrows = 0, g_ncols = 0, gather_rows = 0, in_rows_local = 0;
    // diagonal (DIA) form for the fused sliding-window step kernel: dia[row][dia_lo+dia_hi+1]
    DevBuf dia, tile_ok, slow_ranges;
    int dia_lo = 0, dia_hi = 0;
    bool dia_valid = false;
    std::vector<unsigned char> row_regular;
    // block-CSR form (dense B x B blocks) for the LNS-type operators
    BcsrDev bcsr;
    DevBuf b_rowptr, b_col, b_val;
    bool bcsr_valid = false;
    // tiled block-CSR form (stage_tiled.cuh): the layout the TMA-staged stage kernel streams
    TbcsrDev tb;
    DevBuf t_chunks, t_tileptr, t_order, t_flags, t_winptr, t_win;
    bool tbcsr_valid = false;
    bool stencil_only = false;                    // generated on the device: no CSR / block-CSR copies exist
    std::vector<unsigned char> tile_ghost;        // [ntiles] 1: a block column of the tile is a ghost (row-partitioned runs)
    std::vector<int32_t> base_order;              // processing order of the tiles chosen for L2 locality (empty: row order)
    int plan_nv = -1, plan_variant = -1, plan_mode = -1, ntiles = 0, nslow = 0;
    bool plan_ok = false;
    int plan_TR = 0;                              // rows per tile of the plan; whether its first / last tile is a regular one
    bool first_ok = false, last_ok = false;
    void release()
    {
        rowptr.release(); col.release(); val.release(); dia.release(); tile_ok.release(); slow_ranges.release();
        b_rowptr.release(); b_col.release(); b_val.release(); bcsr = BcsrDev(); bcsr_valid = false;
        t_chunks.release(); t_tileptr.release(); t_order.release(); t_flags.release(); t_winptr.release(); t_win.release();
        tb = TbcsrDev(); tbcsr_valid = false;
        stencil_only = false; tile_ghost.clear(); base_order.clear();
        d = CsrDev(); set = false; dia_valid = false; plan_nv = -1; plan_ok = false; row_regular.clear();
    }
};

struct EventPair { cudaEvent_t a, b; };

}  // namespace

struct rsvdt_group_ctx;

struct rsvdt_context {
    int device = 0;
    rsvdt_group_ctx *group = nullptr;     // != null: this handle stands for P member handles, one per GPU (group.inc)
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;          // runs the few irregular tiles of a fused step next to the main kernel
    cudaStream_t xstream = nullptr;       // layout 1: the halo exchange of step n runs here, under the interior tiles of step n
    cudaEvent_t ev_ex = nullptr;          // ... recorded after it; the main stream waits for it before anything reads ghost rows
    bool ex_pending = false;
    DevBuf bnd_flag;                      // counter of finished boundary tiles (rk4_dia_tma_kernel)
    unsigned long long bnd_target = 0;
    int64_t use_overlap1 = 1;             // layout 1: exchange overlapped with the step kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t synth_stream = nullptr;  // low priority: forcing synthesis of the next chunk, overlapped with the RK4 steps
    cudaEvent_t ev_step0 = nullptr, ev_synth = nullptr;
    std::string err;

    // LNS_vars (Variables.h:59-62).  N = rows owned by this rank (all rows on one GPU); the local arrays hold n_ext rows:
    // the owned ones at [own0, own0 + N) plus halo/ghost copies of the neighbours' rows (row-partitioned runs).
    int64_t N = 0, n_ext = 0, own0 = 0, N_global = 0;
    CsrOwned A, AH;
    bool real_operator = false;

    // row partition over GPUs (PETSc's MPIAIJ layout; one process per GPU): halo messages, peer mailboxes, flags
    struct DistMsg { int peer; int count; int64_t dest_row0; int64_t mb_off_rows; DevBuf rows; std::vector<int32_t> rows_h; };
    struct DistIn { int peer; int count; int64_t dest_row0; int64_t mb_off_rows; };
    struct Dist {
        bool on = false, connected = false;
        int rank = 0, nranks = 1;
        int layout = 0;                       // 0: ghosts appended after the owned rows, 1: halo rows around the owned range
        std::vector<DistMsg> out;             // what I send
        std::vector<DistIn> in;               // what I receive
        int64_t in_rows_total = 0;            // rows per mailbox slot
        int nvec_max = 0;
        DevBuf mailbox, flags, msgs_dev, ins_dev, peer_flags_dev;
        std::vector<void *> peer_mailbox, peer_flags;   // IPC-mapped peer pointers (index = rank)
        bool peer_mapped_directly = false;              // members of a one-process group: plain peer pointers, nothing to close
        unsigned long long epoch = 0;
        int max_out = 0, max_in = 0;
        // interior / boundary split of the owned rows (layout 0): rows [0, lo_end) and [hi_begin, N) are sent to peers
        bool overlap = false;
        int64_t lo_end = 0, hi_begin = 0;
        rsvdt_allreduce_fn allreduce = nullptr;
        void *allreduce_user = nullptr;
        rsvdt_allgather_fn allgather = nullptr;
        void *allgather_user = nullptr;
        DevBuf gather_send, gather_recv;
        // exchange fused into the tiled stage kernel (layout 0; stage_tiled.cuh): three more mailbox slots in ghost order after
        // the two of the unfused path, one more flag per rank, tables of the tiles that send
        bool fz_ready = false;
        unsigned long long fz_epoch = 0;
        const cplx *fz_last_out = nullptr;    // block vector whose halo rows sit in mailbox slot fz_epoch % 3
        int fz_last_nv = 0;
        std::vector<int64_t> peer_in_rows;    // in_rows_total of every peer (mailbox slot size over there)
        DevBuf fz_send_index, fz_send_dest, fz_send_mask, fz_out_tiles, fz_out_done;
    } dist;

    // Weight_matrices (Variables.h:69-80): 0=W_f_sqrt_inv 1=B 2=W_q_sqrt_inv 3=W_q_sqrt 4=C, and their conj-transposes
    CsrOwned W[5], WH[5];

    // RSVD_vars + TS_vars
    rsvdt_grid g{};
    bool grid_set = false;

    // RSVD_matrices / Resolvent_matrices (device stacks [iw][row][ik])
    DevBuf Yhat, Uhat, Vhat, Shat;
    int64_t yhat_rows = 0, uhat_rows = 0, vhat_rows = 0;

    // TS_matrices: work block vectors, forcing slices, snapshots
    DevBuf y, xa, xb, acc, slices[2], coef, snaps, small, smallM, status, gram_partial, scalars;
    bool keep_snapshots = false;
    int64_t snaps_nstore = 0;

    // options
    int64_t chunk_steps = 0;   // RK4 steps per forcing-synthesis chunk; 0 = automatic (up to 64, bounded by free memory)
    int64_t stage_variant = 0;
    int64_t idft_variant = 0;
    int64_t use_banded = -1;  // -1: automatic; 0: stage kernels, 1: fused CSR banded step, 2: fused DIA step, 3: TMA-staged DIA step
    int64_t tma_variant = 0;  // 0: tiles sized for 3 CTAs/SM, 1: for 2 CTAs/SM
    int64_t gram_variant = 1; // Gram matrices of QR / SVD: 1 = FP64 tensor cores (DMMA, dense_k4.cu), 0 = FMA kernel
    int64_t qr_fused = 1;     // CholeskyQR2 in 5 passes over Y_hat (first apply fused with the second Gram, FP64 tensor cores) instead of 6
    int64_t use_bcsr = 1;     // block-CSR stage kernel when the operator has dense B x B blocks
    int64_t use_tiled = 1;    // ... and its TMA-staged tiled form (stage_tiled.cu) when the number of vectors is even
    int64_t use_small_exchange = 1; // halo exchanges of a few rows per neighbour: one launch instead of three
    int64_t tiled_flags = -1; // stage_tiled.cuh TbcsrDev::flags (-1: the default)
    int64_t tiled_order = 8;  // grid lines per group of the z-fastest tile order of 3-D stencil operators (0: row order)
    int64_t tiled_cv = 0;     // test vectors per consumer thread of the tiled stage kernel (0: automatic)
    int64_t tiled_stages = 0; // operator ring depth of the tiled stage kernel; 0 = what shared memory holds
    int64_t tiled_variant = 37; // compile-time variant of the tiled stage kernel (stage_tiled.cu VAR_*; B = 5, k = 8 only): batched epilogue
                                // loads + setmaxnreg + single-copy x windows, the fastest of profiles/r2_lns3d_stage_variants.json
    int64_t tiled_ctas = 2;   // resident CTAs per SM of the tiled stage kernel (ring depth follows from the shared memory left)
    int64_t use_small = 1;    // single-CTA persistent kernel when the whole block vector fits in one SM's shared memory

    // CUDA graphs of a whole snapshot-free chunk of RK4 steps (launch-bound small problems): [direct][slice buffer]
    struct ChunkGraph {
        cudaGraphExec_t exec = nullptr;
        const void *key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        int64_t nsteps = 0, launches = 0;
    } graphs[2][2];
    int64_t use_graph = 1;
    int64_t use_fused_halo = 1;   // row-partitioned runs, layout 0: the tiled stage kernel exchanges the halo rows itself
    int64_t use_overlap = 0;  // row-partitioned runs: hide the halo exchange behind the interior rows (measured: the three
                              // launches per stage cost more than the exchange they hide at plane-sized halos; off by default)
    int64_t use_pipeline = 0; // (measured: no gain, the two kernels compete for SM residency) synthesise the next chunk's forcing on a low-priority stream while the current chunk steps

    // instrumentation
    int64_t launches = 0;
    bool profile = false;
    std::vector<EventPair> ev_stage, ev_idft, ev_dft;
    float last_ms_stage = 0, last_ms_idft = 0, last_ms_dft = 0, last_ms_all = 0;
    int64_t last_n_stage = 0, last_n_idft = 0, last_n_dft = 0;
};

namespace {

int fail(rsvdt_handle h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(h, e__ == cudaErrorMemoryAllocation ? RSVDT_ERR_MEM : RSVDT_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);    \
    } while (0)

#define RS(call)                     \
    do {                             \
        int r__ = (call);            \
        if (r__ != RSVDT_OK) return r__; \
    } while (0)

// the main stream waits for a halo exchange still running on the exchange stream (layout 1, overlapped with the step kernel)
int settle_exchange(rsvdt_handle h)
{
    if (h->ex_pending) {
        CU(cudaStreamWaitEvent(h->stream, h->ev_ex, 0));
        h->ex_pending = false;
    }
    return RSVDT_OK;
}

int bind(rsvdt_handle h)
{
    if (!h) return RSVDT_ERR_ARG;
    CU(cudaSetDevice(h->device));
    return settle_exchange(h);
}

// event bracketing of one kernel family while profiling
struct Bracket {
    rsvdt_handle h;
    std::vector<EventPair> *list;
    EventPair ep{};
    bool on = false;
    cudaStream_t st;
    Bracket(rsvdt_handle h_, std::vector<EventPair> *l, cudaStream_t s = nullptr) : h(h_), list(l), st(s ? s : h_->stream)
    {
        if (h->profile && list->size() < 8192) {
            on = cudaEventCreate(&ep.a) == cudaSuccess && cudaEventCreate(&ep.b) == cudaSuccess;
            if (on) cudaEventRecord(ep.a, st);
        }
    }
    ~Bracket()
    {
        if (on) {
            cudaEventRecord(ep.b, st);
            list->push_back(ep);
        }
    }
};

float drain(std::vector<EventPair> &v, int64_t *n)
{
    float tot = 0;
    *n = (int64_t)v.size();
    for (auto &e : v) {
        float ms = 0;
        cudaEventSynchronize(e.b);
        cudaEventElapsedTime(&ms, e.a, e.b);
        tot += ms;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    v.clear();
    return tot;
}

// ---- CSR upload: 64-bit host CSR -> device CSR with 32-bit columns; optionally conj-transposed -------------
int upload_csr(rsvdt_handle h, CsrOwned &dst, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int64_t *col,
               const cplx *val, bool hermitian_transpose, double shift)
{
    dst.release();
    if (shift != 0.0) {
        // MatShift(A, -beta) (LNSStructure.c:53-56) inserts diagonal entries the matrix does not store: rows without one get a
        // zero entry at its sorted place first (local layouts: the diagonal of owned row r is column r as well)
        int64_t missing = 0;
        for (int64_t r = 0; r < nrows && r < ncols; r++) {
            bool has = false;
            for (int64_t p = rowptr[r]; p < rowptr[r + 1] && !has; p++) has = col[p] == r;
            missing += has ? 0 : 1;
        }
        if (missing) {
            std::vector<int64_t> rp2((size_t)nrows + 1, 0), ci2;
            std::vector<cplx> va2;
            ci2.reserve((size_t)(rowptr[nrows] + missing));
            va2.reserve((size_t)(rowptr[nrows] + missing));
            for (int64_t r = 0; r < nrows; r++) {
                bool has = r >= ncols;
                for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) has = has || col[p] == r;
                bool put = has;
                for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
                    if (!put && col[p] > r) { ci2.push_back(r); va2.push_back(make_double2(0.0, 0.0)); put = true; }
                    ci2.push_back(col[p]);
                    va2.push_back(val[p]);
                }
                if (!put) { ci2.push_back(r); va2.push_back(make_double2(0.0, 0.0)); }
                rp2[(size_t)r + 1] = (int64_t)ci2.size();
            }
            return upload_csr(h, dst, nrows, ncols, rp2.data(), ci2.data(), va2.data(), hermitian_transpose, shift);
        }
    }
    const int64_t nnz = rowptr[nrows];
    if (ncols > (int64_t)INT32_MAX) return fail(h, RSVDT_ERR_ARG, "more than 2^31-1 columns not supported");
    std::vector<int64_t> rp;
    std::vector<int32_t> ci((size_t)std::max<int64_t>(nnz, 1));
    std::vector<cplx> va((size_t)std::max<int64_t>(nnz, 1));
    int64_t out_rows = nrows, out_cols = ncols;
    int64_t lo = 0, hi = 0, maxnnz = 0;
    if (!hermitian_transpose) {
        rp.assign(rowptr, rowptr + nrows + 1);
        for (int64_t r = 0; r < nrows; r++) {
            bool shifted = (shift == 0.0);
            for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
                if (col[p] < 0 || col[p] >= ncols) return fail(h, RSVDT_ERR_ARG, "column index out of range in row %lld", (long long)r);
                ci[p] = (int32_t)col[p];
                va[p] = val[p];
                if (!shifted && col[p] == r) { va[p].x -= shift; shifted = true; }  // MatShift, LNSStructure.c:54
                lo = std::max(lo, r - col[p]);
                hi = std::max(hi, col[p] - r);
            }
            if (!shifted) return fail(h, RSVDT_ERR_ARG, "discounting needs a stored diagonal entry (row %lld)", (long long)r);
            maxnnz = std::max(maxnnz, rowptr[r + 1] - rowptr[r]);
        }
    } else {
        out_rows = ncols;
        out_cols = nrows;
        if (nrows > (int64_t)INT32_MAX) return fail(h, RSVDT_ERR_ARG, "more than 2^31-1 rows not supported");
        rp.assign((size_t)ncols + 1, 0);
        for (int64_t p = 0; p < nnz; p++) {
            if (col[p] < 0 || col[p] >= ncols) return fail(h, RSVDT_ERR_ARG, "column index out of range");
            rp[col[p] + 1]++;
        }
        for (int64_t c = 0; c < ncols; c++) rp[c + 1] += rp[c];
        std::vector<int64_t> next(rp.begin(), rp.end() - 1);
        for (int64_t r = 0; r < nrows; r++)
            for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
                const int64_t q = next[col[p]]++;
                ci[q] = (int32_t)r;
                va[q] = make_double2(val[p].x, -val[p].y);
                if (shift != 0.0 && col[p] == r) va[q].x -= shift;
                lo = std::max(lo, col[p] - r);
                hi = std::max(hi, r - col[p]);
            }
        for (int64_t c = 0; c < ncols; c++) maxnnz = std::max(maxnnz, rp[c + 1] - rp[c]);
    }
    CU(dst.rowptr.alloc((size_t)(out_rows + 1) * sizeof(int64_t)));
    CU(dst.col.alloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    CU(dst.val.alloc((size_t)std::max<int64_t>(nnz, 1) * sizeof(cplx)));
    CU(cudaMemcpyAsync(dst.rowptr.p, rp.data(), (size_t)(out_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    if (nnz > 0) {
        CU(cudaMemcpyAsync(dst.col.p, ci.data(), (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(dst.val.p, va.data(), (size_t)nnz * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    dst.d.nrows = out_rows;
    dst.d.ncols = out_cols;
    dst.d.nnz = nnz;
    dst.d.rowptr = dst.rowptr.as<int64_t>();
    dst.d.col = dst.col.as<int32_t>();
    dst.d.val = dst.val.as<cplx>();
    dst.d.max_row_nnz = maxnnz;
    dst.lo_bw = lo;
    dst.hi_bw = hi;
    dst.set = true;
    // ---- block structure analysis: dense B x B blocks on an aligned block grid (B = 5, 4, 3) ------------------------
    for (int B : {5, 4, 3}) {
        if (out_rows % B || out_cols % B || nnz == 0 || maxnnz < 2 * B) continue;
        bool okb = true;
        int64_t nblocks = 0;
        for (int64_t br = 0; br < out_rows / B && okb; br++) {
            const int64_t p0 = rp[br * B], len = rp[br * B + 1] - p0;
            if (len % B) { okb = false; break; }
            for (int r = 1; r < B && okb; r++) {
                const int64_t q0 = rp[br * B + r];
                if (rp[br * B + r + 1] - q0 != len) { okb = false; break; }
                for (int64_t i = 0; i < len; i++)
                    if (ci[q0 + i] != ci[p0 + i]) { okb = false; break; }
            }
            for (int64_t i = 0; i < len && okb; i += B) {
                if (ci[p0 + i] % B) { okb = false; break; }
                for (int j = 1; j < B; j++)
                    if (ci[p0 + i + j] != ci[p0 + i] + j) { okb = false; break; }
            }
            nblocks += len / B;
        }
        if (!okb) continue;
        const int64_t nbrows = out_rows / B;
        std::vector<int64_t> brp((size_t)nbrows + 1, 0);
        std::vector<int32_t> bc((size_t)nblocks);
        const int BS = (B * B + 1) & ~1;   // padded to an even number of entries: 32-byte aligned blocks
        std::vector<cplx> bv((size_t)nblocks * BS, make_double2(0.0, 0.0));
        int64_t pb = 0;
        for (int64_t br = 0; br < nbrows; br++) {
            const int64_t p0 = rp[br * B], len = rp[br * B + 1] - p0;
            for (int64_t i = 0; i < len; i += B, pb++) {
                bc[(size_t)pb] = (int32_t)(ci[p0 + i] / B);
                for (int r = 0; r < B; r++)
                    for (int j = 0; j < B; j++) bv[(size_t)(pb * BS + j * B + r)] = va[rp[br * B + r] + i + j];   // column-major block
            }
            brp[(size_t)br + 1] = pb;
        }
        CU(dst.b_rowptr.alloc(brp.size() * sizeof(int64_t)));
        CU(dst.b_col.alloc(std::max<size_t>(bc.size(), 1) * sizeof(int32_t)));
        CU(dst.b_val.alloc(std::max<size_t>(bv.size(), 1) * sizeof(cplx)));
        CU(cudaMemcpy(dst.b_rowptr.p, brp.data(), brp.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dst.b_col.p, bc.data(), bc.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dst.b_val.p, bv.data(), bv.size() * sizeof(cplx), cudaMemcpyHostToDevice));
        dst.bcsr.B = B;
        dst.bcsr.nbrows = nbrows;
        dst.bcsr.nblocks = nblocks;
        dst.bcsr.browptr = dst.b_rowptr.as<int64_t>();
        dst.bcsr.bcol = dst.b_col.as<int32_t>();
        dst.bcsr.bval = dst.b_val.as<cplx>();
        dst.bcsr_valid = true;
        // ---- tiled form: chunk p of a tile = the p-th block of each of its TB block rows, entries [e][tb], then columns [tb]
        {
            const int TB = TBCSR_TB;
            const int64_t ntiles = (nbrows + TB - 1) / TB;
            std::vector<int64_t> tptr((size_t)ntiles + 1, 0);
            // Tiles whose blocks line up on few block diagonals (stencils) are cut by diagonal: chunk = one diagonal, rows that
            // lack it padded in place -- the order of the blocks inside a row is kept (ascending columns), and the chunk's
            // columns become affine in the block row, which is what the x windows of the stage kernel need.  Other tiles are
            // cut by position in the row (p-th block of every row).
            std::vector<std::vector<int64_t>> tile_diags((size_t)ntiles);
            for (int64_t t = 0; t < ntiles; t++) {
                int64_t mx = 0;
                std::vector<int64_t> dg;
                bool ascending = true;   // (the diagonal cut keeps the order of a row's blocks only if they ascend)
                for (int64_t br = t * TB; br < std::min(nbrows, (t + 1) * TB); br++) {
                    mx = std::max(mx, brp[(size_t)br + 1] - brp[(size_t)br]);
                    for (int64_t q = brp[(size_t)br]; q < brp[(size_t)br + 1]; q++) {
                        dg.push_back((int64_t)bc[(size_t)q] - br);
                        if (q > brp[(size_t)br] && bc[(size_t)q] <= bc[(size_t)q - 1]) ascending = false;
                    }
                }
                std::sort(dg.begin(), dg.end());
                dg.erase(std::unique(dg.begin(), dg.end()), dg.end());
                if (ascending && (int64_t)dg.size() <= mx + std::max<int64_t>(4, mx / 4)) {
                    mx = (int64_t)dg.size();
                    tile_diags[(size_t)t] = std::move(dg);
                }
                tptr[(size_t)t + 1] = tptr[(size_t)t] + mx;
            }
            const size_t cb = tbcsr_chunk_bytes(B, TB);
            const int64_t nchunks = tptr[(size_t)ntiles];
            dst.tile_ghost.assign((size_t)ntiles, 0);
            for (int64_t br = 0; br < nbrows; br++)
                for (int64_t q = brp[(size_t)br]; q < brp[(size_t)br + 1]; q++)
                    if (bc[(size_t)q] >= nbrows) { dst.tile_ghost[(size_t)(br / TB)] = 1; break; }
            std::vector<unsigned char> chunks;
            try { chunks.resize((size_t)std::max<int64_t>(nchunks, 1) * cb); } catch (const std::bad_alloc &) { return fail(h, RSVDT_ERR_MEM, "host memory for the tiled operator"); }
            parallel_for(ntiles, [&](int64_t t0, int64_t t1) {
                for (int64_t t = t0; t < t1; t++) {
                    const int64_t nch = tptr[(size_t)t + 1] - tptr[(size_t)t];
                    const std::vector<int64_t> &dg = tile_diags[(size_t)t];
                    for (int64_t pch = 0; pch < nch; pch++) {
                        unsigned char *base = chunks.data() + (size_t)(tptr[(size_t)t] + pch) * cb;
                        cplx *cv = reinterpret_cast<cplx *>(base + tbcsr_vals_offset(TB));
                        int32_t *cc = reinterpret_cast<int32_t *>(base + tbcsr_cols_offset());
                        for (int tbi = 0; tbi < TB; tbi++) {
                            const int64_t br = t * TB + tbi;
                            int64_t q = -1;   // the block of row br that belongs to this chunk
                            if (br < nbrows) {
                                if (dg.empty()) { if (pch < brp[(size_t)br + 1] - brp[(size_t)br]) q = brp[(size_t)br] + pch; }
                                else {
                                    const int32_t want = (int32_t)(br + dg[(size_t)pch]);
                                    const int32_t *b0 = bc.data() + brp[(size_t)br], *b1 = bc.data() + brp[(size_t)br + 1];
                                    const int32_t *f = std::lower_bound(b0, b1, want);
                                    if (f != b1 && *f == want) q = f - bc.data();
                                }
                            }
                            cc[tbi] = q >= 0 ? bc[(size_t)q] : -1;
                            for (int e = 0; e < B * B; e++)
                                cv[(size_t)e * TB + tbi] = q >= 0 ? bv[(size_t)(q * BS + e)] : make_double2(0.0, 0.0);
                        }
                        cc[TB] = TBCSR_GATHER; cc[TB + 1] = cc[TB + 2] = cc[TB + 3] = 0;
                    }
                }
            });
            // x windows (stage_tiled.cuh): planned from the chunk columns, written into the chunk headers
            {
                std::vector<int32_t> ccols((size_t)std::max<int64_t>(nchunks, 1) * TB);
                for (int64_t pch = 0; pch < nchunks; pch++)
                    memcpy(ccols.data() + (size_t)pch * TB, chunks.data() + (size_t)pch * cb + tbcsr_cols_offset(), (size_t)TB * sizeof(int32_t));
                TbcsrPlan plan;
                tbcsr_plan_windows(ntiles, tptr.data(), ccols.data(), nbrows, out_cols / B, TB, TBCSR_WMAX, plan);
                for (int64_t pch = 0; pch < nchunks; pch++)
                    memcpy(chunks.data() + (size_t)pch * cb + tbcsr_cols_offset() + (size_t)TB * sizeof(int32_t), plan.hdr.data() + (size_t)pch * 4, 16);
                CU(dst.t_winptr.alloc(plan.win_ptr.size() * sizeof(int64_t)));
                CU(dst.t_win.alloc(std::max<size_t>(plan.win.size(), 1) * sizeof(TbcsrWindow)));
                CU(cudaMemcpy(dst.t_winptr.p, plan.win_ptr.data(), plan.win_ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
                if (!plan.win.empty()) CU(cudaMemcpy(dst.t_win.p, plan.win.data(), plan.win.size() * sizeof(TbcsrWindow), cudaMemcpyHostToDevice));
                dst.tb.win_ptr = dst.t_winptr.as<int64_t>();
                dst.tb.win = dst.t_win.as<TbcsrWindow>();
                dst.tb.wmax = 1;
                for (auto &w : plan.win) dst.tb.wmax = std::max(dst.tb.wmax, (int)w.nblocks);
                dst.tb.nwindows_chunks = plan.window_chunks;
            }
            CU(dst.t_tileptr.alloc(tptr.size() * sizeof(int64_t)));
            CU(dst.t_chunks.alloc(chunks.size()));
            CU(cudaMemcpy(dst.t_tileptr.p, tptr.data(), tptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(dst.t_chunks.p, chunks.data(), chunks.size(), cudaMemcpyHostToDevice));
            dst.tb.B = B;
            dst.tb.TB = TB;
            dst.tb.nbrows = nbrows;
            dst.tb.nbcols_own = nbrows;   // single GPU / in-buffer ghosts: every block column is a row range of the stage vector
            dst.tb.ntiles = ntiles;
            dst.tb.nchunks = nchunks;
            dst.tb.tile_ptr = dst.t_tileptr.as<int64_t>();
            dst.tb.chunks = dst.t_chunks.as<unsigned char>();
            dst.tb.chunk_bytes = cb;
            dst.tbcsr_valid = true;
        }
        break;
    }
    // ---- diagonal pattern analysis (square, narrow-band operators only) ------------------------------------------
    if (out_rows == out_cols && lo <= 8 && hi <= 8 && out_rows >= 64) {
        std::vector<int64_t> cnt(17, 0);
        for (int64_t r = 0; r < out_rows; r++)
            for (int64_t p = rp[r]; p < rp[r + 1]; p++) cnt[(size_t)(ci[p] - r + 8)]++;
        int omin = 99, omax = -99;
        for (int o = -8; o <= 8; o++)
            if (2 * cnt[(size_t)(o + 8)] >= out_rows) { omin = std::min(omin, o); omax = std::max(omax, o); }
        if (omin <= omax) {
            const int LO = std::max(0, -omin), HI = std::max(0, omax), ND = LO + HI + 1;
            std::vector<cplx> dia((size_t)out_rows * ND, make_double2(0.0, 0.0));
            dst.row_regular.assign((size_t)out_rows, 1);
            int64_t nirr = 0;
            for (int64_t r = 0; r < out_rows; r++) {
                bool reg = true;
                int64_t prev = -1;
                for (int64_t p = rp[r]; p < rp[r + 1]; p++) {
                    const int64_t o = ci[p] - r;
                    if (o < -LO || o > HI || ci[p] <= prev) { reg = false; break; }   // outside the window / not ascending
                    prev = ci[p];
                    dia[(size_t)r * ND + (size_t)(o + LO)] = va[p];
                }
                if (!reg) {
                    dst.row_regular[(size_t)r] = 0;
                    nirr++;
                    for (int d = 0; d < ND; d++) dia[(size_t)r * ND + d] = make_double2(0.0, 0.0);
                }
            }
            if (nirr * 100 <= out_rows) {
                CU(dst.dia.alloc(dia.size() * sizeof(cplx)));
                CU(cudaMemcpy(dst.dia.p, dia.data(), dia.size() * sizeof(cplx), cudaMemcpyHostToDevice));
                dst.dia_lo = LO;
                dst.dia_hi = HI;
                dst.dia_valid = true;
            }
        }
    }
    return RSVDT_OK;
}

// tile plan of the DIA step kernel for nv vectors: which tiles are regular, and the row ranges left to the CSR kernel
int plan_dia(rsvdt_handle h, CsrOwned &Op, int nv, int mode)
{
    const int variant = (int)h->stage_variant + 16 * (int)h->tma_variant;
    if (Op.plan_nv == nv && Op.plan_variant == variant && Op.plan_mode == mode) return RSVDT_OK;
    Op.plan_nv = nv;
    Op.plan_variant = variant;
    Op.plan_mode = mode;
    Op.plan_ok = false;
    if (!Op.dia_valid) return RSVDT_OK;
    if (mode == 3 ? !dia_tma_supported(Op.dia_lo, Op.dia_hi, nv, (int)h->tma_variant) : !dia_step_supported(Op.dia_lo, Op.dia_hi, nv))
        return RSVDT_OK;
    const int64_t N = Op.d.nrows;                       // rows of the local arrays (owned + halo)
    const int64_t ob = h->own0, oe = h->own0 + h->N;    // owned rows: the tiles cover exactly these
    if (ob != 0 && mode != 3) return RSVDT_OK;
    const int TR = mode == 3 ? dia_tma_tile_rows(Op.dia_lo, Op.dia_hi, nv, (int)h->tma_variant) : dia_tile_rows(Op.dia_lo, Op.dia_hi, nv);
    const int64_t ntiles = (oe - ob + TR - 1) / TR;
    std::vector<int64_t> irr_prefix((size_t)N + 1, 0);
    for (int64_t r = 0; r < N; r++) irr_prefix[(size_t)r + 1] = irr_prefix[(size_t)r] + (Op.row_regular[(size_t)r] ? 0 : 1);
    std::vector<unsigned char> ok((size_t)ntiles, 1);
    std::vector<int64_t> ranges;
    const bool slow_ok = banded_step_supported(Op.d, Op.lo_bw, Op.hi_bw, nv, (int)h->stage_variant);
    const int TRs = slow_ok ? banded_tile_rows(Op.lo_bw, Op.hi_bw, nv, (int)h->stage_variant) : 0;
    for (int64_t t = 0; t < ntiles; t++) {
        const int64_t r0 = ob + t * TR, r1 = std::min(oe, r0 + TR);
        const int64_t e0 = std::max<int64_t>(0, r0 - 3 * Op.dia_lo), e1 = std::min(N, r1 + 3 * Op.dia_hi);
        if (irr_prefix[(size_t)e1] - irr_prefix[(size_t)e0] > 0) {
            ok[(size_t)t] = 0;
            if (!slow_ok) return RSVDT_OK;   // no way to finish the irregular tiles: stay on the other paths
            for (int64_t a = r0; a < r1; a += TRs) { ranges.push_back(a); ranges.push_back(std::min(r1, a + TRs)); }
        }
    }
    if ((int64_t)ranges.size() / 2 > std::max<int64_t>(8, ntiles / 10)) return RSVDT_OK;   // too irregular to pay off
    CU(Op.tile_ok.alloc((size_t)ntiles));
    CU(cudaMemcpy(Op.tile_ok.p, ok.data(), (size_t)ntiles, cudaMemcpyHostToDevice));
    if (!ranges.empty()) {
        CU(Op.slow_ranges.alloc(ranges.size() * sizeof(int64_t)));
        CU(cudaMemcpy(Op.slow_ranges.p, ranges.data(), ranges.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    Op.ntiles = (int)ntiles;
    Op.nslow = (int)(ranges.size() / 2);
    Op.plan_TR = TR;
    Op.first_ok = ntiles > 0 && ok[0];
    Op.last_ok = ntiles > 0 && ok[(size_t)ntiles - 1];
    Op.plan_ok = true;
    return RSVDT_OK;
}

int ensure(rsvdt_handle h, DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return RSVDT_OK;
    CU(b.alloc(bytes));
    return RSVDT_OK;
}

size_t stack_bytes(int64_t na, int64_t rows, int64_t nb) { return (size_t)na * (size_t)rows * (size_t)nb * sizeof(cplx); }

// host (pageable or pinned) column-major <-> device stack, through a device staging buffer
// srows: rows per slot of the device stack; the `rows` host rows are its rows [roff, roff + rows)
int upload_stack(rsvdt_handle h, DevBuf &dst, const double *host, int64_t rows, int na, int nb, int64_t ca, int64_t cb,
                 int64_t srows = -1, int64_t roff = 0)
{
    if (srows < 0) srows = rows;
    const size_t bytes = stack_bytes(na, rows, nb);
    const bool fresh = dst.bytes < stack_bytes(na, srows, nb) || !dst.p;
    RS(ensure(h, dst, stack_bytes(na, srows, nb)));
    if (fresh && srows != rows) CU(cudaMemsetAsync(dst.p, 0, stack_bytes(na, srows, nb), h->stream));
    DevBuf stage;
    CU(stage.alloc(bytes));
    cudaError_t e = cudaMemcpyAsync(stage.p, host, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = launch_host_to_stack(dst.as<cplx>(), stage.as<cplx>(), rows, na, nb, ca, cb, srows, roff, h->stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    stage.release();
    CU(e);
    return RSVDT_OK;
}

int download_stack(rsvdt_handle h, const DevBuf &src, double *host, int64_t rows, int na, int nb, int64_t ca, int64_t cb,
                   int64_t srows = -1, int64_t roff = 0)
{
    if (srows < 0) srows = rows;
    const size_t bytes = stack_bytes(na, rows, nb);
    if (src.bytes < stack_bytes(na, srows, nb)) return fail(h, RSVDT_ERR_STATE, "device matrix not available");
    DevBuf stage;
    CU(stage.alloc(bytes));
    cudaError_t e = launch_stack_to_host(stage.as<cplx>(), src.as<cplx>(), rows, na, nb, ca, cb, srows, roff, h->stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, stage.p, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    stage.release();
    CU(e);
    return RSVDT_OK;
}

// ---- halo exchange of a block vector [row][nvec] (the VecScatter of PETSc's MPIAIJ mat-vec) ----------------------------
// first half: send my boundary rows (push into the peers' mailboxes, publish the epoch) on stream `s`
int dist_push(rsvdt_handle h, const cplx *buf, int nvec, cudaStream_t s)
{
    auto &D = h->dist;
    if (!D.connected) return fail(h, RSVDT_ERR_STATE, "row-partitioned run: peers not connected (rsvdt_dist_ipc_import)");
    if (nvec > D.nvec_max) return fail(h, RSVDT_ERR_ARG, "halo exchange of %d vectors exceeds the mailbox (%d)", nvec, D.nvec_max);
    RS(settle_exchange(h));
    D.epoch++;
    const int parity = (int)(D.epoch & 1ull);
    CU(launch_halo_push(buf, nvec, D.msgs_dev.p, (int)D.out.size(), D.max_out, parity, D.nvec_max, s));
    CU(launch_halo_signal(D.peer_flags_dev.p, (int)D.out.size(), D.epoch, s));
    h->launches += D.out.empty() ? 0 : 2;
    return RSVDT_OK;
}

// second half: wait for the peers' epoch and copy their rows into my ghost rows (main stream)
int dist_wait_unpack(rsvdt_handle h, cplx *buf, int nvec)
{
    auto &D = h->dist;
    const int parity = (int)(D.epoch & 1ull);
    CU(launch_halo_wait_unpack(buf, nvec, D.ins_dev.p, (int)D.in.size(), D.max_in, parity, D.nvec_max, D.epoch, h->stream));
    h->launches += D.in.empty() ? 0 : 1;
    return RSVDT_OK;
}

int dist_exchange(rsvdt_handle h, cplx *buf, int nvec)
{
    if (!h->dist.on) return RSVDT_OK;
    h->dist.fz_last_out = nullptr;   // the ghost rows of `buf` are (re)filled in place: the next fused stage gathers them from there
    auto &D = h->dist;
    RS(settle_exchange(h));
    if (h->use_small_exchange && D.connected && nvec <= D.nvec_max && (int64_t)std::max(D.max_out, D.max_in) * nvec <= 8192 &&
        D.out.size() <= 16 && D.in.size() <= 16) {
        // a few rows per neighbour (layout 1): the whole exchange in one launch
        D.epoch++;
        unsigned long long *tflag = D.flags.p ? D.flags.as<unsigned long long>() + 2 * D.nranks + 1 : nullptr;
        CU(launch_halo_exchange_small(buf, nvec, D.msgs_dev.p, (int)D.out.size(), D.peer_flags_dev.p, D.ins_dev.p, (int)D.in.size(),
                                      (int)(D.epoch & 1ull), D.nvec_max, D.epoch, tflag, nullptr, 0ull, h->stream));
        h->launches += 1;
        return RSVDT_OK;
    }
    RS(dist_push(h, buf, nvec, h->stream));
    return dist_wait_unpack(h, buf, nvec);
}

// ---- one RK4 step on a block of nv vectors (TSRK4.c:18-47 / TSTransRK4.c:15-36) ---------------------------
// y: state (in/out), xa/xb/acc: work, f_prev/f_half/f_next: forcing slices or nullptr, snap: optional copy of y_{n+1}
bool tiled_active(rsvdt_handle h, const CsrOwned &Op, int nv)
{
    return (Op.stencil_only || (h->use_bcsr && h->use_tiled)) && Op.tbcsr_valid && tbcsr_stage_supported(Op.tb.B, nv);
}
// the stage kernel itself moves the halo rows (no push / signal / unpack launches)
bool fused_active(rsvdt_handle h, const CsrOwned &Op, int nv)
{
    const auto &D = h->dist;
    return D.on && D.layout == 0 && D.fz_ready && h->use_fused_halo && nv <= D.nvec_max && Op.t_flags.p && tiled_active(h, Op, nv);
}

cudaError_t launch_stage_any(rsvdt_handle h, const CsrOwned &Op, int nv, int mode, const StageArgs &a)
{
    if (tiled_active(h, Op, nv)) {
        HaloFuse hf;
        TbcsrDev top = Op.tb;
        if (h->tiled_flags >= 0) top.flags = (int)h->tiled_flags;
        if (fused_active(h, Op, nv)) {
            auto &D = h->dist;
            const unsigned long long E = ++D.fz_epoch;
            const int64_t slot_elems = D.in_rows_total * (int64_t)D.nvec_max;
            unsigned long long *flags = D.flags.as<unsigned long long>();
            hf.enabled = 1;
            hf.epoch = E;
            const bool from_mail = E > 1 && a.xin == D.fz_last_out && nv == D.fz_last_nv;
            hf.wait_epoch = from_mail ? 1 : 0;
            hf.ghost_src = from_mail ? D.mailbox.as<cplx>() + (int64_t)(2 + (E - 1) % 3) * slot_elems
                                     : a.xin + Op.tb.nbcols_own * Op.tb.B * (int64_t)nv;   // in-buffer ghost rows (exchanged by the caller)
            hf.n_in = (int)D.in.size();
            hf.n_out = (int)D.out.size();
            for (int q = 0; q < hf.n_in; q++) hf.in_flag[q] = flags + D.nranks + D.in[(size_t)q].peer;
            for (int i = 0; i < hf.n_out; i++) {
                const auto &m = D.out[(size_t)i];
                const int64_t peer_slot = D.peer_in_rows[(size_t)m.peer] * (int64_t)D.nvec_max;
                hf.out_mail[i] = reinterpret_cast<cplx *>(D.peer_mailbox[(size_t)m.peer]) + (int64_t)(2 + E % 3) * peer_slot;
                hf.out_flag[i] = reinterpret_cast<unsigned long long *>(D.peer_flags[(size_t)m.peer]) + D.nranks + D.rank;
            }
            hf.out_done = D.fz_out_done.as<unsigned long long>();
            hf.out_tiles = D.fz_out_tiles.as<int32_t>();
            hf.tile_flags = Op.t_flags.as<unsigned char>();
            hf.send_index = D.fz_send_index.as<int32_t>();
            hf.send_dest = D.fz_send_dest.as<int64_t>();
            hf.send_mask = D.fz_send_mask.as<int32_t>();
            hf.timeout_flag = flags + 2 * D.nranks + 1;
            D.fz_last_out = mode == 2 ? a.acc_out : a.xout;
            D.fz_last_nv = nv;
        }
        return launch_rk4_stage_tbcsr(top, nv, mode, a, hf, (int)h->tiled_ctas, (int)h->tiled_stages, (int)h->tiled_cv, (int)h->tiled_variant, h->stream);
    }
    if (Op.stencil_only) return cudaErrorNotSupported;   // device-generated operators exist in the tiled form only (even nv)
    if (h->use_bcsr && Op.bcsr_valid && bcsr_stage_supported(Op.bcsr.B, nv))
        return launch_rk4_stage_bcsr(Op.bcsr, nv, mode, a, (int)h->stage_variant, h->stream);
    return launch_rk4_stage(Op.d, nv, mode, a, (int)h->stage_variant, h->stream);
}

// One stage of a row-partitioned run with the halo exchange hidden behind the interior rows (the overlap MPIAIJ's mat-vec
// does between its VecScatter and the diagonal-block product): boundary rows first, their push to the peers on the side
// stream while the interior rows are computed, then the unpack of the peers' rows.
int stage_overlapped(rsvdt_handle h, const CsrOwned &Op, int nv, int mode, const StageArgs &a, cplx *out_buf)
{
    auto &D = h->dist;
    const bool bc = h->use_bcsr && Op.bcsr_valid && bcsr_stage_supported(Op.bcsr.B, nv);
    auto launch = [&](int64_t r0, int64_t cnt) -> cudaError_t {
        if (cnt <= 0) return cudaSuccess;
        return bc ? launch_rk4_stage_bcsr(Op.bcsr, nv, mode, a, (int)h->stage_variant, h->stream, r0, cnt)
                  : launch_rk4_stage(Op.d, nv, mode, a, (int)h->stage_variant, h->stream, r0, cnt);
    };
    CU(launch(0, D.lo_end));
    CU(launch(D.hi_begin, h->N - D.hi_begin));
    CU(cudaEventRecord(h->ev_fork, h->stream));
    CU(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    RS(dist_push(h, out_buf, nv, h->side));
    CU(cudaEventRecord(h->ev_join, h->side));
    CU(launch(D.lo_end, D.hi_begin - D.lo_end));
    CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    return dist_wait_unpack(h, out_buf, nv);
}

int rk4_step(rsvdt_handle h, const CsrOwned &Op, int nv, cplx *y, cplx *xa, cplx *xb, cplx *acc, const cplx *f_prev,
             const cplx *f_half, const cplx *f_next, cplx *snap)
{
    const double dt = h->g.dt;
    StageArgs a;
    const bool fused = fused_active(h, Op, nv);
    if (!fused && h->dist.on && h->dist.overlap && h->use_overlap) {
        Bracket br(h, &h->ev_stage);
        a = StageArgs();
        a.xin = y; a.ybase = y; a.f = f_prev; a.xout = xa; a.acc_out = acc; a.cx = dt / 2; a.ca = dt / 6;
        RS(stage_overlapped(h, Op, nv, 0, a, xa));
        a = StageArgs();
        a.xin = xa; a.ybase = y; a.f = f_half; a.acc_in = acc; a.xout = xb; a.acc_out = acc; a.cx = dt / 2; a.ca = dt / 6 * 2;
        RS(stage_overlapped(h, Op, nv, 1, a, xb));
        a = StageArgs();
        a.xin = xb; a.ybase = y; a.f = f_half; a.acc_in = acc; a.xout = xa; a.acc_out = acc; a.cx = dt; a.ca = dt / 6 * 2;
        RS(stage_overlapped(h, Op, nv, 1, a, xa));
        a = StageArgs();
        a.xin = xa; a.f = f_next; a.acc_in = acc; a.acc_out = y; a.snap = snap; a.ca = dt / 6;
        RS(stage_overlapped(h, Op, nv, 2, a, y));
        h->launches += 12;
        return RSVDT_OK;
    }
    {   // k1 = Op y + F3(prev) ; x2 = y + dt/2 k1 ; acc = y + dt/6 k1            TSRK4.c:18,30,44
        Bracket br(h, &h->ev_stage);
        a = StageArgs();
        a.xin = y; a.ybase = y; a.f = f_prev; a.xout = xa; a.acc_out = acc; a.cx = dt / 2; a.ca = dt / 6;
        CU(launch_stage_any(h, Op, nv, 0, a));
    }
    if (!fused) RS(dist_exchange(h, xa, nv));
    {   // k2 = Op x2 + F2 ; x3 = y + dt/2 k2 ; acc += dt/6*2 k2                   TSRK4.c:31-32,45
        Bracket br(h, &h->ev_stage);
        a = StageArgs();
        a.xin = xa; a.ybase = y; a.f = f_half; a.acc_in = acc; a.xout = xb; a.acc_out = acc; a.cx = dt / 2; a.ca = dt / 6 * 2;
        CU(launch_stage_any(h, Op, nv, 1, a));
    }
    if (!fused) RS(dist_exchange(h, xb, nv));
    {   // k3 = Op x3 + F2 ; x4 = y + dt k3 ; acc += dt/6*2 k3                     TSRK4.c:33-34,46
        Bracket br(h, &h->ev_stage);
        a = StageArgs();
        a.xin = xb; a.ybase = y; a.f = f_half; a.acc_in = acc; a.xout = xa; a.acc_out = acc; a.cx = dt; a.ca = dt / 6 * 2;
        CU(launch_stage_any(h, Op, nv, 1, a));
    }
    if (!fused) RS(dist_exchange(h, xa, nv));
    {   // k4 = Op x4 + F3 ; y = acc + dt/6 k4 ; snapshot                         TSRK4.c:35,47 ; SaveSnapshots.c:26-28
        Bracket br(h, &h->ev_stage);
        a = StageArgs();
        a.xin = xa; a.f = f_next; a.acc_in = acc; a.acc_out = y; a.snap = snap; a.ca = dt / 6;
        CU(launch_stage_any(h, Op, nv, 2, a));
    }
    if (!fused) RS(dist_exchange(h, y, nv));   // (fused: the halo of y_{n+1} is in the peers' mailboxes, read by the next stage 1)
    h->launches += 4;
    return RSVDT_OK;
}

int64_t pos_mod(int64_t a, int64_t m)
{
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

// SaveSnapshots.c:19-28 : slot of step i in Y_all, or -1
int64_t snapshot_slot(const rsvdt_grid &g, bool direct, int64_t Nstore, int64_t i)
{
    if (i <= g.Nt) return -1;
    int64_t pos = i - g.Nt;
    if (pos % g.ResRatio != 0) return -1;
    pos = pos / g.ResRatio - 1;
    pos = pos_mod(pos, Nstore);
    pos = direct ? pos : Nstore - 1 - pos;
    return pos_mod(pos, Nstore);
}

// 0: stage-by-stage CSR kernels, 1: fused CSR banded step, 2: fused DIA step, 3: TMA-staged fused DIA step
// (2 and 3 leave the tiles with irregular rows to the CSR banded kernel)
int step_path(rsvdt_handle h, CsrOwned &Op, int nv, int *path)
{
    *path = 0;
    if (h->dist.on) {   // layout 1 <-> TMA-staged fused step on the halo-extended rows; layout 0 <-> stage kernels
        if (h->dist.layout == 1) {
            RS(plan_dia(h, Op, nv, 3));
            if (!Op.plan_ok) return fail(h, RSVDT_ERR_ARG, "partition layout 1 needs an operator the fused band kernel supports");
            *path = 3;
        }
        return RSVDT_OK;
    }
    const int64_t ub = h->use_banded < 0 ? 3 : h->use_banded;
    if (!ub) return RSVDT_OK;
    if (ub >= 3) {
        RS(plan_dia(h, Op, nv, 3));
        if (Op.plan_ok) { *path = 3; return RSVDT_OK; }
    }
    if (ub >= 2) {
        RS(plan_dia(h, Op, nv, 2));
        if (Op.plan_ok) { *path = 2; return RSVDT_OK; }
    }
    if (banded_step_supported(Op.d, Op.lo_bw, Op.hi_bw, nv, (int)h->stage_variant)) *path = 1;
    return RSVDT_OK;
}

// Layout 1 (narrow band, TMA-staged step kernel): can the exchange of y_{n+1}'s halo rows run under the interior tiles of the
// step that produces them?  Needs the rows every neighbour wants to come out of the first / last tile of the rank, both
// regular tiles; returns the number of tiles that will count themselves (0: not eligible).
int overlap1_tiles(rsvdt_handle h, const CsrOwned &Op, int nv, int path, int nbatch)
{
    const auto &D = h->dist;
    if (!D.on || D.layout != 1 || !D.connected || !h->use_overlap1 || !h->use_small_exchange || path != 3 || nbatch != 1) return 0;
    if (!Op.plan_ok || Op.ntiles < 2 || Op.plan_TR <= 0 || nv > D.nvec_max) return 0;
    if ((int64_t)std::max(D.max_out, D.max_in) * nv > 8192 || D.out.size() > 16 || D.in.size() > 16 || D.out.empty()) return 0;
    const int64_t lo_end = Op.plan_TR, hi_begin = (int64_t)(Op.ntiles - 1) * Op.plan_TR;
    for (const auto &m : D.out)
        for (int32_t r : m.rows_h) {
            const int64_t o = (int64_t)r - h->own0;
            if (o < 0 || o >= h->N) return 0;
            if (o < lo_end) { if (!Op.first_ok) return 0; }
            else if (o >= hi_begin) { if (!Op.last_ok) return 0; }
            else return 0;
        }
    return (Op.first_ok ? 1 : 0) + (Op.last_ok ? 1 : 0);
}

// one whole RK4 step, dispatching between the fused kernels (y ping-pongs between *ycur and *ynext) and the
// stage-by-stage kernels (in place on *ycur)
int rk4_step_any(rsvdt_handle h, const CsrOwned &Op, int nv, int path, cplx **ycur, cplx **ynext, const cplx *f_prev,
                 const cplx *f_half, const cplx *f_next, cplx *snap, int nbatch = 1, int64_t batch_stride = 0)
{
    RS(settle_exchange(h));   // (the ghost rows of *ycur come from the exchange of the previous step)
    if (path == 0)
        return rk4_step(h, Op, nv, *ycur, h->xa.as<cplx>(), h->xb.as<cplx>(), h->acc.as<cplx>(), f_prev, f_half, f_next, snap);
    const int ov_tiles = overlap1_tiles(h, Op, nv, path, nbatch);
    if (ov_tiles && !h->bnd_flag.p) {
        CU(h->bnd_flag.alloc(sizeof(unsigned long long)));
        CU(cudaMemsetAsync(h->bnd_flag.p, 0, sizeof(unsigned long long), h->stream));
        CU(cudaStreamSynchronize(h->stream));   // (once: the exchange stream must never see the counter before it is zero)
        h->bnd_target = 0;
    }
    Bracket br(h, &h->ev_stage);
    BandedStepArgs a;
    a.y_in = *ycur; a.y_out = *ynext; a.f_prev = f_prev; a.f_half = f_half; a.f_next = f_next; a.snap = snap; a.dt = h->g.dt;
    a.own_begin = h->own0;
    a.own_end = h->own0 + h->N;
    a.nbatch = nbatch;
    a.batch_stride = batch_stride;
    a.bnd_flag = ov_tiles ? h->bnd_flag.as<unsigned long long>() : nullptr;
    if (path >= 2) {
        const bool fork = Op.nslow > 0;
        if (Op.nslow > 0) {
            // the tiles holding rows off the diagonal pattern (domain boundaries) go through the CSR kernel -- on a side
            // stream, concurrently with the main kernel, when the grid is large: both read y_n and write disjoint rows
            if (fork) {
                CU(cudaEventRecord(h->ev_fork, h->stream));
                CU(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
            }
            CU(launch_rk4_banded_step(Op.d, Op.lo_bw, Op.hi_bw, nv, a, (int)h->stage_variant, Op.slow_ranges.as<int64_t>(), Op.nslow, fork ? h->side : h->stream));
            if (fork) CU(cudaEventRecord(h->ev_join, h->side));
            h->launches += 1;
        }
        if (path == 3)
            CU(launch_rk4_dia_tma_step(Op.dia.as<cplx>(), Op.dia_lo, Op.dia_hi, Op.tile_ok.as<unsigned char>(), Op.ntiles, Op.d.nrows, nv, (int)h->tma_variant, a, h->stream));
        else
            CU(launch_rk4_dia_step(Op.dia.as<cplx>(), Op.dia_lo, Op.dia_hi, Op.tile_ok.as<unsigned char>(), Op.ntiles, Op.d.nrows, nv, a, h->stream));
        h->launches += 1;
        if (fork) CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    } else {
        CU(launch_rk4_banded_step(Op.d, Op.lo_bw, Op.hi_bw, nv, a, (int)h->stage_variant, nullptr, 0, h->stream));
        h->launches += 1;
    }
    std::swap(*ycur, *ynext);
    if (ov_tiles) {
        // the exchange of y_{n+1}'s halo rows, on its own stream: it starts as soon as the boundary tiles of the step kernel
        // above have counted themselves and runs under the interior tiles; the next step (or whatever reads ghost rows
        // next) waits for ev_ex.  Exchanges follow each other in stream order, as their epochs require.
        auto &D = h->dist;
        D.fz_last_out = nullptr;
        D.epoch++;
        h->bnd_target += (unsigned long long)ov_tiles;
        unsigned long long *tflag = D.flags.p ? D.flags.as<unsigned long long>() + 2 * D.nranks + 1 : nullptr;
        CU(launch_halo_exchange_small(*ycur, nv, D.msgs_dev.p, (int)D.out.size(), D.peer_flags_dev.p, D.ins_dev.p, (int)D.in.size(),
                                      (int)(D.epoch & 1ull), D.nvec_max, D.epoch, tflag, h->bnd_flag.as<unsigned long long>(),
                                      h->bnd_target, h->xstream));
        CU(cudaEventRecord(h->ev_ex, h->xstream));
        h->ex_pending = true;
        h->launches += 1;
        return RSVDT_OK;
    }
    RS(dist_exchange(h, *ycur, nv));
    return RSVDT_OK;
}

// largest |harmonic| of the frequency slots (CreateDFTiDFTMats.c:44-50)
int64_t max_harmonic(const rsvdt_grid &g)
{
    if (g.RealOperator) return g.Nw_eff - 1;
    return (g.Nw_eff % 2 == 0) ? g.Nw_eff / 2 : (g.Nw_eff - 1) / 2;
}

// forcing slices s = 0..Sc-1 at half-step indices j_start + dir*s  (TSRK4.c:21-27, CreateForcingOnFly.c:17-19)
// prev (optional): also produce the slice at j_start - dir (the forcing at t_n of the chunk's first step, i.e. the F3
// carried over from the previous step, TSRK4.c:18 / TSActionRK4.c:103) into *prev
int synth_forcing(rsvdt_handle h, const cplx *Fhat, cplx *slices, int Sc, int64_t j_start, int dir, cplx *prev,
                  cudaStream_t st = nullptr)
{
    const rsvdt_grid &g = h->g;
    const int64_t U = h->n_ext * g.k;
    if (!st) st = h->stream;
    Bracket br(h, &h->ev_idft, st);
    const double PI = 3.14159265358979323846;
    const double hh = 0.5 * (double)Sc;
    const double x = (double)max_harmonic(g) * (PI / (double)g.Ns) * (hh + 1.0);   // |tau| <= 1 + 1/hh with the extra slice
    const int P = (h->idft_variant >= 10 || Sc < 4) ? -1 : idft_taylor_terms(x);
    if (P > 0 && P < Sc) {
        const int64_t J2 = 2 * j_start + (int64_t)dir * (Sc - 1);
        CU(launch_idft_taylor_coef(h->coef.as<cplx>(), P, (int)g.Nw_eff, J2, dir, g.Ns, hh, g.RealOperator, st));
        CU(launch_idft_taylor(slices, U, h->coef.as<cplx>(), Fhat, U, Sc, P, (int)g.Nw_eff, U, prev, st));
    } else {
        if (prev) {
            CU(launch_idft_coef(h->coef.as<cplx>(), 1, (int)g.Nw_eff, j_start - dir, 1, g.Ns, g.RealOperator, st));
            CU(launch_small_cgemm(prev, U, h->coef.as<cplx>(), Fhat, U, 1, (int)g.Nw_eff, U, 2, st));
            h->launches += 2;
        }
        // direct contraction in tiles of at most 32 slices
        for (int s0 = 0; s0 < Sc; s0 += 32) {
            const int Sb = std::min(32, Sc - s0);
            CU(launch_idft_coef(h->coef.as<cplx>() + (int64_t)s0 * g.Nw_eff, Sb, (int)g.Nw_eff, j_start + (int64_t)dir * s0, dir, g.Ns, g.RealOperator, st));
            CU(launch_small_cgemm(slices + (int64_t)s0 * U, U, h->coef.as<cplx>() + (int64_t)s0 * g.Nw_eff, Fhat, U, Sb, (int)g.Nw_eff, U,
                                  (int)(h->idft_variant % 10), st));
            h->launches += 2;
        }
        h->launches -= 2;
    }
    h->launches += 2;
    return RSVDT_OK;
}

// ---- the forced stepping loop of TSActionRK4.c:89-122 with all k test vectors batched -------------------------
// Fhat: stack [Nw_eff][N][k]; y must be zeroed by the caller for a fresh action.  Steps i = i_begin..i_end (1-based).
int forced_steps(rsvdt_handle h, bool direct, const cplx *Fhat, cplx *y, cplx *snaps, int64_t Nstore, int64_t i_begin,
                 int64_t i_end)
{
    const rsvdt_grid &g = h->g;
    CsrOwned &Op = direct ? h->A : h->AH;
    const int nv = (int)g.k;
    const int64_t U = h->n_ext * g.k;
    int64_t Tc = h->chunk_steps;
    if (Tc <= 0) {
        // automatic: the two slice buffers (2 x 2*Tc block vectors) may take a quarter of the free memory, Tc <= 64
        if (h->slices[0].p && h->slices[1].p && h->slices[0].bytes == h->slices[1].bytes && h->slices[0].bytes >= (size_t)16 * U * sizeof(cplx)) {
            Tc = (int64_t)(h->slices[0].bytes / ((size_t)2 * U * sizeof(cplx)));
        } else {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            Tc = (int64_t)((double)free_b * 0.25 / ((double)4 * U * sizeof(cplx)));
        }
        Tc = std::max<int64_t>(8, std::min<int64_t>(64, Tc));
    }
    const int S = (int)(2 * Tc);
    int path = 0;
    RS(step_path(h, Op, nv, &path));
    // launch-bound sizes: four plain stage kernels per step replayed from a CUDA graph (9 us per step at 5,000 units) beat
    // the fused tile kernels (16 us: TMA round trip, mbarrier and fork/join latencies dominate when the grid is a few CTAs)
    if (h->use_banded < 0 && h->use_graph && !h->dist.on && U <= 262144) path = 0;
    const bool banded = path != 0;
    // measured on the tutorial size (5,000 units): one persistent CTA needs 28 us per forced step, the tiled kernels 21 us;
    // the persistent kernel wins when a thread owns one or two units, or when many independent vectors run at once (below)
    const bool small = !h->dist.on && small_persistent_supported(Op.d, nv) &&
                       (h->use_small >= 2 || (h->use_small == 1 && h->n_ext * (int64_t)nv <= 1024));
    for (int b = 0; b < 2; b++) RS(ensure(h, h->slices[b], (size_t)S * U * sizeof(cplx)));
    RS(ensure(h, h->coef, (size_t)std::max<int64_t>((int64_t)128 * g.Nw_eff, g.Nw_eff * g.Nw) * sizeof(cplx)));
    RS(ensure(h, h->xa, (size_t)U * sizeof(cplx)));
    if (!banded) {
        RS(ensure(h, h->xb, (size_t)U * sizeof(cplx)));
        RS(ensure(h, h->acc, (size_t)U * sizeof(cplx)));
    }
    cplx *ycur = y, *ynext = h->xa.as<cplx>();
    if (h->dist.on && h->dist.layout == 1) {
        // halo rows of the forcing: the fused step recomputes the stages on the halo rows, so F_hat is needed there too
        for (int64_t w = 0; w < g.Nw_eff; w++) RS(dist_exchange(h, const_cast<cplx *>(Fhat) + w * U, nv));
    }
    if (h->dist.on) RS(dist_exchange(h, ycur, nv));

    int cur = 0;
    // forcing at t_n of the first step of the chunk (F3 carried over, TSRK4.c:18); for the very first chunk it is the
    // initial forcing of TSActionRK4.c:103, produced together with the chunk's slices
    const cplx *f_carry = nullptr;
    // Pipelining: the synthesis of chunk c+1 (HBM- and FP64-bound) runs on a low-priority stream while chunk c steps (the
    // step kernel is bound by the shared-memory pipe and leaves HBM half idle).  It may start once step 0 of chunk c has
    // read the carried slice that lives in the buffer it overwrites.
    const bool pipe = h->use_pipeline && !small && U > 262144;
    bool next_ready = false;
    for (int64_t i0 = i_begin; i0 <= i_end; i0 += Tc) {
        const int64_t nsteps = std::min(Tc, i_end - i0 + 1);
        const int Sc = (int)(2 * nsteps);
        cplx *sl = h->slices[cur].as<cplx>();
        const int64_t j_start = direct ? 2 * (i0 - 1) + 1 : -1 - 2 * (i0 - 1);
        cplx *first_prev = nullptr;
        if (!f_carry) {
            first_prev = h->slices[1].as<cplx>() + (int64_t)(S - 1) * U;   // a slot of the other buffer, free until chunk 2
            f_carry = first_prev;
        }
        if (next_ready) CU(cudaStreamWaitEvent(h->stream, h->ev_synth, 0));
        else RS(synth_forcing(h, Fhat, sl, Sc, j_start, direct ? 1 : -1, first_prev));
        next_ready = false;
        if (small) {   // the whole chunk in one launch of the persistent single-CTA kernel
            Bracket br(h, &h->ev_stage);
            SmallStepArgs sa;
            sa.y = ycur; sa.nsteps = (int)nsteps; sa.dt = g.dt; sa.slices = sl; sa.f_carry = f_carry; sa.snaps = snaps;
            sa.i0 = i0; sa.Nt = g.Nt; sa.ResRatio = g.ResRatio; sa.Nstore = Nstore; sa.direct = direct ? 1 : 0;
            CU(launch_rk4_small_persistent(Op.d, nv, sa, 1, h->stream));
            h->launches += 1;
            f_carry = sl + (int64_t)(Sc - 1) * U;
            cur ^= 1;
            continue;
        }
        // launch-bound problems: a full chunk without snapshot steps is replayed as one CUDA graph (the kernel arguments of
        // such a chunk depend only on the direction and on which of the two slice buffers it reads)
        bool has_snap = false;
        if (snaps)
            for (int64_t c = 0; c < nsteps; c++) has_snap |= snapshot_slot(g, direct, Nstore, i0 + c) >= 0;
        const bool graph_ok = h->use_graph && !h->dist.on && !has_snap && nsteps == Tc && (Tc % 2 == 0) &&
                              U <= 262144 && f_carry == h->slices[cur ^ 1].as<cplx>() + (int64_t)(S - 1) * U;
        if (graph_ok) {
            auto &G = h->graphs[direct ? 1 : 0][cur];
            const void *key[6] = {sl, f_carry, ycur, ynext, Op.d.val, (const void *)(intptr_t)(path * 1000 + nv)};
            const bool hit = G.exec && G.nsteps == nsteps && !memcmp(G.key, key, sizeof key);
            if (!hit) {
                if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
                const int64_t l0 = h->launches;
                cplx *yc = ycur, *yn = ynext;
                const bool prof = h->profile;
                h->profile = false;   // no timing events inside a capture
                CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                int rc = RSVDT_OK;
                for (int64_t c = 0; c < nsteps && rc == RSVDT_OK; c++) {
                    const cplx *f_prev = (c == 0) ? f_carry : sl + (2 * c - 1) * U;
                    rc = rk4_step_any(h, Op, nv, path, &yc, &yn, f_prev, sl + (2 * c) * U, sl + (2 * c + 1) * U, nullptr);
                }
                cudaGraph_t graph = nullptr;
                cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                h->profile = prof;
                if (rc != RSVDT_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                CU(ce);
                ce = cudaGraphInstantiate(&G.exec, graph, 0);
                cudaGraphDestroy(graph);
                CU(ce);
                memcpy(G.key, key, sizeof key);
                G.nsteps = nsteps;
                G.launches = h->launches - l0;
                h->launches = l0;   // the capture did not run anything
            }
            {
                Bracket br(h, &h->ev_stage);
                CU(cudaGraphLaunch(G.exec, h->stream));
            }
            h->launches += G.launches;
            f_carry = sl + (int64_t)(Sc - 1) * U;
            cur ^= 1;
            continue;
        }
        for (int64_t c = 0; c < nsteps; c++) {
            const int64_t i = i0 + c;
            const cplx *f_prev = (c == 0) ? f_carry : sl + (2 * c - 1) * U;
            const cplx *f_half = sl + (2 * c) * U;
            const cplx *f_next = sl + (2 * c + 1) * U;
            cplx *snap = nullptr;
            if (snaps) {
                const int64_t slot = snapshot_slot(g, direct, Nstore, i);
                if (slot >= 0) snap = snaps + slot * U;
            }
            RS(rk4_step_any(h, Op, nv, path, &ycur, &ynext, f_prev, f_half, f_next, snap));
            if (c == 0 && pipe && i0 + Tc <= i_end) {
                const int64_t i0n = i0 + Tc;
                const int Scn = (int)(2 * std::min(Tc, i_end - i0n + 1));
                const int64_t j_next = direct ? 2 * (i0n - 1) + 1 : -1 - 2 * (i0n - 1);
                CU(cudaEventRecord(h->ev_step0, h->stream));
                CU(cudaStreamWaitEvent(h->synth_stream, h->ev_step0, 0));
                RS(synth_forcing(h, Fhat, h->slices[cur ^ 1].as<cplx>(), Scn, j_next, direct ? 1 : -1, nullptr, h->synth_stream));
                CU(cudaEventRecord(h->ev_synth, h->synth_stream));
                next_ready = true;
            }
        }
        f_carry = sl + (int64_t)(Sc - 1) * U;
        cur ^= 1;
    }
    if (ycur != y) CU(cudaMemcpyAsync(y, ycur, (size_t)U * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    return RSVDT_OK;
}

// unforced steps on a block of nv vectors (TSTransRK4.c:15-36)
int unforced_steps(rsvdt_handle h, bool direct, int nv, cplx *y, int64_t nsteps)
{
    CsrOwned &Op = direct ? h->A : h->AH;
    const int64_t U = h->n_ext * nv;
    if (h->use_small && !h->dist.on && small_persistent_supported(Op.d, nv)) {
        for (int64_t done = 0; done < nsteps; done += 4096) {
            SmallStepArgs sa;
            sa.y = y; sa.nsteps = (int)std::min<int64_t>(4096, nsteps - done); sa.dt = h->g.dt;
            CU(launch_rk4_small_persistent(Op.d, nv, sa, 1, h->stream));
            h->launches += 1;
        }
        return RSVDT_OK;
    }
    int path = 0;
    RS(step_path(h, Op, nv, &path));
    const bool banded = path != 0;
    RS(ensure(h, h->xa, (size_t)U * sizeof(cplx)));
    if (!banded) {
        RS(ensure(h, h->xb, (size_t)U * sizeof(cplx)));
        RS(ensure(h, h->acc, (size_t)U * sizeof(cplx)));
    }
    cplx *ycur = y, *ynext = h->xa.as<cplx>();
    if (h->dist.on) RS(dist_exchange(h, ycur, nv));   // the ghost rows of the initial vector
    for (int64_t i = 0; i < nsteps; i++) RS(rk4_step_any(h, Op, nv, path, &ycur, &ynext, nullptr, nullptr, nullptr, nullptr));
    if (ycur != y) CU(cudaMemcpyAsync(y, ycur, (size_t)U * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    return RSVDT_OK;
}

// unforced steps on `nslots` independent block vectors (stride U) with ONE launch per step when the fused band kernel
// applies and the second buffer is affordable; otherwise slot by slot (TSExpADeltaT.c:25-34 applied to every snapshot slot)
int unforced_steps_slots(rsvdt_handle h, bool direct, int nv, cplx *base, int nslots, int64_t nsteps)
{
    CsrOwned &Op = direct ? h->A : h->AH;
    const int64_t U = h->n_ext * nv;
    if (h->use_small && !h->dist.on && small_persistent_supported(Op.d, nv)) {   // one persistent CTA per slot
        for (int64_t done = 0; done < nsteps; done += 4096) {
            SmallStepArgs sa;
            sa.y = base; sa.batch_stride = U; sa.nsteps = (int)std::min<int64_t>(4096, nsteps - done); sa.dt = h->g.dt;
            CU(launch_rk4_small_persistent(Op.d, nv, sa, nslots, h->stream));
            h->launches += 1;
        }
        return RSVDT_OK;
    }
    int path = 0;
    RS(step_path(h, Op, nv, &path));
    const size_t bytes = (size_t)nslots * U * sizeof(cplx);
    if (path == 3 && !h->dist.on && bytes <= ((size_t)1 << 30)) {
        DevBuf tmp;
        CU(tmp.alloc(bytes));
        cplx *ycur = base, *ynext = tmp.as<cplx>();
        int rc = RSVDT_OK;
        for (int64_t i = 0; i < nsteps && rc == RSVDT_OK; i++)
            rc = rk4_step_any(h, Op, nv, path, &ycur, &ynext, nullptr, nullptr, nullptr, nullptr, nslots, U);
        cudaError_t e = cudaSuccess;
        if (rc == RSVDT_OK && ycur != base) e = cudaMemcpyAsync(base, ycur, bytes, cudaMemcpyDeviceToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        tmp.release();
        if (rc != RSVDT_OK) return rc;
        CU(e);
        return RSVDT_OK;
    }
    for (int a = 0; a < nslots; a++) RS(unforced_steps(h, direct, nv, base + (int64_t)a * U, nsteps));
    return RSVDT_OK;
}

// DFT.c:26-39 : Y_hat[w'][u] = ResRatio * sum_j dft[j][w'] * Y_all[j][u]
int dft_snapshots(rsvdt_handle h, const cplx *snaps, cplx *Yhat_out)
{
    const rsvdt_grid &g = h->g;
    const int64_t U = h->n_ext * g.k;
    RS(ensure(h, h->coef, (size_t)std::max<int64_t>(g.Nw_eff * g.Nw, 128 * g.Nw_eff) * sizeof(cplx)));
    Bracket br(h, &h->ev_dft);
    CU(launch_dft_coef(h->coef.as<cplx>(), (int)g.Nw_eff, (int)g.Nw, g.ResRatio, h->stream));
    CU(launch_small_cgemm(Yhat_out, U, h->coef.as<cplx>(), snaps, U, (int)g.Nw_eff, (int)g.Nw, U, 0, h->stream));
    h->launches += 2;
    return RSVDT_OK;
}

int dist_allreduce(rsvdt_handle h, void *dev_ptr, int64_t n_doubles);
int check_halo_timeout(rsvdt_handle h);
int build_fused(rsvdt_handle h);
int finish_connect(rsvdt_handle h, const int64_t *out_mb_off);

// TransientRemovalStrategy.c:30-40 on the device.  snaps: Y_all as a stack [Nw+1][row][ik]; Yhat_out: stack [Nw_eff][row][ik].
//   CreateMtilde.c:41-54      V = QR(Y_all_k) -> shifted CholeskyQR3 per test vector (the snapshot matrices reach cond ~1e7,
//                             beyond plain CholeskyQR2; any orthonormal basis of the same span gives the same projection),
//                             MV = exp(A dT) V by ResRatio unforced RK4 steps on every slot, Mt = V^H MV
//   GalerkinProj.c:26-74      y1, y2 = ResRatio * Y_all[:, 0:Nw | 1:Nw+1] * dft[:, iw]; b = V^H (alpha y1 - y2);
//                             (alpha I - Mt) c = b;  Y_hat[:, iw] = y1 - V c
int transient_removal(rsvdt_handle h, bool direct, const cplx *snaps, cplx *Yhat_out)
{
    const rsvdt_grid &g = h->g;
    const int m = (int)g.Nw + 1, nv = (int)g.k, nw = (int)g.Nw_eff;
    // row-partitioned runs: the Gram-type products sum over the owned rows and are all-reduced; everything else is row-local
    const int64_t grow0 = h->own0, grows = h->N;
#define ALLRED(buf, n) do { rc = dist_allreduce(h, (buf), (n)); if (rc != RSVDT_OK) { cleanup(); return rc; } } while (0)
    const int64_t rows = h->n_ext, U = rows * nv;
    DevBuf V, W, Y1, Y2, G, Minv, B1, B2, Cc, part, scratch, alpha_d, status;
    int rc = RSVDT_OK;
    auto cleanup = [&]() { for (DevBuf *b : {&V, &W, &Y1, &Y2, &G, &Minv, &B1, &B2, &Cc, &part, &scratch, &alpha_d, &status}) b->release(); };
#define TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(h, e__ == cudaErrorMemoryAllocation ? RSVDT_ERR_MEM : RSVDT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return rc; } } while (0)
    const int chunks = stack_gram2_chunks(grows);
    const int mbmax = std::max(m, nw);
    TRY(V.alloc((size_t)m * U * sizeof(cplx)));
    TRY(W.alloc((size_t)m * U * sizeof(cplx)));
    TRY(G.alloc((size_t)nv * m * mbmax * sizeof(cplx)));
    TRY(Minv.alloc((size_t)nv * m * m * sizeof(cplx)));
    TRY(part.alloc((size_t)chunks * nv * m * mbmax * sizeof(cplx)));
    TRY(status.alloc((size_t)nv * std::max(nw, 1) * sizeof(int)));
    TRY(cudaMemcpyAsync(V.p, snaps, (size_t)m * U * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    std::vector<int> st((size_t)nv * nw);
    // ---- V = QR(Y_all_k): shifted CholeskyQR3 ---------------------------------------------------------------------
    for (int pass = 0; pass < 3; pass++) {
        TRY(launch_stack_gram2(V.as<cplx>(), U, m, V.as<cplx>(), U, m, nv, grow0, grows, part.as<cplx>(), G.as<cplx>(), h->stream));
        ALLRED(G.p, (int64_t)2 * nv * m * m);
        // shift of Fukaya et al. (2020): 11 (m n + n (n+1)) u ||X||_2^2, with ||X||_2^2 <= trace(G)
        const double shift = pass == 0 ? 11.0 * ((double)h->N_global * m + (double)m * (m + 1)) * 1.1e-16 : 0.0;
        TRY(launch_chol_inverse_big(G.as<cplx>(), m, nv, shift, Minv.as<cplx>(), status.as<int>(), h->stream, 1));
        TRY(launch_stack_rmul(W.as<cplx>(), U, nullptr, 0, 0.0, V.as<cplx>(), U, m, Minv.as<cplx>(), m, 1.0, nv, rows, h->stream));
        h->launches += 4;
        TRY(cudaMemcpyAsync(st.data(), status.p, (size_t)nv * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        TRY(cudaStreamSynchronize(h->stream));
        for (int z = 0; z < nv; z++)   // (negative: that many dependent snapshot columns were left out of the basis -- not an error)
            if (st[(size_t)z] > 0) { cleanup(); return fail(h, RSVDT_ERR_NUMERIC, "transient removal: QR breakdown for test vector %d (column %d, pass %d)", z, st[(size_t)z] - 1, pass); }
        V.swap(W);
    }
    // ---- MV = exp(A dT) V (TSExpADeltaT.c:25-34), slot by slot with the fused step kernels ---------------------------
    TRY(cudaMemcpyAsync(W.p, V.p, (size_t)m * U * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
    rc = unforced_steps_slots(h, direct, nv, W.as<cplx>(), m, g.ResRatio);
    if (rc != RSVDT_OK) { cleanup(); return rc; }
    // ---- Mt = V^H MV (CreateMtilde.c:49-51) -> Minv buffer reused -----------------------------------------------------
    TRY(launch_stack_gram2(V.as<cplx>(), U, m, W.as<cplx>(), U, m, nv, grow0, grows, part.as<cplx>(), Minv.as<cplx>(), h->stream));
    ALLRED(Minv.p, (int64_t)2 * nv * m * m);
    W.release();
    // ---- y1, y2 (GalerkinProj.c:30-46) -----------------------------------------------------------------------------------
    TRY(Y1.alloc((size_t)nw * U * sizeof(cplx)));
    TRY(Y2.alloc((size_t)nw * U * sizeof(cplx)));
    rc = ensure(h, h->coef, (size_t)std::max<int64_t>(g.Nw_eff * g.Nw, 128 * g.Nw_eff) * sizeof(cplx));
    if (rc != RSVDT_OK) { cleanup(); return rc; }
    TRY(launch_dft_coef(h->coef.as<cplx>(), nw, (int)g.Nw, g.ResRatio, h->stream));
    const cplx *s1 = snaps + (direct ? 0 : U), *s2 = snaps + (direct ? U : 0);
    TRY(launch_small_cgemm(Y1.as<cplx>(), U, h->coef.as<cplx>(), s1, U, nw, (int)g.Nw, U, 0, h->stream));
    TRY(launch_small_cgemm(Y2.as<cplx>(), U, h->coef.as<cplx>(), s2, U, nw, (int)g.Nw, U, 0, h->stream));
    // ---- b = V^H (alpha y1 - y2) through B1 = V^H Y1, B2 = V^H Y2 (:55) -------------------------------------------------
    TRY(B1.alloc((size_t)nv * m * nw * sizeof(cplx)));
    TRY(B2.alloc((size_t)nv * m * nw * sizeof(cplx)));
    TRY(launch_stack_gram2(V.as<cplx>(), U, m, Y1.as<cplx>(), U, nw, nv, grow0, grows, part.as<cplx>(), B1.as<cplx>(), h->stream));
    ALLRED(B1.p, (int64_t)2 * nv * m * nw);
    TRY(launch_stack_gram2(V.as<cplx>(), U, m, Y2.as<cplx>(), U, nw, nv, grow0, grows, part.as<cplx>(), B2.as<cplx>(), h->stream));
    ALLRED(B2.p, (int64_t)2 * nv * m * nw);
    Y2.release();
    // ---- (alpha I - Mt) c = b (:48-62), alpha = exp(+-i w iw dT), dT = dt*ResRatio (:17,49-54) --------------------------------
    std::vector<cplx> alpha((size_t)nw);
    const double deltaT = g.dt * (double)g.ResRatio;
    for (int iw = 0; iw < nw; iw++) {
        const double ph = (direct ? 1.0 : -1.0) * g.w * (double)iw * deltaT;
        alpha[(size_t)iw] = make_double2(std::cos(ph), std::sin(ph));
    }
    TRY(alpha_d.alloc((size_t)nw * sizeof(cplx)));
    TRY(cudaMemcpyAsync(alpha_d.p, alpha.data(), (size_t)nw * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    TRY(scratch.alloc((size_t)nv * nw * m * (m + 1) * sizeof(cplx)));
    TRY(Cc.alloc((size_t)nv * m * nw * sizeof(cplx)));
    TRY(cudaMemsetAsync(status.p, 0, (size_t)nv * nw * sizeof(int), h->stream));
    TRY(launch_galerkin_solve(Minv.as<cplx>(), B1.as<cplx>(), B2.as<cplx>(), alpha_d.as<cplx>(), m, nw, nv, scratch.as<cplx>(),
                              Cc.as<cplx>(), status.as<int>(), h->stream));
    // ---- Y_hat[:, iw] = y1 - V c (:65-71) ---------------------------------------------------------------------------------
    TRY(launch_stack_rmul(Yhat_out, U, Y1.as<cplx>(), U, 1.0, V.as<cplx>(), U, m, Cc.as<cplx>(), nw, -1.0, nv, rows, h->stream));
    h->launches += 12;
    TRY(cudaMemcpyAsync(st.data(), status.p, (size_t)nv * nw * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < st.size(); i++)
        if (st[i]) { cleanup(); return fail(h, RSVDT_ERR_NUMERIC, "transient removal: singular Galerkin system (test vector %d, frequency %d)", (int)(i / nw), (int)(i % nw)); }
#undef TRY
#undef ALLRED
    cleanup();
    return RSVDT_OK;
}

int check_ready(rsvdt_handle h)
{
    if (!h) return RSVDT_ERR_ARG;
    if (!h->A.set) return fail(h, RSVDT_ERR_STATE, "operator not set (rsvdt_set_operator_csr)");
    if (!h->grid_set) return fail(h, RSVDT_ERR_STATE, "grid not set (rsvdt_set_grid)");
    return RSVDT_OK;
}

// TSActionRK4.c:15-164 on the resident Y_hat (rows must equal N here; weights are applied by the callers)
int ts_action(rsvdt_handle h, bool direct)
{
    RS(check_ready(h));
    const rsvdt_grid &g = h->g;
    if (h->yhat_rows != h->N || !h->Yhat.p) return fail(h, RSVDT_ERR_STATE, "Y_hat not set or wrong row count");
    const int64_t U = h->n_ext * g.k;
    int64_t rend = g.Nt + g.Ns;                                   // TSActionRK4.c:41
    rend += g.TransientRemoval ? g.ResRatio : 0;                  // :42
    const int64_t Nstore = g.TransientRemoval ? g.Nw + 1 : g.Nw;  // :43

    DevBuf Fhat;
    Fhat.swap(h->Yhat);                                           // :49-50 ownership hand-off
    h->yhat_rows = 0;
    int rc = RSVDT_OK;
    do {
        if ((rc = ensure(h, h->y, (size_t)U * sizeof(cplx))) != RSVDT_OK) break;
        if ((rc = ensure(h, h->snaps, (size_t)Nstore * U * sizeof(cplx))) != RSVDT_OK) break;
        h->snaps_nstore = Nstore;
        cudaError_t e = launch_fill_zero(h->y.as<cplx>(), U, h->stream);  // :91
        if (e == cudaSuccess) e = launch_fill_zero(h->snaps.as<cplx>(), Nstore * U, h->stream);
        if (e != cudaSuccess) { rc = fail(h, RSVDT_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e)); break; }
        if ((rc = forced_steps(h, direct, Fhat.as<cplx>(), h->y.as<cplx>(), h->snaps.as<cplx>(), Nstore, 1, rend)) != RSVDT_OK) break;
        // the forcing is no longer needed (:139); its storage becomes the new Y_hat
        if (g.TransientRemoval) {
            h->slices[0].release();   // the forcing slices are dead: make room for the snapshot-sized work matrices
            h->slices[1].release();
            if ((rc = transient_removal(h, direct, h->snaps.as<cplx>(), Fhat.as<cplx>())) != RSVDT_OK) break;   // :147
        } else if ((rc = dft_snapshots(h, h->snaps.as<cplx>(), Fhat.as<cplx>())) != RSVDT_OK) break;       // :148
        cudaError_t es = cudaStreamSynchronize(h->stream);
        if (es != cudaSuccess) { rc = fail(h, RSVDT_ERR_CUDA, "action failed: %s", cudaGetErrorString(es)); break; }
        if ((rc = check_halo_timeout(h)) != RSVDT_OK) break;
    } while (0);
    if (rc == RSVDT_OK) {
        h->Yhat.swap(Fhat);
        h->yhat_rows = h->N;
        if (!h->keep_snapshots) h->snaps.release();
    }
    Fhat.release();
    return rc;
}

// ApplyWeightMats.c:16-80 : X <- M * X for sparse M on a device stack [iw][row][ik] (rows may change).
// Row-partitioned runs: M holds this rank's rows with columns indexing the all-gathered input; one frequency block at a time
// the owned rows of X are all-gathered (NCCL through the registered callback) and the local rows of the product formed.
// The weights are applied once per action, so this costs one pass over X per matrix against ~1e4 RK4 steps in between.
void stack_geom(rsvdt_handle h, int64_t rows, int64_t *srows, int64_t *roff);
int apply_sparse_to(rsvdt_handle h, const CsrOwned &M, DevBuf &X, int64_t &xrows)
{
    const rsvdt_grid &g = h->g;
    const bool part = h->dist.on;
    const int64_t want = part ? M.in_rows_local : M.d.ncols;
    if (xrows != want) return fail(h, RSVDT_ERR_ARG, "weight matrix expects %lld rows, the matrix it multiplies has %lld", (long long)want, (long long)xrows);
    DevBuf out;
    struct Release { DevBuf &b; ~Release() { b.release(); } } release_out{out};   // the product on failure, the old X on success
    if (!part) {
        CU(out.alloc(stack_bytes(g.Nw_eff, M.d.nrows, g.k)));
        CU(launch_spmm_stack(M.d, X.as<cplx>(), out.as<cplx>(), (int)g.Nw_eff, (int)g.k, h->stream));
        h->launches++;
    } else {
        if (h->dist.nranks > 1 && !h->dist.allgather) return fail(h, RSVDT_ERR_STATE, "row-partitioned run: no all-gather registered (rsvdt_set_allgather)");
        int64_t srows_in, roff_in, srows_out, roff_out;
        stack_geom(h, xrows, &srows_in, &roff_in);
        stack_geom(h, M.d.nrows, &srows_out, &roff_out);
        CU(out.alloc(stack_bytes(g.Nw_eff, srows_out, g.k)));
        if (srows_out != M.d.nrows) CU(cudaMemsetAsync(out.p, 0, stack_bytes(g.Nw_eff, srows_out, g.k), h->stream));
        const size_t blk = (size_t)M.gather_rows * g.k * sizeof(cplx);
        RS(ensure(h, h->dist.gather_send, blk));
        RS(ensure(h, h->dist.gather_recv, blk * (size_t)h->dist.nranks));
        CU(cudaMemsetAsync(h->dist.gather_send.p, 0, blk, h->stream));
        for (int64_t a = 0; a < g.Nw_eff; a++) {
            const cplx *src = X.as<cplx>() + ((size_t)a * srows_in + roff_in) * g.k;
            CU(cudaMemcpyAsync(h->dist.gather_send.p, src, (size_t)xrows * g.k * sizeof(cplx), cudaMemcpyDeviceToDevice, h->stream));
            if (h->dist.nranks > 1) {
                const int rc = h->dist.allgather(h->dist.gather_send.p, h->dist.gather_recv.p, (int64_t)(blk / sizeof(double)), (void *)h->stream,
                                                 h->dist.allgather_user);
                if (rc) return fail(h, RSVDT_ERR_CUDA, "all-gather callback failed (%d)", rc);
            } else {
                CU(cudaMemcpyAsync(h->dist.gather_recv.p, h->dist.gather_send.p, blk, cudaMemcpyDeviceToDevice, h->stream));
            }
            cplx *dst = out.as<cplx>() + ((size_t)a * srows_out + roff_out) * g.k;
            CU(launch_spmm_stack(M.d, h->dist.gather_recv.as<cplx>(), dst, 1, (int)g.k, h->stream));
            h->launches++;
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    X.swap(out);
    xrows = M.d.nrows;
    return RSVDT_OK;
}

int apply_sparse(rsvdt_handle h, const CsrOwned &M) { return apply_sparse_to(h, M, h->Yhat, h->yhat_rows); }

int action_with_weights(rsvdt_handle h, bool direct)
{
    RS(check_ready(h));
    if (direct) {  // ApplyWeightMats.c:18-28 then :29-41
        if (h->W[0].set) RS(apply_sparse(h, h->W[0]));   // W_f_sqrt_inv
        if (h->W[1].set) RS(apply_sparse(h, h->W[1]));   // B
        RS(ts_action(h, true));
        if (h->W[4].set) RS(apply_sparse(h, h->W[4]));   // C
        if (h->W[3].set) RS(apply_sparse(h, h->W[3]));   // W_q_sqrt
    } else {       // ApplyWeightMats.c:43-60 then :61-78
        if (h->W[3].set) RS(apply_sparse(h, h->WH[3]));  // W_q_sqrt^H
        if (h->W[4].set) RS(apply_sparse(h, h->WH[4]));  // C^H
        RS(ts_action(h, false));
        if (h->W[1].set) RS(apply_sparse(h, h->WH[1]));  // B^H
        if (h->W[0].set) RS(apply_sparse(h, h->WH[0]));  // W_f_sqrt_inv^H
    }
    return RSVDT_OK;
}

int dist_allreduce(rsvdt_handle h, void *dev_ptr, int64_t n_doubles)
{
    if (!h->dist.on || h->dist.nranks == 1) return RSVDT_OK;
    if (!h->dist.allreduce) return fail(h, RSVDT_ERR_STATE, "row-partitioned run: no all-reduce registered (rsvdt_set_allreduce)");
    const int rc = h->dist.allreduce(dev_ptr, n_doubles, (void *)h->stream, h->dist.allreduce_user);
    if (rc) return fail(h, RSVDT_ERR_CUDA, "all-reduce callback failed (%d)", rc);
    return RSVDT_OK;
}

// Gram matrices of the per-frequency blocks of a stack with `srows` rows per slot, summed over the owned rows
// [roff, roff + rows) and over the ranks
// transform (optional, device, Nw_eff x k x k): the Gram matrices of Y_w * transform_w, formed on the fly (dense_k4.cu)
int gram_all(rsvdt_handle h, const cplx *Y, int64_t rows, int64_t srows = -1, int64_t roff = 0, const cplx *transform = nullptr)
{
    const rsvdt_grid &g = h->g;
    if (srows < 0) srows = rows;
    const int nblk = gram_num_blocks(rows);
    RS(ensure(h, h->small, (size_t)g.Nw_eff * g.k * g.k * sizeof(cplx)));
    RS(ensure(h, h->smallM, (size_t)g.Nw_eff * g.k * g.k * sizeof(cplx)));
    RS(ensure(h, h->status, (size_t)g.Nw_eff * sizeof(int)));
    // FP64 tensor cores (DMMA) unless switched off: ncu (profiles/r2_dense_phases_ncu.txt) has the comparison with the FMA kernel
    if (h->gram_variant != 0 || transform) {
        if (transform && g.k > 16) return fail(h, RSVDT_ERR_ARG, "fused CholeskyQR2 pass needs k <= 16");
        RS(ensure(h, h->gram_partial, gram_dmma_partial_doubles(rows, (int)g.k, (int)g.Nw_eff) * sizeof(double)));
        CU(launch_gram_dmma(Y + roff * g.k, rows, (int)g.k, (int)g.Nw_eff, srows * g.k, transform, h->gram_partial.as<double>(),
                            h->small.as<cplx>(), h->stream));
    } else {
        RS(ensure(h, h->gram_partial, (size_t)nblk * g.Nw_eff * g.k * g.k * sizeof(cplx)));
        CU(launch_gram(Y + roff * g.k, rows, (int)g.k, (int)g.Nw_eff, srows * g.k, h->gram_partial.as<cplx>(), h->small.as<cplx>(), h->stream));
    }
    h->launches += 2;
    RS(dist_allreduce(h, h->small.p, 2 * g.Nw_eff * g.k * g.k));
    return RSVDT_OK;
}

// slot geometry of a stack holding `rows` owned rows: N-row matrices live on the halo-extended rows
void stack_geom(rsvdt_handle h, int64_t rows, int64_t *srows, int64_t *roff)
{
    if (rows == h->N) { *srows = h->n_ext; *roff = h->own0; }
    else { *srows = rows; *roff = 0; }
}

// per-frequency cross-product SVD: Y_w <- left singular vectors; sigma -> dst_sigma (device, k x Nw_eff) if not null
int svd_all(rsvdt_handle h, DevBuf &Y, int64_t rows, double *dst_sigma)
{
    const rsvdt_grid &g = h->g;
    int64_t srows, roff;
    stack_geom(h, rows, &srows, &roff);
    RS(gram_all(h, Y.as<cplx>(), rows, srows, roff));
    DevBuf sig;
    CU(sig.alloc((size_t)g.Nw_eff * g.k * sizeof(double)));
    CU(launch_cross_svd_small(h->small.as<cplx>(), (int)g.k, (int)g.Nw_eff, h->smallM.as<cplx>(), sig.as<double>(), h->stream));
    if (h->gram_variant != 0 && g.k <= 16 && g.k % 2 == 0)
        CU(launch_apply_dmma(Y.as<cplx>(), srows, (int)g.k, (int)g.Nw_eff, srows * g.k, h->smallM.as<cplx>(), nullptr, nullptr, h->stream));
    else
        CU(launch_apply_right(Y.as<cplx>(), srows, (int)g.k, (int)g.Nw_eff, h->smallM.as<cplx>(), h->stream));
    h->launches += 2;
    if (dst_sigma) CU(cudaMemcpyAsync(dst_sigma, sig.p, (size_t)g.Nw_eff * g.k * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    sig.release();
    return RSVDT_OK;
}

}  // namespace

// =====================================================================================================
namespace {

// Tables of the exchange fused into the tiled stage kernel (stage_tiled.cuh, HaloFuse).  Not an error if the partition does
// not qualify (rows wanted by more than two peers, requests that are not whole block rows, more than 8 peers): the
// unfused exchange stays in use.
int build_fused(rsvdt_handle h)
{
    auto &D = h->dist;
    D.fz_ready = false;
    if (D.layout != 0 || !h->A.tbcsr_valid) return RSVDT_OK;
    if (D.out.size() > (size_t)TBCSR_MAX_PEERS || D.in.size() > (size_t)TBCSR_MAX_PEERS) return RSVDT_OK;
    const int B = h->A.tb.B, TB = h->A.tb.TB;
    const int64_t ntiles = h->A.tb.ntiles;
    if (h->AH.set && (!h->AH.tbcsr_valid || h->AH.tb.B != B || h->AH.tb.ntiles != ntiles)) return RSVDT_OK;
    // mailbox rows of the fused slots are ghost rows: incoming message i must fill ghost rows [mb_off_rows, +count)
    for (auto &m : D.in)
        if (m.dest_row0 - h->N != m.mb_off_rows) return RSVDT_OK;
    std::vector<int32_t> send_index((size_t)ntiles, -1);
    std::vector<int64_t> send_dest;
    std::vector<int32_t> out_tiles(std::max<size_t>(D.out.size(), 1), 0);
    for (size_t i = 0; i < D.out.size(); i++) {
        const auto &m = D.out[i];
        if (m.count % B) return RSVDT_OK;
        std::vector<unsigned char> seen((size_t)ntiles, 0);
        for (int j = 0; j < m.count; j += B) {
            const int64_t r0 = m.rows_h[(size_t)j];
            if (r0 % B || r0 >= h->N) return RSVDT_OK;
            for (int r = 1; r < B; r++)
                if (m.rows_h[(size_t)(j + r)] != r0 + r) return RSVDT_OK;
            const int64_t br = r0 / B, t = br / TB;
            const int tbi = (int)(br % TB);
            if (send_index[(size_t)t] < 0) {
                send_index[(size_t)t] = (int32_t)(send_dest.size() / (size_t)(2 * TB));
                send_dest.resize(send_dest.size() + (size_t)(2 * TB), -1);
            }
            int64_t *d = send_dest.data() + (size_t)send_index[(size_t)t] * 2 * TB;
            const int64_t v = ((int64_t)i << 48) | (m.mb_off_rows + j);
            if (d[tbi] < 0) d[tbi] = v;
            else if (d[TB + tbi] < 0) d[TB + tbi] = v;
            else return RSVDT_OK;
            if (!seen[(size_t)t]) { seen[(size_t)t] = 1; out_tiles[i]++; }
        }
    }
    std::vector<int32_t> send_mask(send_dest.size() / (size_t)(2 * TB), 0);
    for (size_t si = 0; si < send_mask.size(); si++)
        for (int j = 0; j < 2 * TB; j++) {
            const int64_t dd = send_dest[si * 2 * TB + (size_t)j];
            if (dd >= 0) send_mask[si] |= 1 << (int)(dd >> 48);
        }
    CU(D.fz_send_mask.alloc(std::max<size_t>(send_mask.size(), 1) * sizeof(int32_t)));
    if (!send_mask.empty()) CU(cudaMemcpy(D.fz_send_mask.p, send_mask.data(), send_mask.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(D.fz_send_index.alloc(std::max<size_t>(send_index.size(), 1) * sizeof(int32_t)));
    CU(D.fz_send_dest.alloc(std::max<size_t>(send_dest.size(), 1) * sizeof(int64_t)));
    CU(D.fz_out_tiles.alloc(out_tiles.size() * sizeof(int32_t)));
    CU(D.fz_out_done.alloc(out_tiles.size() * sizeof(unsigned long long)));
    if (!send_index.empty()) CU(cudaMemcpy(D.fz_send_index.p, send_index.data(), send_index.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!send_dest.empty()) CU(cudaMemcpy(D.fz_send_dest.p, send_dest.data(), send_dest.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(D.fz_out_tiles.p, out_tiles.data(), out_tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(cudaMemset(D.fz_out_done.p, 0, out_tiles.size() * sizeof(unsigned long long)));
    // per operator: tile flags (bit0 reads ghosts, bit1 sends) and the processing order (sending tiles, then ghost readers, then the rest)
    for (CsrOwned *Op : {&h->A, &h->AH}) {
        if (!Op->set || !Op->tbcsr_valid) continue;
        std::vector<unsigned char> fl((size_t)ntiles, 0);
        for (int64_t t = 0; t < ntiles; t++)
            fl[(size_t)t] = (unsigned char)((!Op->tile_ghost.empty() && Op->tile_ghost[(size_t)t] ? 1 : 0) | (send_index[(size_t)t] >= 0 ? 2 : 0));
        std::vector<int32_t> order;
        order.reserve((size_t)ntiles);
        for (int pass = 0; pass < 3; pass++)
            for (int64_t i = 0; i < ntiles; i++) {
                const int64_t t = Op->base_order.empty() ? i : Op->base_order[(size_t)i];
                const int cls = (fl[(size_t)t] & 2) ? 0 : (fl[(size_t)t] ? 1 : 2);
                if (cls == pass) order.push_back((int32_t)t);
            }
        CU(Op->t_flags.alloc(std::max<size_t>(fl.size(), 1)));
        CU(Op->t_order.alloc(std::max<size_t>(order.size(), 1) * sizeof(int32_t)));
        CU(cudaMemcpy(Op->t_flags.p, fl.data(), fl.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(Op->t_order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        Op->tb.tile_order = Op->t_order.as<int32_t>();
    }
    D.fz_ready = true;
    return RSVDT_OK;
}

// a stage kernel gave up waiting for a peer's halo rows (stage_tiled.cu): report it instead of returning garbage
int check_halo_timeout(rsvdt_handle h)
{
    auto &D = h->dist;
    if (!D.on || !D.flags.p) return RSVDT_OK;
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, D.flags.as<unsigned long long>() + 2 * D.nranks + 1, sizeof v, cudaMemcpyDeviceToHost));
    if (v) return fail(h, RSVDT_ERR_CUDA, "halo exchange: a peer GPU did not publish its boundary rows (rank %d gave up waiting)", D.rank);
    return RSVDT_OK;
}

// one operator (A or A^H) of a synthetic block-stencil generated on the device in the tiled block-CSR layout
int build_stencil_operator(rsvdt_handle h, CsrOwned &Op, StencilSpec sp, int64_t n_ext)
{
    Op.release();
    const int TB = TBCSR_TB, B = sp.B;
    const int64_t nbrows = sp.pt1 - sp.pt0;
    const int64_t ntiles = (nbrows + TB - 1) / TB;
    DevBuf counts, masks, hasg;
    CU(counts.alloc((size_t)ntiles * sizeof(int32_t)));
    CU(masks.alloc((size_t)ntiles * sizeof(uint32_t)));
    CU(stencil_tile_counts(sp, TB, ntiles, counts.as<int32_t>(), masks.as<uint32_t>(), h->stream));
    std::vector<int32_t> cnt((size_t)ntiles);
    CU(cudaMemcpyAsync(cnt.data(), counts.p, (size_t)ntiles * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<int64_t> tptr((size_t)ntiles + 1, 0);
    for (int64_t t = 0; t < ntiles; t++) tptr[(size_t)t + 1] = tptr[(size_t)t] + cnt[(size_t)t];
    const size_t cb = tbcsr_chunk_bytes(B, TB);
    const int64_t nchunks = tptr[(size_t)ntiles];
    CU(Op.t_tileptr.alloc(tptr.size() * sizeof(int64_t)));
    CU(Op.t_chunks.alloc((size_t)std::max<int64_t>(nchunks, 1) * cb));
    CU(hasg.alloc((size_t)ntiles));
    CU(cudaMemsetAsync(hasg.p, 0, (size_t)ntiles, h->stream));
    CU(cudaMemcpyAsync(Op.t_tileptr.p, tptr.data(), tptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CU(stencil_fill(sp, TB, ntiles, Op.t_tileptr.as<int64_t>(), masks.as<uint32_t>(), Op.t_chunks.as<unsigned char>(), hasg.as<unsigned char>(), h->stream));
    Op.tile_ghost.assign((size_t)ntiles, 0);
    CU(cudaMemcpyAsync(Op.tile_ghost.data(), hasg.p, (size_t)ntiles, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    {   // x windows: the chunk columns come back (4 B per block), the plan goes into the chunk headers
        std::vector<int32_t> ccols((size_t)std::max<int64_t>(nchunks, 1) * TB);
        const size_t coff = tbcsr_cols_offset();
        if (nchunks) CU(cudaMemcpy2D(ccols.data(), (size_t)TB * sizeof(int32_t), Op.t_chunks.as<unsigned char>() + coff, cb,
                                     (size_t)TB * sizeof(int32_t), (size_t)nchunks, cudaMemcpyDeviceToHost));
        TbcsrPlan plan;
        tbcsr_plan_windows(ntiles, tptr.data(), ccols.data(), nbrows, n_ext / B, TB, TBCSR_WMAX, plan);
        if (nchunks) CU(cudaMemcpy2D(Op.t_chunks.as<unsigned char>() + coff + (size_t)TB * sizeof(int32_t), cb, plan.hdr.data(), 16, 16,
                                     (size_t)nchunks, cudaMemcpyHostToDevice));
        CU(Op.t_winptr.alloc(plan.win_ptr.size() * sizeof(int64_t)));
        CU(Op.t_win.alloc(std::max<size_t>(plan.win.size(), 1) * sizeof(TbcsrWindow)));
        CU(cudaMemcpy(Op.t_winptr.p, plan.win_ptr.data(), plan.win_ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        if (!plan.win.empty()) CU(cudaMemcpy(Op.t_win.p, plan.win.data(), plan.win.size() * sizeof(TbcsrWindow), cudaMemcpyHostToDevice));
        Op.tb.win_ptr = Op.t_winptr.as<int64_t>();
        Op.tb.win = Op.t_win.as<TbcsrWindow>();
        Op.tb.wmax = 1;
        for (auto &w : plan.win) Op.tb.wmax = std::max(Op.tb.wmax, (int)w.nblocks);
        Op.tb.nwindows_chunks = plan.window_chunks;
    }
    h->launches += 2;
    Op.tb.B = B;
    Op.tb.TB = TB;
    Op.tb.nbrows = nbrows;
    Op.tb.nbcols_own = nbrows;
    Op.tb.ntiles = ntiles;
    Op.tb.nchunks = nchunks;
    Op.tb.tile_ptr = Op.t_tileptr.as<int64_t>();
    Op.tb.chunks = Op.t_chunks.as<unsigned char>();
    Op.tb.chunk_bytes = cb;
    Op.tbcsr_valid = true;
    Op.stencil_only = true;
    if (sp.d == 3 && h->tiled_order > 0 && ntiles > 1) {
        // L2 locality: row order re-reads a row of the stage vector one and two grid planes later (51 MB apart at 200 x 200 x
        // 5 x 8 vectors), which the operator stream flushes out of L2 in between.  Walk the planes fastest instead: groups of
        // `tiled_order` grid lines, each group through all local planes, so that the reuse distance is a few hundred KB.
        const int64_t G = h->tiled_order, ny = sp.shape[1], nx = sp.shape[2];
        std::vector<std::pair<int64_t, int32_t>> key((size_t)ntiles);
        for (int64_t t = 0; t < ntiles; t++) {
            const int64_t q = sp.pt0 + t * TB, z = q / (ny * nx), y = (q / nx) % ny;
            key[(size_t)t] = {((y / G) * (sp.shape[0] + 1) + z) * (G * nx + TB) + (y % G) * nx + q % nx, (int32_t)t};
        }
        std::sort(key.begin(), key.end());
        Op.base_order.resize((size_t)ntiles);
        for (int64_t t = 0; t < ntiles; t++) Op.base_order[(size_t)t] = key[(size_t)t].second;
        CU(Op.t_order.alloc((size_t)ntiles * sizeof(int32_t)));
        CU(cudaMemcpy(Op.t_order.p, Op.base_order.data(), (size_t)ntiles * sizeof(int32_t), cudaMemcpyHostToDevice));
        Op.tb.tile_order = Op.t_order.as<int32_t>();
    }
    int nb = 1;
    int64_t band = 1;
    for (int i = 0; i < sp.d; i++) nb *= 3;
    for (int i = sp.d - 1, st = 1; i >= 0; st *= (int)sp.shape[i], i--) band += st;
    Op.d = CsrDev();
    Op.d.nrows = nbrows * B;
    Op.d.ncols = n_ext;
    Op.d.nnz = nchunks * (int64_t)TB * B * B;
    Op.d.max_row_nnz = (int64_t)nb * B;
    Op.lo_bw = Op.hi_bw = band * B;   // (keeps the band kernels away from an operator that has no CSR form)
    Op.set = true;
    return RSVDT_OK;
}

}  // namespace

#include "group.inc"

extern "C" {

const char *rsvdt_version(void) { return "rsvdt-b200 0.1 (sm_100a)"; }

int rsvdt_create(int device, rsvdt_handle *hout)
{
    if (!hout) return RSVDT_ERR_ARG;
    *hout = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return RSVDT_ERR_CUDA;  // no CPU fallback: fail loudly
    if (device < 0 || device >= ndev) return RSVDT_ERR_ARG;
    rsvdt_handle h = new (std::nothrow) rsvdt_context();
    if (!h) return RSVDT_ERR_MEM;
    h->device = device;
    if (const char *e = getenv("RSVDT_TILED_ORDER")) h->tiled_order = atoll(e);   // measurement hook (set before the operator)
    if (const char *e = getenv("RSVDT_TILED_VARIANT")) h->tiled_variant = atoll(e);   // (the same, for runs of the test suite under a variant)
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_ex, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_step0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_synth, cudaEventDisableTiming) != cudaSuccess) {
        delete h;
        return RSVDT_ERR_CUDA;
    }
    {
        int lo = 0, hi = 0;   // numerically lower = higher priority; the synthesis stream gets the lowest
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // ... and the exchange stream the highest: its one CTA must get the first slot a step-kernel CTA frees, not the last
        if (cudaStreamCreateWithPriority(&h->synth_stream, cudaStreamNonBlocking, lo) != cudaSuccess ||
            cudaStreamCreateWithPriority(&h->xstream, cudaStreamNonBlocking, hi) != cudaSuccess) {
            delete h;
            return RSVDT_ERR_CUDA;
        }
    }
    *hout = h;
    return RSVDT_OK;
}

int rsvdt_create_group(int ngpus, const int *devices, rsvdt_handle *hout)
{
    if (!hout || ngpus < 1 || ngpus > 64) return RSVDT_ERR_ARG;
    *hout = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < ngpus) return RSVDT_ERR_CUDA;
    rsvdt_handle root = nullptr;
    int rc = rsvdt_create(devices ? devices[0] : 0, &root);
    if (rc) return rc;
    rsvdt_group_ctx *G = new (std::nothrow) rsvdt_group_ctx();
    if (!G) { rsvdt_destroy(root); return RSVDT_ERR_MEM; }
    G->P = ngpus;
    root->group = G;
    for (int p = 0; p < ngpus; p++) {
        rsvdt_handle mh = nullptr;
        rc = rsvdt_create(devices ? devices[p] : p, &mh);
        if (rc) { rsvdt_destroy(root); return rc; }
        G->m.push_back(mh);
    }
    *hout = root;
    return RSVDT_OK;
}

int rsvdt_destroy(rsvdt_handle h)
{
    if (!h) return RSVDT_OK;
    if (h->group) {
        rsvdt_group_ctx *G = h->group;
        for (size_t p = 0; p < G->comm.size(); p++)
            if (G->comm[p]) { cudaSetDevice(G->m[p]->device); nccl().CommDestroy(G->comm[p]); }
        for (rsvdt_handle mh : G->m) rsvdt_destroy(mh);
        delete G;
        h->group = nullptr;
    }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    int64_t n;
    drain(h->ev_stage, &n); drain(h->ev_idft, &n); drain(h->ev_dft, &n);
    h->A.release(); h->AH.release();
    for (int i = 0; i < 5; i++) { h->W[i].release(); h->WH[i].release(); }
    for (auto &m : h->dist.out) m.rows.release();
    for (int p = 0; p < (int)h->dist.peer_mailbox.size() && !h->dist.peer_mapped_directly; p++) {
        if (p == h->dist.rank) continue;
        if (h->dist.peer_mailbox[(size_t)p]) cudaIpcCloseMemHandle(h->dist.peer_mailbox[(size_t)p]);
        if (h->dist.peer_flags[(size_t)p]) cudaIpcCloseMemHandle(h->dist.peer_flags[(size_t)p]);
    }
    DevBuf *dbufs[] = {&h->dist.mailbox, &h->dist.flags, &h->dist.msgs_dev, &h->dist.ins_dev, &h->dist.peer_flags_dev,
                       &h->dist.fz_send_index, &h->dist.fz_send_dest, &h->dist.fz_send_mask, &h->dist.fz_out_tiles, &h->dist.fz_out_done};
    for (DevBuf *b : dbufs) b->release();
    DevBuf *bufs[] = {&h->Yhat, &h->Uhat, &h->Vhat, &h->Shat, &h->y, &h->xa, &h->xb, &h->acc, &h->slices[0], &h->slices[1],
                      &h->coef, &h->snaps, &h->small, &h->smallM, &h->status, &h->gram_partial, &h->scalars};
    for (DevBuf *b : bufs) b->release();
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            if (h->graphs[a][b].exec) cudaGraphExecDestroy(h->graphs[a][b].exec);
    cudaEventDestroy(h->ev_fork);
    cudaEventDestroy(h->ev_join);
    cudaEventDestroy(h->ev_step0);
    cudaEventDestroy(h->ev_synth);
    cudaStreamDestroy(h->synth_stream);
    cudaStreamDestroy(h->side);
    cudaEventDestroy(h->ev_ex);
    cudaStreamDestroy(h->xstream);
    cudaStreamDestroy(h->stream);
    delete h;
    return RSVDT_OK;
}

const char *rsvdt_last_error(rsvdt_handle h) { return h ? h->err.c_str() : "null handle"; }

int rsvdt_setup_time_freq_grid(double w, int64_t Nw, double dt_in, double TransientLength, double *dt, int64_t *Ns,
                               int64_t *Nt, int64_t *ResRatio)
{
    if (!dt || !Ns || !Nt || !ResRatio || Nw <= 0 || !(w > 0) || !(dt_in > 0)) return RSVDT_ERR_ARG;
    const double PI = 3.14159265358979323846264338327950288;  // PETSC_PI
    const double T_ss = 2 * PI / w;                           // SetupTimeFreqGrid.c:16
    const double dt_w = T_ss / (double)Nw;                    // :17
    *dt = dt_w / std::ceil(dt_w / dt_in);                     // :18
    *Ns = (int64_t)std::round(T_ss / *dt);                    // :19
    *Nt = (int64_t)std::round(TransientLength / *dt);         // :20
    *ResRatio = *Ns / Nw;                                     // :21
    return RSVDT_OK;
}

int rsvdt_set_operator_csr(rsvdt_handle h, int64_t N, const int64_t *rowptr, const int64_t *col, const double *val,
                           int disc_flg, double beta, int *real_operator)
{
    if (h && h->group) return group_set_operator_csr(h, N, rowptr, col, val, disc_flg, beta, real_operator);
    RS(bind(h));
    if (N <= 0 || !rowptr || !col || !val) return fail(h, RSVDT_ERR_ARG, "bad operator arguments");
    if (disc_flg && beta < 0) return fail(h, RSVDT_ERR_USER, "'beta' must be positive, current value: %g", beta);  // ReadUserInput.c:133
    const cplx *v = reinterpret_cast<const cplx *>(val);
    const int64_t nnz = rowptr[N];
    double im2 = 0.0;                                         // LNSStructure.c:37-40
    for (int64_t p = 0; p < nnz; p++) im2 += v[p].y * v[p].y;
    h->real_operator = std::sqrt(im2) < 1e-12;
    if (real_operator) *real_operator = h->real_operator ? 1 : 0;
    const double shift = disc_flg ? beta : 0.0;               // LNSStructure.c:53-56
    h->N = N;
    h->n_ext = N;
    h->own0 = 0;
    h->N_global = N;
    h->dist = rsvdt_context::Dist();
    RS(upload_csr(h, h->A, N, N, rowptr, col, v, false, shift));
    RS(upload_csr(h, h->AH, N, N, rowptr, col, v, true, shift));
    return RSVDT_OK;
}

int rsvdt_set_grid(rsvdt_handle h, const rsvdt_grid *g)
{
    if (h && h->group) return (g ? group_set_grid(h, g) : RSVDT_ERR_ARG);
    RS(bind(h));
    if (!g) return fail(h, RSVDT_ERR_ARG, "null grid");
    if (g->k <= 0 || g->Nw <= 0 || g->Nw_eff <= 0 || g->Ns <= 0 || g->Nt < 0 || g->ResRatio <= 0 || !(g->dt > 0))
        return fail(h, RSVDT_ERR_ARG, "grid parameters out of range");
    if (g->k > RSVDT_MAX_K) return fail(h, RSVDT_ERR_ARG, "k > %d is not supported by the per-frequency QR/SVD kernels", RSVDT_MAX_K);
    if (g->Nw_eff != (g->RealOperator ? g->Nw / 2 : g->Nw)) return fail(h, RSVDT_ERR_ARG, "Nw_eff inconsistent with RealOperator (LNSStructure.c:44)");
    h->g = *g;
    h->grid_set = true;
    return RSVDT_OK;
}

int rsvdt_set_weight_csr(rsvdt_handle h, int which, int64_t nrows, int64_t ncols, const int64_t *rowptr,
                         const int64_t *col, const double *val)
{
    if (h && h->group) return group_set_weight_csr(h, which, nrows, ncols, rowptr, col, val);
    RS(bind(h));
    if (which < 0 || which > 4 || !rowptr || !col || !val) return fail(h, RSVDT_ERR_ARG, "bad weight arguments");
    if (!h->A.set) return fail(h, RSVDT_ERR_STATE, "set the operator before the weights");
    if (h->dist.on) return fail(h, RSVDT_ERR_ARG, "row-partitioned run: pass the owned rows with rsvdt_set_weight_csr_local");
    static const char *names[5] = {"W_f_sqrt_inv", "B", "W_q_sqrt_inv", "W_q_sqrt", "C"};
    if ((which == 0 || which == 2 || which == 3) && nrows != ncols)   // ReadWeightMats.c:34,64,75
        return fail(h, RSVDT_ERR_USER, "weight matrix (%s) must be square, current size = %lld x %lld", names[which], (long long)nrows, (long long)ncols);
    if (which == 1 && nrows != h->N) return fail(h, RSVDT_ERR_USER, "Input matrix (B) must have %lld rows", (long long)h->N);        // :49
    if (which == 4 && ncols != h->N) return fail(h, RSVDT_ERR_USER, "Output matrix (C) must have %lld columns", (long long)h->N);   // :89
    const cplx *v = reinterpret_cast<const cplx *>(val);
    RS(upload_csr(h, h->W[which], nrows, ncols, rowptr, col, v, false, 0.0));
    RS(upload_csr(h, h->WH[which], nrows, ncols, rowptr, col, v, true, 0.0));
    return RSVDT_OK;
}

// PETSC_DECIDE split of n rows over nranks: first row of `rank`
static int64_t split_begin(int64_t n, int nranks, int rank) { return (int64_t)rank * (n / nranks) + std::min<int64_t>(rank, n % nranks); }

// columns of a row-partitioned weight matrix: global column -> position in the all-gathered block [rank][gather_rows]
static int remap_gathered(rsvdt_handle h, int64_t ncols_g, int64_t nnz, const int64_t *col, std::vector<int64_t> &out, int64_t *gather_rows)
{
    const int P = h->dist.nranks;
    const int64_t q = ncols_g / P, r = ncols_g % P, big = q + (r ? 1 : 0);
    *gather_rows = std::max<int64_t>(big, 1);
    out.resize((size_t)std::max<int64_t>(nnz, 1));
    for (int64_t p = 0; p < nnz; p++) {
        const int64_t c = col[p];
        if (c < 0 || c >= ncols_g) return fail(h, RSVDT_ERR_ARG, "weight matrix: column index %lld out of range", (long long)c);
        const int64_t owner = (c < r * (q + 1)) ? c / (q + 1) : r + (q ? (c - r * (q + 1)) / q : 0);
        out[(size_t)p] = owner * (*gather_rows) + (c - split_begin(ncols_g, P, (int)owner));
    }
    return RSVDT_OK;
}

int rsvdt_set_weight_csr_local(rsvdt_handle h, int which, int64_t nrows_global, int64_t ncols_global, int64_t n_local,
                               const int64_t *rowptr, const int64_t *col_global, const double *val, int64_t nH_local,
                               const int64_t *rowptr_H, const int64_t *col_global_H, const double *val_H)
{
    RS(bind(h));
    if (which < 0 || which > 4 || !rowptr || !col_global || !val || !rowptr_H || !col_global_H || !val_H) return fail(h, RSVDT_ERR_ARG, "bad weight arguments");
    if (!h->dist.on) return fail(h, RSVDT_ERR_STATE, "rsvdt_set_weight_csr_local needs rsvdt_dist_init; use rsvdt_set_weight_csr otherwise");
    if (!h->A.set) return fail(h, RSVDT_ERR_STATE, "set the operator before the weights");
    static const char *names[5] = {"W_f_sqrt_inv", "B", "W_q_sqrt_inv", "W_q_sqrt", "C"};
    if ((which == 0 || which == 2 || which == 3) && nrows_global != ncols_global)   // ReadWeightMats.c:34,64,75
        return fail(h, RSVDT_ERR_USER, "weight matrix (%s) must be square, current size = %lld x %lld", names[which], (long long)nrows_global, (long long)ncols_global);
    if (which == 1 && nrows_global != h->N_global) return fail(h, RSVDT_ERR_USER, "Input matrix (B) must have %lld rows", (long long)h->N_global);        // :49
    if (which == 4 && ncols_global != h->N_global) return fail(h, RSVDT_ERR_USER, "Output matrix (C) must have %lld columns", (long long)h->N_global);   // :89
    const int P = h->dist.nranks, me = h->dist.rank;
    const int64_t own_r = split_begin(nrows_global, P, me + 1) - split_begin(nrows_global, P, me);
    const int64_t own_c = split_begin(ncols_global, P, me + 1) - split_begin(ncols_global, P, me);
    if (n_local != own_r || nH_local != own_c)
        return fail(h, RSVDT_ERR_ARG, "%s: rank %d must pass %lld rows of M and %lld rows of M^H (PETSC_DECIDE split), got %lld and %lld", names[which], me,
                    (long long)own_r, (long long)own_c, (long long)n_local, (long long)nH_local);
    std::vector<int64_t> cm;
    int64_t gr = 0;
    RS(remap_gathered(h, ncols_global, rowptr[n_local], col_global, cm, &gr));
    RS(upload_csr(h, h->W[which], n_local, gr * P, rowptr, cm.data(), reinterpret_cast<const cplx *>(val), false, 0.0));
    h->W[which].g_nrows = nrows_global; h->W[which].g_ncols = ncols_global; h->W[which].gather_rows = gr; h->W[which].in_rows_local = own_c;
    RS(remap_gathered(h, nrows_global, rowptr_H[nH_local], col_global_H, cm, &gr));
    RS(upload_csr(h, h->WH[which], nH_local, gr * P, rowptr_H, cm.data(), reinterpret_cast<const cplx *>(val_H), false, 0.0));
    h->WH[which].g_nrows = ncols_global; h->WH[which].g_ncols = nrows_global; h->WH[which].gather_rows = gr; h->WH[which].in_rows_local = own_r;
    return RSVDT_OK;
}

// Row-partitioned runs in local layout 1 advance with the fused band kernel on halo-extended rows; whether this rank's rows of
// A and A^H qualify (diagonal pattern, tile regularity) is a purely local, host-side question -- answered here without
// touching the exchange, so that the ranks can agree (all-reduce of *ok) BEFORE the first collective step.
int rsvdt_dist_layout_supported(rsvdt_handle h, int nvec, int *ok)
{
    RS(bind(h));
    if (!ok || nvec < 1) return fail(h, RSVDT_ERR_ARG, "bad arguments");
    *ok = 1;
    if (!h->dist.on || h->dist.layout != 1) return RSVDT_OK;
    if (!h->A.set || !h->AH.set) return fail(h, RSVDT_ERR_STATE, "operator not set");
    RS(plan_dia(h, h->A, nvec, 3));
    const bool a = h->A.plan_ok;
    RS(plan_dia(h, h->AH, nvec, 3));
    *ok = (a && h->AH.plan_ok) ? 1 : 0;
    return RSVDT_OK;
}

int rsvdt_set_allgather(rsvdt_handle h, rsvdt_allgather_fn fn, void *user)
{
    RS(bind(h));
    h->dist.allgather = fn;
    h->dist.allgather_user = user;
    return RSVDT_OK;
}

int64_t rsvdt_yhat_rows(rsvdt_handle h) { return h ? h->yhat_rows : -1; }

int rsvdt_set_yhat(rsvdt_handle h, const double *Y_host, int64_t nrows)
{
    if (h && h->group) return group_set_yhat(h, Y_host, nrows);
    RS(bind(h));
    if (!h->grid_set) return fail(h, RSVDT_ERR_STATE, "grid not set");
    if (!Y_host || nrows <= 0) return fail(h, RSVDT_ERR_ARG, "bad Y_hat arguments");
    int64_t srows, roff;
    stack_geom(h, nrows, &srows, &roff);
    const bool weighted = h->W[0].set || h->W[1].set || h->W[3].set || h->W[4].set;   // then Y_hat lives in the N_b / N_c row space
    if (h->dist.on && nrows != h->N && !weighted) return fail(h, RSVDT_ERR_ARG, "row-partitioned runs take Y_hat with the %lld owned rows", (long long)h->N);
    RS(upload_stack(h, h->Yhat, Y_host, nrows, (int)h->g.Nw_eff, (int)h->g.k, h->g.k, 1, srows, roff));
    h->yhat_rows = nrows;
    return RSVDT_OK;
}

int rsvdt_get_yhat(rsvdt_handle h, double *Y_host, int64_t nrows)
{
    if (h && h->group) return group_get_matrix(h, 0, Y_host, nrows);
    RS(bind(h));
    if (!Y_host || nrows != h->yhat_rows || !h->Yhat.p) return fail(h, RSVDT_ERR_STATE, "Y_hat not available with %lld rows", (long long)nrows);
    int64_t srows, roff;
    stack_geom(h, nrows, &srows, &roff);
    return download_stack(h, h->Yhat, Y_host, nrows, (int)h->g.Nw_eff, (int)h->g.k, h->g.k, 1, srows, roff);
}

int rsvdt_direct_action_rk4(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_direct_action_rk4);
    RS(bind(h));
    return action_with_weights(h, true);
}
int rsvdt_adjoint_action_rk4(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_adjoint_action_rk4);
    RS(bind(h));
    return action_with_weights(h, false);
}

int rsvdt_ts_action_rk4(rsvdt_handle h, int direct, const double *F_hat_host, double *Y_hat_host)
{
    RS(bind(h));
    RS(check_ready(h));
    if (!F_hat_host || !Y_hat_host) return fail(h, RSVDT_ERR_ARG, "null host buffer");
    RS(rsvdt_set_yhat(h, F_hat_host, h->N));
    RS(ts_action(h, direct != 0));
    return rsvdt_get_yhat(h, Y_hat_host, h->N);
}

int rsvdt_qr_all_freqs(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_qr_all_freqs);
    RS(bind(h));
    RS(check_ready(h));
    if (!h->Yhat.p || h->yhat_rows <= 0) return fail(h, RSVDT_ERR_STATE, "Y_hat not set");
    const rsvdt_grid &g = h->g;
    std::vector<int> st((size_t)g.Nw_eff);
    int64_t srows, roff;
    stack_geom(h, h->yhat_rows, &srows, &roff);
    // CholeskyQR2 (QRAllFreqs.c:22-32 semantics: thin Q with diag(R) > 0) in five passes over Y_hat instead of six, all on FP64
    // tensor cores: G1 = Y^H Y (read); Y <- Y M1 and, in the same pass, G2 = Gram of the rounded product (read + write);
    // Y <- Y M2 (read + write).  (Folding M1 M2 into ONE apply would save another pass but loses orthogonality to
    // cond(Y) * eps -- measured 8e-13 at cond 1e4 -- so the second factor is applied to the orthonormalised block, as
    // CholeskyQR2 prescribes.)  Any breakdown falls through to the general loop below, on the untouched or once-scaled Y.
    if (h->qr_fused && g.k <= 16 && g.k % 2 == 0) {
        auto chol_ok = [&](bool *ok) -> int {
            CU(launch_chol_inverse_big(h->small.as<cplx>(), (int)g.k, (int)g.Nw_eff, 0.0, h->smallM.as<cplx>(), h->status.as<int>(), h->stream));
            h->launches += 1;
            CU(cudaMemcpyAsync(st.data(), h->status.p, st.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            *ok = true;
            for (int64_t w = 0; w < g.Nw_eff; w++) *ok = *ok && st[(size_t)w] == 0;
            return RSVDT_OK;
        };
        bool ok = false;
        RS(gram_all(h, h->Yhat.as<cplx>(), h->yhat_rows, srows, roff));
        RS(chol_ok(&ok));
        if (ok) {
            cplx *Yo = h->Yhat.as<cplx>() + roff * g.k;
            RS(ensure(h, h->gram_partial, gram_dmma_partial_doubles(h->yhat_rows, (int)g.k, (int)g.Nw_eff) * sizeof(double)));
            CU(launch_apply_dmma(Yo, h->yhat_rows, (int)g.k, (int)g.Nw_eff, srows * g.k, h->smallM.as<cplx>(), h->gram_partial.as<double>(),
                                 h->small.as<cplx>(), h->stream));
            h->launches += 2;
            RS(dist_allreduce(h, h->small.p, 2 * g.Nw_eff * g.k * g.k));
            RS(chol_ok(&ok));
            if (!ok) return fail(h, RSVDT_ERR_NUMERIC, "QR breakdown in the second CholeskyQR pass: numerically rank-deficient block");
            CU(launch_apply_dmma(Yo, h->yhat_rows, (int)g.k, (int)g.Nw_eff, srows * g.k, h->smallM.as<cplx>(), nullptr, nullptr, h->stream));
            h->launches += 1;
            CU(cudaStreamSynchronize(h->stream));
            return RSVDT_OK;
        }
    }
    int npass = 2;
    bool shifted = false;
    for (int pass = 0; pass < npass; pass++) {
        double shift = 0.0;
        for (int attempt = 0;; attempt++) {
            RS(gram_all(h, h->Yhat.as<cplx>(), h->yhat_rows, srows, roff));
            CU(launch_chol_inverse_big(h->small.as<cplx>(), (int)g.k, (int)g.Nw_eff, shift, h->smallM.as<cplx>(), h->status.as<int>(), h->stream));
            h->launches += 1;
            CU(cudaMemcpyAsync(st.data(), h->status.p, st.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            int64_t badw = -1;
            for (int64_t w = 0; w < g.Nw_eff && badw < 0; w++)
                if (st[(size_t)w] != 0) badw = w;
            if (badw < 0) break;
            if (attempt > 0 || shifted)
                return fail(h, RSVDT_ERR_NUMERIC, "QR breakdown at frequency %lld (column %d, pass %d): numerically rank-deficient block",
                            (long long)badw, st[(size_t)badw] - 1, pass);
            shifted = true;
            npass = pass + 3;
            shift = 11.0 * ((double)h->N_global * g.k + (double)g.k * (g.k + 1)) * 1.1e-16;
        }
        CU(launch_apply_right(h->Yhat.as<cplx>(), srows, (int)g.k, (int)g.Nw_eff, h->smallM.as<cplx>(), h->stream));
        h->launches += 1;
    }
    CU(cudaStreamSynchronize(h->stream));
    return RSVDT_OK;
}

int rsvdt_power_iteration_rk4(rsvdt_handle h, int64_t q)
{
    if (h && h->group) {
        rsvdt_group_ctx *G = h->group;
        return group_run(h, [&](int p) -> int { return rsvdt_power_iteration_rk4(G->m[(size_t)p], q); });
    }
    RS(bind(h));
    for (int64_t iq = 0; iq < q; iq++) {  // PowerIterationRK4.c:20-29
        RS(rsvdt_qr_all_freqs(h));
        RS(action_with_weights(h, false));
        RS(rsvdt_qr_all_freqs(h));
        RS(action_with_weights(h, true));
    }
    return RSVDT_OK;
}

int rsvdt_svd_all_freqs_before_adjoint(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_svd_all_freqs_before_adjoint);
    RS(bind(h));
    RS(check_ready(h));
    if (!h->Yhat.p || h->yhat_rows <= 0) return fail(h, RSVDT_ERR_STATE, "Y_hat not set");
    return svd_all(h, h->Yhat, h->yhat_rows, nullptr);
}

int rsvdt_store_u(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_store_u);
    RS(bind(h));
    if (!h->Yhat.p || h->yhat_rows <= 0) return fail(h, RSVDT_ERR_STATE, "Y_hat not set");
    int64_t srows, roff;
    stack_geom(h, h->yhat_rows, &srows, &roff);
    const size_t bytes = stack_bytes(h->g.Nw_eff, srows, h->g.k);
    RS(ensure(h, h->Uhat, bytes));
    CU(cudaMemcpyAsync(h->Uhat.p, h->Yhat.p, bytes, cudaMemcpyDeviceToDevice, h->stream));  // StoreU.c:15
    CU(cudaStreamSynchronize(h->stream));
    h->uhat_rows = h->yhat_rows;
    return RSVDT_OK;
}

int rsvdt_svd_all_freqs(rsvdt_handle h)
{
    if (h && h->group) return group_call(h, rsvdt_svd_all_freqs);
    RS(bind(h));
    RS(check_ready(h));
    if (!h->Yhat.p || h->yhat_rows <= 0) return fail(h, RSVDT_ERR_STATE, "Y_hat not set");
    const rsvdt_grid &g = h->g;
    RS(ensure(h, h->Shat, (size_t)g.Nw_eff * g.k * sizeof(double)));
    RS(svd_all(h, h->Yhat, h->yhat_rows, h->Shat.as<double>()));
    h->Vhat.release();
    h->Vhat.swap(h->Yhat);                                    // SVDAllFreqs.c:58 MatDestroy(Y_hat)
    h->vhat_rows = h->yhat_rows;
    h->yhat_rows = 0;
    // SVDAllFreqs.c:63-64 optional W^{-1/2} scalings
    if (h->W[2].set && h->Uhat.p) RS(apply_sparse_to(h, h->W[2], h->Uhat, h->uhat_rows));
    if (h->W[0].set) RS(apply_sparse_to(h, h->W[0], h->Vhat, h->vhat_rows));
    return RSVDT_OK;
}

int rsvdt_get_u_hat(rsvdt_handle h, double *U_host)
{
    if (h && h->group) return group_get_matrix(h, 1, U_host, -1);
    RS(bind(h));
    if (!U_host || !h->Uhat.p) return fail(h, RSVDT_ERR_STATE, "U_hat not available");
    int64_t srows, roff;
    stack_geom(h, h->uhat_rows, &srows, &roff);
    return download_stack(h, h->Uhat, U_host, h->uhat_rows, (int)h->g.Nw_eff, (int)h->g.k, h->g.k, 1, srows, roff);
}

int rsvdt_get_v_hat(rsvdt_handle h, double *V_host)
{
    if (h && h->group) return group_get_matrix(h, 2, V_host, -1);
    RS(bind(h));
    if (!V_host || !h->Vhat.p) return fail(h, RSVDT_ERR_STATE, "V_hat not available");
    int64_t srows, roff;
    stack_geom(h, h->vhat_rows, &srows, &roff);
    return download_stack(h, h->Vhat, V_host, h->vhat_rows, (int)h->g.Nw_eff, (int)h->g.k, h->g.k, 1, srows, roff);
}

int rsvdt_get_s_hat(rsvdt_handle h, double *S_host)
{
    if (h && h->group) return rsvdt_get_s_hat(h->group->m[0], S_host);
    RS(bind(h));
    if (!S_host || !h->Shat.p) return fail(h, RSVDT_ERR_STATE, "S_hat not available");
    // device layout [iw][m] == column-major k x Nw_eff
    CU(cudaMemcpyAsync(S_host, h->Shat.p, (size_t)h->g.Nw_eff * h->g.k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return RSVDT_OK;
}

int rsvdt_ts_exp_a_deltat(rsvdt_handle h, int direct, double *U_host, int64_t ncols)
{
    RS(bind(h));
    RS(check_ready(h));
    if (!U_host || ncols <= 0) return fail(h, RSVDT_ERR_ARG, "bad arguments");
    // columns batched: block vector [row][col] ; host column-major N x ncols = stack with na=1, nb=ncols, cb=1
    DevBuf blk;
    RS(upload_stack(h, blk, U_host, h->N, 1, (int)ncols, 0, 1, h->n_ext, h->own0));
    int rc = dist_exchange(h, blk.as<cplx>(), (int)ncols);
    if (rc == RSVDT_OK) rc = unforced_steps(h, direct != 0, (int)ncols, blk.as<cplx>(), h->g.ResRatio);   // TSExpADeltaT.c:25-34
    if (rc == RSVDT_OK) rc = download_stack(h, blk, U_host, h->N, 1, (int)ncols, 0, 1, h->n_ext, h->own0);
    blk.release();
    return rc;
}

int rsvdt_transient_run_rk4(rsvdt_handle h, int direct, double *q0_host, int64_t rend, int64_t TransSaveMod,
                            double TransDivVal, double TransConVal, double *norms_out, int64_t *steps_done, int *reason,
                            rsvdt_snapshot_cb cb, void *user)
{
    return rsvdt_transient_run_rk4_est(h, direct, q0_host, rend, TransSaveMod, TransDivVal, TransConVal, norms_out, steps_done, reason,
                                       cb, user, 0, nullptr, nullptr);
}

int rsvdt_transient_run_rk4_est(rsvdt_handle h, int direct, double *q0_host, int64_t rend, int64_t TransSaveMod,
                                double TransDivVal, double TransConVal, double *norms_out, int64_t *steps_done, int *reason,
                                rsvdt_snapshot_cb cb, void *user, int64_t est_mod, rsvdt_snapshot_cb est_cb, void *est_user)
{
    if (!est_cb || est_mod <= 0) { est_cb = nullptr; est_mod = 0; }
    if (h && h->group)
        return group_transient_run(h, direct, q0_host, rend, TransSaveMod, TransDivVal, TransConVal, norms_out, steps_done, reason, cb,
                                   user, est_mod, est_cb, est_user);
    RS(bind(h));
    if (!h->A.set) return fail(h, RSVDT_ERR_STATE, "operator not set");
    if (!h->grid_set) return fail(h, RSVDT_ERR_STATE, "grid not set (dt)");
    if (!q0_host || rend < 0 || TransSaveMod <= 0 || !norms_out) return fail(h, RSVDT_ERR_ARG, "bad arguments");
    const int64_t N = h->N;
    DevBuf q;
    CU(q.alloc((size_t)h->n_ext * sizeof(cplx)));
    CU(cudaMemsetAsync(q.p, 0, (size_t)h->n_ext * sizeof(cplx), h->stream));
    RS(ensure(h, h->scalars, 512 * sizeof(double)));
    double *partial = h->scalars.as<double>();
    double *d_norm = partial + 400;
    cplx *q_own = q.as<cplx>() + h->own0;
    CU(cudaMemcpyAsync(q_own, q0_host, (size_t)N * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
    RS(dist_exchange(h, q.as<cplx>(), 1));
    double n2 = 0;
    CU(launch_norm2(q_own, N, partial, d_norm, h->stream));
    RS(dist_allreduce(h, d_norm, 1));
    CU(cudaMemcpyAsync(&n2, d_norm, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    norms_out[0] = std::sqrt(n2);                              // TransientRunRK4.c:106-108
    int why = 0;
    int64_t i = 1;
    std::vector<double> snap_host;
    for (; i <= rend;) {
        int64_t next_stop = std::min(rend, ((i - 1) / TransSaveMod + 1) * TransSaveMod);
        if (est_mod > 0) next_stop = std::min(next_stop, ((i - 1) / est_mod + 1) * est_mod);
        int rc = unforced_steps(h, direct != 0, 1, q.as<cplx>(), next_stop - i + 1);   // :128-130
        if (rc != RSVDT_OK) { q.release(); return rc; }
        i = next_stop + 1;
        if (next_stop % TransSaveMod == 0) {                   // :137
            const int64_t pos = next_stop / TransSaveMod;
            CU(launch_norm2(q_own, N, partial, d_norm, h->stream));
            h->launches += 2;
            RS(dist_allreduce(h, d_norm, 1));
            CU(cudaMemcpyAsync(&n2, d_norm, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            const double norm = std::sqrt(n2);
            if (std::isnan(norm)) { why = 1; i = next_stop; break; }         // :143-146
            if (norm > TransDivVal) { why = 2; i = next_stop; break; }       // :147-151
            if (norm < TransConVal) { why = 3; i = next_stop; break; }       // :152-156
            if (cb) {                                                        // :160-164
                snap_host.resize((size_t)2 * N);
                CU(cudaMemcpyAsync(snap_host.data(), q_own, (size_t)N * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                int crc = cb(pos, snap_host.data(), N, user);
                if (crc != 0) { q.release(); return fail(h, RSVDT_ERR_USER, "snapshot callback failed (%d)", crc); }
            }
            norms_out[pos] = norm;                                           // :166
        }
        if (est_mod > 0 && next_stop % est_mod == 0) {                       // :174-184 (after the stop tests of this step)
            snap_host.resize((size_t)2 * N);
            CU(cudaMemcpyAsync(snap_host.data(), q_own, (size_t)N * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            int crc = est_cb(next_stop / est_mod, snap_host.data(), N, est_user);
            if (crc != 0) { q.release(); return fail(h, RSVDT_ERR_USER, "transient-removal estimate callback failed (%d)", crc); }
        }
    }
    if (steps_done) *steps_done = why ? i : rend;
    if (reason) *reason = why;
    CU(cudaMemcpyAsync(q0_host, q_own, (size_t)N * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    q.release();
    return RSVDT_OK;
}

// ---- row partition across GPUs ---------------------------------------------------------------------------------------
int rsvdt_plan_partition(int64_t N, int nranks, int rank, int64_t *row_begin, int64_t *row_end)
{
    if (N < 0 || nranks <= 0 || rank < 0 || rank >= nranks || !row_begin || !row_end) return RSVDT_ERR_ARG;
    // PetscSplitOwnership (PETSC_DECIDE): n = N/P + (rank < N%P)
    const int64_t q = N / nranks, r = N % nranks;
    *row_begin = rank * q + std::min<int64_t>(rank, r);
    *row_end = *row_begin + q + (rank < r ? 1 : 0);
    return RSVDT_OK;
}

int rsvdt_dist_init(rsvdt_handle h, int rank, int nranks, int layout)
{
    RS(bind(h));
    if (nranks <= 0 || rank < 0 || rank >= nranks || layout < 0 || layout > 1) return fail(h, RSVDT_ERR_ARG, "bad partition arguments");
    h->dist = rsvdt_context::Dist();
    h->dist.on = true;
    h->dist.rank = rank;
    h->dist.nranks = nranks;
    h->dist.layout = layout;
    return RSVDT_OK;
}

int rsvdt_set_operator_csr_local(rsvdt_handle h, int64_t N_global, int64_t n_own, int64_t own0, int64_t n_ext, int64_t n_rows,
                                 const int64_t *rowptr, const int64_t *col, const double *val, const int64_t *rowptr_H,
                                 const int64_t *col_H, const double *val_H, int disc_flg, double beta, int real_operator)
{
    RS(bind(h));
    if (!h->dist.on) return fail(h, RSVDT_ERR_STATE, "call rsvdt_dist_init first");
    if (n_own <= 0 || own0 < 0 || own0 + n_own > n_ext || !rowptr || !col || !val || !rowptr_H || !col_H || !val_H)
        return fail(h, RSVDT_ERR_ARG, "bad local operator arguments");
    if (h->dist.layout == 0 && (own0 != 0 || n_rows != n_own)) return fail(h, RSVDT_ERR_ARG, "layout 0 stores exactly the owned rows");
    if (h->dist.layout == 1 && n_rows != n_ext) return fail(h, RSVDT_ERR_ARG, "layout 1 stores the halo-extended rows");
    if (disc_flg && beta < 0) return fail(h, RSVDT_ERR_USER, "'beta' must be positive, current value: %g", beta);
    h->N = n_own;
    h->n_ext = n_ext;
    h->own0 = own0;
    h->N_global = N_global;
    h->real_operator = real_operator != 0;
    // A - beta*I: the local diagonal of stored row r is local column r (layout 1) or r (layout 0, owned rows first)
    const double shift = disc_flg ? beta : 0.0;
    RS(upload_csr(h, h->A, n_rows, n_ext, rowptr, col, reinterpret_cast<const cplx *>(val), false, shift));
    RS(upload_csr(h, h->AH, n_rows, n_ext, rowptr_H, col_H, reinterpret_cast<const cplx *>(val_H), false, shift));
    return RSVDT_OK;
}

int rsvdt_set_operator_stencil(rsvdt_handle h, int ndim, const int64_t *shape, int nvar, const double *blocks_re_im,
                               const double *mod_tab, const double *diag_tab_re_im, int64_t pt0, int64_t pt1,
                               const int64_t *ghost_pts, int64_t n_ghost, int with_adjoint, int disc_flg, double beta,
                               int real_operator)
{
    RS(bind(h));
    if (ndim < 1 || ndim > 3 || !shape || !(nvar == 3 || nvar == 4 || nvar == 5) || !blocks_re_im || !mod_tab || !diag_tab_re_im)
        return fail(h, RSVDT_ERR_ARG, "bad stencil operator arguments (1-3 grid dimensions, 3-5 variables per point)");
    int64_t npts = 1;
    int nb = 1;
    for (int i = 0; i < ndim; i++) { if (shape[i] <= 0) return fail(h, RSVDT_ERR_ARG, "bad grid shape"); npts *= shape[i]; nb *= 3; }
    if (pt0 < 0 || pt1 <= pt0 || pt1 > npts || n_ghost < 0 || (n_ghost > 0 && !ghost_pts)) return fail(h, RSVDT_ERR_ARG, "bad point range");
    if ((npts + n_ghost) * nvar > (int64_t)INT32_MAX) return fail(h, RSVDT_ERR_ARG, "more than 2^31-1 local columns not supported");
    if (disc_flg && beta < 0) return fail(h, RSVDT_ERR_USER, "'beta' must be positive, current value: %g", beta);
    if (h->dist.on) {
        if (h->dist.layout != 0) return fail(h, RSVDT_ERR_ARG, "stencil operators are partitioned in layout 0");
    } else {
        if (pt0 != 0 || pt1 != npts || n_ghost != 0) return fail(h, RSVDT_ERR_ARG, "single-GPU run: the whole grid, no ghosts");
        h->dist = rsvdt_context::Dist();
    }
    h->N = (pt1 - pt0) * nvar;
    h->n_ext = h->N + n_ghost * nvar;
    h->own0 = 0;
    h->N_global = npts * nvar;
    h->real_operator = real_operator != 0;
    DevBuf dblocks, dmod, ddiag, dghost;
    CU(dblocks.alloc((size_t)nb * nvar * nvar * sizeof(cplx)));
    CU(dmod.alloc(997 * sizeof(double)));
    CU(ddiag.alloc(499 * sizeof(cplx)));
    CU(dghost.alloc(std::max<size_t>((size_t)n_ghost, 1) * sizeof(int64_t)));
    CU(cudaMemcpy(dblocks.p, blocks_re_im, (size_t)nb * nvar * nvar * sizeof(cplx), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dmod.p, mod_tab, 997 * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ddiag.p, diag_tab_re_im, 499 * sizeof(cplx), cudaMemcpyHostToDevice));
    if (n_ghost) CU(cudaMemcpy(dghost.p, ghost_pts, (size_t)n_ghost * sizeof(int64_t), cudaMemcpyHostToDevice));
    StencilSpec sp;
    sp.d = ndim;
    for (int i = 0; i < ndim; i++) sp.shape[i] = shape[i];
    sp.B = nvar;
    sp.pt0 = pt0;
    sp.pt1 = pt1;
    sp.blocks = dblocks.as<cplx>();
    sp.mod_tab = dmod.as<double>();
    sp.diag_tab = ddiag.as<cplx>();
    sp.ghost_pts = dghost.as<int64_t>();
    sp.n_ghost = n_ghost;
    sp.shift = disc_flg ? beta : 0.0;
    sp.hermitian = 0;
    RS(build_stencil_operator(h, h->A, sp, h->n_ext));
    h->AH.release();
    if (with_adjoint) {
        sp.hermitian = 1;
        RS(build_stencil_operator(h, h->AH, sp, h->n_ext));
    }
    return RSVDT_OK;
}

int rsvdt_create_random_mat(rsvdt_handle h, int64_t nrows, int64_t global_row0, uint64_t seed)
{
    RS(bind(h));
    if (!h->grid_set) return fail(h, RSVDT_ERR_STATE, "grid not set");
    if (nrows <= 0 || global_row0 < 0) return fail(h, RSVDT_ERR_ARG, "bad Y_hat arguments");
    int64_t srows, roff;
    stack_geom(h, nrows, &srows, &roff);
    const size_t bytes = stack_bytes(h->g.Nw_eff, srows, h->g.k);
    RS(ensure(h, h->Yhat, bytes));
    if (srows != nrows) CU(cudaMemsetAsync(h->Yhat.p, 0, bytes, h->stream));
    CU(launch_random_stack(h->Yhat.as<cplx>(), nrows, srows, roff, (int)h->g.Nw_eff, (int)h->g.k, global_row0, (unsigned long long)seed, h->stream));
    h->launches++;
    CU(cudaStreamSynchronize(h->stream));
    h->yhat_rows = nrows;
    return RSVDT_OK;
}

int rsvdt_get_state(rsvdt_handle h, double *y_host, int64_t nrows)
{
    RS(bind(h));
    if (!y_host || nrows != h->N || !h->grid_set || h->y.bytes < stack_bytes(1, h->n_ext, h->g.k))
        return fail(h, RSVDT_ERR_STATE, "no RK4 state with %lld rows on the device", (long long)nrows);
    return download_stack(h, h->y, y_host, nrows, 1, (int)h->g.k, 0, 1, h->n_ext, h->own0);
}

int rsvdt_dist_set_messages(rsvdt_handle h, int nout, const int *out_peer, const int *out_count, const int64_t *out_rows,
                            const int64_t *out_dest_row0, int nin, const int *in_peer, const int *in_count,
                            const int64_t *in_dest_row0)
{
    RS(bind(h));
    auto &D = h->dist;
    if (!D.on) return fail(h, RSVDT_ERR_STATE, "call rsvdt_dist_init first");
    if (!h->grid_set) return fail(h, RSVDT_ERR_STATE, "set the grid before the halo messages (the mailbox is sized for k)");
    D.out.clear();
    D.in.clear();
    D.connected = false;
    D.nvec_max = (int)std::max<int64_t>(h->g.k, 1);
    int64_t off = 0;
    D.max_out = D.max_in = 0;
    for (int i = 0; i < nout; i++) {
        rsvdt_context::DistMsg m;
        m.peer = out_peer[i];
        m.count = out_count[i];
        m.dest_row0 = out_dest_row0[i];
        m.mb_off_rows = -1;
        std::vector<int32_t> rows((size_t)m.count);
        for (int j = 0; j < m.count; j++) {
            const int64_t r = out_rows[off + j];
            if (r < 0 || r >= h->n_ext) return fail(h, RSVDT_ERR_ARG, "halo send row out of range");
            rows[(size_t)j] = (int32_t)r;
        }
        off += m.count;
        CU(m.rows.alloc(std::max<size_t>(rows.size(), 1) * sizeof(int32_t)));
        if (m.count) CU(cudaMemcpy(m.rows.p, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        m.rows_h = rows;
        D.max_out = std::max(D.max_out, m.count);
        D.out.push_back(std::move(m));
    }
    D.in_rows_total = 0;
    for (int i = 0; i < nin; i++) {
        rsvdt_context::DistIn m;
        m.peer = in_peer[i];
        m.count = in_count[i];
        m.dest_row0 = in_dest_row0[i];
        m.mb_off_rows = D.in_rows_total;
        if (m.dest_row0 < 0 || m.dest_row0 + m.count > h->n_ext) return fail(h, RSVDT_ERR_ARG, "ghost rows out of range");
        D.in_rows_total += m.count;
        D.max_in = std::max(D.max_in, m.count);
        D.in.push_back(m);
    }
    // interior / boundary split (layout 0): the rows sent to peers sit at the two ends of the owned range for slab partitions
    D.overlap = false;
    if (D.layout == 0 && nout > 0) {
        int64_t lo_end = 0, hi_begin = h->N, o2 = 0;
        for (int i = 0; i < nout; o2 += out_count[i], i++)
            for (int j = 0; j < out_count[i]; j++) {
                const int64_t r = out_rows[o2 + j];
                if (r < h->N / 2) lo_end = std::max(lo_end, r + 1); else hi_begin = std::min(hi_begin, r);
            }
        lo_end = std::min(h->N, (lo_end + 59) / 60 * 60);   // whole block rows for every supported block size (3, 4, 5)
        hi_begin = hi_begin / 60 * 60;
        if (hi_begin > lo_end && 2 * (lo_end + h->N - hi_begin) <= h->N) {
            D.overlap = true;
            D.lo_end = lo_end;
            D.hi_begin = hi_begin;
        }
    }
    // mailbox: two epoch-parity slots per incoming message (unfused exchange) + three slots in ghost order (exchange fused
    // into the stage kernel); flags: [0, P) epoch of the unfused exchange per peer, [P, 2P) epoch of the fused one,
    // [2P] rows per mailbox slot (read by the peers at import), [2P+1] set by a kernel whose wait for a peer gave up
    CU(D.mailbox.alloc(std::max<size_t>((size_t)5 * D.in_rows_total * D.nvec_max * sizeof(cplx), 256)));
    CU(D.flags.alloc((size_t)(2 * D.nranks + 2) * sizeof(unsigned long long)));
    CU(cudaMemset(D.flags.p, 0, (size_t)(2 * D.nranks + 2) * sizeof(unsigned long long)));
    {
        const unsigned long long rows_per_slot = (unsigned long long)D.in_rows_total;
        CU(cudaMemcpy(D.flags.as<unsigned long long>() + 2 * D.nranks, &rows_per_slot, sizeof rows_per_slot, cudaMemcpyHostToDevice));
    }
    D.epoch = 0;
    D.fz_epoch = 0;
    D.fz_ready = false;
    D.fz_last_out = nullptr;
    return RSVDT_OK;
}

int rsvdt_dist_get_in_offsets(rsvdt_handle h, int64_t *in_off)
{
    if (!h || !in_off) return RSVDT_ERR_ARG;
    for (size_t i = 0; i < h->dist.in.size(); i++) in_off[i] = h->dist.in[i].mb_off_rows;
    return RSVDT_OK;
}

int rsvdt_dist_ipc_export(rsvdt_handle h, unsigned char *bytes128)
{
    RS(bind(h));
    auto &D = h->dist;
    if (!D.on || !D.mailbox.p || !D.flags.p) return fail(h, RSVDT_ERR_STATE, "halo messages not set");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    cudaIpcMemHandle_t hm, hf;
    CU(cudaIpcGetMemHandle(&hm, D.mailbox.p));
    CU(cudaIpcGetMemHandle(&hf, D.flags.p));
    memcpy(bytes128, &hm, 64);
    memcpy(bytes128 + 64, &hf, 64);
    return RSVDT_OK;
}

int rsvdt_dist_ipc_import(rsvdt_handle h, const unsigned char *all_bytes, const int64_t *out_mb_off)
{
    RS(bind(h));
    auto &D = h->dist;
    if (!D.on || !D.mailbox.p) return fail(h, RSVDT_ERR_STATE, "halo messages not set");
    D.peer_mailbox.assign((size_t)D.nranks, nullptr);
    D.peer_flags.assign((size_t)D.nranks, nullptr);
    for (int p = 0; p < D.nranks; p++) {
        if (p == D.rank) { D.peer_mailbox[(size_t)p] = D.mailbox.p; D.peer_flags[(size_t)p] = D.flags.p; continue; }
        bool needed = false;
        for (auto &m : D.out) needed |= (m.peer == p);
        if (!needed) continue;
        cudaIpcMemHandle_t hm, hf;
        memcpy(&hm, all_bytes + (size_t)p * 128, 64);
        memcpy(&hf, all_bytes + (size_t)p * 128 + 64, 64);
        CU(cudaIpcOpenMemHandle(&D.peer_mailbox[(size_t)p], hm, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&D.peer_flags[(size_t)p], hf, cudaIpcMemLazyEnablePeerAccess));
    }
    return finish_connect(h, out_mb_off);
}

}  // extern "C"

namespace {
// second half of the bootstrap, common to CUDA IPC (one process per GPU) and to direct peer pointers (one process, P GPUs):
// device-side message tables, the peers' mailbox geometry, the tables of the fused exchange
int finish_connect(rsvdt_handle h, const int64_t *out_mb_off)
{
    auto &D = h->dist;
    std::vector<HaloMsgHost> mh(D.out.size());
    std::vector<unsigned long long *> pf(D.out.size());
    for (size_t i = 0; i < D.out.size(); i++) {
        auto &m = D.out[i];
        m.mb_off_rows = out_mb_off[i];
        mh[i].peer_mailbox = reinterpret_cast<cplx *>(D.peer_mailbox[(size_t)m.peer]) + (int64_t)2 * m.mb_off_rows * D.nvec_max;
        mh[i].rows = m.rows.as<int32_t>();
        mh[i].count = m.count;
        mh[i].pad = 0;
        pf[i] = reinterpret_cast<unsigned long long *>(D.peer_flags[(size_t)m.peer]) + D.rank;
    }
    std::vector<HaloInHost> ih(D.in.size());
    for (size_t i = 0; i < D.in.size(); i++) {
        auto &m = D.in[i];
        ih[i].mailbox = D.mailbox.as<cplx>() + (int64_t)2 * m.mb_off_rows * D.nvec_max;
        ih[i].flag = D.flags.as<unsigned long long>() + m.peer;
        ih[i].dest_row0 = m.dest_row0;
        ih[i].count = m.count;
        ih[i].pad = 0;
    }
    CU(D.msgs_dev.alloc(std::max<size_t>(mh.size(), 1) * sizeof(HaloMsgHost)));
    CU(D.peer_flags_dev.alloc(std::max<size_t>(pf.size(), 1) * sizeof(void *)));
    CU(D.ins_dev.alloc(std::max<size_t>(ih.size(), 1) * sizeof(HaloInHost)));
    if (!mh.empty()) {
        CU(cudaMemcpy(D.msgs_dev.p, mh.data(), mh.size() * sizeof(HaloMsgHost), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(D.peer_flags_dev.p, pf.data(), pf.size() * sizeof(void *), cudaMemcpyHostToDevice));
    }
    if (!ih.empty()) CU(cudaMemcpy(D.ins_dev.p, ih.data(), ih.size() * sizeof(HaloInHost), cudaMemcpyHostToDevice));
    D.peer_in_rows.assign((size_t)D.nranks, 0);
    for (auto &m : D.out) {
        unsigned long long v = 0;
        CU(cudaMemcpy(&v, reinterpret_cast<unsigned long long *>(D.peer_flags[(size_t)m.peer]) + 2 * D.nranks, sizeof v, cudaMemcpyDeviceToHost));
        D.peer_in_rows[(size_t)m.peer] = (int64_t)v;
    }
    D.connected = true;
    return build_fused(h);
}
}  // namespace

extern "C" {

int rsvdt_set_allreduce(rsvdt_handle h, rsvdt_allreduce_fn fn, void *user)
{
    if (!h) return RSVDT_ERR_ARG;
    h->dist.allreduce = fn;
    h->dist.allreduce_user = user;
    return RSVDT_OK;
}

void *rsvdt_cuda_stream(rsvdt_handle h) { return h ? (void *)h->stream : nullptr; }

int rsvdt_keep_snapshots(rsvdt_handle h, int keep)
{
    if (!h) return RSVDT_ERR_ARG;
    h->keep_snapshots = keep != 0;
    return RSVDT_OK;
}

int rsvdt_get_snapshots(rsvdt_handle h, double *Yall_host)
{
    RS(bind(h));
    if (!Yall_host || !h->snaps.p || h->snaps_nstore <= 0) return fail(h, RSVDT_ERR_STATE, "snapshots not kept (rsvdt_keep_snapshots)");
    // device [pos][row][ik] -> reference test-vector-major column ik*Nstore + pos  (TSActionRK4.c:126-130)
    return download_stack(h, h->snaps, Yall_host, h->N, (int)h->snaps_nstore, (int)h->g.k, 1, h->snaps_nstore, h->n_ext, h->own0);
}

int rsvdt_time_forced_steps(rsvdt_handle h, int direct, int64_t nsteps, float *ms_total)
{
    RS(bind(h));
    RS(check_ready(h));
    if (nsteps <= 0 || !ms_total) return fail(h, RSVDT_ERR_ARG, "bad arguments");
    if (!h->Yhat.p || h->yhat_rows != h->N) return fail(h, RSVDT_ERR_STATE, "Y_hat (forcing) not set");
    const int64_t U = h->n_ext * h->g.k;
    RS(ensure(h, h->y, (size_t)U * sizeof(cplx)));
    CU(launch_fill_zero(h->y.as<cplx>(), U, h->stream));
    // allocate work buffers outside the timed region
    RS(forced_steps(h, direct != 0, h->Yhat.as<cplx>(), h->y.as<cplx>(), nullptr, h->g.Nw, 1, std::min<int64_t>(nsteps, 64)));
    CU(cudaStreamSynchronize(h->stream));
    CU(launch_fill_zero(h->y.as<cplx>(), U, h->stream));
    int64_t n;
    drain(h->ev_stage, &n); drain(h->ev_idft, &n); drain(h->ev_dft, &n);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    h->profile = true;
    CU(cudaEventRecord(e0, h->stream));
    int rc = forced_steps(h, direct != 0, h->Yhat.as<cplx>(), h->y.as<cplx>(), nullptr, h->g.Nw, 1, nsteps);
    cudaEventRecord(e1, h->stream);
    h->profile = false;
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->last_ms_stage = drain(h->ev_stage, &h->last_n_stage);
    h->last_ms_idft = drain(h->ev_idft, &h->last_n_idft);
    h->last_ms_dft = drain(h->ev_dft, &h->last_n_dft);
    h->last_ms_all = ms;
    if (rc != RSVDT_OK) return rc;
    CU(e);
    RS(check_halo_timeout(h));
    *ms_total = ms;
    return RSVDT_OK;
}

int rsvdt_operator_info(rsvdt_handle h, int direct, int nvec, int64_t *out4)
{
    if (!h || !out4) return RSVDT_ERR_ARG;
    const CsrOwned &Op = direct ? h->A : h->AH;
    if (!Op.set) return fail(h, RSVDT_ERR_STATE, "operator not set");
    out4[3] = h->N;
    if (tiled_active(h, Op, nvec)) {
        out4[0] = 2;
        out4[1] = (int64_t)((size_t)Op.tb.nchunks * Op.tb.chunk_bytes + (size_t)(Op.tb.ntiles + 1) * sizeof(int64_t));
        out4[2] = Op.tb.nchunks * (int64_t)Op.tb.TB * Op.tb.B * Op.tb.B;
    } else if (h->use_bcsr && Op.bcsr_valid && bcsr_stage_supported(Op.bcsr.B, nvec)) {
        const int64_t BS = (Op.bcsr.B * Op.bcsr.B + 1) & ~1;
        out4[0] = 1;
        out4[1] = Op.bcsr.nblocks * (BS * (int64_t)sizeof(cplx) + 4) + (Op.bcsr.nbrows + 1) * 8;
        out4[2] = Op.bcsr.nblocks * (int64_t)Op.bcsr.B * Op.bcsr.B;
    } else {
        out4[0] = 0;
        out4[1] = Op.d.nnz * (int64_t)(sizeof(cplx) + 4) + (Op.d.nrows + 1) * 8;
        out4[2] = Op.d.nnz;
    }
    return RSVDT_OK;
}

int64_t rsvdt_kernel_launches(rsvdt_handle h, int reset)
{
    if (!h) return -1;
    const int64_t n = h->launches;
    if (reset) h->launches = 0;
    return n;
}

int rsvdt_set_option(rsvdt_handle h, const char *name, int64_t value)
{
    if (!h || !name) return RSVDT_ERR_ARG;
    if (!strcmp(name, "chunk_steps")) { if (value < 0 || value > 64) return fail(h, RSVDT_ERR_ARG, "chunk_steps out of range"); h->chunk_steps = value; }
    else if (!strcmp(name, "stage_variant")) h->stage_variant = value;
    else if (!strcmp(name, "idft_variant")) h->idft_variant = value;
    else if (!strcmp(name, "use_banded")) h->use_banded = value;
    else if (!strcmp(name, "tma_variant")) h->tma_variant = value;
    else if (!strcmp(name, "use_bcsr")) h->use_bcsr = value;
    else if (!strcmp(name, "gram_variant")) h->gram_variant = value;
    else if (!strcmp(name, "qr_fused")) h->qr_fused = value;
    else if (!strcmp(name, "use_tiled")) h->use_tiled = value;
    else if (!strcmp(name, "tiled_stages")) h->tiled_stages = value;
    else if (!strcmp(name, "tiled_cv")) h->tiled_cv = value;
    else if (!strcmp(name, "use_small_exchange")) h->use_small_exchange = value;
    else if (!strcmp(name, "use_overlap1")) h->use_overlap1 = value;
    else if (!strcmp(name, "tiled_variant")) h->tiled_variant = value;
    else if (!strcmp(name, "tiled_flags")) h->tiled_flags = value;
    else if (!strcmp(name, "tiled_order")) h->tiled_order = value;
    else if (!strcmp(name, "tiled_ctas")) { if (value < 1 || value > 3) return fail(h, RSVDT_ERR_ARG, "tiled_ctas out of range"); h->tiled_ctas = value; }
    else if (!strcmp(name, "use_small")) h->use_small = value;
    else if (!strcmp(name, "use_graph")) h->use_graph = value;
    else if (!strcmp(name, "use_overlap")) h->use_overlap = value;
    else if (!strcmp(name, "use_fused_halo")) h->use_fused_halo = value;
    else if (!strcmp(name, "use_pipeline")) h->use_pipeline = value;
    else return fail(h, RSVDT_ERR_ARG, "unknown option '%s'", name);
    return RSVDT_OK;
}

int rsvdt_last_kernel_ms(rsvdt_handle h, const char *which, float *ms, int64_t *launches)
{
    if (!h || !which || !ms) return RSVDT_ERR_ARG;
    int64_t n = 0;
    if (!strcmp(which, "stage")) { *ms = h->last_ms_stage; n = h->last_n_stage; }
    else if (!strcmp(which, "idft")) { *ms = h->last_ms_idft; n = h->last_n_idft; }
    else if (!strcmp(which, "dft")) { *ms = h->last_ms_dft; n = h->last_n_dft; }
    else if (!strcmp(which, "all")) { *ms = h->last_ms_all; n = h->last_n_stage + h->last_n_idft + h->last_n_dft; }
    else return fail(h, RSVDT_ERR_ARG, "unknown kernel family '%s'", which);
    if (launches) *launches = n;
    return RSVDT_OK;
}

}  // extern "C"
