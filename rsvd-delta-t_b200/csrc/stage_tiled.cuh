// stage_tiled.cuh -- K1f: the RK4 stage on a TILED block-CSR operator, operator entries staged through shared memory by
// cp.async.bulk (TMA) + mbarrier ring, halo rows exchanged by the kernel itself (peer stores over NVLink).
//
// Tiled block-CSR ("TBCSR"): the block rows are cut into tiles of TB consecutive block rows.  Tile t owns the chunks
// [tile_ptr[t], tile_ptr[t+1]); chunk p of a tile holds the p-th block of every block row of the tile (block rows with
// fewer blocks are padded, column -1), laid out for one bulk copy and conflict-free shared-memory reads:
//     int32 col[TB]           block column of that block, or -1 (padding: skipped, no arithmetic)
//     int32 hdr[4]            {xoff, flags, 0, 0}: where the chunk's rows of the stage vector sit in the x window (below)
//     cplx  val[B*B][TB]      entry e = j*B + r of the block (column-major inside the block), block row tb fastest
// (columns and header first: they ride with block column 0 when the kernel brings a chunk in piece by piece)
// chunk_bytes = B*B*TB*16 + TB*4 + 16.  Per nonzero this is 16 B + ~4/(B*B) B -- the floor of a complex-double block format.
// When the blocks of a tile line up on block diagonals (col - block row constant along a chunk: stencils), the chunks are
// cut by diagonal instead of by position in the row (missing neighbours padded in place), and runs of neighbouring
// diagonals share an "x window": one contiguous range of block rows of the stage vector, staged in shared memory once.
// Summation order per row = ascending block columns, ascending column inside the block: the same as the CSR and block-CSR
// stage kernels, so the three agree bit for bit.
#pragma once

#include "kernels.cuh"

#include <vector>

constexpr int TBCSR_TB = 32;          // block rows per tile
constexpr int TBCSR_MAX_PEERS = 8;    // halo peers a rank can send to / receive from in the fused exchange
constexpr int TBCSR_GATHER = -(1 << 30);   // chunk header xoff: no window, gather the rows from global memory
constexpr int TBCSR_WMAX = 40;        // longest x window in block rows (TB + 8: diagonals up to 8 blocks apart share one)

struct TbcsrWindow { int32_t start, nblocks, first_chunk, pad; };   // block rows [start, start + nblocks) of the stage vector

struct TbcsrDev {
    int B = 0, TB = TBCSR_TB;
    int64_t nbrows = 0;               // block rows computed by this rank
    int64_t nbcols_own = 0;           // block columns < nbcols_own are rows of the stage vector itself; the others are ghosts
    int64_t ntiles = 0, nchunks = 0;
    const int64_t *tile_ptr = nullptr;        // [ntiles + 1]
    const unsigned char *chunks = nullptr;    // nchunks * chunk_bytes
    const int32_t *tile_order = nullptr;      // optional [ntiles]: processing order (halo tiles first)
    size_t chunk_bytes = 0;
    const int64_t *win_ptr = nullptr;         // [ntiles + 1] windows of tile t, in the order their chunks come
    const TbcsrWindow *win = nullptr;
    int wmax = 0;                             // longest window (block rows); 0: no windows
    int64_t nwindows_chunks = 0;              // chunks that need a window (if shared memory cannot hold the x ring the launch fails)
    int flags = 35;                           // 1: operator chunks L2 evict-first, 2: x windows L2 evict-last, 32: L2 prefetch of the
                                              // tile's block vectors; measurement only: 8 no FMAs, 16 no block-vector loads / stores
};
__host__ __device__ inline size_t tbcsr_chunk_bytes(int B, int TB) { return (size_t)B * B * TB * sizeof(cplx) + (size_t)TB * sizeof(int32_t) + 16; }
__host__ __device__ inline size_t tbcsr_cols_offset() { return 0; }                                             // col[TB], then hdr[4]
__host__ __device__ inline size_t tbcsr_vals_offset(int TB) { return (size_t)TB * sizeof(int32_t) + 16; }      // val[B*B][TB]

// The halo exchange done by the stage kernel itself (row-partitioned runs, layout 0).
//   Every fused stage has an epoch E (1, 2, ...).  Tiles whose rows a peer needs store them, next to the local store,
//   straight into the peer's mailbox slot E % 3 (ghost order: mailbox row j = ghost row j of the peer) and, when the last
//   such tile of the launch is done, publish E in the peer's flag with a system-scope release.  Tiles that read ghost
//   columns (or send) first wait until every incoming peer has published E - 1 and then gather the ghost rows from
//   mailbox slot (E - 1) % 3 -- there is no unpack copy and no extra launch.  Three slots + "senders wait for E - 1"
//   make the slot reuse safe: a peer that has published E - 1 has finished the kernel that read slot (E - 3) % 3.
struct HaloFuse {
    int enabled = 0;
    unsigned long long epoch = 0;             // E of this launch
    const cplx *ghost_src = nullptr;          // base of ghost block column 0 for this launch (mailbox slot or in-buffer rows)
    int wait_epoch = 0;                       // 1: ghost_src is mailbox slot (E-1)%3 -> readers wait for E-1
    int n_in = 0, n_out = 0;
    const unsigned long long *in_flag[TBCSR_MAX_PEERS];      // my flags, one per incoming peer
    cplx *out_mail[TBCSR_MAX_PEERS];                         // peer's mailbox slot E%3 (base of the peer's ghost row 0)
    unsigned long long *out_flag[TBCSR_MAX_PEERS];           // peer's flag for me
    unsigned long long *out_done = nullptr;                  // local counters [n_out]: sending tiles finished (monotonic)
    const int32_t *out_tiles = nullptr;                      // [n_out] sending tiles per launch and peer
    const unsigned char *tile_flags = nullptr;               // [ntiles] bit0: reads ghosts, bit1: sends
    const int32_t *send_index = nullptr;                     // [ntiles] -> row of send_dest, or -1
    const int64_t *send_dest = nullptr;                      // [nsendtiles][2][TB]: (peer slot << 48) | ghost row on the peer, or -1
    const int32_t *send_mask = nullptr;                      // [nsendtiles] bit p: the tile sends to outgoing peer p
    unsigned long long *timeout_flag = nullptr;              // set to 1 if a wait gave up (peer dead): the host reports it
};

bool tbcsr_stage_supported(int B, int nv);
// mode / StageArgs as for the other stage kernels (kernels.cuh)
// ring_stages: 0 = as many operator chunks as shared memory holds (up to 12), else at most that many;
// vectors_per_thread: 0 = automatic, 1 or 2 test vectors per consumer thread;
// variant: compile-time variant bits of the kernel (stage_tiled.cu, VAR_*), honoured for the B = 5, k = 8 shape
cudaError_t launch_rk4_stage_tbcsr(const TbcsrDev &Op, int nv, int mode, const StageArgs &a, const HaloFuse &hf, int ctas_per_sm,
                                   int ring_stages, int vectors_per_thread, int variant, cudaStream_t s);

// Host-side plan of the x windows from the chunk columns (cols: [nchunks][TB] as stored): fills hdr ([nchunks][4]) and the
// window lists.  nb_own: block columns below it are rows of the stage vector, the others ghost rows (separate source).
struct TbcsrPlan {
    std::vector<int32_t> hdr;
    std::vector<int64_t> win_ptr;
    std::vector<TbcsrWindow> win;
    int64_t window_chunks = 0, gather_chunks = 0;
};
void tbcsr_plan_windows(int64_t ntiles, const int64_t *tile_ptr, const int32_t *cols, int64_t nb_own, int64_t nb_total, int TB,
                        int wmax, TbcsrPlan &out);

// Synthetic LNS block-stencil operator generated on the device, straight into the tiled layout (measurement aid for the
// BASELINE.json configurations: the 2e7-DOF operator is 65 GB as a file).  Same entries, bit for bit, as
// synthetic.lns_stencil on the host: blocks[b][v][u] * mod[pt % 997], diagonal diag_tab[pt % 499].
struct StencilSpec {
    int d = 3;                 // grid dimensions
    int64_t shape[3] = {1, 1, 1};   // slowest first
    int B = 5;
    int hermitian = 0;         // 1: generate A^H (conjugate transposed blocks of the mirrored neighbour)
    int64_t pt0 = 0, pt1 = 0;  // grid points (= block rows) of this rank
    const cplx *blocks = nullptr;     // device [3^d][B][B]
    const double *mod_tab = nullptr;  // device [997]
    const cplx *diag_tab = nullptr;   // device [499]
    // ghost numbering (layout 0): block columns outside [pt0, pt1) -> nbcols_own + position in the sorted ghost list
    const int64_t *ghost_pts = nullptr;
    int64_t n_ghost = 0;
    double shift = 0.0;        // discounting: subtracted from the diagonal (LNSStructure.c:53-56)
};
// counts_dev[t] = chunks of tile t = stencil offsets that occur among its points (masks_dev[t]: which); then, with the prefix
// sums of the counts as tile_ptr, the fill kernel writes every chunk (one per block diagonal; the chunk headers are written
// by the caller from tbcsr_plan_windows); tile_has_ghost[t] (optional, zeroed by the caller) is set when a block column of
// the tile lies outside [pt0, pt1)
cudaError_t stencil_tile_counts(const StencilSpec &sp, int TB, int64_t ntiles, int32_t *counts_dev, uint32_t *masks_dev, cudaStream_t s);
cudaError_t stencil_fill(const StencilSpec &sp, int TB, int64_t ntiles, const int64_t *tile_ptr_dev, const uint32_t *masks_dev,
                         unsigned char *chunks, unsigned char *tile_has_ghost, cudaStream_t s);
// stack[a][roff + r][b] = uniform [0,1) + i [0,1), r < rows (CreateRandomMat.c:36-39 on the device)
cudaError_t launch_random_stack(cplx *stack, int64_t rows, int64_t srows, int64_t roff, int na, int nb, int64_t grow0,
                                unsigned long long seed, cudaStream_t s);
