// stencil_gen.cu -- the synthetic LNS block-stencil operators of BASELINE.json generated on the device, straight into the
// tiled block-CSR layout the stage kernel streams (stage_tiled.cuh).  Measurement aid: the 2e7-DOF configuration is a 65 GB
// PETSc-binary file and its host-side CSR -> device conversion takes minutes; the entries written here are, bit for bit,
// those synthetic.lns_stencil produces on the host (tested), so the timed operator is the one the parity tests check.
//
//   entry (row point q, row variable v; neighbour p' = q + offset_b, column variable u):
//       A   : blocks[b][v][u] * mod[q % 997]              (b = centre, u = v:  diag_tab[q % 499])
//       A^H : conj(blocks[nb-1-b][u][v] * mod[p' % 997])  (b = centre, u = v:  conj(diag_tab[q % 499]))
//   minus `shift` on the diagonal (discounting, LNSStructure.c:53-56).
#include "stage_tiled.cuh"

namespace {

struct Geo {
    int d, nb;
    int64_t shape[3], stride[3];
};

__device__ __forceinline__ int valid_neighbours(const Geo &g, int64_t q, int64_t *nbr /*[27] or null*/, int *bidx /*[27] or null*/)
{
    int64_t c[3], rem = q;
    for (int i = g.d - 1; i >= 0; i--) { c[i] = rem % g.shape[i]; rem /= g.shape[i]; }
    int n = 0;
    for (int b = 0; b < g.nb; b++) {
        int t = b, off[3];
        for (int i = g.d - 1; i >= 0; i--) { off[i] = t % 3 - 1; t /= 3; }
        bool ok = true;
        int64_t p = q;
        for (int i = 0; i < g.d; i++) {
            const int64_t ci = c[i] + off[i];
            ok = ok && ci >= 0 && ci < g.shape[i];
            p += off[i] * g.stride[i];
        }
        if (ok) {
            if (nbr) { nbr[n] = p; bidx[n] = b; }
            n++;
        }
    }
    return n;
}

// which stencil offsets occur in a tile (bit b of mask) -- its chunks, one per block diagonal, in ascending column order
__global__ void stencil_tile_counts_kernel(Geo g, int64_t pt0, int64_t pt1, int TB, int64_t ntiles, int32_t *counts, uint32_t *masks)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint32_t m = 0;
    for (int tb = 0; tb < TB; tb++) {
        const int64_t q = pt0 + t * TB + tb;
        if (q >= pt1) break;
        int64_t nbr[27];
        int bidx[27];
        const int n = valid_neighbours(g, q, nbr, bidx);
        for (int i = 0; i < n; i++) m |= 1u << bidx[i];
    }
    counts[t] = __popc(m);
    masks[t] = m;
}

// one thread per (tile, block row of the tile): writes its column of every chunk of the tile.  Chunk of stencil offset b =
// the (number of offsets < b present in the tile)-th chunk; a point without that neighbour gets column -1 there.
template <int B>
__global__ void stencil_fill_kernel(Geo g, StencilSpec sp, int TB, int64_t ntiles, const int64_t *tile_ptr, const uint32_t *masks,
                                    unsigned char *chunks, unsigned char *tile_has_ghost)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = gid / TB;
    const int tb = (int)(gid - t * TB);
    if (t >= ntiles) return;
    const size_t cb = tbcsr_chunk_bytes(B, TB);
    const uint32_t mask = masks[t];
    const int64_t q = sp.pt0 + t * TB + tb;
    int64_t nbr[27];
    int bidx[27];
    const int n = q < sp.pt1 ? valid_neighbours(g, q, nbr, bidx) : 0;
    const int centre = (g.nb - 1) / 2;
    const int64_t nown = sp.pt1 - sp.pt0;
    bool ghost = false;
    int i = 0;   // next valid neighbour of this point
    int p = 0;   // chunk index inside the tile
    for (int b = 0; b < g.nb; b++) {
        if (!(mask & (1u << b))) continue;
        unsigned char *base = chunks + (size_t)(tile_ptr[t] + p) * cb;
        p++;
        cplx *cv = reinterpret_cast<cplx *>(base + tbcsr_vals_offset(TB));
        int32_t *cc = reinterpret_cast<int32_t *>(base + tbcsr_cols_offset());
        if (i >= n || bidx[i] != b) {
            cc[tb] = -1;
            for (int e = 0; e < B * B; e++) cv[(size_t)e * TB + tb] = make_double2(0.0, 0.0);
            continue;
        }
        const int64_t pn = nbr[i];
        i++;
        int64_t lc;
        if (pn >= sp.pt0 && pn < sp.pt1) lc = pn - sp.pt0;
        else {   // ghost block column: position in the sorted ghost list
            int64_t lo = 0, hi = sp.n_ghost;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sp.ghost_pts[mid] < pn) lo = mid + 1; else hi = mid; }
            lc = nown + lo;
            ghost = true;
        }
        cc[tb] = (int32_t)lc;
        const double m = sp.hermitian ? sp.mod_tab[pn % 997] : sp.mod_tab[q % 997];
        const cplx *blk = sp.blocks + (size_t)(sp.hermitian ? g.nb - 1 - b : b) * B * B;
        for (int u = 0; u < B; u++)        // column variable
            for (int v = 0; v < B; v++) {  // row variable
                cplx x;
                if (b == centre && u == v) {
                    x = sp.diag_tab[q % 499];
                    x.x -= sp.shift;
                } else {
                    const cplx e = sp.hermitian ? blk[u * B + v] : blk[v * B + u];
                    // numpy: complex * real is carried out as (a + bi)(m + 0i) = (a*m - b*0) + (a*0 + b*m)i
                    x = make_double2(e.x * m - e.y * 0.0, e.x * 0.0 + e.y * m);
                }
                if (sp.hermitian) x.y = -x.y;
                cv[(size_t)(u * B + v) * TB + tb] = x;
            }
    }
    if (ghost && tile_has_ghost) tile_has_ghost[t] = 1;
}

// CreateRandomMat.c:36-39 on the device: entries uniform in [0,1) + i[0,1) (PETSc's default complex MatSetRandom), a counter-based
// generator keyed by (seed, global row, column) so that a row-partitioned run fills the same matrix as one GPU
__device__ __forceinline__ double u01(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;                       // splitmix64
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void random_stack_kernel(cplx *stack, int64_t rows, int64_t srows, int64_t roff, int na, int nb, int64_t grow0,
                                    unsigned long long seed)
{
    const int64_t n = rows * na * nb;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(e % nb);
        const int64_t r = (e / nb) % rows;
        const int a = (int)(e / (nb * rows));
        const unsigned long long idx = ((unsigned long long)(grow0 + r) * (unsigned long long)(na * nb) + (unsigned long long)(a * nb + b)) * 2ull;
        const unsigned long long key = seed * 0xD1342543DE82EF95ull;
        stack[((int64_t)a * srows + roff + r) * nb + b] = make_double2(u01(idx ^ key), u01((idx + 1ull) ^ key));
    }
}

}  // namespace

cudaError_t launch_random_stack(cplx *stack, int64_t rows, int64_t srows, int64_t roff, int na, int nb, int64_t grow0,
                                unsigned long long seed, cudaStream_t s)
{
    random_stack_kernel<<<148 * 8, 256, 0, s>>>(stack, rows, srows, roff, na, nb, grow0, seed);
    return cudaGetLastError();
}

cudaError_t stencil_tile_counts(const StencilSpec &sp, int TB, int64_t ntiles, int32_t *counts_dev, uint32_t *masks_dev, cudaStream_t s)
{
    Geo g;
    g.d = sp.d;
    g.nb = 1;
    for (int i = 0; i < sp.d; i++) { g.nb *= 3; g.shape[i] = sp.shape[i]; }
    for (int i = 0; i < sp.d; i++) {
        g.stride[i] = 1;
        for (int j = i + 1; j < sp.d; j++) g.stride[i] *= sp.shape[j];
    }
    stencil_tile_counts_kernel<<<(unsigned)((ntiles + 127) / 128), 128, 0, s>>>(g, sp.pt0, sp.pt1, TB, ntiles, counts_dev, masks_dev);
    return cudaGetLastError();
}

cudaError_t stencil_fill(const StencilSpec &sp, int TB, int64_t ntiles, const int64_t *tile_ptr_dev, const uint32_t *masks_dev,
                         unsigned char *chunks, unsigned char *tile_has_ghost, cudaStream_t s)
{
    Geo g;
    g.d = sp.d;
    g.nb = 1;
    for (int i = 0; i < sp.d; i++) { g.nb *= 3; g.shape[i] = sp.shape[i]; }
    for (int i = 0; i < sp.d; i++) {
        g.stride[i] = 1;
        for (int j = i + 1; j < sp.d; j++) g.stride[i] *= sp.shape[j];
    }
    const int64_t nthr = ntiles * TB;
    const unsigned blocks = (unsigned)((nthr + 127) / 128);
    switch (sp.B) {
        case 3: stencil_fill_kernel<3><<<blocks, 128, 0, s>>>(g, sp, TB, ntiles, tile_ptr_dev, masks_dev, chunks, tile_has_ghost); break;
        case 4: stencil_fill_kernel<4><<<blocks, 128, 0, s>>>(g, sp, TB, ntiles, tile_ptr_dev, masks_dev, chunks, tile_has_ghost); break;
        case 5: stencil_fill_kernel<5><<<blocks, 128, 0, s>>>(g, sp, TB, ntiles, tile_ptr_dev, masks_dev, chunks, tile_has_ghost); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
