// kc_merge.cu -- GPU merge-path merge of two sorted key-unique runs, summing the
// counts of equal keys.
//
// Replaces KMerFileMerger::Merge (KMerFileMerger.cpp:49-96: a serial O(M) min-scan
// per output record over file cursors) and the cursor class under it
// (SortedKMerFile.cpp).  M runs are merged as a pairwise tree by the caller.
//
// Kernel 1 cuts the merged sequence into equal diagonals (merge path, ties go
// to run A) and nudges a cut that would separate an equal (A,B) pair.  Kernel 2
// loads one tile of A and B into shared memory, ranks every element against the
// other run (binary search in shared memory), drops the B half of equal pairs
// into its A partner, and writes the surviving records in order; the output
// offset of a tile comes from a decoupled look-back.
#include "kc_internal.h"

namespace kc {

namespace {

constexpr int kMergeThreads = 256;

template <int W> struct MergeCfg { static constexpr int TILE = 2048 / W; };

template <int W>
__global__ void merge_partition_kernel(const uint64_t *__restrict__ ka, uint64_t na, const uint64_t *__restrict__ kb,
                                       uint64_t nb, uint32_t n_tiles, uint64_t *__restrict__ cut_a,
                                       uint64_t *__restrict__ cut_b) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    uint64_t diag = (uint64_t)t * MergeCfg<W>::TILE;
    if (diag > na + nb) diag = na + nb;
    uint64_t lo = diag > nb ? diag - nb : 0, hi = diag < na ? diag : na;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const Key<W> a = ld_key<W>(ka, mid), b = ld_key<W>(kb, diag - 1 - mid);
        if (!key_lt<W>(b, a)) lo = mid + 1;   // a <= b: a goes first
        else hi = mid;
    }
    uint64_t a = lo, b = diag - lo;
    if (a > 0 && b < nb && key_eq<W>(ld_key<W>(ka, a - 1), ld_key<W>(kb, b))) b++;   // keep equal pairs together
    cut_a[t] = a;
    cut_b[t] = b;
}

template <int W>
__global__ void __launch_bounds__(kMergeThreads) merge_tile_kernel(
    const uint64_t *__restrict__ ka, const uint32_t *__restrict__ ca, const uint64_t *__restrict__ kb,
    const uint32_t *__restrict__ cb, const uint64_t *__restrict__ cut_a, const uint64_t *__restrict__ cut_b,
    uint32_t n_tiles, uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_counts,
    unsigned long long *__restrict__ d_num_out, unsigned long long *ticket, uint64_t *status) {
    constexpr int TILE = MergeCfg<W>::TILE;
    constexpr int CAP = TILE + 1;
    constexpr int IPT = (CAP + kMergeThreads - 1) / kMergeThreads;
    constexpr int WARPS = kMergeThreads / 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Key<W> *s_in = reinterpret_cast<Key<W> *>(smem_raw);                       // A then B, CAP keys
    Key<W> *s_mk = s_in + CAP;                                                 // merged keys
    uint32_t *s_inc = reinterpret_cast<uint32_t *>(s_mk + CAP);                // input counts
    uint32_t *s_mc = s_inc + CAP;                                              // merged counts | dup flag (bit 32 kept apart)
    uint8_t *s_dup = reinterpret_cast<uint8_t *>(s_mc + CAP);
    __shared__ uint32_t s_tile, s_warp[WARPS];
    __shared__ uint64_t s_base;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (uint32_t)atomicAdd(ticket, 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t a0 = cut_a[tile], a1 = cut_a[tile + 1], b0 = cut_b[tile], b1 = cut_b[tile + 1];
    const uint32_t na = (uint32_t)(a1 - a0), nb = (uint32_t)(b1 - b0), m = na + nb;
    Key<W> *sA = s_in, *sB = s_in + na;
    for (uint32_t i = tid; i < m; i += kMergeThreads) {
        if (i < na) { s_in[i] = ld_key<W>(ka, a0 + i); s_inc[i] = ca[a0 + i]; }
        else { s_in[i] = ld_key<W>(kb, b0 + (i - na)); s_inc[i] = cb[b0 + (i - na)]; }
    }
    __syncthreads();
    for (uint32_t i = tid; i < m; i += kMergeThreads) {
        const Key<W> x = s_in[i];
        uint32_t pos;
        uint8_t dup = 0;
        if (i < na) {           // # of B strictly below x
            uint32_t lo = 0, hi = nb;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (key_lt<W>(sB[mid], x)) lo = mid + 1; else hi = mid; }
            pos = i + lo;
        } else {                // # of A at or below x
            uint32_t lo = 0, hi = na;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (!key_lt<W>(x, sA[mid])) lo = mid + 1; else hi = mid; }
            pos = (i - na) + lo;
            dup = (lo > 0 && key_eq<W>(sA[lo - 1], x)) ? 1 : 0;
        }
        s_mk[pos] = x;
        s_mc[pos] = s_inc[i];
        s_dup[pos] = dup;
    }
    __syncthreads();
    // survivors = non-dup positions; a survivor followed by a dup absorbs its count
    const uint32_t p0 = tid * IPT;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const uint32_t p = p0 + j;
        if (p < m && !s_dup[p]) mine++;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
        if (w < (int)warp) woff += s_warp[w];
        total += s_warp[w];
    }
    if (tid == 0) {
        const uint64_t excl = lookback_exclusive(status, tile, total);
        s_base = excl;
        if (tile + 1 == n_tiles) *d_num_out = excl + total;
    }
    __syncthreads();
    uint64_t o = s_base + woff + incl - mine;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const uint32_t p = p0 + j;
        if (p < m && !s_dup[p]) {
            uint32_t c = s_mc[p];
            if (p + 1 < m && s_dup[p + 1]) c += s_mc[p + 1];
            st_key<W>(out_keys, o, s_mk[p]);
            out_counts[o] = c;
            o++;
        }
    }
}

template <int W>
cudaError_t merge_pair_w(const uint64_t *ka, const uint32_t *ca, uint64_t na, const uint64_t *kb, const uint32_t *cb,
                         uint64_t nb, uint64_t *out_keys, uint32_t *out_counts, unsigned long long *d_num_out,
                         void *ws, cudaStream_t s, int *n_launches) {
    constexpr int TILE = MergeCfg<W>::TILE;
    constexpr int CAP = TILE + 1;
    const uint64_t total = na + nb;
    const uint32_t n_tiles = (uint32_t)div_up(total, (uint64_t)TILE);
    uint8_t *wsb = static_cast<uint8_t *>(ws);
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(wsb);
    uint64_t *status = reinterpret_cast<uint64_t *>(wsb + 256);
    uint64_t *cut_a = status + n_tiles;
    uint64_t *cut_b = cut_a + n_tiles + 1;
    cudaError_t e;
    if ((e = cudaMemsetAsync(ws, 0, 256 + (size_t)n_tiles * 8, s)) != cudaSuccess) return e;
    merge_partition_kernel<W><<<(n_tiles + 1 + 127) / 128, 128, 0, s>>>(ka, na, kb, nb, n_tiles, cut_a, cut_b);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const size_t smem = (size_t)CAP * (2 * sizeof(Key<W>) + 2 * 4 + 1) + 16;
    auto kern = merge_tile_kernel<W>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    kern<<<n_tiles, kMergeThreads, smem, s>>>(ka, ca, kb, cb, cut_a, cut_b, n_tiles, out_keys, out_counts, d_num_out,
                                              ticket, status);
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

}  // namespace

uint64_t merge_workspace_bytes(uint64_t na, uint64_t nb) {
    const uint64_t tiles = div_up(na + nb + 1, (uint64_t)MergeCfg<4>::TILE) + 2;
    return 256 + tiles * 8 * 3 + 256;
}

cudaError_t merge_pair(const uint64_t *ka, const uint32_t *ca, uint64_t na, const uint64_t *kb, const uint32_t *cb,
                       uint64_t nb, int W, uint64_t *out_keys, uint32_t *out_counts, unsigned long long *d_num_out,
                       void *ws, cudaStream_t s, int *n_launches) {
    if (na + nb == 0) return cudaMemsetAsync(d_num_out, 0, 8, s);
    switch (W) {
        case 1: return merge_pair_w<1>(ka, ca, na, kb, cb, nb, out_keys, out_counts, d_num_out, ws, s, n_launches);
        case 2: return merge_pair_w<2>(ka, ca, na, kb, cb, nb, out_keys, out_counts, d_num_out, ws, s, n_launches);
        case 3: return merge_pair_w<3>(ka, ca, na, kb, cb, nb, out_keys, out_counts, d_num_out, ws, s, n_launches);
        case 4: return merge_pair_w<4>(ka, ca, na, kb, cb, nb, out_keys, out_counts, d_num_out, ws, s, n_launches);
    }
    return cudaErrorInvalidValue;
}

}  // namespace kc
