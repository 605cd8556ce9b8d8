// kc_reduce.cu -- run-length reduction of sorted keys and record (un)packing.
//
// Replaces reduceKMers (GPUHandler.cu:329-360, a serial host loop over the
// D2H'd raw records) and the byte-level record handling of FileDump /
// SortedKMerFile.  Sorted keys never leave the device: a single-pass kernel
// flags run heads, ranks them with a decoupled look-back and writes each
// distinct key once together with the index where its run starts; counts are
// differences of neighbouring starts (uint32 wrap).
#include "kc_internal.h"

namespace kc {

namespace {

constexpr int kRleThreads = 256;
constexpr int kRleItems = 8;
constexpr int kRleTile = kRleThreads * kRleItems;

struct RleWs {
    unsigned long long ticket;
    unsigned long long pad[31];
    // followed by status[n_tiles]
};

template <int W>
__global__ void __launch_bounds__(kRleThreads) rle_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                          uint64_t *__restrict__ out_keys,
                                                          uint32_t *__restrict__ out_starts,
                                                          unsigned long long *__restrict__ d_num_unique,
                                                          unsigned long long *ticket, uint64_t *status) {
    constexpr int WARPS = kRleThreads / 32;
    __shared__ uint32_t s_cnt[kRleItems * WARPS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (uint32_t)atomicAdd(ticket, 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * kRleTile;

    Key<W> key[kRleItems];
    uint32_t ballots[kRleItems];
    uint32_t heads = 0;
#pragma unroll
    for (int j = 0; j < kRleItems; j++) {
        const uint64_t i = base + (uint64_t)j * kRleThreads + tid;
        bool head = false;
        if (i < n) {
            key[j] = ld_key<W>(keys, i);
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) key[j].w[w] = 0;
        }
        // previous key: neighbour lane, or one extra load at the warp edge
        Key<W> prev;
#pragma unroll
        for (int w = 0; w < W; w++) prev.w[w] = __shfl_up_sync(0xffffffffu, key[j].w[w], 1);
        if (lane == 0 && i > 0 && i < n) prev = ld_key<W>(keys, i - 1);
        if (i < n) head = (i == 0) || !key_eq<W>(key[j], prev);
        ballots[j] = __ballot_sync(0xffffffffu, head);
        if (head) heads |= 1u << j;
        if (lane == 0) s_cnt[j * WARPS + warp] = __popc(ballots[j]);
    }
    __syncthreads();
    // exclusive scan of the ITEMS x WARPS head counts (index order = j-major), one warp
    if (warp == 0) {
        uint32_t total = 0;
        for (int c0 = 0; c0 < kRleItems * WARPS; c0 += 32) {
            uint32_t v = (c0 + lane < kRleItems * WARPS) ? s_cnt[c0 + lane] : 0;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            if (c0 + lane < kRleItems * WARPS) s_cnt[c0 + lane] = total + incl - v;
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
        const uint64_t excl = lookback_exclusive_warp(status, tile, total);      // (32 predecessors per round)
        if (lane == 0) {
            s_base = excl;
            if (base + kRleTile >= n) {          // the tile holding the last key closes the run list
                const uint64_t u = excl + total;
                *d_num_unique = u;
                out_starts[u] = (uint32_t)n;
            }
        }
    }
    __syncthreads();
    const uint64_t obase = s_base;
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < kRleItems; j++) {
        if (heads & (1u << j)) {
            const uint64_t i = base + (uint64_t)j * kRleThreads + tid;
            const uint64_t o = obase + s_cnt[j * WARPS + warp] + __popc(ballots[j] & lt);
            st_key<W>(out_keys, o, key[j]);
            out_starts[o] = (uint32_t)i;
        }
    }
}

template <int W>
__global__ void starts_to_counts_kernel(const uint32_t *__restrict__ starts, const uint64_t *__restrict__ keys,
                                        uint64_t n_unique, const unsigned long long *__restrict__ d_n_invalid,
                                        uint32_t *__restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_unique) return;
    uint32_t c = starts[i + 1] - starts[i];
    if (i == 0 && d_n_invalid) {
        Key<W> k0 = ld_key<W>(keys, 0);
        if (key_is_zero<W>(k0)) c -= (uint32_t)(*d_n_invalid);   // empty slots sorted in as key 0
    }
    counts[i] = c;
}

// counts_out[i] = sum of counts_in over run i; the number of runs lives on the device
__global__ void segsum_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts_in,
                              const unsigned long long *__restrict__ d_n_unique, uint32_t *__restrict__ counts_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *d_n_unique) return;
    uint32_t s = 0;
    for (uint32_t j = starts[i]; j < starts[i + 1]; j++) s += counts_in[j];
    counts_out[i] = s;
}

// records as a stream of uint32: 2W key halves then the count
__global__ void pack_kernel(const uint32_t *__restrict__ keys32, const uint32_t *__restrict__ counts, uint64_t n,
                            uint32_t words_per_rec, uint32_t *__restrict__ out) {
    const uint64_t total = n * words_per_rec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t rec = t / words_per_rec;
        const uint32_t part = (uint32_t)(t - rec * words_per_rec);
        out[t] = (part + 1 < words_per_rec) ? keys32[rec * (words_per_rec - 1) + part] : counts[rec];
    }
}

__global__ void unpack_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t words_per_rec,
                              uint32_t *__restrict__ keys32, uint32_t *__restrict__ counts) {
    const uint64_t total = n * words_per_rec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t rec = t / words_per_rec;
        const uint32_t part = (uint32_t)(t - rec * words_per_rec);
        const uint32_t v = in[t];
        if (part + 1 < words_per_rec) keys32[rec * (words_per_rec - 1) + part] = v;
        else counts[rec] = v;
    }
}

template <int W>
__global__ void lower_bound_kernel(const uint64_t *__restrict__ keys, uint64_t n, const uint64_t *__restrict__ q,
                                   uint32_t n_q, unsigned long long *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_q) return;
    Key<W> x;
#pragma unroll
    for (int w = 0; w < W; w++) x.w[w] = q[(uint64_t)t * W + w];
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (key_lt<W>(ld_key<W>(keys, mid), x)) lo = mid + 1;
        else hi = mid;
    }
    out[t] = lo;
}

inline uint32_t grid_for(uint64_t n, int threads, uint32_t cap = 0) {
    if (cap == 0) cap = (uint32_t)sm_count() * 16;
    uint64_t g = div_up(n ? n : 1, (uint64_t)threads);
    return (uint32_t)(g < cap ? g : cap);
}

}  // namespace

// ---- text dump of a run: replaces KMerPrinter::print / printKmer (KMerPrinter.cpp:35-91)
// One line per record: the 32 W letters of the key words (A, C, G, T for codes 0..3, most
// significant pair first), a blank, the count in decimal, a newline. Lines differ in length by the
// count's digits, so a tile of records ranks its bytes with a block scan and a decoupled look-back.
constexpr int kPrintThreads = 256;
constexpr int kPrintItems = 4;                      // consecutive records per thread
constexpr int kPrintTile = kPrintThreads * kPrintItems;

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10u) { v /= 10u; d++; }
    return d;
}

template <int W>
__global__ void __launch_bounds__(kPrintThreads) print_kernel(const uint64_t *__restrict__ keys,
                                                              const uint32_t *__restrict__ counts, uint64_t n,
                                                              char *__restrict__ out, unsigned long long *d_bytes,
                                                              unsigned long long *ticket, uint64_t *status) {
    constexpr int WARPS = kPrintThreads / 32;
    constexpr uint32_t kFixed = 32u * W + 2u;       // letters + blank + newline
    __shared__ uint32_t s_warp[WARPS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (uint32_t)atomicAdd(ticket, 1ull);     // tiles in the order CTAs start: a look-back never
    __syncthreads();                                              // waits for a tile that has not been scheduled
    const uint32_t tile = s_tile;
    const uint64_t first = (uint64_t)tile * kPrintTile + (uint64_t)tid * kPrintItems;
    uint32_t cnt[kPrintItems], len[kPrintItems], mine = 0;
#pragma unroll
    for (int j = 0; j < kPrintItems; j++) {
        cnt[j] = 0;
        len[j] = 0;
        if (first + j < n) {
            cnt[j] = counts[first + j];
            len[j] = kFixed + dec_digits(cnt[j]);
        }
        mine += len[j];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
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
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive_warp(status, tile, total);
        if (lane == 0) {
            s_base = excl;
            if (tile + 1 == gridDim.x) *d_bytes = excl + total;     // the last tile closes the text
        }
    }
    __syncthreads();
    uint64_t o = s_base + woff + incl - mine;
#pragma unroll
    for (int j = 0; j < kPrintItems; j++) {
        if (first + j >= n) break;
        const Key<W> k = ld_key<W>(keys, first + j);
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint64_t v = k.w[w];
#pragma unroll
            for (int b = 0; b < 32; b++) out[o + 32 * w + b] = "ACGT"[(v >> (62 - 2 * b)) & 3ull];
        }
        out[o + 32 * W] = ' ';
        const uint32_t nd = len[j] - kFixed;
        uint32_t v = cnt[j];
        for (uint32_t d = 0; d < nd; d++) {
            out[o + 32 * W + nd - d] = (char)('0' + v % 10u);
            v /= 10u;
        }
        out[o + len[j] - 1] = '\n';
        o += len[j];
    }
}

uint64_t print_workspace_bytes(uint64_t n) { return 256 + div_up(n ? n : 1, (uint64_t)kPrintTile) * 8 + 256; }
uint64_t print_max_bytes(uint64_t n, int W) { return n * (32ull * W + 12ull); }

cudaError_t print_records_text(const uint64_t *keys, const uint32_t *counts, uint64_t n, int W, char *d_text,
                               unsigned long long *d_bytes, void *ws, cudaStream_t s) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_bytes, 0, 8, s)) != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    const uint64_t tiles = div_up(n, (uint64_t)kPrintTile);
    if (tiles >= (1ull << 31)) return cudaErrorInvalidValue;
    if ((e = cudaMemsetAsync(ws, 0, 256 + tiles * 8, s)) != cudaSuccess) return e;
    unsigned long long *ticket = static_cast<unsigned long long *>(ws);
    uint64_t *status = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ws) + 256);
    switch (W) {
        case 1: print_kernel<1><<<(uint32_t)tiles, kPrintThreads, 0, s>>>(keys, counts, n, d_text, d_bytes, ticket, status); break;
        case 2: print_kernel<2><<<(uint32_t)tiles, kPrintThreads, 0, s>>>(keys, counts, n, d_text, d_bytes, ticket, status); break;
        case 3: print_kernel<3><<<(uint32_t)tiles, kPrintThreads, 0, s>>>(keys, counts, n, d_text, d_bytes, ticket, status); break;
        case 4: print_kernel<4><<<(uint32_t)tiles, kPrintThreads, 0, s>>>(keys, counts, n, d_text, d_bytes, ticket, status); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

uint64_t rle_workspace_bytes(uint64_t n) {
    return 256 + div_up(n ? n : 1, (uint64_t)kRleTile) * 8 + 256;
}

cudaError_t rle_unique(const uint64_t *sorted_keys, uint64_t n, int W, uint64_t *out_keys, uint32_t *out_starts,
                       unsigned long long *d_num_unique, void *ws, cudaStream_t s, int *n_launches) {
    cudaError_t e;
    if (n == 0) {
        if ((e = cudaMemsetAsync(d_num_unique, 0, 8, s)) != cudaSuccess) return e;
        return cudaMemsetAsync(out_starts, 0, 4, s);
    }
    if (n >= (1ull << 32)) return cudaErrorInvalidValue;
    const uint64_t tiles = div_up(n, (uint64_t)kRleTile);
    if ((e = cudaMemsetAsync(ws, 0, 256 + tiles * 8, s)) != cudaSuccess) return e;
    unsigned long long *ticket = static_cast<unsigned long long *>(ws);
    uint64_t *status = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ws) + 256);
    switch (W) {
        case 1: rle_kernel<1><<<(uint32_t)tiles, kRleThreads, 0, s>>>(sorted_keys, n, out_keys, out_starts, d_num_unique, ticket, status); break;
        case 2: rle_kernel<2><<<(uint32_t)tiles, kRleThreads, 0, s>>>(sorted_keys, n, out_keys, out_starts, d_num_unique, ticket, status); break;
        case 3: rle_kernel<3><<<(uint32_t)tiles, kRleThreads, 0, s>>>(sorted_keys, n, out_keys, out_starts, d_num_unique, ticket, status); break;
        case 4: rle_kernel<4><<<(uint32_t)tiles, kRleThreads, 0, s>>>(sorted_keys, n, out_keys, out_starts, d_num_unique, ticket, status); break;
        default: return cudaErrorInvalidValue;
    }
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t starts_to_counts(const uint32_t *starts, const uint64_t *keys, int W, uint64_t n_unique,
                             const unsigned long long *d_n_invalid, uint32_t *counts, cudaStream_t s) {
    if (n_unique == 0) return cudaSuccess;
    const uint32_t g = (uint32_t)div_up(n_unique, 256);
    switch (W) {
        case 1: starts_to_counts_kernel<1><<<g, 256, 0, s>>>(starts, keys, n_unique, d_n_invalid, counts); break;
        case 2: starts_to_counts_kernel<2><<<g, 256, 0, s>>>(starts, keys, n_unique, d_n_invalid, counts); break;
        case 3: starts_to_counts_kernel<3><<<g, 256, 0, s>>>(starts, keys, n_unique, d_n_invalid, counts); break;
        case 4: starts_to_counts_kernel<4><<<g, 256, 0, s>>>(starts, keys, n_unique, d_n_invalid, counts); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t pack_records(const uint64_t *keys, const uint32_t *counts, uint64_t n, int W, void *records,
                         cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t wpr = 2 * W + 1;
    pack_kernel<<<grid_for(n * wpr, 256), 256, 0, s>>>(reinterpret_cast<const uint32_t *>(keys), counts, n, wpr,
                                                       static_cast<uint32_t *>(records));
    return cudaGetLastError();
}

cudaError_t unpack_records(const void *records, uint64_t n, int W, uint64_t *keys, uint32_t *counts,
                           cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t wpr = 2 * W + 1;
    unpack_kernel<<<grid_for(n * wpr, 256), 256, 0, s>>>(static_cast<const uint32_t *>(records), n, wpr,
                                                         reinterpret_cast<uint32_t *>(keys), counts);
    return cudaGetLastError();
}

cudaError_t fold_sorted_pairs(const uint64_t *keys, const uint32_t *counts, uint64_t n, int W, uint64_t *out_keys,
                              uint32_t *out_counts, unsigned long long *d_num_unique, void *ws, cudaStream_t s,
                              int *n_launches) {
    // ws: [rle workspace][starts (n+1) x uint32]
    uint8_t *wsb = static_cast<uint8_t *>(ws);
    uint32_t *starts = reinterpret_cast<uint32_t *>(wsb + ((rle_workspace_bytes(n) + 255) & ~255ull));
    cudaError_t e = rle_unique(keys, n, W, out_keys, starts, d_num_unique, ws, s, n_launches);
    if (e != cudaSuccess || n == 0) return e;
    const uint32_t g = (uint32_t)div_up(n, 256);   // worst case: every key distinct
    segsum_kernel<<<g, 256, 0, s>>>(starts, counts, d_num_unique, out_counts);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t lower_bounds(const uint64_t *keys, uint64_t n, int W, const uint64_t *d_queries, uint32_t n_q,
                         unsigned long long *d_out, cudaStream_t s) {
    if (n_q == 0) return cudaSuccess;
    const uint32_t g = (n_q + 63) / 64;
    switch (W) {
        case 1: lower_bound_kernel<1><<<g, 64, 0, s>>>(keys, n, d_queries, n_q, d_out); break;
        case 2: lower_bound_kernel<2><<<g, 64, 0, s>>>(keys, n, d_queries, n_q, d_out); break;
        case 3: lower_bound_kernel<3><<<g, 64, 0, s>>>(keys, n, d_queries, n_q, d_out); break;
        case 4: lower_bound_kernel<4><<<g, 64, 0, s>>>(keys, n, d_queries, n_q, d_out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace kc
