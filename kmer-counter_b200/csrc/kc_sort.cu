// kc_sort.cu -- hand-written LSD radix sort for 64..256-bit keys (sm_100a).
//
// Replaces sortKmers (GPUHandler.cu:300-327, thrust::sort with the comparators
// at :247-298).  Order: words compared most-significant first, unsigned.
//
// Structure ("onesweep"): one histogram kernel counts every 8-bit digit of every
// pass in a single read of the keys; each pass is then ONE kernel that reads a
// tile, ranks it, and scatters it, with the cross-tile digit offsets resolved by
// a decoupled look-back over per-tile status words (no separate scan kernel, no
// second read of the keys).  Per pass a key moves HBM -> SM -> HBM exactly once:
// 2 * 8W bytes of algorithmic traffic per key per pass.
//
// Ranking uses warp-private digit counters and ballot-based peer matching, so
// there are no shared-memory atomics in the scatter kernel.
#include "kc_internal.h"

namespace kc {

namespace {

constexpr int kMaxPasses = 8 * kMaxWords;   // 32
constexpr uint32_t kStEmpty = 0u, kStAggregate = 1u << 30, kStInclusive = 2u << 30, kStMask = (1u << 30) - 1;

struct PassList {
    int n;
    uint8_t word[kMaxPasses];
    uint8_t shift[kMaxPasses];
};

template <int W>
__device__ __forceinline__ uint32_t digit_of(const Key<W> &k, int word, int shift) {
    uint64_t w = k.w[0];
#pragma unroll
    for (int i = 1; i < W; i++) w = (word == i) ? k.w[i] : w;
    return (uint32_t)(w >> shift) & 0xFFu;
}

// ------------------------------------------------------------------ histogram
constexpr int kHistThreads = 512;

template <int W>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const uint64_t *__restrict__ keys, uint32_t n,
                                                            int lo_bit, int n_passes, uint32_t *__restrict__ ghist) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < n_passes * kRadixBins; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * kHistThreads;
    for (uint32_t i = blockIdx.x * kHistThreads + threadIdx.x; i < n; i += stride) {
        Key<W> k = ld_key<W>(keys, i);
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint64_t x = k.w[w];
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int bit = (W - 1 - w) * 64 + 8 * b;      // pass index = (bit - lo_bit) / 8
                if (bit >= lo_bit)
                    atomicAdd(&sh[((bit - lo_bit) >> 3) * kRadixBins + ((uint32_t)(x >> (8 * b)) & 0xFFu)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * kRadixBins; i += kHistThreads) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// one block of 256 threads per pass: exclusive scan of the 256 bins
__global__ void __launch_bounds__(kRadixBins) bin_scan_kernel(const uint32_t *__restrict__ ghist,
                                                              uint32_t *__restrict__ bin_base) {
    __shared__ uint32_t warp_tot[8];
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    uint32_t c = ghist[blockIdx.x * kRadixBins + d];
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; w++) off += warp_tot[w];
    bin_base[blockIdx.x * kRadixBins + d] = off + incl - c;
}

// --------------------------------------------------------------- scatter pass
struct PassParams {
    const uint64_t *keys_in;
    uint64_t *keys_out;
    const uint32_t *vals_in;
    uint32_t *vals_out;
    uint32_t n;
    int word, shift;
    const uint32_t *bin_base;   // [256] global exclusive offsets of this pass
    uint32_t *status;           // [n_tiles][256], zeroed
    uint32_t *ticket;           // zeroed
};

template <int W, bool HAS_VAL, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) onesweep_kernel(PassParams p) {
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    static_assert(THREADS >= kRadixBins, "one thread per bin needed");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *whist = reinterpret_cast<uint32_t *>(smem_raw);                 // [WARPS][256]
    Key<W> *ex_keys = reinterpret_cast<Key<W> *>(smem_raw + WARPS * kRadixBins * 4);   // [TILE]
    uint32_t *ex_vals = reinterpret_cast<uint32_t *>(ex_keys);                // reused after the key scatter
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_tile_start[kRadixBins];
    __shared__ uint32_t s_scatter[kRadixBins];
    __shared__ uint32_t s_warp_tot[kRadixBins / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
    for (uint32_t i = tid; i < WARPS * kRadixBins; i += THREADS) whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * TILE;
    const uint32_t n_valid = (p.n - base) < (uint32_t)TILE ? (p.n - base) : (uint32_t)TILE;

    // ---- load, warp-striped: item i of lane l is element warp_base + i*32 + l
    Key<W> key[ITEMS];
    uint32_t val[HAS_VAL ? ITEMS : 1];
    const uint32_t warp_base = base + warp * 32 * ITEMS;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        uint32_t idx = warp_base + i * 32 + lane;
        if (idx < p.n) {
            key[i] = ld_key<W>(p.keys_in, idx);
            if constexpr (HAS_VAL) val[i] = p.vals_in[idx];
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) key[i].w[w] = ~0ull;   // pads rank after every real key of bin 255
            if constexpr (HAS_VAL) val[i] = 0;
        }
    }

    // ---- rank inside the warp: peers = lanes holding the same digit
    uint32_t *wh = whist + warp * kRadixBins;
    uint16_t rank[ITEMS];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = digit_of<W>(key[i], p.word, p.shift);
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < kRadixBits; b++) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if ((int)lane == leader) {
            prev = wh[d];
            wh[d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = (uint16_t)(prev + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive scan over warps, tile totals, scan over digits, look-back
    uint32_t tile_count = 0;
    if (tid < kRadixBins) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t c = whist[w * kRadixBins + tid];
            whist[w * kRadixBins + tid] = tile_count;
            tile_count += c;
        }
        uint32_t incl = tile_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        s_tile_start[tid] = incl - tile_count;   // completed below
    }
    __syncthreads();
    if (tid < kRadixBins) {
        uint32_t off = 0;
        for (uint32_t w = 0; w < warp; w++) off += s_warp_tot[w];
        const uint32_t start = s_tile_start[tid] + off;
        s_tile_start[tid] = start;
        // real keys only: the pads all sit in bin 255
        const uint32_t count = tile_count - ((tid == kRadixBins - 1) ? ((uint32_t)TILE - n_valid) : 0u);
        uint32_t *st = p.status + (size_t)tile * kRadixBins + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_release_u32(st, kStInclusive | count);
        } else {
            st_release_u32(st, kStAggregate | count);
            const uint32_t *q = st - kRadixBins;
            while (true) {
                const uint32_t s = ld_acquire_u32(q);
                const uint32_t fl = s & ~kStMask;
                if (fl == kStEmpty) { __nanosleep(32); continue; }
                excl += s & kStMask;
                if (fl == kStInclusive) break;
                q -= kRadixBins;
            }
            st_release_u32(st, kStInclusive | (excl + count));
        }
        s_scatter[tid] = p.bin_base[tid] + excl - start;   // dst = s_scatter[d] + slot
    }
    __syncthreads();

    // ---- exchange through shared memory so that the global writes are runs of equal digits
    uint16_t pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t d = digit_of<W>(key[i], p.word, p.shift);
        pos[i] = (uint16_t)(s_tile_start[d] + wh[d] + rank[i]);
        ex_keys[pos[i]] = key[i];
    }
    __syncthreads();
    uint32_t dst[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const uint32_t slot = j * THREADS + tid;
        if (slot < n_valid) {
            const Key<W> k = ex_keys[slot];
            const uint32_t d = digit_of<W>(k, p.word, p.shift);
            dst[j] = s_scatter[d] + slot;
            st_key<W>(p.keys_out, dst[j], k);
        }
    }
    if constexpr (HAS_VAL) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) ex_vals[pos[i]] = val[i];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const uint32_t slot = j * THREADS + tid;
            if (slot < n_valid) p.vals_out[dst[j]] = ex_vals[slot];
        }
    }
}

template <int W> struct SortCfg;
template <> struct SortCfg<1> { static constexpr int THREADS = 384, ITEMS = 16; };
template <> struct SortCfg<2> { static constexpr int THREADS = 384, ITEMS = 8; };
template <> struct SortCfg<3> { static constexpr int THREADS = 256, ITEMS = 8; };
template <> struct SortCfg<4> { static constexpr int THREADS = 256, ITEMS = 6; };

inline int tile_of(int W) {
    switch (W) {
        case 1: return SortCfg<1>::THREADS * SortCfg<1>::ITEMS;
        case 2: return SortCfg<2>::THREADS * SortCfg<2>::ITEMS;
        case 3: return SortCfg<3>::THREADS * SortCfg<3>::ITEMS;
        default: return SortCfg<4>::THREADS * SortCfg<4>::ITEMS;
    }
}

template <int W, bool HAS_VAL>
cudaError_t launch_pass(const PassParams &pp, uint32_t n_tiles, cudaStream_t s) {
    constexpr int THREADS = SortCfg<W>::THREADS, ITEMS = SortCfg<W>::ITEMS;
    auto kern = onesweep_kernel<W, HAS_VAL, THREADS, ITEMS>;
    const size_t smem = (THREADS / 32) * kRadixBins * 4 + (size_t)THREADS * ITEMS * sizeof(Key<W>);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    kern<<<n_tiles, THREADS, smem, s>>>(pp);
    return cudaGetLastError();
}

template <int W>
cudaError_t launch_hist(const uint64_t *keys, uint32_t n, int lo_bit, int n_passes, uint32_t *ghist, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t grid = (uint32_t)sms * 2;
    uint32_t need = (n + kHistThreads - 1) / kHistThreads;
    if (grid > need) grid = need ? need : 1;
    hist_kernel<W><<<grid, kHistThreads, n_passes * kRadixBins * 4, s>>>(keys, n, lo_bit, n_passes, ghist);
    return cudaGetLastError();
}

}  // namespace

uint64_t sort_workspace_bytes(uint64_t n, int W) {
    uint64_t tiles = div_up(n ? n : 1, (uint64_t)tile_of(W));
    uint64_t b = 0;
    b += (uint64_t)kMaxPasses * kRadixBins * 4 * 2;   // ghist + bin_base
    b += 256;                                         // tickets
    b += tiles * kRadixBins * 4;                      // status (reused by every pass)
    return (b + 255) & ~255ull;
}

cudaError_t radix_sort(uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, uint64_t n, int W,
                       int lo_bit, SortWorkspace ws, cudaStream_t s, uint64_t **sorted_keys, uint32_t **sorted_vals,
                       int *n_launches, cudaEvent_t ev_pass_begin, cudaEvent_t ev_pass_end) {
    *sorted_keys = keys_a;
    if (sorted_vals) *sorted_vals = vals_a;
    if (ev_pass_begin && n <= 1) {
        cudaEventRecord(ev_pass_begin, s);
        cudaEventRecord(ev_pass_end, s);
    }
    if (n <= 1) return cudaSuccess;
    if (n >= (1ull << 30) || W < 1 || W > kMaxWords) return cudaErrorInvalidValue;
    if (ws.bytes < sort_workspace_bytes(n, W)) return cudaErrorInvalidValue;
    const bool has_val = vals_a != nullptr;

    PassList pl{};
    lo_bit &= ~7;
    for (int bit = lo_bit; bit < 64 * W; bit += kRadixBits) {
        pl.word[pl.n] = (uint8_t)(W - 1 - bit / 64);
        pl.shift[pl.n] = (uint8_t)(bit % 64);
        pl.n++;
    }
    uint8_t *wsb = static_cast<uint8_t *>(ws.base);
    uint32_t *ghist = reinterpret_cast<uint32_t *>(wsb);
    uint32_t *bin_base = ghist + kMaxPasses * kRadixBins;
    uint32_t *tickets = bin_base + kMaxPasses * kRadixBins;
    uint32_t *status = tickets + 64;
    const uint32_t n_tiles = (uint32_t)div_up(n, (uint64_t)tile_of(W));
    const size_t status_bytes = (size_t)n_tiles * kRadixBins * 4;

    cudaError_t e;
    if ((e = cudaMemsetAsync(ghist, 0, (size_t)kMaxPasses * kRadixBins * 4 * 2 + 256, s)) != cudaSuccess) return e;
    switch (W) {
        case 1: e = launch_hist<1>(keys_a, (uint32_t)n, lo_bit, pl.n, ghist, s); break;
        case 2: e = launch_hist<2>(keys_a, (uint32_t)n, lo_bit, pl.n, ghist, s); break;
        case 3: e = launch_hist<3>(keys_a, (uint32_t)n, lo_bit, pl.n, ghist, s); break;
        default: e = launch_hist<4>(keys_a, (uint32_t)n, lo_bit, pl.n, ghist, s); break;
    }
    if (e != cudaSuccess) return e;
    bin_scan_kernel<<<pl.n, kRadixBins, 0, s>>>(ghist, bin_base);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (n_launches) *n_launches += 2;

    if (ev_pass_begin) cudaEventRecord(ev_pass_begin, s);
    uint64_t *kin = keys_a, *kout = keys_b;
    uint32_t *vin = vals_a, *vout = vals_b;
    for (int ps = 0; ps < pl.n; ps++) {
        if ((e = cudaMemsetAsync(status, 0, status_bytes, s)) != cudaSuccess) return e;
        PassParams pp{kin, kout, vin, vout, (uint32_t)n, pl.word[ps], pl.shift[ps], bin_base + ps * kRadixBins,
                      status, tickets + ps};
        if (has_val) {
            switch (W) {
                case 1: e = launch_pass<1, true>(pp, n_tiles, s); break;
                case 2: e = launch_pass<2, true>(pp, n_tiles, s); break;
                case 3: e = launch_pass<3, true>(pp, n_tiles, s); break;
                default: e = launch_pass<4, true>(pp, n_tiles, s); break;
            }
        } else {
            switch (W) {
                case 1: e = launch_pass<1, false>(pp, n_tiles, s); break;
                case 2: e = launch_pass<2, false>(pp, n_tiles, s); break;
                case 3: e = launch_pass<3, false>(pp, n_tiles, s); break;
                default: e = launch_pass<4, false>(pp, n_tiles, s); break;
            }
        }
        if (e != cudaSuccess) return e;
        if (n_launches) *n_launches += 1;
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (ev_pass_end) cudaEventRecord(ev_pass_end, s);
    *sorted_keys = kin;
    if (sorted_vals) *sorted_vals = vin;
    return cudaSuccess;
}

}  // namespace kc
