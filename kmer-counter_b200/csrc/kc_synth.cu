// kc_synth.cu -- deterministic synthetic reads generated on the device.
//
// Measurement infrastructure for bench.py: a counter-based
// generator, bit-identical to the host generator the tests use, so that inputs of
// the benchmark shapes can be produced in HBM (and copied to pinned host memory
// for the end-to-end leg) without a multi-second host loop.  Not part of the
// reference; reads are "packed lines" (stride L, no separators).
#include "../../include/kc_api.h"
#include "kc_common.cuh"

namespace kc {
namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr uint64_t kSeedGenome = 0x67656E6F6D650000ull, kSeedStart = 0x7374617274000000ull,
                   kSeedError = 0x6572726F72000000ull, kSeedLocus = 0x6C6F637573000000ull,
                   kSeedPick = 0x7069636B00000000ull;

__device__ __forceinline__ uint32_t genome_code(uint64_t seed, uint64_t g) {
    const uint64_t w = splitmix64((seed * 0x100000001B3ull) ^ kSeedGenome ^ (g >> 5));
    return (uint32_t)(w >> (2 * (g & 31))) & 3u;
}

__device__ __forceinline__ uint64_t zipf_rank(uint64_t h, uint64_t M) {
    uint32_t levels = 0;
    while ((1ull << levels) < M + 1 && levels < 63) levels++;
    if (levels == 0) return 0;
    const uint32_t j = (uint32_t)((h >> 40) % levels);
    const uint64_t r = ((1ull << j) - 1) + ((h & 0xFFFFFFFFFFull) & ((1ull << j) - 1));
    return r < M ? r : r % M;
}

// one thread per base
__global__ void synth_kernel(uint8_t *out, uint64_t first_read, uint64_t n_reads, uint32_t L, uint64_t genome_len,
                             double sub_rate, double n_rate, uint64_t seed, uint64_t zipf_loci) {
    const uint64_t total = n_reads * L;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool noisy = sub_rate > 0.0 || n_rate > 0.0;
    const double inv53 = 1.0 / 9007199254740992.0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint64_t rd = t / L;
        const uint32_t j = (uint32_t)(t - rd * L);
        const uint64_t i = first_read + rd;
        uint32_t code;
        if (genome_len) {
            const uint64_t span = genome_len - L + 1;
            uint64_t start = splitmix64(seed ^ kSeedStart ^ (i * 0x9E3779B97F4A7C15ull)) % span;
            if (zipf_loci) {
                const uint64_t h = splitmix64(seed ^ kSeedPick ^ i);
                if (h >> 63) start = splitmix64(seed ^ kSeedLocus ^ zipf_rank(h, zipf_loci)) % span;
            }
            code = genome_code(seed, start + j);
        } else {
            code = genome_code(seed, i * (uint64_t)L + j);
        }
        uint8_t c = (uint8_t)("ACGT"[code]);
        if (noisy) {
            const uint64_t h = splitmix64(seed ^ kSeedError ^ (i * (uint64_t)L + j));
            const double u = (double)(h >> 11) * inv53;
            if (u < n_rate) c = 'N';
            else if (u < n_rate + sub_rate) c = (uint8_t)("ACGT"[(code + 1 + (uint32_t)(h & 0x3ff) % 3) & 3]);
        }
        out[t] = c;
    }
}

}  // namespace
}  // namespace kc

extern "C" int kc_synth_reads(void *d_out, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                              uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed, uint64_t zipf_loci,
                              void *stream) {
    if (!d_out || read_len == 0 || (genome_len && genome_len < read_len)) return KC_ERR_ARG;
    if (n_reads == 0) return KC_OK;
    kc::synth_kernel<<<148 * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint8_t *>(d_out), first_read, n_reads, read_len, genome_len, sub_rate, n_rate, seed, zipf_loci);
    return cudaGetLastError() == cudaSuccess ? KC_OK : KC_ERR_CUDA;
}


// Development aid: can a kernel on the current device read n 64-bit words at d_ptr (e.g. a
// peer rank's buffer mapped through CUDA IPC)? Returns their sum in *out.
namespace kc { namespace {
__global__ void probe_sum_kernel(const unsigned long long *p, uint64_t n, unsigned long long *out) {
    unsigned long long s = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += p[i];
    atomicAdd(out, s);
}
} }
extern "C" int kc_debug_probe_read(const void *d_ptr, uint64_t n_words, unsigned long long *out) {
    unsigned long long *d = nullptr;
    if (cudaMalloc((void **)&d, 8) != cudaSuccess) return KC_ERR_NOMEM;
    cudaMemset(d, 0, 8);
    kc::probe_sum_kernel<<<64, 256>>>(static_cast<const unsigned long long *>(d_ptr), n_words, d);
    cudaError_t e = cudaMemcpy(out, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { fprintf(stderr, "kc_debug_probe_read: %s\n", cudaGetErrorString(e)); return KC_ERR_CUDA; }
    return KC_OK;
}
