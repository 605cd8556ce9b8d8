// kc_fastq.cu -- raw FASTQ text -> packed fixed-length reads, on the device.
//
// Replaces the host parser FASTQFileReader::readData (FASTQFileReader.cpp:49-89: two
// getline calls per line, ~0.5 GB/s on one core) for well-formed input: four lines per
// record, the sequence is the line before the line that starts with '+', every sequence
// has the read length L taken from line 2 of the first file (FASTQFileReader.cpp:30-35).
// The chunk must start at a record boundary; the number of bytes consumed (whole records
// only) is reported so that the host can carry the tail into the next chunk.
//
// Anything else -- a record whose third line does not start with '+', a sequence whose
// length is not L -- sets a flag and nothing is trusted: the caller then falls back to the
// host chunker (host/FastqChunker.cpp), which implements the reference's general rule.
//
//   K1 count newlines per tile  ->  scan  ->  K2 write newline positions (ordered)
//   K3 one warp per record: validate, copy the L sequence bytes to reads[r*L ...)
#include "../../include/kc_api.h"
#include "kc_internal.h"

namespace kc {
namespace {

constexpr int kFqThreads = 256, kFqBytesPerThread = 64, kFqTile = kFqThreads * kFqBytesPerThread;

__device__ __forceinline__ uint32_t count_nl_16(uint4 v) {
    // bytes equal to '\n' (0x0A) in 16 bytes: x ^ 0x0A..0A has a zero byte there
    uint32_t c = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t x = w[i] ^ 0x0A0A0A0Au;
        c += ((x & 0xFFu) == 0) + ((x & 0xFF00u) == 0) + ((x & 0xFF0000u) == 0) + ((x & 0xFF000000u) == 0);
    }
    return c;
}

__global__ void __launch_bounds__(kFqThreads) nl_count_kernel(const uint8_t *__restrict__ text, uint64_t n,
                                                              uint32_t *__restrict__ tile_counts) {
    __shared__ uint32_t s_w[kFqThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kFqTile + (uint64_t)threadIdx.x * kFqBytesPerThread;
    uint32_t c = 0;
    if (base + kFqBytesPerThread <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(text + base);
#pragma unroll
        for (int i = 0; i < kFqBytesPerThread / 16; i++) c += count_nl_16(p[i]);
    } else {
        for (uint64_t i = base; i < n && i < base + kFqBytesPerThread; i++) c += text[i] == '\n';
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kFqThreads / 32; w++) t += s_w[w];
        tile_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kFqThreads) nl_write_kernel(const uint8_t *__restrict__ text, uint64_t n,
                                                              const uint32_t *__restrict__ tile_base,
                                                              uint32_t *__restrict__ nl_pos, uint32_t nl_cap) {
    __shared__ uint32_t s_w[kFqThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kFqTile + (uint64_t)tid * kFqBytesPerThread;
    uint8_t buf[kFqBytesPerThread];
    uint32_t c = 0;
    const bool full = base + kFqBytesPerThread <= n;
    if (full) {
        const uint4 *p = reinterpret_cast<const uint4 *>(text + base);
#pragma unroll
        for (int i = 0; i < kFqBytesPerThread / 16; i++) *reinterpret_cast<uint4 *>(buf + 16 * i) = p[i];
    } else {
        for (int i = 0; i < kFqBytesPerThread; i++) buf[i] = (base + i < n) ? text[base + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < kFqBytesPerThread; i++) c += buf[i] == '\n';
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t off = tile_base[blockIdx.x];
    for (uint32_t w = 0; w < warp; w++) off += s_w[w];
    uint32_t o = off + incl - c;
#pragma unroll
    for (int i = 0; i < kFqBytesPerThread; i++)
        if (buf[i] == '\n') { if (o < nl_cap) nl_pos[o] = (uint32_t)(base + i); o++; }
}

// flags: bit 0 = a record's third line does not start with '+', bit 1 = a sequence is not L long
__global__ void __launch_bounds__(256) fastq_pack_kernel(const uint8_t *__restrict__ text,
                                                         const uint32_t *__restrict__ nl_pos,
                                                         const uint32_t *__restrict__ n_lines_ptr, uint32_t nl_cap,
                                                         uint32_t L, uint64_t max_reads, uint8_t *__restrict__ reads,
                                                         unsigned long long *__restrict__ out /* [0] reads [1] consumed [2] flags */) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t n_lines = *n_lines_ptr;
    uint64_t n_rec = (n_lines < nl_cap ? n_lines : nl_cap) / 4;
    if (n_rec > max_reads) n_rec = max_reads;
    uint32_t bad = n_lines > nl_cap ? 1u : 0u;                // more newlines than any FASTQ of this size can have
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rec; r += warps) {
        const uint32_t s = nl_pos[4 * r] + 1;                 // sequence line: after the header's newline
        uint32_t e = nl_pos[4 * r + 1];
        if (e > s && text[e - 1] == '\r') e--;
        if (text[nl_pos[4 * r + 1] + 1] != '+') bad |= 1u;
        if (e - s != L) { bad |= 2u; continue; }
        uint8_t *dst = reads + r * L;
        for (uint32_t i = lane; i < L; i += 32) dst[i] = text[s + i];
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicOr(&out[2], (unsigned long long)bad);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = n_rec;
        out[1] = n_rec ? (unsigned long long)nl_pos[4 * n_rec - 1] + 1 : 0ull;
    }
}

// generic single-block exclusive scan (tile totals -> tile bases, total at [n])
__global__ void __launch_bounds__(1024) scan_u32_kernel(const uint32_t *__restrict__ in, uint32_t n,
                                                        uint32_t *__restrict__ out) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < n; t0 += 1024) {
        const uint32_t i = t0 + tid;
        const uint32_t v = i < n ? in[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t off = s_carry, tot = 0;
        for (uint32_t w = 0; w < 32; w++) { const uint32_t x = s_w[w]; if (w < warp) off += x; tot += x; }
        if (i < n) out[i] = off + incl - v;
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    if (tid == 0) out[n] = s_carry;
}

}  // namespace

uint64_t fastq_workspace_bytes(uint64_t n_bytes) {
    const uint64_t tiles = div_up(n_bytes ? n_bytes : 1, (uint64_t)kFqTile);
    // tile counts + tile bases (+1) + newline positions (at most one per 2 bytes of well-formed text; sized n/2 + slack)
    return (2 * tiles + 16) * 4 + (n_bytes / 2 + 1024) * 4 + 64;
}

// d_text: n_bytes of FASTQ (16-byte aligned, < 4 GiB). d_out: [0] reads, [1] consumed bytes, [2] flags.
cudaError_t fastq_parse(const void *d_text, uint64_t n_bytes, uint32_t L, void *d_reads, uint64_t max_reads,
                        unsigned long long *d_out, void *ws, cudaStream_t s, int *n_launches) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_out, 0, 24, s)) != cudaSuccess) return e;
    if (n_bytes == 0) return cudaSuccess;
    if (n_bytes >= (1ull << 32)) return cudaErrorInvalidValue;
    const uint32_t tiles = (uint32_t)div_up(n_bytes, (uint64_t)kFqTile);
    uint32_t *tile_counts = static_cast<uint32_t *>(ws);
    uint32_t *tile_base = tile_counts + tiles + 8;
    uint32_t *nl_pos = tile_base + tiles + 8;
    const uint8_t *text = static_cast<const uint8_t *>(d_text);
    nl_count_kernel<<<tiles, kFqThreads, 0, s>>>(text, n_bytes, tile_counts);
    scan_u32_kernel<<<1, 1024, 0, s>>>(tile_counts, tiles, tile_base);
    const uint32_t nl_cap = (uint32_t)(n_bytes / 2 + 1024);
    nl_write_kernel<<<tiles, kFqThreads, 0, s>>>(text, n_bytes, tile_base, nl_pos, nl_cap);
    fastq_pack_kernel<<<(uint32_t)sm_count() * 8, 256, 0, s>>>(text, nl_pos, tile_base + tiles, nl_cap, L, max_reads,
                                              static_cast<uint8_t *>(d_reads), d_out);
    if (n_launches) *n_launches += 4;
    return cudaGetLastError();
}

}  // namespace kc
