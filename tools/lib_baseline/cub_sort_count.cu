// cub_sort_count.cu -- the LIBRARY baseline of SURVEY 8(d) / K3: what the reference's sortKmers +
// reduceKMers (GPUHandler.cu:300-360: thrust::sort, then adjacent-equal reduce) cost on this GPU when
// written with today's CUB primitives, on the key volume of configs[1] (7e8 64-bit keys, ~1.29e8
// distinct). Not product code and not linked into libkc_b200.so: a number to hold the hand-written
// path against (profiles/r2/lib_baseline_cub.json).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cub_sort_count cub_sort_count.cu
//   ./cub_sort_count [n_keys] [n_distinct] [key_bits]
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
// occurrence i is a copy of distinct key (hash(i) mod n_distinct): keys repeat like a 10x coverage does
__global__ void fill(uint64_t *k, uint64_t n, uint64_t n_distinct, int key_bits) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        k[i] = mix64(mix64(i) % n_distinct + 1) >> (64 - key_bits) << (64 - key_bits);
}

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 700000000ull;
    const uint64_t nd = argc > 2 ? strtoull(argv[2], 0, 10) : 128845341ull;
    const int key_bits = argc > 3 ? atoi(argv[3]) : 62;          // k=31: 62 significant bits, left-aligned
    uint64_t *in, *out, *uniq; uint32_t *cnt; uint64_t *n_runs;
    CK(cudaMalloc(&in, n * 8)); CK(cudaMalloc(&out, n * 8));
    CK(cudaMalloc(&uniq, (nd + 1024) * 8)); CK(cudaMalloc(&cnt, (nd + 1024) * 4)); CK(cudaMalloc(&n_runs, 8));
    size_t t1 = 0, t2 = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, t1, in, out, n, 64 - key_bits, 64));
    CK(cub::DeviceRunLengthEncode::Encode(nullptr, t2, out, uniq, cnt, n_runs, n));
    void *tmp; CK(cudaMalloc(&tmp, t1 > t2 ? t1 : t2));
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float best_sort = 1e30f, best_rle = 1e30f;
    uint64_t runs = 0;
    for (int it = 0; it < 6; it++) {
        fill<<<148 * 8, 256>>>(in, n, nd, key_bits);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        CK(cub::DeviceRadixSort::SortKeys(tmp, t1, in, out, n, 64 - key_bits, 64));
        cudaEventRecord(e1);
        CK(cub::DeviceRunLengthEncode::Encode(tmp, t2, out, uniq, cnt, n_runs, n));
        cudaEventRecord(e2);
        CK(cudaDeviceSynchronize());
        float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
        if (it >= 2) { best_sort = a < best_sort ? a : best_sort; best_rle = b < best_rle ? b : best_rle; }
        CK(cudaMemcpy(&runs, n_runs, 8, cudaMemcpyDeviceToHost));
    }
    printf("{\"baseline\": \"CUB DeviceRadixSort::SortKeys + DeviceRunLengthEncode::Encode (CUDA 12.9 toolkit)\", \"keys\": %llu, "
           "\"key_bits\": %d, \"distinct\": %llu, \"sort_ms\": %.3f, \"rle_ms\": %.3f, \"total_ms\": %.3f, "
           "\"sort_gbs_algorithmic\": %.1f, \"note\": \"keys already in HBM; extraction (a key write of 8 B per occurrence) not included\"}\n",
           (unsigned long long)n, key_bits, (unsigned long long)runs, best_sort, best_rle, best_sort + best_rle,
           (double)n * 16.0 * ((key_bits + 7) / 8) / (best_sort * 1e-3) / 1e9);
    return 0;
}
