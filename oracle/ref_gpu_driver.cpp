// oracle/ref_gpu_driver.cpp -- times the REFERENCE'S OWN GPU seam on this GPU.
//
// TEST / BENCH INFRASTRUCTURE ONLY (SURVEY.md F10, 8(d) "reference-on-B200" row). This file is
// ours; GPUHandler.cu and FileDump.cpp are compiled UNMODIFIED from /root/reference where they lie
// (oracle/build_ref.sh, nvcc for sm_100a) into oracle/_ref/ref_gpu. It drives the seam the way
// KMerCounter::dispatchWork does (KMerCounter.cpp:61-82): PrepareGPU once (GPUHandler.cu:479-508),
// then processKMers per chunk (GPUHandler.cu:397-477: H2D, bitEncode + extractKMers with a stream
// sync after each, D2H of the 8.4x record block into pageable memory, reduceKMers on the host).
// What HEAD does NEXT with every record -- the TBB concurrent_hash_map accumulate -- is not part of
// the seam and not timed here, so the figure is an upper bound on the reference's GPU pipeline.
//
//   ref_gpu <packed_reads.bin> <read_len> <k> <chunk_reads>
// prints one JSON line on stderr (processKMers prints progress on stdout; it is sent to /dev/null).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <vector>

#include "GPUHandler.h"

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: ref_gpu reads.bin read_len k chunk_reads\n"); return 2; }
    const int64_t L = atoll(argv[2]), k = atoll(argv[3]), chunk_reads = atoll(argv[4]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    const int64_t bytes = ftell(f) / L * L;
    fseek(f, 0, SEEK_SET);
    std::vector<char> reads((size_t)bytes);
    if (fread(reads.data(), 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);
    if (!freopen("/dev/null", "w", stdout)) return 2;
    const int64_t chunk_bytes = chunk_reads * L;
    FileDump dump;
    GPUStream **streams = PrepareGPU(1, (uint64_t)chunk_bytes, (uint64_t)L, k);
    // one untimed chunk: context creation, first-launch costs
    processKMers(streams[0], reads.data(), k, bytes < chunk_bytes ? bytes : chunk_bytes, L, 0, dump);
    const auto t0 = std::chrono::steady_clock::now();
    int64_t out_bytes = 0, chunks = 0;
    for (int64_t off = 0; off < bytes; off += chunk_bytes, chunks++) {
        const int64_t n = bytes - off < chunk_bytes ? bytes - off : chunk_bytes;
        out_bytes += processKMers(streams[0], reads.data() + off, k, n, L, (uint32_t)chunks, dump);
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int64_t n_reads = bytes / L, kmers = n_reads * (L - k + 1);
    fprintf(stderr, "{\"reads\": %lld, \"read_len\": %lld, \"k\": %lld, \"chunk_reads\": %lld, \"chunks\": %lld, \"kmers\": %lld, "
                    "\"seconds\": %.6f, \"kmers_per_s\": %.1f, \"reduced_bytes\": %lld}\n",
            (long long)n_reads, (long long)L, (long long)k, (long long)chunk_reads, (long long)chunks, (long long)kmers, s,
            (double)kmers / s, (long long)out_bytes);
    FreeGPU(streams, 1);
    return 0;
}
