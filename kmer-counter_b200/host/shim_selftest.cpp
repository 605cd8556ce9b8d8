// shim_selftest.cpp -- drives the reference-shaped seam the way KMerCounter::Start and
// dispatchWork do (KMerCounter.cpp:51-89,108-161): PrepareGPU once, processKMers per chunk
// from several host threads with one GPUStream each, FreeGPU; concatenated runs go to a file
// as <uint64 n_bytes><records> per chunk.   usage: shim_selftest <packed reads> <L> <k> <chunk_reads> <out>
#include <stdio.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "GPUHandler.h"

int main(int argc, char **argv) {
    if (argc != 6) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> reads((size_t)size);
    if (fread(reads.data(), 1, (size_t)size, f) != (size_t)size) return 1;
    fclose(f);
    const int64_t L = atoll(argv[2]), k = atoll(argv[3]), chunk_reads = atoll(argv[4]);
    const int64_t chunk_bytes = chunk_reads * L;
    const uint32_t n_streams = 4;
    GPUStream **streams = PrepareGPU(n_streams, (uint64_t)chunk_bytes, (uint64_t)L, k);
    if (!streams) return 1;
    const int64_t n_chunks = (size + chunk_bytes - 1) / chunk_bytes;
    std::vector<std::vector<char>> runs((size_t)n_chunks);
    std::vector<std::thread> th;
    for (uint32_t s = 0; s < n_streams; s++) {
        th.emplace_back([&, s]() {
            for (int64_t c = s; c < n_chunks; c += n_streams) {
                const int64_t off = c * chunk_bytes;
                const int64_t n = size - off < chunk_bytes ? size - off : chunk_bytes;
                int64_t bytes = processKMers(streams[s], reads.data() + off, k, n, L, (uint32_t)(c + 1));
                if (bytes < 0) exit(3);
                runs[(size_t)c].assign(streams[s]->_h_output, streams[s]->_h_output + bytes);
            }
        });
    }
    for (auto &t : th) t.join();
    FreeGPU(streams, n_streams);
    for (uint32_t s = 0; s < n_streams; s++) {            // as KMerCounter.cpp:153-161 does
        for (char *line : streams[s]->_kmer_db) delete[] line;
        delete streams[s];
    }
    delete[] streams;
    FILE *o = fopen(argv[5], "wb");
    if (!o) return 1;
    for (auto &r : runs) {
        uint64_t nb = r.size();
        fwrite(&nb, 8, 1, o);
        fwrite(r.data(), 1, r.size(), o);
    }
    fclose(o);
    return 0;
}
