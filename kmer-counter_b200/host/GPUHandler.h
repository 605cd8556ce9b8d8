// GPUHandler.h -- the reference's seam (GPUHandler.h:36-65) re-provided over libkc_b200.
//
// A caller written against the reference (KMerCounter::Start / dispatchWork,
// KMerCounter.cpp:51-89,108-161) keeps its three calls and the GPUStream fields it
// touches (_h_output, _id, _kmer_db, _kmer_db_line_length, _kmer_db_line_index -- checked by
// tests/test_reference_seam.py, which compiles the reference's KMerCounter.cpp against this
// header); everything behind them is the B200 path.  processKMers
// returns the chunk's SORTED, key-unique records (the result the reference gets with its
// commented-out sort step enabled, GPUHandler.cu:455-458) in gpuStream->_h_output.
#pragma once

#include <stdint.h>

#include <list>

struct kc_ctx;
class FileDump;   // accepted for signature compatibility; runs are dumped via kc_run_write

struct GPUStream {
    uint32_t _id;            // 1-based, as PrepareGPU numbers them (GPUHandler.cu:500)
    char *_h_output;         // pinned; sized calculateOutputSize(inputSize, ...) worst case
    // key arenas the caller's accumulate loop fills and frees (KMerCounter.cpp:61-82,153-161); set up
    // as the reference's constructor does (GPUHandler.h:36-47): one 250 MiB line, index 0
    std::list<char *> _kmer_db;
    uint64_t _kmer_db_line_length;
    uint64_t _kmer_db_line_index;
    // ours
    uint64_t _h_output_capacity;
    kc_ctx *_ctx;            // shared by the streams of one PrepareGPU call
    uint32_t _slot;
};

GPUStream **PrepareGPU(uint32_t streamCount, uint64_t inputSize, uint64_t lineLength, int64_t kmerLength);
void FreeGPU(GPUStream **streams, uint32_t streamCount);
int64_t processKMers(GPUStream *gpuStream, const char *input, int64_t kmerLength, int64_t inputSize,
                     int64_t lineLength, uint32_t readId, FileDump &fileDump);
// same, for callers that have no FileDump object at hand
int64_t processKMers(GPUStream *gpuStream, const char *input, int64_t kmerLength, int64_t inputSize,
                     int64_t lineLength, uint32_t readId);
