#include "GPUHandler.h"

#include <stdio.h>
#include <stdlib.h>

#include "../../include/kc_api.h"

// One kc_ctx per PrepareGPU call, one slot per GPUStream (the reference allocates three
// device buffers and a stream per GPUStream, GPUHandler.cu:479-509).
GPUStream **PrepareGPU(uint32_t streamCount, uint64_t inputSize, uint64_t lineLength, int64_t kmerLength) {
    kc_config cfg = {};
    cfg.struct_size = sizeof cfg;
    cfg.k = (uint32_t)kmerLength;
    cfg.read_len = (uint32_t)lineLength;
    cfg.device = 0;                                  // the reference uses the implicit device 0
    cfg.n_slots = streamCount;
    cfg.max_chunk_bytes = inputSize;
    kc_ctx *ctx = nullptr;
    if (kc_create(&cfg, &ctx) != KC_OK) {
        fprintf(stderr, "PrepareGPU: %s\n", kc_last_error(nullptr));
        return nullptr;                              // the reference exit()s here (GPUHandler.h:27-34); we do not
    }
    const uint64_t cap = kc_output_size(inputSize, (uint32_t)lineLength, (uint32_t)kmerLength) +
                         kc_record_size((uint32_t)kmerLength);
    GPUStream **streams = new GPUStream *[streamCount];
    for (uint32_t i = 0; i < streamCount; i++) {
        GPUStream *s = new GPUStream();
        s->_id = i + 1;
        s->_kmer_db_line_length = 250ull * 1024 * 1024;
        s->_kmer_db.push_front(new char[s->_kmer_db_line_length]);       // untouched until the caller writes keys into it
        s->_kmer_db_line_index = 0;
        s->_ctx = ctx;
        s->_slot = i;
        s->_h_output_capacity = cap;
        void *p = nullptr;
        if (kc_host_alloc(ctx, cap, &p) != KC_OK) {
            fprintf(stderr, "PrepareGPU: %s\n", kc_last_error(ctx));
            p = nullptr;
        }
        s->_h_output = static_cast<char *>(p);
        streams[i] = s;
    }
    return streams;
}

// Like the reference, FreeGPU releases the buffers; the caller deletes the GPUStream
// objects and the array (KMerCounter.cpp:153-161).
void FreeGPU(GPUStream **streams, uint32_t streamCount) {
    if (!streams || streamCount == 0) return;
    kc_ctx *ctx = streams[0]->_ctx;
    for (uint32_t i = 0; i < streamCount; i++) {
        if (streams[i]->_h_output) kc_host_free(ctx, streams[i]->_h_output);
        streams[i]->_h_output = nullptr;
        streams[i]->_ctx = nullptr;
    }
    kc_destroy(ctx);
}

int64_t processKMers(GPUStream *gpuStream, const char *input, int64_t kmerLength, int64_t inputSize,
                     int64_t lineLength, uint32_t readId) {
    (void)kmerLength; (void)lineLength; (void)readId;      // fixed at PrepareGPU, like the reference's buffers
    uint64_t bytes = 0;
    int rc = kc_process_chunk(gpuStream->_ctx, gpuStream->_slot, input, (uint64_t)inputSize, gpuStream->_h_output,
                              gpuStream->_h_output_capacity, &bytes);
    if (rc != KC_OK) {
        fprintf(stderr, "processKMers: %s\n", kc_last_error(gpuStream->_ctx));
        return -1;
    }
    return (int64_t)bytes;
}

int64_t processKMers(GPUStream *gpuStream, const char *input, int64_t kmerLength, int64_t inputSize,
                     int64_t lineLength, uint32_t readId, FileDump &) {
    return processKMers(gpuStream, input, kmerLength, inputSize, lineLength, readId);
}
