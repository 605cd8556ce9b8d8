// kc_main.cpp -- command line with the reference's flags (main.cpp:25-89).
//
//   kmer_counter_b200 kmerLength=31 inputFileLocation=<dir> outputFile=<file> [gpuMemoryLimit=..]
//                     [tempFileLocation=..] [noOfMergersAtOnce=..] [noOfMergeThreads=..]
//                     [method=auto|sort|hash] [compat=ref|strict] [device=N] [keepRuns=1] [parser=gpu|host]
//                     [runBudget=BYTES] [gpus=N] [mode=auto|accumulate|runs] [expectedReads=N]
//   kmer_counter_b200 print <record file> <ignored> <k>       (KMerPrinter, main.cpp:78-82)
//
// What KMerCounter::Start does (KMerCounter.cpp:108-191), on the B200 path: read fixed
// length reads into pinned chunks, count each chunk on the GPU into a sorted run
// (double-buffered: the next chunk is parsed while the previous one is copied and
// counted), merge the runs on the GPU, write the sorted unique records.
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/kc_api.h"
#include "FastqChunker.h"
#include "Options.h"
#include "RunMerger.h"

// KMerPrinter::print (KMerPrinter.cpp:35-91): every word as 32 letters, then " count".
static int print_records(const char *path, uint64_t k) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return 1; }
    const uint32_t W = kc_key_words((uint32_t)k), S = kc_record_size((uint32_t)k);
    std::vector<unsigned char> buf((size_t)S * 10000);
    std::string line;
    size_t n;
    while ((n = fread(buf.data(), 1, buf.size(), f)) >= S) {
        for (size_t off = 0; off + S <= n; off += S) {
            line.clear();
            for (uint32_t w = 0; w < W; w++) {
                uint64_t v;
                memcpy(&v, &buf[off + 8 * w], 8);
                for (int b = 0; b < 32; b++) line.push_back("ACGT"[(v >> (62 - 2 * b)) & 3]);
            }
            uint32_t c;
            memcpy(&c, &buf[off + 8 * W], 4);
            printf("%s %u\n", line.c_str(), c);
        }
    }
    fclose(f);
    return 0;
}

#include <sys/stat.h>

// mode=accumulate: every chunk is only packed into super-window records (kc_accum_*); the count
// happens once, over everything -- on one GPU, or on gpus=N with the two exchanges fused into the
// counting kernels (kc_xchg_*). Returns 1 if this input / shape is not for this mode (the caller
// then counts chunk by chunk into runs), 0 on success, -1 on error.
static int run_accumulate(const Options &opt, FastqChunker &reader, int64_t L, int64_t chunk, uint32_t method,
                          uint64_t *n_records_out) {
    const int G = opt.gpus;
    // reads to plan for: a FASTQ record is 2L + header + '+' line + newlines
    uint64_t expected = opt.expectedReads > 0 ? (uint64_t)opt.expectedReads : 0;
    if (!expected) {
        uint64_t bytes = 0;
        for (const std::string &f : reader.files()) {
            struct stat st;
            if (stat(f.c_str(), &st) == 0) bytes += (uint64_t)st.st_size;
        }
        expected = bytes / (uint64_t)(2 * L + 6) + 1024;
    }
    const uint64_t per_gpu = (expected + G - 1) / G + (uint64_t)(chunk / L) + 16;
    std::vector<kc_ctx *> ctx(G, nullptr);
    auto destroy_all = [&]() { for (kc_ctx *c : ctx) if (c) kc_destroy(c); };
    for (int g = 0; g < G; g++) {
        kc_config cfg = {};
        cfg.struct_size = sizeof cfg;
        cfg.k = (uint32_t)opt.kmerLength;
        cfg.read_len = (uint32_t)L;
        cfg.device = getenv("KC_CLI_SAME_DEVICE") ? opt.device : opt.device + g;   // (test knob: all ranks on one device)
        cfg.flags = opt.compat == "strict" ? KC_COMPAT_STRICT : KC_COMPAT_REF;
        cfg.method = method;
        cfg.n_slots = 2;
        cfg.max_chunk_bytes = (uint64_t)chunk;
        if (kc_create(&cfg, &ctx[g]) != KC_OK) { fprintf(stderr, "kc_create(device %d): %s\n", cfg.device, kc_last_error(nullptr)); destroy_all(); return -1; }
        const int rc = G > 1 ? kc_xchg_begin(ctx[g], (uint32_t)g, (uint32_t)G, per_gpu) : kc_accum_begin(ctx[g], per_gpu);
        if (rc == KC_ERR_ARG) { destroy_all(); return 1; }            // k or read length not for the super-window path
        if (rc != KC_OK) { fprintf(stderr, "accumulator: %s\n", kc_last_error(ctx[g])); destroy_all(); return -1; }
    }
    int rc = KC_OK;
    uint64_t n_chunks = 0;
    std::vector<bool> busy(2 * G, false);
    auto settle = [&](uint64_t i) -> int {                          // (gpu, slot) of chunk i is free again
        const int g = (int)(i % G), sl = (int)((i / G) % 2);
        if (!busy[2 * g + sl]) return KC_OK;
        busy[2 * g + sl] = false;
        return kc_accum_wait(ctx[g], (uint32_t)sl);
    };
    bool refused = false;
    if (opt.parser != "host") {
        // FASTQ text goes to the GPUs as it is and is parsed there
        const uint64_t raw_cap = (uint64_t)chunk * 23 / 10 + (1u << 20);
        std::vector<void *> raw(2 * G, nullptr);
        for (int i = 0; i < 2 * G && rc == KC_OK; i++) rc = kc_host_alloc(ctx[i / 2], raw_cap, &raw[i]);
        uint64_t carry = 0;
        std::vector<char> carry_buf;
        for (const std::string &path : reader.files()) {
            if (rc != KC_OK || refused) break;
            FILE *f = fopen(path.c_str(), "rb");
            if (!f) continue;
            carry = 0;
            bool eof = false;
            while (rc == KC_OK && !refused) {
                const int g = (int)(n_chunks % G), sl = (int)((n_chunks / G) % 2);
                if ((rc = settle(n_chunks)) != KC_OK) break;
                char *buf = static_cast<char *>(raw[2 * g + sl]);
                if (carry) memcpy(buf, carry_buf.data(), carry);
                uint64_t total = carry;
                if (!eof) {
                    const size_t n = fread(buf + carry, 1, raw_cap - carry - 1, f);
                    total += n;
                    eof = n < raw_cap - carry - 1;
                }
                if (eof && total > 0 && buf[total - 1] != '\n') buf[total++] = '\n';
                if (total == 0) break;
                uint64_t used = 0;
                uint32_t flags = 0;
                if ((rc = kc_accum_submit_fastq(ctx[g], (uint32_t)sl, buf, total, &used, &flags)) != KC_OK) break;
                if (flags) { refused = true; break; }
                if (used == 0) {
                    if (eof) break;                                 // a truncated last record: ignored
                    rc = KC_ERR_CAPACITY;
                    fprintf(stderr, "a FASTQ record does not fit %llu bytes\n", (unsigned long long)raw_cap);
                    break;
                }
                busy[2 * g + sl] = true;
                n_chunks++;
                carry = total - used;
                carry_buf.assign(buf + used, buf + used + carry);
                if (eof && carry == 0) break;
            }
            fclose(f);
        }
        for (int g = 0; g < G; g++)
            for (int sl = 0; sl < 2; sl++)
                if (busy[2 * g + sl]) { kc_accum_wait(ctx[g], (uint32_t)sl); busy[2 * g + sl] = false; }
        for (int i = 0; i < 2 * G; i++)
            if (raw[i]) kc_host_free(ctx[i / 2], raw[i]);
        if (refused && rc == KC_OK) {
            // what has been accumulated is dropped with the contexts; the caller starts over with the host parser
            fprintf(stderr, "input is not plain 4-line fixed-length FASTQ: parsing on the host instead\n");
            destroy_all();
            Options o2 = opt;
            o2.parser = "host";
            FastqChunker again(opt.inputFileDirectory);
            return run_accumulate(o2, again, L, chunk, method, n_records_out);
        }
    } else {
        while (rc == KC_OK) {
            const int g = (int)(n_chunks % G), sl = (int)((n_chunks / G) % 2);
            if ((rc = settle(n_chunks)) != KC_OK) break;
            void *buf = nullptr;
            uint64_t cap = 0;
            if ((rc = kc_slot_buffer(ctx[g], (uint32_t)sl, &buf, &cap)) != KC_OK) break;
            const int64_t n = reader.read(static_cast<char *>(buf), (int64_t)cap);
            if (n == 0) break;
            if ((rc = kc_accum_submit(ctx[g], (uint32_t)sl, (uint64_t)n)) != KC_OK) break;
            busy[2 * g + sl] = true;
            n_chunks++;
        }
        for (int g = 0; g < G; g++)
            for (int sl = 0; sl < 2; sl++)
                if (busy[2 * g + sl]) { kc_accum_wait(ctx[g], (uint32_t)sl); busy[2 * g + sl] = false; }
    }
    // the count, and the artefact: rank r holds the r-th key range, the file is their concatenation
    std::vector<kc_run *> runs(G, nullptr);
    if (rc == KC_OK) rc = G > 1 ? kc_xchg_run_all(ctx.data(), (uint32_t)G, runs.data()) : kc_accum_flush(ctx[0], &runs[0]);
    uint64_t n_records = 0;
    for (int g = 0; g < G && rc == KC_OK; g++) {
        rc = kc_run_write(ctx[g], runs[g], opt.outputFile.c_str(), g == 0 ? 0 : 1);      // truncates, then appends
        n_records += kc_run_records(runs[g]);
    }
    if (rc != KC_OK) {
        for (int g = 0; g < G; g++) if (kc_last_error(ctx[g])[0]) { fprintf(stderr, "kmer_counter_b200: %s\n", kc_last_error(ctx[g])); break; }
    } else {
        kc_stats st;
        kc_stats_get(ctx[0], &st);
        for (int g = 1; g < G; g++) { kc_stats sg; kc_stats_get(ctx[g], &sg); st.reads += sg.reads; }
        fprintf(stderr, "mode=accumulate gpus=%d parser=%s reads=%llu skipped=%llu chunks=%llu records=%llu\n", G,
                opt.parser != "host" ? "gpu" : "host", (unsigned long long)(opt.parser != "host" ? st.reads : reader.totalReads()),
                (unsigned long long)reader.skippedReads(), (unsigned long long)n_chunks, (unsigned long long)n_records);
    }
    for (int g = 0; g < G; g++) if (runs[g]) kc_run_free(ctx[g], runs[g]);
    destroy_all();
    if (n_records_out) *n_records_out = n_records;
    return rc == KC_OK ? 0 : -1;
}

static uint32_t method_of(const std::string &m) {
    if (m == "sort") return KC_COUNT_SORT;
    if (m == "hash") return KC_COUNT_HASH;
    if (m == "hash_global") return KC_COUNT_HASH_GLOBAL;
    if (m == "super") return KC_COUNT_SUPER;
    if (m == "place") return KC_COUNT_PLACE;
    return KC_COUNT_AUTO;
}

int main(int argc, char **argv) {
    if (argc == 5 && strncmp(argv[1], "print", 5) == 0) return print_records(argv[2], strtoull(argv[4], nullptr, 10));
    Options opt = Options::parse(argc, argv);
    if (opt.inputFileDirectory.empty()) {
        fprintf(stderr, "usage: %s kmerLength=K inputFileLocation=DIR outputFile=FILE [gpuMemoryLimit=BYTES] ...\n", argv[0]);
        return 2;
    }
    FastqChunker reader(opt.inputFileDirectory);
    if (!reader.ok()) { fprintf(stderr, "no FASTQ input in %s\n", opt.inputFileDirectory.c_str()); return 1; }
    const int64_t L = reader.getLineLength();
    int64_t chunk = opt.chunkSize(L);
    const int64_t max_chunk = (int64_t)((((1ull << 30) - 1) / (uint64_t)(L - opt.kmerLength + 1)) * (uint64_t)L);
    if (chunk > max_chunk) chunk = max_chunk - max_chunk % (16 * L);

    // one count over everything (default where the super-window path applies), unless the reference's
    // run-per-chunk shape is asked for (mode=runs, or its options keepRuns / runBudget)
    const bool want_runs = opt.mode == "runs" || opt.keepRuns || opt.runBudget > 0 ||
                           (opt.mode == "auto" && (method_of(opt.method) != KC_COUNT_AUTO && method_of(opt.method) != KC_COUNT_SUPER));
    if (!want_runs || opt.gpus > 1) {
        uint64_t n_rec = 0;
        const int r = run_accumulate(opt, reader, L, chunk, method_of(opt.method), &n_rec);
        if (r == 0) { fprintf(stderr, "records=%" PRIu64 " -> %s\n", n_rec, opt.outputFile.c_str()); return 0; }
        if (r < 0) return 1;
        if (opt.gpus > 1) { fprintf(stderr, "gpus=%d needs k <= 64 with windows of at least 22 bases\n", opt.gpus); return 1; }
        // not a shape for the accumulator: fall through to runs
    }

    kc_config cfg = {};
    cfg.struct_size = sizeof cfg;
    cfg.k = (uint32_t)opt.kmerLength;
    cfg.read_len = (uint32_t)L;
    cfg.device = opt.device;
    cfg.flags = opt.compat == "strict" ? KC_COMPAT_STRICT : KC_COMPAT_REF;
    cfg.method = method_of(opt.method);
    cfg.n_slots = 2;
    cfg.max_chunk_bytes = (uint64_t)chunk;
    kc_ctx *ctx = nullptr;
    if (kc_create(&cfg, &ctx) != KC_OK) { fprintf(stderr, "kc_create: %s\n", kc_last_error(nullptr)); return 1; }

    int rc = KC_OK;
    uint32_t chunk_id = 0;
    uint64_t gpu_parsed_bytes = 0;
    uint64_t n_records = 0;
    // attempt 0: FASTQ text goes to the GPU as it is and is parsed there (kc_submit_fastq);
    // attempt 1: the host chunker, for input the device parser refuses (or parser=host)
    for (int attempt = opt.parser == "host" ? 1 : 0; attempt < 2; attempt++) {
        RunMerger merger(ctx, opt.noOfMergersAtOnce, (uint32_t)opt.kmerLength, opt.runBudget > 0 ? (uint64_t)opt.runBudget : 0,
                         opt.noOfMergeThreads);
        bool busy[2] = {false, false};
        bool refused = false;
        chunk_id = 0;
        auto collect = [&](uint32_t slot) -> int {
            kc_run *run = nullptr;
            int r = kc_wait(ctx, slot, &run);
            busy[slot] = false;
            if (r != KC_OK) return r;
            if (opt.keepRuns) {                                 // FileDump::dumpKmersToFile naming (FileDump.cpp:51-58)
                std::string path = opt.tempFileLocation + "/" + std::to_string(++chunk_id);
                if ((r = kc_run_write(ctx, run, path.c_str(), 0)) != KC_OK) return r;
            }
            return merger.AddRun(run);
        };
        if (attempt == 0) {
            const uint64_t raw_cap = (uint64_t)chunk * 23 / 10 + (1u << 20);
            void *raw[2] = {nullptr, nullptr};
            if (kc_host_alloc(ctx, raw_cap, &raw[0]) != KC_OK || kc_host_alloc(ctx, raw_cap, &raw[1]) != KC_OK) rc = KC_ERR_NOMEM;
            uint32_t slot = 0;
            for (const std::string &path : reader.files()) {
                if (rc != KC_OK || refused) break;
                FILE *f = fopen(path.c_str(), "rb");
                if (!f) continue;                               // the reference ignores unreadable files too
                uint64_t carry = 0;
                bool eof = false;
                while (rc == KC_OK && !refused) {
                    char *buf = static_cast<char *>(raw[slot]);
                    uint64_t total = carry;
                    if (!eof) {
                        const size_t n = fread(buf + carry, 1, raw_cap - carry - 1, f);
                        total += n;
                        eof = n < raw_cap - carry - 1;
                    }
                    if (eof && total > 0 && buf[total - 1] != '\n') buf[total++] = '\n';
                    if (total == 0) break;
                    if (busy[slot] && (rc = collect(slot)) != KC_OK) break;
                    uint64_t used = 0;
                    uint32_t flags = 0;
                    if ((rc = kc_submit_fastq(ctx, slot, buf, total, &used, &flags)) != KC_OK) break;
                    if (flags) { refused = true; break; }
                    if (used == 0) {
                        // fewer than four lines are left: at end of file that is a truncated record, ignored
                        if (eof) { kc_run *empty = nullptr; rc = kc_wait(ctx, slot, &empty); if (empty) kc_run_free(ctx, empty); break; }
                        rc = KC_ERR_CAPACITY;
                        fprintf(stderr, "a FASTQ record does not fit %llu bytes\n", (unsigned long long)raw_cap);
                        break;
                    }
                    busy[slot] = true;
                    gpu_parsed_bytes += used;
                    carry = total - used;
                    memcpy(raw[slot ^ 1], buf + used, carry);   // the tail opens the next block
                    slot ^= 1;
                    if (eof && carry == 0) break;
                }
                fclose(f);
            }
            for (uint32_t sl = 0; sl < 2 && rc == KC_OK; sl++)
                if (busy[sl]) rc = collect(sl);
            if (raw[0]) kc_host_free(ctx, raw[0]);
            if (raw[1]) kc_host_free(ctx, raw[1]);
            if (refused && rc == KC_OK) {
                for (uint32_t sl = 0; sl < 2; sl++)
                    if (busy[sl]) collect(sl);
                fprintf(stderr, "input is not plain 4-line fixed-length FASTQ: parsing on the host instead\n");
                gpu_parsed_bytes = 0;
                continue;                                       // merger (and its runs) is dropped, start over
            }
        } else {
            for (uint32_t slot = 0; rc == KC_OK; slot ^= 1) {
                if (busy[slot]) rc = collect(slot);             // the other slot keeps the GPU busy meanwhile
                if (rc != KC_OK) break;
                void *buf = nullptr;
                uint64_t cap = 0;
                if ((rc = kc_slot_buffer(ctx, slot, &buf, &cap)) != KC_OK) break;
                const int64_t n = reader.read(static_cast<char *>(buf), (int64_t)cap);
                if (n == 0) break;
                if ((rc = kc_submit(ctx, slot, (uint64_t)n)) != KC_OK) break;
                busy[slot] = true;
            }
            for (uint32_t slot = 0; slot < 2 && rc == KC_OK; slot++)
                if (busy[slot]) rc = collect(slot);
        }
        if (rc == KC_OK) rc = merger.Finish(opt.outputFile.c_str(), &n_records);     // truncates (KMerFileMerger.cpp:129 appends)
        if (rc == KC_OK) {
            kc_stats st;
            kc_stats_get(ctx, &st);
            fprintf(stderr, "parser=%s reads=%" PRIu64 " skipped=%" PRIu64 " kmers=%" PRIu64 " chunks=%" PRIu64
                            " merges=%" PRIu64 " spills=%" PRIu64 " ranges=%" PRIu64 "\n",
                    attempt == 0 ? "gpu" : "host", attempt == 0 ? st.reads : reader.totalReads(), reader.skippedReads(),
                    st.kmers_valid, st.chunks, merger.merges(), merger.spills(), merger.ranges());
        }
        break;
    }
    if (rc != KC_OK) fprintf(stderr, "kmer_counter_b200: %s\n", kc_last_error(ctx));
    else fprintf(stderr, "records=%" PRIu64 " -> %s\n", n_records, opt.outputFile.c_str());
    kc_destroy(ctx);
    return rc == KC_OK ? 0 : 1;
}
