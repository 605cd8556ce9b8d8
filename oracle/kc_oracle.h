/*
 * kc_oracle.h -- CPU restatement of the reference's k-mer counting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Parity pinning: the reference has no tests, golden vectors or fixtures of
 * its own.  This restatement is pinned instead against
 * (1) the known-answer vectors of SURVEY.md Appendix A.4 (tests/golden/kat.json),
 * and (2) oracle/_ref -- the reference's own sources compiled where they lie
 * (oracle/build_ref.sh) -- on seeded inputs (tests/test_oracle.py).
 *
 * All citations are file:line into /root/reference.
 */
#ifndef KC_ORACLE_H
#define KC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* words per key and bytes per packed record (GPUHandler.cu:235-245, KMerSizes.h:10-28) */
uint32_t kco_words(uint32_t k);
uint32_t kco_record_size(uint32_t k);
/* R * (L-k+1) * S (GPUHandler.cu:235-245) */
uint64_t kco_output_size(uint64_t input_size, uint32_t L, uint32_t k);

/* bitEncode for one read (GPUHandler.cu:10-111).  words: ceil(L/32) entries,
 * filter: ceil(L/64) entries.  Returns 0, or -1 if L is a case the reference
 * corrupts (L % 32 == 0). */
int kco_encode_read(const char *read, uint32_t L, uint64_t *words, uint64_t *filter);

/* extractKMers for one read (GPUHandler.cu:113-233).  out must hold
 * (L-k+1)*S zero-initialised bytes (GPUHandler.cu:408).  Returns the number of
 * records written (the remaining slots stay zero: the phantom of SURVEY F7). */
uint32_t kco_extract_read(const uint64_t *words, const uint64_t *filter, uint32_t L,
                          uint32_t k, unsigned char *out);

/* comparator order of GPUHandler.cu:247-298 applied by sortKmers (:300-327) */
void kco_sort_records(unsigned char *records, uint64_t n_records, uint32_t k);

/* reduceKMers (GPUHandler.cu:329-360): adjacent-equal fold, returns valid bytes.
 * (Returns 0 for an empty input; the reference returns S there, SURVEY App. D.) */
uint64_t kco_reduce(unsigned char *records, uint32_t k, uint64_t n_bytes);

/* processKMers (GPUHandler.cu:397-477) for one chunk of packed reads.
 * do_sort != 0 runs the sort step at :455-458 ("Pipeline B", the parity
 * artefact's producer); do_sort == 0 is HEAD ("Pipeline A").
 * out must hold kco_output_size() bytes.  Returns valid bytes in out, <0 on error. */
int64_t kco_process_chunk(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                          int do_sort, unsigned char *out);

/* KMerFileMerger::Merge over in-memory runs (KMerFileMerger.cpp:49-135) with
 * SortedKMerFile::ReadKmer's intra-run fold (SortedKMerFile.cpp:57-82).
 * out must hold sum(run_bytes).  Returns bytes written. */
uint64_t kco_merge_runs(const unsigned char *const *runs, const uint64_t *run_bytes,
                        uint32_t n_runs, uint32_t k, unsigned char *out);

/* Whole path: chunk the packed reads every chunk_reads reads, process each
 * chunk (sorted), merge all runs.  *out is malloc'd; caller frees with kco_free.
 * Returns bytes, <0 on error. */
int64_t kco_count(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                  uint64_t chunk_reads, unsigned char **out);

/* Same, chunks processed by n_threads worker threads (mirrors the reference's
 * 8 stream threads, KMerCounter.cpp:117-139); the merge stays single-threaded
 * like KMerFileMerger.  Used by bench.py's CPU legs only. */
int64_t kco_count_mt(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                     uint64_t chunk_reads, uint32_t n_threads, unsigned char **out);

/* Independent model of SURVEY Appendix A.2's "window" statement (NOT a
 * restatement of the reference's control flow): used for the three-way
 * differential test.  *out malloc'd.  strict != 0 gives true masked k-mers
 * with no phantom record. */
int64_t kco_naive_count(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                        int strict, unsigned char **out);

void kco_free(void *p);

/* FASTQFileReader::readData semantics (FASTQFileReader.cpp:49-89) on an
 * in-memory FASTQ text: "sequence = the line before a line starting with '+'".
 * Copies sequences back to back into out (cap bytes). Returns bytes written. */
int64_t kco_parse_fastq(const char *text, uint64_t n, char *out, uint64_t cap);

/* KMerPrinter::printKmer (KMerPrinter.cpp:68-91): 32 letters per word */
void kco_print_word(uint64_t w, char out32[32]);

/* ---- deterministic synthetic data; not from the reference ---- */
uint64_t kco_splitmix64(uint64_t x);
/* Packed reads (stride L, no separators). genome_len==0 -> iid random reads. */
void kco_gen_reads(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                   uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed);
/* zipf_loci > 0: half of the reads start at one of zipf_loci hot loci drawn
 * with probability ~ 1/rank^zipf_s (config 5). */
void kco_gen_reads_zipf(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                        uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed,
                        uint64_t zipf_loci, double zipf_s);
/* FASTQ text for reads [first, first+n): "@SYN.%010llu\n" + bases + "\n+\n" + L*'I' + "\n" */
uint64_t kco_fastq_bytes(uint64_t n_reads, uint32_t L);
void kco_gen_fastq(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                   uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
