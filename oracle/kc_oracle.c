/*
 * kc_oracle.c -- CPU restatement of the reference's k-mer counting path.
 *
 * TEST INFRASTRUCTURE ONLY (see kc_oracle.h).  Plain C, single translation
 * unit, no dependency on the CUDA product.  Each function cites the reference
 * file:line it follows; citations are into /root/reference.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this file
 * is pinned against SURVEY.md Appendix A.4's known answers and against
 * oracle/_ref (the reference's own sources compiled in place); see
 * tests/test_oracle.py.
 */
#include "kc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ sizes */

uint32_t kco_words(uint32_t k) { return (k + 31) / 32; }

/* 8 bytes per started 32 bases + 4 bytes of count (GPUHandler.cu:237-243) */
uint32_t kco_record_size(uint32_t k) { return 8 * kco_words(k) + 4; }

uint64_t kco_output_size(uint64_t input_size, uint32_t L, uint32_t k) {
    if (L == 0 || k == 0 || k > L) return 0;
    uint64_t reads = input_size / L;              /* :236 */
    uint64_t per_read = (uint64_t)L - k + 1;      /* :237 */
    return reads * per_read * kco_record_size(k); /* :244 */
}

static uint64_t ld64(const unsigned char *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t ld32(const unsigned char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static void st64(unsigned char *p, uint64_t v) { memcpy(p, &v, 8); }
static void st32(unsigned char *p, uint32_t v) { memcpy(p, &v, 4); }

/* ----------------------------------------------------------------- encode */

/* bitEncode, GPUHandler.cu:10-111.  The reference keeps one 64-bit shift
 * register for the 2-bit codes and one for the invalid bits, stores the code
 * register every 32 bases (:25-34) and the filter register every 64 (:36-40),
 * then left-aligns and stores the partial tail (:94-109).  Letters other than
 * A/C/G/T take code 3 and set the filter bit (:79-87). */
int kco_encode_read(const char *read, uint32_t L, uint64_t *words, uint64_t *filter) {
    if (L == 0 || L % 32 == 0) return -1; /* tail store is broken there */
    uint64_t acc = 0, facc = 0;
    for (uint32_t j = 0; j < L; j++) {
        if (j > 0 && j % 32 == 0) { words[j / 32 - 1] = acc; acc = 0; }
        if (j > 0 && j % 64 == 0) { filter[j / 64 - 1] = facc; facc = 0; }
        uint64_t code, bad = 0;
        switch (read[j]) {
            case 'A': code = 0; break;
            case 'C': code = 1; break;
            case 'G': code = 2; break;
            case 'T': code = 3; break;
            default:  code = 3; bad = 1; break;
        }
        acc = (acc << 2) | code;
        facc = (facc << 1) | bad;
    }
    /* :94-109 (runs whenever L % 64 > 0, which L % 32 != 0 implies) */
    acc <<= (32 - (L % 32)) * 2;
    words[L / 32] = acc;
    facc <<= 64 - (L % 64);
    filter[L / 64] = facc;
    return 0;
}

/* checkBit, GPUHandler.cu:113-121: bit 0 is the most significant */
static int filter_bit(uint64_t f, uint32_t bit) { return (f >> (63 - bit)) & 1u; }

/* extractKMers, GPUHandler.cu:129-233 */
uint32_t kco_extract_read(const uint64_t *words, const uint64_t *filter, uint32_t L,
                          uint32_t k, unsigned char *out) {
    uint32_t S = kco_record_size(k);
    uint32_t emitted = 0;
    int64_t run = 0;            /* filterReadLength :147 */
    uint64_t o = 0;             /* outputIndex :149 */
    for (uint32_t i = 0; i < L; i++) {
        uint64_t f = filter[i / 64];                    /* :155-161 */
        if (!filter_bit(f, (uint8_t)i % 64)) {          /* :165 */
            run++;
            if (run >= (int64_t)k) {                    /* :168 */
                uint32_t first = i - k + 1;             /* :171 */
                uint32_t lsh = (first % 32) * 2;        /* :172 */
                uint32_t rsh = (32 - (k % 32)) * 2;     /* :173 */
                uint64_t x0 = (uint64_t)(first / 32) * 8;       /* :175 */
                uint64_t nbytes = k / 4 + ((k % 4) ? 1 : 0);    /* :177-182 */
                for (uint64_t x = x0; x < x0 + nbytes; x += 8) {    /* :187 */
                    uint64_t v = words[x / 8];
                    if (lsh > 0) {                               /* :193 */
                        v <<= lsh;
                        uint64_t v2 = 0;
                        if ((x + 8) * 4 < L)                     /* :196 */
                            v2 = words[x / 8 + 1] >> (64 - lsh);
                        v |= v2;
                    }
                    if (x + 8 > x0 + nbytes) {                   /* :206 tail mask */
                        v >>= rsh;
                        v <<= rsh;
                    }
                    st64(out + o, v);
                    o += 8;
                }
                st32(out + o, 1u);                               /* :217-219 */
                o += 4;
                emitted++;
                run--;                                           /* :221 */
            }
        } else {
            run = 0;                                             /* :225 */
        }
    }
    (void)S;
    return emitted;
}

/* ------------------------------------------------------------ sort/reduce */

static uint32_t g_cmp_words; /* qsort has no context argument */

/* KMer32/64/96/128Comparator, GPUHandler.cu:247-298: words compared in order,
 * unsigned, most significant word first */
static int cmp_records(const void *a, const void *b) {
    const unsigned char *pa = a, *pb = b;
    for (uint32_t w = 0; w < g_cmp_words; w++) {
        uint64_t x = ld64(pa + 8 * w), y = ld64(pb + 8 * w);
        if (x < y) return -1;
        if (x > y) return 1;
    }
    return 0;
}

/* LSD byte-radix over the key words, stable; same order as cmp_records.
 * (sortKmers, GPUHandler.cu:300-327, is thrust::sort with those comparators;
 * tied records differ at most in count and are summed by the reduce step, so
 * any correct sort gives the same artefact.) */
static void radix_sort_records(unsigned char *rec, uint64_t n, uint32_t W) {
    uint32_t S = 8 * W + 4;
    unsigned char *tmp = malloc((size_t)n * S);
    if (!tmp) { g_cmp_words = W; qsort(rec, n, S, cmp_records); return; }
    unsigned char *src = rec, *dst = tmp;
    for (int w = (int)W - 1; w >= 0; w--) {
        for (int b = 0; b < 8; b++) {
            uint64_t hist[256] = {0};
            const unsigned char *p = src + 8 * w + b;
            for (uint64_t i = 0; i < n; i++) hist[p[i * S]]++;
            int trivial = 0;
            for (int d = 0; d < 256; d++) if (hist[d] == n) trivial = 1;
            if (trivial) continue;
            uint64_t sum = 0;
            for (int d = 0; d < 256; d++) { uint64_t c = hist[d]; hist[d] = sum; sum += c; }
            for (uint64_t i = 0; i < n; i++) {
                unsigned d = p[i * S];
                memcpy(dst + hist[d]++ * S, src + i * S, S);
            }
            unsigned char *t = src; src = dst; dst = t;
        }
    }
    if (src != rec) memcpy(rec, src, (size_t)n * S);
    free(tmp);
}

void kco_sort_records(unsigned char *records, uint64_t n_records, uint32_t k) {
    uint32_t W = kco_words(k);
    if (n_records < 4096) {
        g_cmp_words = W;
        qsort(records, n_records, kco_record_size(k), cmp_records);
    } else {
        radix_sort_records(records, n_records, W);
    }
}

/* CheckEquals, GPUHandler.cu:329-338 */
static int keys_equal(const unsigned char *a, const unsigned char *b, uint32_t S) {
    for (uint32_t off = 0; off + 4 < S; off += 8)
        if (ld64(a + off) != ld64(b + off)) return 0;
    return 1;
}

static int key_less(const unsigned char *a, const unsigned char *b, uint32_t S) {
    for (uint32_t off = 0; off + 4 < S; off += 8) {
        uint64_t x = ld64(a + off), y = ld64(b + off);
        if (x < y) return 1;
        if (x > y) return 0;
    }
    return 0;
}

/* reduceKMers, GPUHandler.cu:340-360: the survivor of an adjacent-equal group
 * is its first record; counts add in uint32 (wrap) */
uint64_t kco_reduce(unsigned char *rec, uint32_t k, uint64_t n_bytes) {
    uint32_t S = kco_record_size(k);
    if (n_bytes < S) return 0;
    uint64_t keep = 0;
    for (uint64_t i = S; i + S <= n_bytes; i += S) {
        if (keys_equal(rec + keep, rec + i, S)) {
            st32(rec + keep + S - 4, ld32(rec + keep + S - 4) + ld32(rec + i + S - 4));
        } else {
            if (i - keep > S) memcpy(rec + keep + S, rec + i, S);
            keep += S;
        }
    }
    return keep + S;
}

/* ------------------------------------------------------------ chunk driver */

/* processKMers, GPUHandler.cu:397-477: zero the per-read slots (:406-409),
 * encode + extract every whole read (a partial tail is ignored, :13,:134),
 * [sort :455-458], reduce (:466). */
int64_t kco_process_chunk(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                          int do_sort, unsigned char *out) {
    if (k == 0 || k > 128 || k > L || L % 32 == 0 || L > 32767) return -1;
    uint32_t S = kco_record_size(k);
    uint64_t reads = input_size / L;
    uint64_t slots = (uint64_t)L - k + 1;
    uint64_t out_bytes = reads * slots * S;
    if (out_bytes == 0) return 0;
    memset(out, 0, out_bytes);
    uint32_t nw = (L + 31) / 32, nf = (L + 63) / 64;
    uint64_t words[1024 + 2], filter[512 + 1];
    for (uint64_t r = 0; r < reads; r++) {
        memset(words, 0, sizeof(uint64_t) * (nw + 1));
        if (kco_encode_read(input + r * L, L, words, filter) != 0) return -1;
        (void)nf;
        kco_extract_read(words, filter, L, k, out + r * slots * S);
    }
    if (do_sort) kco_sort_records(out, reads * slots, k);
    return (int64_t)kco_reduce(out, k, out_bytes);
}

/* ------------------------------------------------------------------ merge */

typedef struct {
    unsigned char *buf;   /* mutable copy of the run (the cursor's cache) */
    uint64_t pos, end;    /* byte offsets */
} cursor_t;

/* SortedKMerFile::ReadKmer, SortedKMerFile.cpp:57-82: equal neighbours inside
 * a run are folded forward -- the LATER record receives the sum */
static unsigned char *cursor_read(cursor_t *c, uint32_t S) {
    if (c->pos + S > c->end) return NULL;
    while (c->pos + 2 * S <= c->end && keys_equal(c->buf + c->pos, c->buf + c->pos + S, S)) {
        unsigned char *cur = c->buf + c->pos, *nxt = cur + S;
        st32(nxt + S - 4, ld32(nxt + S - 4) + ld32(cur + S - 4));
        c->pos += S;                                    /* PopKmer :84-87 */
    }
    return c->buf + c->pos;
}

/* KMerFileMerger::Merge, KMerFileMerger.cpp:49-96: linear min-scan over the
 * open cursors collecting every cursor tied at the minimum (:52-66), pop them
 * all and add their counts into the first (:68-82), emit (:85). */
uint64_t kco_merge_runs(const unsigned char *const *runs, const uint64_t *run_bytes,
                        uint32_t n_runs, uint32_t k, unsigned char *out) {
    uint32_t S = kco_record_size(k);
    cursor_t *cur = calloc(n_runs ? n_runs : 1, sizeof(cursor_t));
    uint32_t *open = malloc(sizeof(uint32_t) * (n_runs ? n_runs : 1));
    uint32_t *tied = malloc(sizeof(uint32_t) * (n_runs ? n_runs : 1));
    uint32_t n_open = 0;
    for (uint32_t i = 0; i < n_runs; i++) {
        uint64_t nb = run_bytes[i] - run_bytes[i] % S;
        cur[i].buf = malloc(nb ? nb : 1);
        memcpy(cur[i].buf, runs[i], nb);
        cur[i].end = nb;
        if (nb) open[n_open++] = i;
    }
    uint64_t w = 0;
    while (n_open) {
        uint32_t n_tied = 0;
        uint32_t best = open[0];
        for (uint32_t j = 1; j < n_open; j++) {
            uint32_t c = open[j];
            unsigned char *a = cursor_read(&cur[c], S), *b = cursor_read(&cur[best], S);
            if (key_less(a, b, S)) { best = c; n_tied = 0; }
            else if (keys_equal(a, b, S)) { tied[n_tied++] = best; best = c; }
        }
        tied[n_tied++] = best;
        unsigned char *first = cursor_read(&cur[tied[0]], S);
        uint32_t total = ld32(first + S - 4);
        for (uint32_t t = 1; t < n_tied; t++)
            total += ld32(cursor_read(&cur[tied[t]], S) + S - 4);
        memcpy(out + w, first, S - 4);
        st32(out + w + S - 4, total);
        w += S;
        for (uint32_t t = 0; t < n_tied; t++) {
            cursor_t *c = &cur[tied[t]];
            c->pos += S;
            if (cursor_read(c, S) == NULL) {    /* exhausted: drop from the open list, keep order */
                uint32_t j = 0;
                while (open[j] != tied[t]) j++;
                memmove(open + j, open + j + 1, sizeof(uint32_t) * (n_open - j - 1));
                n_open--;
            }
        }
    }
    for (uint32_t i = 0; i < n_runs; i++) free(cur[i].buf);
    free(cur); free(open); free(tied);
    return w;
}

/* ------------------------------------------------------------ whole path */

typedef struct {
    const char *input; uint64_t input_size; uint32_t L, k; uint64_t chunk_bytes;
    uint64_t n_chunks; unsigned char **runs; uint64_t *run_bytes;
    volatile uint64_t next; int failed;
    pthread_mutex_t mu;
} job_t;

static void *chunk_worker(void *arg) {
    job_t *j = arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint64_t c = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (c >= j->n_chunks) break;
        uint64_t off = c * j->chunk_bytes;
        uint64_t sz = j->input_size - off < j->chunk_bytes ? j->input_size - off : j->chunk_bytes;
        uint64_t cap = kco_output_size(sz, j->L, j->k);
        j->runs[c] = malloc(cap ? cap : 1);
        int64_t nb = kco_process_chunk(j->input + off, sz, j->L, j->k, 1, j->runs[c]);
        if (nb < 0) { j->failed = 1; nb = 0; }
        j->run_bytes[c] = (uint64_t)nb;
    }
    return NULL;
}

int64_t kco_count_mt(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                     uint64_t chunk_reads, uint32_t n_threads, unsigned char **out) {
    *out = NULL;
    if (k == 0 || k > 128 || k > L || L % 32 == 0 || L > 32767) return -1;
    uint64_t reads = input_size / L;
    input_size = reads * L;
    if (chunk_reads == 0) chunk_reads = reads ? reads : 1;
    job_t j;
    memset(&j, 0, sizeof j);
    j.input = input; j.input_size = input_size; j.L = L; j.k = k;
    j.chunk_bytes = chunk_reads * L;
    j.n_chunks = (input_size + j.chunk_bytes - 1) / j.chunk_bytes;
    j.runs = calloc(j.n_chunks ? j.n_chunks : 1, sizeof(*j.runs));
    j.run_bytes = calloc(j.n_chunks ? j.n_chunks : 1, sizeof(*j.run_bytes));
    pthread_mutex_init(&j.mu, NULL);
    if (n_threads <= 1) {
        chunk_worker(&j);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
        for (uint32_t t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, chunk_worker, &j);
        for (uint32_t t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    uint64_t total = 0;
    for (uint64_t c = 0; c < j.n_chunks; c++) total += j.run_bytes[c];
    int64_t ret = -1;
    if (!j.failed) {
        *out = malloc(total ? total : 1);
        ret = (int64_t)kco_merge_runs((const unsigned char *const *)j.runs, j.run_bytes,
                                      (uint32_t)j.n_chunks, k, *out);
    }
    for (uint64_t c = 0; c < j.n_chunks; c++) free(j.runs[c]);
    free(j.runs); free(j.run_bytes);
    pthread_mutex_destroy(&j.mu);
    return ret;
}

int64_t kco_count(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                  uint64_t chunk_reads, unsigned char **out) {
    return kco_count_mt(input, input_size, L, k, chunk_reads, 1, out);
}

void kco_free(void *p) { free(p); }

/* --------------------------------------------- independent window model */

/* SURVEY Appendix A.2, "equivalent statement": a key is the 2-bit codes of the
 * span bases from p, span = 32W when k%32 is 0,29,30,31 else k; code 0 past the
 * read end; code 3 for non-ACGT letters inside the overhang. */
static int base_code(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}

int64_t kco_naive_count(const char *input, uint64_t input_size, uint32_t L, uint32_t k,
                        int strict, unsigned char **out) {
    *out = NULL;
    if (k == 0 || k > 128 || k > L) return -1;
    uint32_t W = kco_words(k), S = 8 * W + 4;
    uint32_t m = k % 32;
    uint32_t span = (!strict && (m == 0 || m >= 29)) ? 32 * W : k;
    uint64_t reads = input_size / L, slots = (uint64_t)L - k + 1;
    unsigned char *rec = malloc((size_t)(reads * slots + 1) * S);
    uint64_t n = 0, phantom = 0;
    for (uint64_t r = 0; r < reads; r++) {
        const char *s = input + r * L;
        uint32_t run = 0, got = 0;
        for (uint32_t j = 0; j < L; j++) {
            run = base_code(s[j]) >= 0 ? run + 1 : 0;
            if (run < k) continue;
            uint32_t p = j + 1 - k;
            unsigned char *o = rec + n * S;
            for (uint32_t w = 0; w < W; w++) {
                uint64_t v = 0;
                for (uint32_t b = 0; b < 32; b++) {
                    uint32_t q = 32 * w + b;
                    uint64_t code = 0;
                    if (q < span && p + q < L) {
                        int c = base_code(s[p + q]);
                        code = c < 0 ? 3 : (uint64_t)c;
                    }
                    v = (v << 2) | code;
                }
                st64(o + 8 * w, v);
            }
            st32(o + 8 * W, 1);
            n++; got++;
        }
        phantom += slots - got;
    }
    if (phantom && !strict) { memset(rec + n * S, 0, S); n++; }
    g_cmp_words = W;
    qsort(rec, n, S, cmp_records);
    uint64_t nb = kco_reduce(rec, k, n * S);
    *out = rec;
    return (int64_t)nb;
}

/* ------------------------------------------------------------ FASTQ text */

typedef struct { const char *p; uint64_t n, pos; int eof, fail; } tstream;
typedef struct { char *s; uint64_t len, cap; } tline;

/* std::getline on an ifstream: a stream that is no longer good() leaves the
 * string untouched; otherwise the string is cleared and filled up to '\n' */
static void t_getline(tstream *st, tline *ln) {
    if (st->eof || st->fail) { st->fail = 1; return; }
    ln->len = 0;
    uint64_t b = st->pos;
    while (st->pos < st->n && st->p[st->pos] != '\n') st->pos++;
    uint64_t len = st->pos - b;
    if (len + 1 > ln->cap) { ln->cap = 2 * (len + 1); ln->s = realloc(ln->s, ln->cap); }
    memcpy(ln->s, st->p + b, len);
    ln->len = len;
    if (st->pos < st->n) st->pos++;          /* consume the newline */
    else { st->eof = 1; if (len == 0) st->fail = 1; }
}

/* FASTQFileReader::readData, FASTQFileReader.cpp:49-89 */
int64_t kco_parse_fastq(const char *text, uint64_t n, char *out, uint64_t cap) {
    tstream st = { text, n, 0, 0, 0 };
    tline temp = {0}, line = {0};
    temp.cap = line.cap = 512; temp.s = malloc(512); line.s = malloc(512);
    uint64_t off = 0;
    t_getline(&st, &temp);                                   /* :54 */
    t_getline(&st, &line);                                   /* :55 */
    while (line.len != 0 && off + temp.len < cap) {          /* :57 */
        if (line.s[0] == '+') {                              /* :59 */
            memcpy(out + off, temp.s, temp.len);             /* :62-63 */
            off += temp.len;
            t_getline(&st, &temp);                           /* :68-69 */
            t_getline(&st, &line);
            /* a file without a final newline leaves temp/line unchanged here, so
             * the last quality line is re-copied until the chunk is full (SURVEY
             * App. B) -- reproduced; only the zero-progress case is cut short,
             * where the reference would spin forever */
            if (st.fail && temp.len == 0) break;
        } else {
            /* :76 temp = line (copy: a failed getline must leave line as it was) */
            if (line.len + 1 > temp.cap) { temp.cap = 2 * (line.len + 1); temp.s = realloc(temp.s, temp.cap); }
            memcpy(temp.s, line.s, line.len);
            temp.len = line.len;
            t_getline(&st, &line);                           /* :77 */
            if (st.fail) break;                              /* reference spins here */
        }
    }
    free(temp.s); free(line.s);
    return (int64_t)off;
}

/* KMerPrinter::printKmer, KMerPrinter.cpp:68-91 */
void kco_print_word(uint64_t w, char out32[32]) {
    static const char letters[4] = { 'A', 'C', 'G', 'T' };
    for (int i = 0; i < 32; i++) out32[i] = letters[(w >> (62 - 2 * i)) & 3];
}

/* ----------------------------------------------------------- synthetic data */

uint64_t kco_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

#define SEED_GENOME 0x67656E6F6D650000ull
#define SEED_START  0x7374617274000000ull
#define SEED_ERROR  0x6572726F72000000ull
#define SEED_LOCUS  0x6C6F637573000000ull
#define SEED_PICK   0x7069636B00000000ull

static inline unsigned genome_code(uint64_t seed, uint64_t g) {
    uint64_t w = kco_splitmix64((seed * 0x100000001B3ull) ^ SEED_GENOME ^ (g >> 5));
    return (unsigned)(w >> (2 * (g & 31))) & 3u;
}

/* Zipf-like rank in [0, M) using integers only (so a device twin can be bit-identical):
 * pick an octave uniformly, then a rank uniformly inside it => P(rank) ~ 1/rank (s = 1). */
static uint64_t zipf_rank(uint64_t h, uint64_t M) {
    uint32_t levels = 0;
    while ((1ull << levels) < M + 1 && levels < 63) levels++;       /* octaves covering [0, M) */
    if (levels == 0) return 0;
    uint32_t j = (uint32_t)((h >> 40) % levels);
    uint64_t r = ((1ull << j) - 1) + ((h & 0xFFFFFFFFFFull) & ((1ull << j) - 1));
    return r < M ? r : r % M;
}

void kco_gen_reads_zipf(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                        uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed,
                        uint64_t zipf_loci, double zipf_s) {
    (void)zipf_s;   /* the octave sampler realises s = 1; kept in the signature for config files */
    static const char letters[4] = { 'A', 'C', 'G', 'T' };
    const double inv53 = 1.0 / 9007199254740992.0;
    int noisy = (sub_rate > 0.0) || (n_rate > 0.0);
    for (uint64_t t = 0; t < n_reads; t++) {
        uint64_t i = first_read + t;
        char *o = out + t * L;
        uint64_t start = 0;
        if (genome_len) {
            uint64_t span = genome_len - L + 1;
            start = kco_splitmix64(seed ^ SEED_START ^ (i * 0x9E3779B97F4A7C15ull)) % span;
            if (zipf_loci) {
                uint64_t h = kco_splitmix64(seed ^ SEED_PICK ^ i);
                if (h >> 63) {
                    uint64_t rank = zipf_rank(h, zipf_loci);
                    start = kco_splitmix64(seed ^ SEED_LOCUS ^ rank) % span;
                }
            }
        }
        for (uint32_t j = 0; j < L; j++) {
            unsigned code = genome_len ? genome_code(seed, start + j)
                                       : genome_code(seed, i * (uint64_t)L + j);
            char c = letters[code];
            if (noisy) {
                uint64_t h = kco_splitmix64(seed ^ SEED_ERROR ^ (i * (uint64_t)L + j));
                double u = (double)(h >> 11) * inv53;
                if (u < n_rate) c = 'N';
                else if (u < n_rate + sub_rate) c = letters[(code + 1 + (unsigned)(h & 0x3ff) % 3) & 3];
            }
            o[j] = c;
        }
    }
}

void kco_gen_reads(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                   uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed) {
    kco_gen_reads_zipf(out, first_read, n_reads, L, genome_len, sub_rate, n_rate, seed, 0, 0.0);
}

uint64_t kco_fastq_bytes(uint64_t n_reads, uint32_t L) {
    /* "@SYN." + 10 digits + '\n' = 16, bases + '\n', "+\n", quals + '\n' */
    return n_reads * (16 + (uint64_t)L + 1 + 2 + L + 1);
}

void kco_gen_fastq(char *out, uint64_t first_read, uint64_t n_reads, uint32_t L,
                   uint64_t genome_len, double sub_rate, double n_rate, uint64_t seed) {
    uint64_t rec = kco_fastq_bytes(1, L);
    for (uint64_t t = 0; t < n_reads; t++) {
        char *o = out + t * rec;
        char hdr[32];
        snprintf(hdr, sizeof hdr, "@SYN.%010llu\n", (unsigned long long)(first_read + t));
        memcpy(o, hdr, 16);
        kco_gen_reads(o + 16, first_read + t, 1, L, genome_len, sub_rate, n_rate, seed);
        o[16 + L] = '\n';
        o[17 + L] = '+';
        o[18 + L] = '\n';
        memset(o + 19 + L, 'I', L);
        o[19 + 2 * (uint64_t)L] = '\n';
    }
}
