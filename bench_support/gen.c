/*
 * gen.c -- seeded synthetic workloads of BASELINE.json / SURVEY.md 8(d), written straight into Arrow
 * Utf8View buffers (16-byte views, one data buffer, optional validity bitmap) in caller-provided
 * memory (so bench.py can hand in pinned host memory).  Bench/test support, not product code.
 *
 *   config 2 (C2): ASCII names, length U{4..24}, chars U{a..z}; name_b = name_a with k~U{0..3} random
 *                  single-codepoint edits (substitute / insert / delete / adjacent swap) w.p. 0.8,
 *                  an independent draw w.p. 0.2
 *   config 3 (C3): 70 % Latin rows (4..24 codepoints, each w.p. 0.15 from U+00C0..U+00FF except
 *                  U+00D7/U+00F7, else a..z), 30 % CJK rows (2..6 codepoints from U+4E00..U+9FFF);
 *                  name_b mutated as in C2 from the same script
 *   config 6 (L1): every row a Latin row of config 3, no nulls (not a BASELINE config: shows the
 *                  Latin-1 bit-plane path of the general kernel on its own)
 *   config 8 (N1): C2 with real-name spelling: words start with a capital, positions inside a name are a
 *                  space / hyphen / apostrophe w.p. 0.08 (any ASCII: the 7-plane kernel instantiation)
 *   config 9 (T1): config 7 (20..60-character ASCII strings) with a tail: one row in ten has 100..300
 *                  characters (free-text fields next to names: the rows above 64 bytes take the long-row kernels)
 *   config 4 (C4): long text, length U{200..4000} codepoints, 90 % a..z/space, 10 % two- and
 *                  three-byte codepoints; b = a with ~10 % random edits w.p. 0.5, else independent
 * Every row is generated from splitmix64(seed, row), so output is independent of the thread count.
 * Two passes: gen_pairs(..., data == NULL) only sizes the data buffers.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define MAXCP 4600

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint32_t rnd(rng_t *r, uint32_t n) { /* uniform in [0, n) */
    return (uint32_t)(((splitmix(&r->s) >> 32) * (uint64_t)n) >> 32);
}
static inline double rndf(rng_t *r) { return (double)(splitmix(&r->s) >> 11) * (1.0 / 9007199254740992.0); }

enum { S_ASCII, S_LATIN, S_CJK, S_TEXT, S_NAME };

static uint32_t draw_char(rng_t *r, int script) {
    switch (script) {
        case S_ASCII: return 'a' + rnd(r, 26);
        case S_LATIN:
            if (rndf(r) < 0.15) {
                uint32_t c;
                do c = 0xC0 + rnd(r, 64);
                while (c == 0xD7 || c == 0xF7);
                return c;
            }
            return 'a' + rnd(r, 26);
        case S_CJK: return 0x4E00 + rnd(r, 0x9FFF - 0x4E00 + 1);
        case S_NAME: {
            double u = rndf(r);
            if (u < 0.80) return 'a' + rnd(r, 26);
            if (u < 0.92) return 'A' + rnd(r, 26);
            return " -'"[rnd(r, 3)];
        }
        default: {
            double u = rndf(r);
            if (u < 0.75) return 'a' + rnd(r, 26);
            if (u < 0.90) return ' ';
            if (u < 0.95) return 0xC0 + rnd(r, 0x250 - 0xC0);    /* 2-byte */
            return 0x4E00 + rnd(r, 0x9FFF - 0x4E00 + 1);         /* 3-byte */
        }
    }
}

static int draw_string(rng_t *r, int script, int lo, int hi, uint32_t *out) {
    int n = lo + (int)rnd(r, (uint32_t)(hi - lo + 1));
    if (script == S_NAME) { /* "Mary-Ann O'Neil": capital after every separator */
        int start = 1;
        for (int i = 0; i < n; i++) {
            if (!start && i + 1 < n && rndf(r) < 0.08) {
                out[i] = " -'"[rnd(r, 3)];
                start = 1;
            } else {
                out[i] = (start ? 'A' : 'a') + rnd(r, 26);
                start = 0;
            }
        }
        return n;
    }
    for (int i = 0; i < n; i++) out[i] = draw_char(r, script);
    return n;
}

static int mutate(rng_t *r, int script, const uint32_t *a, int la, int k, uint32_t *b) {
    int lb = la;
    memcpy(b, a, sizeof(uint32_t) * (size_t)la);
    for (int e = 0; e < k; e++) {
        int op = (int)rnd(r, 4);
        if (op == 0 && lb > 0) {
            b[rnd(r, (uint32_t)lb)] = draw_char(r, script);
        } else if (op == 1 && lb < MAXCP - 1) {
            int p = (int)rnd(r, (uint32_t)lb + 1);
            memmove(b + p + 1, b + p, sizeof(uint32_t) * (size_t)(lb - p));
            b[p] = draw_char(r, script);
            lb++;
        } else if (op == 2 && lb > 1) {
            int p = (int)rnd(r, (uint32_t)lb);
            memmove(b + p, b + p + 1, sizeof(uint32_t) * (size_t)(lb - p - 1));
            lb--;
        } else if (op == 3 && lb > 1) {
            int p = (int)rnd(r, (uint32_t)lb - 1);
            uint32_t t = b[p];
            b[p] = b[p + 1];
            b[p + 1] = t;
        }
    }
    return lb;
}

static int encode(const uint32_t *cp, int n, uint8_t *out) {
    int k = 0;
    for (int i = 0; i < n; i++) {
        uint32_t c = cp[i];
        if (c < 0x80) {
            out[k++] = (uint8_t)c;
        } else if (c < 0x800) {
            out[k++] = (uint8_t)(0xC0 | (c >> 6));
            out[k++] = (uint8_t)(0x80 | (c & 0x3F));
        } else if (c < 0x10000) {
            out[k++] = (uint8_t)(0xE0 | (c >> 12));
            out[k++] = (uint8_t)(0x80 | ((c >> 6) & 0x3F));
            out[k++] = (uint8_t)(0x80 | (c & 0x3F));
        } else {
            out[k++] = (uint8_t)(0xF0 | (c >> 18));
            out[k++] = (uint8_t)(0x80 | ((c >> 12) & 0x3F));
            out[k++] = (uint8_t)(0x80 | ((c >> 6) & 0x3F));
            out[k++] = (uint8_t)(0x80 | (c & 0x3F));
        }
    }
    return k;
}

typedef struct {
    uint32_t a[MAXCP], b[MAXCP];
    uint8_t ea[4 * MAXCP], eb[4 * MAXCP];
} rowbuf_t;

/* generates row `row`; returns byte lengths and null flags */
static void gen_row(int config, uint64_t seed, int64_t row, double null_p, rowbuf_t *rb, int *na, int *nb,
                    int *null_a, int *null_b) {
    rng_t r = {seed * 0xD1342543DE82EF95ULL + (uint64_t)row * 0x2545F4914F6CDD1DULL + 1};
    splitmix(&r.s);
    int la, lb;
    if (config == 4) {
        la = draw_string(&r, S_TEXT, 200, 4000, rb->a);
        if (rndf(&r) < 0.5) {
            lb = mutate(&r, S_TEXT, rb->a, la, la / 10, rb->b);
        } else {
            lb = draw_string(&r, S_TEXT, 200, 4000, rb->b);
        }
    } else {
        int script = S_ASCII, lo = 4, hi = 24;
        if (config == 6) {
            script = S_LATIN; /* the Latin rows of C3 alone: a column of names with diacritics */
        } else if (config == 8) {
            script = S_NAME;
        } else if (config == 7 || config == 9) {
            lo = 20; /* medium ASCII strings (street addresses): most rows leave the 32-byte kernels */
            hi = 60;
            if (config == 9 && rndf(&r) < 0.1) {
                lo = 100;
                hi = 300;
            }
        } else if (config == 3) {
            if (rndf(&r) < 0.7) {
                script = S_LATIN;
            } else {
                script = S_CJK;
                lo = 2;
                hi = 6;
            }
        }
        la = draw_string(&r, script, lo, hi, rb->a);
        if (rndf(&r) < 0.8) {
            lb = mutate(&r, script, rb->a, la, (int)rnd(&r, 4), rb->b);
        } else {
            lb = draw_string(&r, script, lo, hi, rb->b);
        }
    }
    *na = encode(rb->a, la, rb->ea);
    *nb = encode(rb->b, lb, rb->eb);
    *null_a = null_p > 0 && rndf(&r) < null_p;
    *null_b = null_p > 0 && rndf(&r) < null_p;
}

typedef struct {
    int config;
    uint64_t seed;
    double null_p;
    int64_t lo, hi;           /* row range of this job */
    int64_t row_base;         /* global row number of local row 0 (multi-GPU shards) */
    int64_t bytes_a, bytes_b; /* out-of-line bytes of the range (count pass) */
    int64_t off_a, off_b;     /* starting data offsets (fill pass) */
    uint8_t *views_a, *views_b, *data_a, *data_b, *valid_a, *valid_b;
    int fill;
} job_t;

static void put_view(uint8_t *views, int64_t row, const uint8_t *bytes, int n, uint8_t *data, int64_t *off) {
    uint8_t *v = views + 16 * row;
    int32_t len = n;
    memset(v, 0, 16);
    memcpy(v, &len, 4);
    if (n <= 12) {
        memcpy(v + 4, bytes, (size_t)n);
    } else {
        int32_t zero = 0, o = (int32_t)*off;
        memcpy(v + 4, bytes, 4);
        memcpy(v + 8, &zero, 4);
        memcpy(v + 12, &o, 4);
        memcpy(data + *off, bytes, (size_t)n);
        *off += n;
    }
}

static void *run_job(void *arg) {
    job_t *j = arg;
    rowbuf_t *rb = malloc(sizeof *rb);
    int64_t oa = j->off_a, ob = j->off_b;
    for (int64_t row = j->lo; row < j->hi; row++) {
        int na, nb, xa, xb;
        gen_row(j->config, j->seed, j->row_base + row, j->null_p, rb, &na, &nb, &xa, &xb);
        if (xa) na = 0; /* null slots carry an empty view */
        if (xb) nb = 0;
        if (!j->fill) {
            if (na > 12) j->bytes_a += na;
            if (nb > 12) j->bytes_b += nb;
            continue;
        }
        put_view(j->views_a, row, rb->ea, na, j->data_a, &oa);
        put_view(j->views_b, row, rb->eb, nb, j->data_b, &ob);
        /* validity bytes are owned by exactly one job: ranges are multiples of 8 rows */
        if (j->valid_a) {
            if (xa) j->valid_a[row >> 3] &= (uint8_t)~(1u << (row & 7));
        }
        if (j->valid_b) {
            if (xb) j->valid_b[row >> 3] &= (uint8_t)~(1u << (row & 7));
        }
    }
    free(rb);
    return NULL;
}

/* views_*: 16*n bytes; valid_*: (n+7)/8 bytes or NULL (required when null_p > 0); data_*: NULL for the
 * sizing pass.  sizes[0..1] receive / must hold the data buffer sizes.  Data offsets must fit int32
 * (split n into several chunks for more than 2 GiB of out-of-line bytes).  Returns 0 or -1. */
API int gen_pairs(int config, uint64_t seed, int64_t row_base, int64_t n, double null_p, int n_threads,
                  uint8_t *views_a,
                  uint8_t *data_a, uint8_t *valid_a, uint8_t *views_b, uint8_t *data_b, uint8_t *valid_b,
                  int64_t sizes[2]) {
    if (n_threads < 1) n_threads = 1;
    int64_t block = ((n + n_threads - 1) / n_threads + 7) & ~7ll;
    if (block < 8) block = 8;
    int nj = (int)((n + block - 1) / block);
    if (nj < 1) nj = 1;
    job_t *jobs = calloc((size_t)nj, sizeof(job_t));
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nj);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !data_a) break;
        if (pass == 1) {
            if (valid_a) memset(valid_a, 0xFF, (size_t)((n + 7) / 8));
            if (valid_b) memset(valid_b, 0xFF, (size_t)((n + 7) / 8));
        }
        int64_t oa = 0, ob = 0;
        for (int p = 0; p < nj; p++) {
            job_t *j = &jobs[p];
            j->config = config;
            j->seed = seed;
            j->null_p = null_p;
            j->row_base = row_base;
            j->lo = p * block;
            j->hi = j->lo + block < n ? j->lo + block : n;
            j->fill = pass;
            j->views_a = views_a;
            j->views_b = views_b;
            j->data_a = data_a;
            j->data_b = data_b;
            j->valid_a = valid_a;
            j->valid_b = valid_b;
            if (pass == 1) {
                j->off_a = oa;
                j->off_b = ob;
                oa += j->bytes_a;
                ob += j->bytes_b;
            } else {
                j->bytes_a = j->bytes_b = 0;
            }
            pthread_create(&th[p], NULL, run_job, j);
        }
        for (int p = 0; p < nj; p++) pthread_join(th[p], NULL);
        if (pass == 0) {
            int64_t ta = 0, tb = 0;
            for (int p = 0; p < nj; p++) {
                ta += jobs[p].bytes_a;
                tb += jobs[p].bytes_b;
            }
            sizes[0] = ta;
            sizes[1] = tb;
            if (ta > 0x7FFFFFFFll || tb > 0x7FFFFFFFll) {
                free(jobs);
                free(th);
                return -1;
            }
        }
    }
    free(jobs);
    free(th);
    return 0;
}
