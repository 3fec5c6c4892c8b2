/*
 * strsim_oracle.c -- CPU restatement of polars-strsim's five similarity measures.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The CUDA library (polars-strsim_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  This restatement is checked against all 1115 known-answer vectors of the
 * reference's own unit tests (/root/reference/src/expressions/strsim.rs:371-1534, tolerance 1e-8,
 * strsim.rs:349) and the README table (/root/reference/README.md:65-70) by tests/test_oracle.py.
 * The reference itself (Rust, needs cargo + polars 0.43.1 + pyo3) cannot be compiled in this image
 * (no rustc/cargo, no network), so there is no oracle/_ref binary; see DESIGN.md.
 *
 * It deliberately keeps the reference's ALGORITHMS (so that it doubles as the CPU baseline,
 * "kind": "port"): two-row Wagner-Fischer with 8-byte cells, greedy windowed Jaro with a flag
 * vector, a SipHash-1-3 hash map for the character multisets, one contiguous row range per thread.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { M_LEVENSHTEIN = 0, M_JARO = 1, M_JARO_WINKLER = 2, M_JACCARD = 3, M_SORENSEN_DICE = 4 };

/* Integer intermediates reported per row (all int64):
 *   [0] flag: 0 = general path, 1 = byte-equal short circuit (1.0), 2 = exactly-one-empty short
 *       circuit (0.0; not taken by Levenshtein), 3 = Jaro single-char branch
 *   [1] la, [2] lb  (codepoints; 0 when flag is 1 or 2 because the reference never decodes there)
 *   [3..5] measure specific: lev: d,0,0 | jaro: m,t,0 | jw: m,t,l | jaccard: inter,uni,0 |
 *       dice: inter,la+lb,0                                                                       */
#define N_INTS 6

/* ---- scratch, reused across rows like the reference's per-object Vec/HashMap (strsim.rs:109-123,
 *      164-178, 274-284); INITIAL_BUFFER_LENGTH = 50 (strsim.rs:7) ------------------------------- */
typedef struct {
    uint32_t key;
    uint8_t used;
    uint64_t cnt[2];
} slot_t;

typedef struct {
    uint32_t *a, *b;
    size_t cap_a, cap_b;
    uint64_t *matrix; /* [(lb+1)][2] like Vec<[usize;2]> */
    size_t cap_m;
    uint8_t *flagged; /* [max(la,lb)][2] like Vec<[bool;2]> */
    size_t cap_f;
    slot_t *slots; /* open addressing, power of two */
    size_t n_slots, n_used;
    uint32_t *touched; /* slots to reset on clear() */
    size_t n_touched, cap_touched;
} scratch_t;

static void scratch_init(scratch_t *s) {
    memset(s, 0, sizeof *s);
    s->cap_a = s->cap_b = s->cap_m = s->cap_f = 50;
    s->a = malloc(s->cap_a * sizeof(uint32_t));
    s->b = malloc(s->cap_b * sizeof(uint32_t));
    s->matrix = malloc(s->cap_m * 2 * sizeof(uint64_t));
    s->flagged = malloc(s->cap_f * 2);
    s->n_slots = 128;
    s->slots = calloc(s->n_slots, sizeof(slot_t));
    s->cap_touched = 128;
    s->touched = malloc(s->cap_touched * sizeof(uint32_t));
}

static void scratch_free(scratch_t *s) {
    free(s->a);
    free(s->b);
    free(s->matrix);
    free(s->flagged);
    free(s->slots);
    free(s->touched);
}

#define ENSURE(ptr, cap, need, elt)                 \
    do {                                            \
        if ((need) > (cap)) {                       \
            while ((cap) < (need)) (cap) *= 2;      \
            (ptr) = realloc((ptr), (cap) * (elt)); \
        }                                           \
    } while (0)

/* ---- UTF-8 -> Unicode scalar values, the unit of every measure (str::chars(), strsim.rs:133,138,
 *      189,194,261-262,297-300,333-336).  Input is valid UTF-8 (Polars guarantees it); malformed
 *      bytes are handled deterministically and memory-safely (sequence clamped to the slice). ---- */
static size_t decode_utf8(const uint8_t *s, size_t n, uint32_t *out) {
    size_t i = 0, k = 0;
    while (i < n) {
        uint8_t c = s[i];
        size_t len = c < 0x80 ? 1 : c < 0xC0 ? 1 : c < 0xE0 ? 2 : c < 0xF0 ? 3 : 4;
        if (len > n - i) len = n - i;
        uint32_t cp;
        if (len == 1) {
            cp = c;
        } else {
            cp = c & (0xFFu >> (len + 1));
            for (size_t j = 1; j < len; j++) cp = (cp << 6) | (s[i + j] & 0x3F);
        }
        out[k++] = cp;
        i += len;
    }
    return k;
}

static int bytes_equal(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    return na == nb && (na == 0 || memcmp(a, b, na) == 0);
}

/* ---- Levenshtein, strsim.rs:125-162 ------------------------------------------------------------ */
static double lev_compute(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb, size_t nb,
                          int64_t *ints) {
    if (bytes_equal(pa, na, pb, nb)) { /* strsim.rs:128 */
        ints[0] = 1;
        return 1.0;
    }
    ENSURE(s->a, s->cap_a, na, sizeof(uint32_t));
    ENSURE(s->b, s->cap_b, nb, sizeof(uint32_t));
    size_t la = decode_utf8(pa, na, s->a); /* strsim.rs:131-135 */
    size_t lb = decode_utf8(pb, nb, s->b); /* strsim.rs:136-140 */
    ENSURE(s->matrix, s->cap_m, lb + 1, 2 * sizeof(uint64_t));
    uint64_t(*m)[2] = (uint64_t(*)[2])s->matrix;
    for (size_t j = 0; j <= lb; j++) { /* strsim.rs:141-145 */
        m[j][0] = j;
        m[j][1] = 0;
    }
    const uint32_t *a = s->a, *b = s->b;
    for (size_t i = 0; i < la; i++) { /* strsim.rs:146-159 */
        size_t v0 = i % 2, v1 = (i + 1) % 2;
        m[0][v1] = i + 1;
        uint32_t ai = a[i];
        for (size_t j = 0; j < lb; j++) {
            uint64_t sub = ai == b[j] ? m[j][v0] : m[j][v0] + 1;
            uint64_t del = m[j + 1][v0] + 1;
            uint64_t ins = m[j][v1] + 1;
            uint64_t x = sub < del ? sub : del;
            m[j + 1][v1] = x < ins ? x : ins;
        }
    }
    uint64_t d = m[lb][la % 2];
    size_t mx = la > lb ? la : lb;
    ints[1] = (int64_t)la;
    ints[2] = (int64_t)lb;
    ints[3] = (int64_t)d;
    return 1.0 - ((double)d / (double)mx); /* strsim.rs:160 */
}

/* ---- Jaro, strsim.rs:180-245 ------------------------------------------------------------------- */
static double jaro_compute(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb, size_t nb,
                           int64_t *ints) {
    if (bytes_equal(pa, na, pb, nb)) { /* strsim.rs:182-183 */
        ints[0] = 1;
        return 1.0;
    }
    if (na == 0 || nb == 0) { /* strsim.rs:184-186 */
        ints[0] = 2;
        return 0.0;
    }
    ENSURE(s->a, s->cap_a, na, sizeof(uint32_t));
    ENSURE(s->b, s->cap_b, nb, sizeof(uint32_t));
    size_t la = decode_utf8(pa, na, s->a);
    size_t lb = decode_utf8(pb, nb, s->b);
    const uint32_t *a = s->a, *b = s->b;
    ints[1] = (int64_t)la;
    ints[2] = (int64_t)lb;
    if (la == 1 && lb == 1) { /* strsim.rs:197-199 */
        ints[0] = 3;
        return a[0] == b[0] ? 1.0 : 0.0;
    }
    size_t mx = la > lb ? la : lb;
    size_t bound = mx / 2 - 1; /* strsim.rs:200 */
    size_t m = 0;
    ENSURE(s->flagged, s->cap_f, mx, 2);
    uint8_t(*fl)[2] = (uint8_t(*)[2])s->flagged;
    memset(fl, 0, mx * 2); /* strsim.rs:202-207 */
    size_t outer = la < lb + bound ? la : lb + bound; /* .take(b.len() + bound), strsim.rs:208 */
    for (size_t i = 0; i < outer; i++) {
        size_t lo = bound > i ? 0 : i - bound;                   /* strsim.rs:209 */
        size_t hi = i + bound < lb - 1 ? i + bound : lb - 1;     /* strsim.rs:210 */
        for (size_t j = lo; j <= hi; j++) {                      /* strsim.rs:211-218 */
            if (a[i] == b[j] && !fl[j][1]) {
                m++;
                fl[i][0] = 1;
                fl[j][1] = 1;
                break;
            }
        }
    }
    /* strsim.rs:220-237: zip the flagged positions of a and of b in index order, count mismatches */
    size_t t = 0, j = 0;
    for (size_t i = 0; i < mx; i++) {
        if (!fl[i][0]) continue;
        while (j < mx && !fl[j][1]) j++;
        if (j >= mx) break;
        if (a[i] != b[j]) t++;
        j++;
    }
    ints[3] = (int64_t)m;
    ints[4] = (int64_t)t;
    if (m == 0) return 0.0; /* strsim.rs:238-239 */
    /* strsim.rs:241-242: left-to-right, integer t/2 */
    return ((double)m / (double)la + (double)m / (double)lb + (double)(m - t / 2) / (double)m) / 3.0;
}

/* ---- Jaro-Winkler, strsim.rs:257-272 ----------------------------------------------------------- */
static double jw_compute(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb, size_t nb,
                         int64_t *ints) {
    double js = jaro_compute(s, pa, na, pb, nb, ints);
    if (js > 0.7) { /* strsim.rs:260 */
        /* a.chars().zip(b.chars()).take(4).take_while(eq).count(), strsim.rs:261-266 (re-decodes) */
        ENSURE(s->a, s->cap_a, na, sizeof(uint32_t));
        ENSURE(s->b, s->cap_b, nb, sizeof(uint32_t));
        size_t la = decode_utf8(pa, na, s->a);
        size_t lb = decode_utf8(pb, nb, s->b);
        size_t l = 0;
        while (l < 4 && l < la && l < lb && s->a[l] == s->b[l]) l++;
        if (ints[0] == 0) ints[5] = (int64_t)l; /* reported for general-path rows only */
        return js + ((double)l * 0.1 * (1.0 - js)); /* strsim.rs:267 */
    }
    return js;
}

/* ---- character multiset, HashMap<char,[usize;2]> with std's SipHash-1-3 (strsim.rs:274-284) ----- */
#define ROTL(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND           \
    do {                   \
        v0 += v1;          \
        v1 = ROTL(v1, 13); \
        v1 ^= v0;          \
        v0 = ROTL(v0, 32); \
        v2 += v3;          \
        v3 = ROTL(v3, 16); \
        v3 ^= v2;          \
        v0 += v3;          \
        v3 = ROTL(v3, 21); \
        v3 ^= v0;          \
        v2 += v1;          \
        v1 = ROTL(v1, 17); \
        v1 ^= v2;          \
        v2 = ROTL(v2, 32); \
    } while (0)

static uint64_t siphash13_u32(uint32_t key) {
    const uint64_t k0 = 0x0706050403020100ULL, k1 = 0x0f0e0d0c0b0a0908ULL;
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
    uint64_t b = ((uint64_t)4 << 56) | key;
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

static void map_clear(scratch_t *s) {
    for (size_t i = 0; i < s->n_touched; i++) s->slots[s->touched[i]].used = 0;
    s->n_touched = 0;
    s->n_used = 0;
}

static void map_grow(scratch_t *s);

static uint64_t *map_entry(scratch_t *s, uint32_t key) {
    if ((s->n_used + 1) * 8 > s->n_slots * 7) map_grow(s);
    size_t mask = s->n_slots - 1;
    size_t i = (size_t)siphash13_u32(key) & mask;
    while (s->slots[i].used && s->slots[i].key != key) i = (i + 1) & mask;
    slot_t *e = &s->slots[i];
    if (!e->used) {
        e->used = 1;
        e->key = key;
        e->cnt[0] = e->cnt[1] = 0;
        s->n_used++;
        ENSURE(s->touched, s->cap_touched, s->n_touched + 1, sizeof(uint32_t));
        s->touched[s->n_touched++] = (uint32_t)i;
    }
    return e->cnt;
}

static void map_grow(scratch_t *s) {
    slot_t *old = s->slots;
    size_t old_touched = s->n_touched;
    uint32_t *old_idx = malloc((old_touched + 1) * sizeof(uint32_t));
    memcpy(old_idx, s->touched, old_touched * sizeof(uint32_t));
    s->n_slots *= 2;
    s->slots = calloc(s->n_slots, sizeof(slot_t));
    s->n_touched = 0;
    s->n_used = 0;
    for (size_t k = 0; k < old_touched; k++) {
        slot_t *o = &old[old_idx[k]];
        uint64_t *c = map_entry(s, o->key);
        c[0] = o->cnt[0];
        c[1] = o->cnt[1];
    }
    free(old_idx);
    free(old);
}

/* counts chars of a into slot 0 and of b into slot 1 (strsim.rs:297-300 / 333-336), then folds
 * min / max / sum over the distinct characters (strsim.rs:301-305 / 337-342) */
static void multiset_fold(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb, size_t nb,
                          uint64_t *inter, uint64_t *uni, uint64_t *total, size_t *la_out,
                          size_t *lb_out) {
    ENSURE(s->a, s->cap_a, na, sizeof(uint32_t));
    ENSURE(s->b, s->cap_b, nb, sizeof(uint32_t));
    size_t la = decode_utf8(pa, na, s->a);
    size_t lb = decode_utf8(pb, nb, s->b);
    map_clear(s);
    for (size_t i = 0; i < la; i++) map_entry(s, s->a[i])[0] += 1;
    for (size_t i = 0; i < lb; i++) map_entry(s, s->b[i])[1] += 1;
    uint64_t fi = 0, fu = 0, ft = 0;
    for (size_t k = 0; k < s->n_touched; k++) {
        const uint64_t *v = s->slots[s->touched[k]].cnt;
        fi += v[0] < v[1] ? v[0] : v[1];
        fu += v[0] > v[1] ? v[0] : v[1];
        ft += v[0];
        ft += v[1];
    }
    *inter = fi;
    *uni = fu;
    *total = ft;
    *la_out = la;
    *lb_out = lb;
}

/* ---- Jaccard, strsim.rs:286-308 ---------------------------------------------------------------- */
static double jaccard_compute(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb,
                              size_t nb, int64_t *ints) {
    if (bytes_equal(pa, na, pb, nb)) {
        ints[0] = 1;
        return 1.0;
    }
    if (na == 0 || nb == 0) {
        ints[0] = 2;
        return 0.0;
    }
    uint64_t inter, uni, total;
    size_t la, lb;
    multiset_fold(s, pa, na, pb, nb, &inter, &uni, &total, &la, &lb);
    ints[1] = (int64_t)la;
    ints[2] = (int64_t)lb;
    ints[3] = (int64_t)inter;
    ints[4] = (int64_t)uni;
    return (double)inter / (double)uni; /* strsim.rs:306 */
}

/* ---- Sorensen-Dice, strsim.rs:322-345 ---------------------------------------------------------- */
static double dice_compute(scratch_t *s, const uint8_t *pa, size_t na, const uint8_t *pb, size_t nb,
                           int64_t *ints) {
    if (bytes_equal(pa, na, pb, nb)) {
        ints[0] = 1;
        return 1.0;
    }
    if (na == 0 || nb == 0) {
        ints[0] = 2;
        return 0.0;
    }
    uint64_t inter, uni, total;
    size_t la, lb;
    multiset_fold(s, pa, na, pb, nb, &inter, &uni, &total, &la, &lb);
    ints[1] = (int64_t)la;
    ints[2] = (int64_t)lb;
    ints[3] = (int64_t)inter;
    ints[4] = (int64_t)total;
    return 2.0 * (double)inter / (double)(total + 0); /* strsim.rs:343: f[2] stays 0 */
}

static double compute_one(int measure, scratch_t *s, const uint8_t *a, size_t na, const uint8_t *b,
                          size_t nb, int64_t *ints) {
    memset(ints, 0, N_INTS * sizeof(int64_t));
    switch (measure) {
        case M_LEVENSHTEIN: return lev_compute(s, a, na, b, nb, ints);
        case M_JARO: return jaro_compute(s, a, na, b, nb, ints);
        case M_JARO_WINKLER: return jw_compute(s, a, na, b, nb, ints);
        case M_JACCARD: return jaccard_compute(s, a, na, b, nb, ints);
        default: return dice_compute(s, a, na, b, nb, ints);
    }
}

/* single pair; ints may be NULL */
ORACLE_API double oracle_pair(int measure, const uint8_t *a, size_t na, const uint8_t *b, size_t nb,
                              int64_t *ints) {
    scratch_t s;
    int64_t local[N_INTS];
    scratch_init(&s);
    double r = compute_one(measure, &s, a, na, b, nb, ints ? ints : local);
    scratch_free(&s);
    return r;
}

/* ---- Arrow Utf8View column (one chunk), the layout Polars hands the plugin ----- */
typedef struct {
    const uint8_t *views;        /* 16 B per row, already at row 0 of the logical array */
    const uint8_t *const *bufs;  /* variadic data buffers */
    const uint8_t *validity;     /* LSB-first bitmap or NULL */
    int64_t validity_offset;     /* bit offset of row 0 */
    int64_t length;              /* rows; 1 = scalar broadcast (strsim.rs:61-66) */
} oracle_column;

static inline const uint8_t *view_bytes(const oracle_column *c, int64_t row, size_t *len) {
    const uint8_t *v = c->views + 16 * row;
    int32_t n, buf, off;
    memcpy(&n, v, 4);
    *len = (size_t)n;
    if (n <= 12) return v + 4;
    memcpy(&buf, v + 8, 4);
    memcpy(&off, v + 12, 4);
    return c->bufs[buf] + off;
}

static inline int col_valid(const oracle_column *c, int64_t row) {
    if (!c->validity) return 1;
    int64_t bit = c->validity_offset + row;
    return (c->validity[bit >> 3] >> (bit & 7)) & 1;
}

typedef struct {
    int measure;
    const oracle_column *a, *b;
    int64_t lo, hi;
    double *out;
    uint8_t *out_valid; /* one byte per row */
    int64_t *ints;      /* may be NULL */
} job_t;

static void *run_range(void *arg) {
    job_t *j = arg;
    scratch_t s; /* one measure object per task, strsim.rs:78-84 */
    scratch_init(&s);
    int64_t local[N_INTS];
    for (int64_t r = j->lo; r < j->hi; r++) {
        int64_t ra = j->a->length == 1 ? 0 : r, rb = j->b->length == 1 ? 0 : r;
        int valid = col_valid(j->a, ra) && col_valid(j->b, rb); /* arity kernels: AND of validities */
        j->out_valid[r] = (uint8_t)valid;
        int64_t *ints = j->ints ? j->ints + N_INTS * r : local;
        if (!valid) {
            j->out[r] = 0.0;
            memset(ints, 0, N_INTS * sizeof(int64_t));
            continue;
        }
        size_t na, nb;
        const uint8_t *pa = view_bytes(j->a, ra, &na);
        const uint8_t *pb = view_bytes(j->b, rb, &nb);
        j->out[r] = compute_one(j->measure, &s, pa, na, pb, nb, ints);
    }
    scratch_free(&s);
    return NULL;
}

/* Row-wise evaluation over two Utf8View columns with the reference's static partition: n contiguous
 * ranges of len/n rows, the last takes the remainder (split_offsets, strsim.rs:21-39), one thread
 * each (strsim.rs:72-100).  Returns 0, or -1 on the reference's ShapeMismatch (strsim.rs:48-52). */
ORACLE_API int oracle_batch_views(int measure, const oracle_column *a, const oracle_column *b,
                                  double *out, uint8_t *out_valid, int64_t *ints, int n_threads) {
    if (a->length != b->length && a->length != 1 && b->length != 1) return -1;
    int64_t n = a->length > b->length ? a->length : b->length;
    if (a->length == 1 && b->length == 1) n = 1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1 || n < n_threads) {
        job_t j = {measure, a, b, 0, n, out, out_valid, ints};
        run_range(&j);
        return 0;
    }
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    job_t *jobs = malloc(sizeof(job_t) * n_threads);
    int64_t chunk = n / n_threads;
    for (int p = 0; p < n_threads; p++) {
        int64_t off = p * chunk;
        int64_t len = p == n_threads - 1 ? n - off : chunk;
        jobs[p] = (job_t){measure, a, b, off, off + len, out, out_valid, ints};
        pthread_create(&th[p], NULL, run_range, &jobs[p]);
    }
    for (int p = 0; p < n_threads; p++) pthread_join(th[p], NULL);
    free(th);
    free(jobs);
    return 0;
}

/* Convenience layout for tests: concatenated bytes + int64 offsets (n+1) + one validity byte per
 * row (or NULL).  Single thread. */
ORACLE_API void oracle_batch_offsets(int measure, int64_t n, const uint8_t *a_data,
                                     const int64_t *a_off, const uint8_t *a_valid,
                                     const uint8_t *b_data, const int64_t *b_off,
                                     const uint8_t *b_valid, double *out, uint8_t *out_valid,
                                     int64_t *ints) {
    scratch_t s;
    scratch_init(&s);
    int64_t local[N_INTS];
    for (int64_t r = 0; r < n; r++) {
        int valid = (!a_valid || a_valid[r]) && (!b_valid || b_valid[r]);
        out_valid[r] = (uint8_t)valid;
        int64_t *pi = ints ? ints + N_INTS * r : local;
        if (!valid) {
            out[r] = 0.0;
            memset(pi, 0, N_INTS * sizeof(int64_t));
            continue;
        }
        out[r] = compute_one(measure, &s, a_data + a_off[r], (size_t)(a_off[r + 1] - a_off[r]),
                             b_data + b_off[r], (size_t)(b_off[r + 1] - b_off[r]), pi);
    }
    scratch_free(&s);
}

ORACLE_API int oracle_n_ints(void) { return N_INTS; }
