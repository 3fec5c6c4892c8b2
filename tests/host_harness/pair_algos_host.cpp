// Host build of the product's per-pair templates (polars-strsim_b200/csrc/row_short.cuh) so that
// tests/test_pair_algos.py can check them against the oracle without a GPU.  Test code only.
#include <cstdint>
#include <cstring>

#include "row_short.cuh"

using namespace strsim;

template <class M>
struct HostStore {
    using mask_type = M;
    static constexpr int CAP = (int)sizeof(M) * 8;
    M table[128];
    uint32_t keys[2 * CAP];
    M masks[2 * CAP];
    uint32_t a_words[CAP / 4], b_words[CAP / 4];
    M& tab(uint32_t c) { return table[c]; }
    const M& tab(uint32_t c) const { return table[c]; }
    uint32_t& hkey(int s) { return keys[s]; }
    const uint32_t& hkey(int s) const { return keys[s]; }
    M& hmask(int s) { return masks[s]; }
    const M& hmask(int s) const { return masks[s]; }
    uint32_t wa(int k) const { return a_words[k]; }
    uint32_t wb(int k) const { return b_words[k]; }
};

template <class M>
static int run(int measure, const uint8_t* a, int na, const uint8_t* b, int nb, int force_unicode,
               int* ints, double* value) {
    static thread_local HostStore<M> s;  // table must stay all-zero between rows
    if (na > HostStore<M>::CAP || nb > HostStore<M>::CAP) return -2;
    std::memset(s.a_words, 0, sizeof s.a_words);
    std::memset(s.b_words, 0, sizeof s.b_words);
    std::memcpy(s.a_words, a, na);
    std::memcpy(s.b_words, b, nb);
    bool equal = na == nb && std::memcmp(a, b, na) == 0;
    bool ascii = !force_unicode;
    for (int i = 0; i < na; i++) ascii = ascii && a[i] < 0x80;
    for (int i = 0; i < nb; i++) ascii = ascii && b[i] < 0x80;
    PairInts pi;
    *value = row_short<M>(measure, s, na, nb, equal, ascii, pi);
    ints[0] = pi.flag; ints[1] = pi.la; ints[2] = pi.lb; ints[3] = pi.x0; ints[4] = pi.x1; ints[5] = pi.x2;
    for (int c = 0; c < 128; c++)
        if (s.table[c] != 0) return -1;  // invariant broken
    for (int c = 0; c < 2 * HostStore<M>::CAP; c++)
        if (s.keys[c] != 0 || s.masks[c] != 0) return -1;
    return 0;
}

extern "C" int algos_row(int measure, int bits, const uint8_t* a, int na, const uint8_t* b, int nb,
                         int force_unicode, int* ints, double* value) {
    return bits == 32 ? run<uint32_t>(measure, a, na, b, nb, force_unicode, ints, value)
                      : run<uint64_t>(measure, a, na, b, nb, force_unicode, ints, value);
}

// batch form: concatenated bytes + offsets; returns the number of rows whose invariant broke
extern "C" int algos_batch(int measure, int bits, int64_t n, const uint8_t* ad, const int64_t* ao,
                           const uint8_t* bd, const int64_t* bo, int force_unicode, int* ints,
                           double* values) {
    int bad = 0;
    for (int64_t r = 0; r < n; r++) {
        int rc = algos_row(measure, bits, ad + ao[r], (int)(ao[r + 1] - ao[r]), bd + bo[r],
                           (int)(bo[r + 1] - bo[r]), force_unicode, ints + 6 * r, values + r);
        if (rc != 0) bad++;
    }
    return bad;
}

// multi-word Myers with the product's block step, blocks chained sequentially per column
#include <unordered_map>
#include <vector>
extern "C" int algos_myers_multiword(const uint32_t* a, int la, const uint32_t* b, int lb) {
    const uint32_t* pat = la <= lb ? a : b;
    const uint32_t* txt = la <= lb ? b : a;
    const int m = la <= lb ? la : lb, n = la <= lb ? lb : la;
    if (m == 0) return n;
    const int W = (m + 63) / 64;
    std::unordered_map<uint32_t, std::vector<uint64_t>> peq;
    for (int i = 0; i < m; i++) {
        auto& row = peq[pat[i]];
        if (row.empty()) row.assign(W, 0);
        row[i / 64] |= 1ull << (i % 64);
    }
    std::vector<uint64_t> Pv(W, ~0ull), Mv(W, 0), zero(W, 0);
    for (int j = 0; j < n; j++) {
        auto it = peq.find(txt[j]);
        const std::vector<uint64_t>& eq = it == peq.end() ? zero : it->second;
        uint32_t hp = 1, hm = 0;
        for (int w = 0; w < W; w++) myers_block(Pv[w], Mv[w], eq[w], hp, hm);
    }
    int d = n;
    for (int w = 0; w < W; w++) d += myers_block_score(Pv[w], Mv[w], m - 64 * w);
    return d;
}

// register-resident ASCII path (row_ascii_reg.cuh): planes instead of tables
#include "row_ascii_reg.cuh"
struct HostByteAt {  // stands in for short_kernel.cuh's SmemByteAt
    const uint8_t* p;
    uint32_t operator()(int i) const { return p[i]; }
};
typedef SlabSrc<1, HostByteAt> HostSrc;
static HostSrc host_src(const uint32_t* w, int len) { return HostSrc{w, len, HostByteAt{reinterpret_cast<const uint8_t*>(w)}}; }
template <int MEASURE>
static double planes_dispatch(int nbits, const uint32_t* a, const uint32_t* b, int na, int nb, PairInts& pi) {
    const HostSrc A = host_src(a, na), B = host_src(b, nb);
    switch (nbits) {
        case 5: return row_planes<MEASURE, 5>(A, B, pi);
        case 6: return row_planes<MEASURE, 6>(A, B, pi);
        case 7: return row_planes<MEASURE, 7>(A, B, pi);
        default: return row_planes<MEASURE, 8>(A, B, pi);
    }
}
template <int MEASURE>
static double planes_dispatch64(const uint32_t* a, const uint32_t* b, int na, int nb, PairInts& pi) {
    return row_planes<MEASURE, 7, uint64_t>(host_src(a, na), host_src(b, nb), pi);  // strings of up to 64 characters
}
extern "C" int algos_batch_reg(int measure, int nbits, int64_t n, const uint8_t* ad, const int64_t* ao,
                               const uint8_t* bd, const int64_t* bo, int* ints, double* values) {
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        const bool wide64 = nbits >= 100;  // 107: 64-bit masks, 7 planes, strings of up to 64 characters
        if (na > (wide64 ? 64 : 32) || nb > (wide64 ? 64 : 32)) return -2;
        uint32_t a[2 * REG_WORDS] = {0}, b[2 * REG_WORDS] = {0};
        std::memcpy(a, ad + ao[r], na);
        std::memcpy(b, bd + bo[r], nb);
        PairInts pi;
        double v;
        {
            // the kernels' form: strings behind sources, pair known not to be byte-equal (row_planes)
            const bool equal = na == nb && std::memcmp(a, b, sizeof a) == 0;
            if (equal) {
                pi = {F_EQUAL, 0, 0, 0, 0, 0};
                v = 1.0;
            } else {
                // bytes past the string are arbitrary for a source: poison them
                std::memset(reinterpret_cast<uint8_t*>(a) + na, 0xA5, sizeof a - na);
                std::memset(reinterpret_cast<uint8_t*>(b) + nb, 0x5A, sizeof b - nb);
                if (wide64) switch (measure) {
                    case 0: v = planes_dispatch64<0>(a, b, na, nb, pi); break;
                    case 1: v = planes_dispatch64<1>(a, b, na, nb, pi); break;
                    case 2: v = planes_dispatch64<2>(a, b, na, nb, pi); break;
                    case 3: v = planes_dispatch64<3>(a, b, na, nb, pi); break;
                    default: v = planes_dispatch64<4>(a, b, na, nb, pi); break;
                } else
                switch (measure) {
                    case 0: v = planes_dispatch<0>(nbits, a, b, na, nb, pi); break;
                    case 1: v = planes_dispatch<1>(nbits, a, b, na, nb, pi); break;
                    case 2: v = planes_dispatch<2>(nbits, a, b, na, nb, pi); break;
                    case 3: v = planes_dispatch<3>(nbits, a, b, na, nb, pi); break;
                    default: v = planes_dispatch<4>(nbits, a, b, na, nb, pi); break;
                }
            }
        }
        values[r] = v;
        int* o = ints + 6 * r;
        o[0] = pi.flag; o[1] = pi.la; o[2] = pi.lb; o[3] = pi.x0; o[4] = pi.x1; o[5] = pi.x2;
    }
    return 0;
}

// register-compare path for any script (row_unicode_reg.cuh)
#include "row_unicode_reg.cuh"
extern "C" int algos_batch_ureg(int measure, int64_t n, const uint8_t* ad, const int64_t* ao, const uint8_t* bd,
                                const int64_t* bo, int* ints, double* values) {
    static thread_local HostStore<uint32_t> s;
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        if (na > 32 || nb > 32) return -2;
        std::memset(s.a_words, 0, sizeof s.a_words);
        std::memset(s.b_words, 0, sizeof s.b_words);
        std::memcpy(s.a_words, ad + ao[r], na);
        std::memcpy(s.b_words, bd + bo[r], nb);
        const bool equal = na == nb && std::memcmp(ad + ao[r], bd + bo[r], na) == 0;
        PairInts pi;
        HostWarpMax wm;
        double v;
        switch (measure) {
            case 0: v = row_unicode_reg<0>(s, na, nb, equal, wm, pi); break;
            case 1: v = row_unicode_reg<1>(s, na, nb, equal, wm, pi); break;
            case 2: v = row_unicode_reg<2>(s, na, nb, equal, wm, pi); break;
            case 3: v = row_unicode_reg<3>(s, na, nb, equal, wm, pi); break;
            default: v = row_unicode_reg<4>(s, na, nb, equal, wm, pi); break;
        }
        values[r] = v;
        int* o = ints + 6 * r;
        o[0] = pi.flag; o[1] = pi.la; o[2] = pi.lb; o[3] = pi.x0; o[4] = pi.x1; o[5] = pi.x2;
    }
    return 0;
}

// fused evaluation of several measures (row_*_multi): values / ints are [5][n] / [5][n][6], measures
// outside `groups` are left untouched
struct HostEmit {
    int64_t r, n;
    int* ints;
    double* values;
    void operator()(int measure, double v, const PairInts& pi) {
        values[measure * n + r] = v;
        int* o = ints + (measure * n + r) * 6;
        o[0] = pi.flag; o[1] = pi.la; o[2] = pi.lb; o[3] = pi.x0; o[4] = pi.x1; o[5] = pi.x2;
    }
};

template <int GROUPS>
static void reg_multi_dispatch(int nbits, uint32_t (&a)[2 * REG_WORDS], uint32_t (&b)[2 * REG_WORDS], int na,
                               int nb, HostEmit& e) {
    // the kernels' form (row_planes_multi over sources), equal pairs settled before
    if (na == nb && std::memcmp(a, b, sizeof a) == 0) {
        const PairInts o = {F_EQUAL, 0, 0, 0, 0, 0};
        emit_groups<GROUPS>(e, 1.0, o);
        return;
    }
    std::memset(reinterpret_cast<uint8_t*>(a) + na, 0xA5, sizeof a - na);
    std::memset(reinterpret_cast<uint8_t*>(b) + nb, 0x5A, sizeof b - nb);
    const HostSrc A = host_src(a, na), B = host_src(b, nb);
    if (nbits >= 100) {
        row_planes_multi<GROUPS, 7, uint64_t>(A, B, e);
        return;
    }
    switch (nbits) {
        case 5: row_planes_multi<GROUPS, 5>(A, B, e); break;
        case 6: row_planes_multi<GROUPS, 6>(A, B, e); break;
        case 7: row_planes_multi<GROUPS, 7>(A, B, e); break;
        default: row_planes_multi<GROUPS, 8>(A, B, e); break;
    }
}

extern "C" int algos_batch_reg_multi(int groups, int nbits, int64_t n, const uint8_t* ad, const int64_t* ao,
                                     const uint8_t* bd, const int64_t* bo, int* ints, double* values) {
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        if (na > (nbits >= 100 ? 64 : 32) || nb > (nbits >= 100 ? 64 : 32)) return -2;
        uint32_t a[2 * REG_WORDS] = {0}, b[2 * REG_WORDS] = {0};
        std::memcpy(a, ad + ao[r], na);
        std::memcpy(b, bd + bo[r], nb);
        HostEmit e{r, n, ints, values};
        switch (groups) {
            case 1: reg_multi_dispatch<1>(nbits, a, b, na, nb, e); break;
            case 2: reg_multi_dispatch<2>(nbits, a, b, na, nb, e); break;
            case 3: reg_multi_dispatch<3>(nbits, a, b, na, nb, e); break;
            case 4: reg_multi_dispatch<4>(nbits, a, b, na, nb, e); break;
            case 5: reg_multi_dispatch<5>(nbits, a, b, na, nb, e); break;
            case 6: reg_multi_dispatch<6>(nbits, a, b, na, nb, e); break;
            case 7: reg_multi_dispatch<7>(nbits, a, b, na, nb, e); break;
            default: return -3;
        }
    }
    return 0;
}

extern "C" int algos_batch_ureg_multi(int groups, int64_t n, const uint8_t* ad, const int64_t* ao, const uint8_t* bd,
                                      const int64_t* bo, int* ints, double* values) {
    static thread_local HostStore<uint32_t> s;
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        if (na > 32 || nb > 32) return -2;
        std::memset(s.a_words, 0, sizeof s.a_words);
        std::memset(s.b_words, 0, sizeof s.b_words);
        std::memcpy(s.a_words, ad + ao[r], na);
        std::memcpy(s.b_words, bd + bo[r], nb);
        const bool equal = na == nb && std::memcmp(ad + ao[r], bd + bo[r], na) == 0;
        HostWarpMax wm;
        HostEmit e{r, n, ints, values};
        switch (groups) {
            case 1: row_unicode_reg_multi<1>(s, na, nb, equal, wm, e); break;
            case 2: row_unicode_reg_multi<2>(s, na, nb, equal, wm, e); break;
            case 3: row_unicode_reg_multi<3>(s, na, nb, equal, wm, e); break;
            case 4: row_unicode_reg_multi<4>(s, na, nb, equal, wm, e); break;
            case 5: row_unicode_reg_multi<5>(s, na, nb, equal, wm, e); break;
            case 6: row_unicode_reg_multi<6>(s, na, nb, equal, wm, e); break;
            case 7: row_unicode_reg_multi<7>(s, na, nb, equal, wm, e); break;
            default: return -3;
        }
    }
    return 0;
}

// Latin-1 rows: transcode to one byte per character, then the 8-plane path (row_ascii_reg.cuh)
struct HostSlab {
    uint32_t* w;
    uint32_t rd(int i) const { return w[i]; }
    void wr(int i, uint32_t v) { w[i] = v; }
};

// returns -4 when a row is not Latin-1 (a byte >= 0xC4)
extern "C" int algos_batch_latin1_multi(int groups, int64_t n, const uint8_t* ad, const int64_t* ao, const uint8_t* bd,
                                        const int64_t* bo, int* ints, double* values) {
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        if (na > 32 || nb > 32) return -2;
        uint32_t a[2 * REG_WORDS] = {0}, b[2 * REG_WORDS] = {0};
        std::memcpy(a, ad + ao[r], na);
        std::memcpy(b, bd + bo[r], nb);
        uint32_t wide = 0;
        for (int w = 0; w < REG_WORDS; w++) wide |= wide_bytes(a[w]) | wide_bytes(b[w]);
        if (wide) return -4;
        const bool equal = na == nb && std::memcmp(a, b, sizeof a) == 0;
        int ca = na, cb = nb;
        if (!equal) {
            HostSlab sa{a}, sb{b};
            ca = transcode_latin1(sa, na);
            cb = transcode_latin1(sb, nb);
            for (int w = 0; w < REG_WORDS; w++) {  // what the kernel does when it loads the registers
                if (4 * w >= ca) a[w] = 0;
                if (4 * w >= cb) b[w] = 0;
            }
        }
        HostEmit e{r, n, ints, values};
        switch (groups) {
            case 1: reg_multi_dispatch<1>(8, a, b, ca, cb, e); break;
            case 2: reg_multi_dispatch<2>(8, a, b, ca, cb, e); break;
            case 3: reg_multi_dispatch<3>(8, a, b, ca, cb, e); break;
            case 4: reg_multi_dispatch<4>(8, a, b, ca, cb, e); break;
            case 5: reg_multi_dispatch<5>(8, a, b, ca, cb, e); break;
            case 6: reg_multi_dispatch<6>(8, a, b, ca, cb, e); break;
            case 7: reg_multi_dispatch<7>(8, a, b, ca, cb, e); break;
            default: return -3;
        }
    }
    return 0;
}

// masks of several words (wide_mask.cuh): ASCII strings of up to 32 * N characters, one pair per thread on
// the device; here the same templates with N = 10 (320 characters) and N = 5, 7 planes
template <int N, int MEASURE>
static double wide_single(const uint32_t* a, const uint32_t* b, int na, int nb, PairInts& pi) {
    return row_planes<MEASURE, 7, Wide<N>>(host_src(a, na), host_src(b, nb), pi);
}
template <int N>
static int wide_batch(int measure, int groups, int64_t n, const uint8_t* ad, const int64_t* ao, const uint8_t* bd,
                      const int64_t* bo, int* ints, double* values) {
    constexpr int CAPB = 32 * N;
    for (int64_t r = 0; r < n; r++) {
        const int na = (int)(ao[r + 1] - ao[r]), nb = (int)(bo[r + 1] - bo[r]);
        if (na > CAPB || nb > CAPB) return -2;
        uint32_t a[CAPB / 4 + 2], b[CAPB / 4 + 2];
        std::memset(a, 0xA5, sizeof a);  // bytes past the string are arbitrary for a source
        std::memset(b, 0x5A, sizeof b);
        std::memcpy(a, ad + ao[r], na);
        std::memcpy(b, bd + bo[r], nb);
        const bool equal = na == nb && std::memcmp(ad + ao[r], bd + bo[r], na) == 0;
        if (groups) {
            HostEmit e{r, n, ints, values};
            const PairInts o = {F_EQUAL, 0, 0, 0, 0, 0};
            const HostSrc A = host_src(a, na), B = host_src(b, nb);
            switch (groups) {
                case 1: if (equal) emit_groups<1>(e, 1.0, o); else row_planes_multi<1, 7, Wide<N>>(A, B, e); break;
                case 2: if (equal) emit_groups<2>(e, 1.0, o); else row_planes_multi<2, 7, Wide<N>>(A, B, e); break;
                case 3: if (equal) emit_groups<3>(e, 1.0, o); else row_planes_multi<3, 7, Wide<N>>(A, B, e); break;
                case 4: if (equal) emit_groups<4>(e, 1.0, o); else row_planes_multi<4, 7, Wide<N>>(A, B, e); break;
                case 5: if (equal) emit_groups<5>(e, 1.0, o); else row_planes_multi<5, 7, Wide<N>>(A, B, e); break;
                case 6: if (equal) emit_groups<6>(e, 1.0, o); else row_planes_multi<6, 7, Wide<N>>(A, B, e); break;
                case 7: if (equal) emit_groups<7>(e, 1.0, o); else row_planes_multi<7, 7, Wide<N>>(A, B, e); break;
                default: return -3;
            }
            continue;
        }
        PairInts pi;
        double v;
        if (equal) {
            pi = {F_EQUAL, 0, 0, 0, 0, 0};
            v = 1.0;
        } else
            switch (measure) {
                case 0: v = wide_single<N, 0>(a, b, na, nb, pi); break;
                case 1: v = wide_single<N, 1>(a, b, na, nb, pi); break;
                case 2: v = wide_single<N, 2>(a, b, na, nb, pi); break;
                case 3: v = wide_single<N, 3>(a, b, na, nb, pi); break;
                default: v = wide_single<N, 4>(a, b, na, nb, pi); break;
            }
        values[r] = v;
        int* o = ints + 6 * r;
        o[0] = pi.flag; o[1] = pi.la; o[2] = pi.lb; o[3] = pi.x0; o[4] = pi.x1; o[5] = pi.x2;
    }
    return 0;
}
// groups == 0: one measure (values[n], ints[n][6]); else the fused form (values[5][n], ints[5][n][6])
extern "C" int algos_batch_wide(int words, int measure, int groups, int64_t n, const uint8_t* ad, const int64_t* ao,
                                const uint8_t* bd, const int64_t* bo, int* ints, double* values) {
    return words == 5 ? wide_batch<5>(measure, groups, n, ad, ao, bd, bo, ints, values)
                      : wide_batch<10>(measure, groups, n, ad, ao, bd, bo, ints, values);
}
