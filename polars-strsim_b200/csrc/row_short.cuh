// row_short.cuh -- one (a, b) pair of short strings -> f64 similarity + integer intermediates.
//
// "Short" = both strings at most bits(M) BYTES (hence at most bits(M) codepoints), M = uint32_t or
// uint64_t.  The pair's bytes arrive zero-padded in 4-byte words through a `Store`, which also
// provides the scratch the algorithms need (all-zero between rows):
//     wa(k), wb(k)   k in [0,bits(M)/4)   the bytes of a and b, little-endian words
//     tab(c)                              position-mask table indexed by an ASCII byte
//     hkey(s), hmask(s)  s in [0,2*bits(M))  open-addressing hash slots (key+1, position mask) for
//                                         non-ASCII pairs; they may alias the table's memory
// short_kernel.cuh backs Store with per-thread shared-memory slabs laid out [slot][thread] so that a
// warp's accesses never bank-conflict; tests back it with plain arrays on the host.
//
// Two code paths, chosen per pair, sharing the step functors of pair_algos.cuh:
//   * ASCII (every byte < 0x80): characters are bytes, a position mask is one table load.
//   * Unicode: characters are decoded on the fly and keyed by their packed UTF-8 bytes (injective for
//     valid UTF-8, so key equality == equality of Unicode scalar values, which is all the measures
//     use; strsim.rs:133,189,261,297 iterate `chars()`); position masks live in a small per-pair hash
//     table; codepoint counts come from a SWAR count of continuation bytes.
//
// Row rules follow /root/reference/src/expressions/strsim.rs:128,182-186,197-199,260-270,288-292.
#pragma once
#include "pair_algos.cuh"

namespace strsim {

// ---- adapters ---------------------------------------------------------------------------------------
template <class Store>
struct StoreWords {
    const Store& s;
    bool second;  // false: a, true: b
    SS_HD StoreWords(const Store& s_, bool second_) : s(s_), second(second_) {}
    SS_HD uint32_t operator()(int w) const { return second ? s.wb(w) : s.wa(w); }
};

template <class M, class Store>
struct StoreTable {
    Store& s;
    SS_HD explicit StoreTable(Store& s_) : s(s_) {}
    SS_HD M& operator()(uint32_t c) { return s.tab(c); }
    SS_HD const M& operator()(uint32_t c) const { return s.tab(c); }
};

// ---- UTF-8 helpers ------------------------------------------------------------------------------------
// number of characters = bytes that are not continuation bytes (10xxxxxx); padding bytes are zero
template <class Src>
SS_HD int count_chars(const Src& src, int nbytes) {
    int cont = 0;
    for (int w = 0; 4 * w < nbytes; w++) {
        const uint32_t x = src(w);
        cont += popc(((x >> 7) & ~(x >> 6)) & 0x01010101u);
    }
    return nbytes - cont;
}

// Calls f(key) for the first max_chars characters of a UTF-8 string of nbytes bytes; key = the
// character's 1-4 bytes packed little-endian (any injective packing would do).  Branch-free per
// character so that the 32 pairs of a warp stay converged: a 4-byte window at the current byte
// position is funnel-shifted out of two consecutive words, the sequence length comes from a 2-bit
// lookup on the lead byte's high nibble.  last_word: highest word index that may be read.
template <class Src, class F>
SS_HD void for_each_char(const Src& src, int nbytes, int max_chars, int last_word, F& f) {
    int pos = 0, done = 0;
    while (pos < nbytes && done < max_chars) {
        const int wi = pos >> 2;
        const uint32_t w0 = src(wi);
        const uint32_t w1 = src(wi < last_word ? wi + 1 : last_word);
        const int sh = (pos & 3) * 8;
#if defined(__CUDA_ARCH__)
        const uint32_t win = __funnelshift_r(w0, w1, sh);
#else
        const uint32_t win = (uint32_t)((((uint64_t)w1 << 32) | w0) >> sh);
#endif
        const uint32_t lead = win & 0xFFu;
        // extra bytes: 0 for 0x00-0xBF, 1 for 0xC0-0xDF, 2 for 0xE0-0xEF, 3 for 0xF0-0xFF
        uint32_t extra = (0xE5000000u >> ((lead >> 4) * 2)) & 3u;
        if ((int)extra > nbytes - pos - 1) extra = (uint32_t)(nbytes - pos - 1);  // malformed tail
        const uint32_t key = win & (0xFFFFFFFFu >> (24 - 8 * extra));
        pos += 1 + (int)extra;
        f(key);
        done++;
    }
}

// ---- per-pair hash table of position masks (Unicode path) ------------------------------------------
// 2*bits(M) slots, linear probing, at most bits(M) distinct keys => load <= 1/2.  `used` remembers
// the occupied slots so that clearing touches only those.
template <class M, class Store>
struct HashTab {
    static constexpr int CAP = (int)sizeof(M) * 8;
    static constexpr int SLOTS = 2 * CAP;
    static constexpr int SHIFT = CAP == 32 ? 26 : 25;  // 32 - log2(SLOTS)
    Store& s;
    uint64_t used_lo, used_hi;  // bitmap of occupied slots (used_hi only for 128 slots)
    SS_HD explicit HashTab(Store& s_) : s(s_), used_lo(0), used_hi(0) {}
    SS_HD static uint32_t home(uint32_t key) { return (key * 0x9E3779B1u) >> SHIFT; }
    SS_HD M operator()(uint32_t key) const {  // position mask of `key` (0 when absent)
        uint32_t slot = home(key);
        for (;;) {
            const uint32_t k = s.hkey(slot);
            if (k == key + 1u) return s.hmask(slot);
            if (k == 0u) return M(0);
            slot = (slot + 1u) & (uint32_t)(SLOTS - 1);
        }
    }
    SS_HD void add(uint32_t key, M bit) {
        uint32_t slot = home(key);
        for (;;) {
            const uint32_t k = s.hkey(slot);
            if (k == key + 1u) {
                s.hmask(slot) = s.hmask(slot) | bit;
                return;
            }
            if (k == 0u) {
                s.hkey(slot) = key + 1u;
                s.hmask(slot) = bit;
                if (slot < 64u)
                    used_lo |= 1ull << slot;
                else
                    used_hi |= 1ull << (slot - 64u);
                return;
            }
            slot = (slot + 1u) & (uint32_t)(SLOTS - 1);
        }
    }
    SS_HD void clear() {
        while (used_lo) {
            const int slot = ctz64(used_lo);
            used_lo &= used_lo - 1ull;
            s.hkey(slot) = 0u;
            s.hmask(slot) = M(0);
        }
        while (used_hi) {
            const int slot = 64 + ctz64(used_hi);
            used_hi &= used_hi - 1ull;
            s.hkey(slot) = 0u;
            s.hmask(slot) = M(0);
        }
    }
};

template <class M, class Tab>
struct HashBuild {
    Tab& tab;
    M bit;
    SS_HD explicit HashBuild(Tab& t) : tab(t), bit(M(1)) {}
    SS_HD void operator()(uint32_t key) {
        tab.add(key, bit);
        bit = bit << 1;
    }
};

struct PrefixKeys {  // first (up to) four character keys of a string
    uint32_t k[4];
    int n;
    SS_HD PrefixKeys() : n(0) { k[0] = k[1] = k[2] = k[3] = 0; }
    SS_HD void operator()(uint32_t key) {
        if (n < 4) k[n] = key;
        n++;
    }
};

// ---- the measure-specific part, shared by both paths -----------------------------------------------------
// each_streamed(n, f): applies f to the first n characters of the streamed string (ASCII: bytes;
// Unicode: decoded keys).  `tab`: position masks of the tabled string (Levenshtein: the shorter
// one; the other measures: b, with a streamed as in strsim.rs:208).
template <class M, bool SMALL = false, class Tab, class Each>
SS_HD double measure_body(int measure, const Tab& tab, const Each& each_streamed, int la, int lb, int n_tab,
                          int n_str, PairInts& out) {
    switch (measure) {
        case LEVENSHTEIN: {
            int d = n_str;
            if (n_tab > 0) {
                MyersStep<M, Tab> step(tab);
                each_streamed(n_str, step);
                d = step.distance(n_tab, n_str);
            }
            out.x0 = d;
            return lev_value<SMALL>(d, la, lb);
        }
        case JARO:
        case JARO_WINKLER: {
            const int mx = la > lb ? la : lb;
            const int bound = mx / 2 - 1;  // strsim.rs:200
            const int outer = la < lb + bound ? la : lb + bound;
            JaroMatchStep<M, Tab> match(tab, lb, bound);
            each_streamed(outer, match);
            match.finish(outer);
            JaroTransStep<M, Tab> trans(tab, match.flag_a, match.flag_b);
            if (match.m > 0) each_streamed(outer, trans);
            out.x0 = match.m;
            out.x1 = trans.t;
            return match.m == 0 ? 0.0 : jaro_value<SMALL>(match.m, trans.t, la, lb);
        }
        default: {
            MultisetStep<M, Tab> ms(tab, lb);
            each_streamed(la, ms);
            ms.finish();
            out.x0 = ms.inter;
            if (measure == JACCARD) {
                out.x1 = la + lb - ms.inter;  // sum_c max = la + lb - sum_c min
                return jaccard_value<SMALL>(ms.inter, la + lb - ms.inter);
            }
            out.x1 = la + lb;
            return dice_value<SMALL>(ms.inter, la + lb);
        }
    }
}

// ---- fused evaluation (pair_algos.cuh: FusedStep) ----------------------------------------------------------
// emit(measure, value, ints) is called once for every measure of GROUPS; the caller decides which of
// them it keeps (short_kernel.cuh stores those whose output pointer is set).
template <int GROUPS, class Emit>
SS_HD void emit_groups(Emit& emit, double v, const PairInts& o) {
    if (GROUPS & G_LEV) emit(LEVENSHTEIN, v, o);
    if (GROUPS & G_JARO) {
        emit(JARO, v, o);
        emit(JARO_WINKLER, v, o);
    }
    if (GROUPS & G_SET) {
        emit(JACCARD, v, o);
        emit(SORENSEN_DICE, v, o);
    }
}

// The pair is NOT byte-equal.  tab: position masks of b (lb characters); each_a(n, f) applies f to the
// first n characters of a; prefix() = length of the common character prefix, capped at 4
// (strsim.rs:261-266); one_empty: either string is empty (strsim.rs:184,290,326).  The integer records
// and values are those of measure_body() / row_short() for each single measure.
template <int GROUPS, class M, class Tab, class Each, class Prefix, class Trans, class Emit>
SS_HD void multi_body(const Tab& tab, const Each& each_a, int la, int lb, bool one_empty, const Prefix& prefix,
                      const Trans& trans_count, Emit& emit) {
    const int mx = la > lb ? la : lb;
    int bound = mx / 2 - 1;  // strsim.rs:200
    if (bound < 0) bound = 0;  // both strings of at most one character: settled by the row rules below
    FusedStep<GROUPS, M, Tab> f(tab, lb, bound);
    each_a(la, f);  // Jaro's outer limit min(la, lb+bound) needs no cut: later windows lie beyond b
    if (GROUPS & G_JARO) f.jm.finish(la);
    if (GROUPS & G_SET) f.ms.finish();
    PairInts o;
    o.flag = F_GENERAL;
    o.la = la;
    o.lb = lb;
    o.x0 = o.x1 = o.x2 = 0;
    if (GROUPS & G_LEV) {
        // b tabled whatever the lengths; an empty side needs no special case: no column (la = 0) leaves
        // Pv all ones -> d = lb, an empty pattern (lb = 0) scores no vertical delta -> d = la
        const int d = f.my.distance(lb, la);
        o.x0 = d;
        emit(LEVENSHTEIN, lev_value<(sizeof(M) <= 8)>(d, la, lb), o);  // (the quotient table ends at 64)
    }
    PairInts z;
    z.flag = F_ONE_EMPTY;
    z.la = z.lb = z.x0 = z.x1 = z.x2 = 0;
    if (GROUPS & G_JARO) {
        if (one_empty) {
            emit(JARO, 0.0, z);
            emit(JARO_WINKLER, 0.0, z);
        } else if (la == 1 && lb == 1) {  // strsim.rs:197; the bytes differ here
            o.flag = F_SINGLE_CHAR;
            o.x0 = 0;
            emit(JARO, 0.0, o);
            emit(JARO_WINKLER, 0.0, o);
            o.flag = F_GENERAL;
        } else {
            const int t = f.jm.m > 0 ? trans_count(tab, each_a, la, f.jm.flag_a, f.jm.flag_b) : 0;
            o.x0 = f.jm.m;
            o.x1 = t;
            double v = f.jm.m == 0 ? 0.0 : jaro_value<(sizeof(M) <= 8)>(f.jm.m, t, la, lb);
            emit(JARO, v, o);
            if (v > 0.7) {  // strsim.rs:260-267
                const int l = prefix();
                o.x2 = l;
                v = winkler_value(v, l);
            }
            emit(JARO_WINKLER, v, o);
            o.x2 = 0;
        }
    }
    if (GROUPS & G_SET) {
        if (one_empty) {
            emit(JACCARD, 0.0, z);
            emit(SORENSEN_DICE, 0.0, z);
        } else {
            const int inter = f.ms.inter;
            o.x0 = inter;
            o.x1 = la + lb - inter;
            emit(JACCARD, jaccard_value<sizeof(M) == 4>(inter, la + lb - inter), o);
            o.x1 = la + lb;
            emit(SORENSEN_DICE, dice_value<sizeof(M) == 4>(inter, la + lb), o);
        }
    }
}

template <class Store>
struct EachByte {
    StoreWords<Store> src;
    SS_HD EachByte(const Store& s, bool second) : src(s, second) {}
    template <class F>
    SS_HD void operator()(int n, F& f) const {
        for_each_byte(src, n, f);
    }
};

template <class Store>
struct EachChar {
    StoreWords<Store> src;
    int nbytes;
    SS_HD EachChar(const Store& s, bool second, int nbytes_) : src(s, second), nbytes(nbytes_) {}
    template <class F>
    SS_HD void operator()(int n, F& f) const {
        for_each_char(src, nbytes, n, (int)sizeof(typename Store::mask_type) * 2 - 1, f);
    }
};

// na, nb: byte lengths (<= bits(M)); equal: bytes identical; ascii: no byte >= 0x80 in either.
template <class M, class Store>
SS_HD double row_short(int measure, Store& s, int na, int nb, bool equal, bool ascii, PairInts& out) {
    out.flag = F_GENERAL;
    out.la = out.lb = out.x0 = out.x1 = out.x2 = 0;
    if (equal) {  // strsim.rs:128,182,288,324 (covers empty/empty)
        out.flag = F_EQUAL;
        return 1.0;
    }
    if (measure != LEVENSHTEIN && (na == 0 || nb == 0)) {  // strsim.rs:184,290,326
        out.flag = F_ONE_EMPTY;
        return 0.0;
    }
    const bool is_jaro = measure == JARO || measure == JARO_WINKLER;
    constexpr int LAST_WORD = (int)sizeof(M) * 2 - 1;  // bits(M)/4 words per string
    double v;
    if (ascii) {
        const int la = na, lb = nb;
        out.la = la;
        out.lb = lb;
        if (is_jaro && la == 1 && lb == 1) {  // strsim.rs:197
            out.flag = F_SINGLE_CHAR;
            return 0.0;  // the bytes differ here
        }
        // Levenshtein tables the shorter string (fewer table writes; extra text steps are cheap)
        const bool table_b = measure != LEVENSHTEIN || lb <= la;
        const int n_tab = table_b ? lb : la;
        StoreWords<Store> tabled(s, table_b);
        StoreTable<M, Store> tab(s);
        {
            BuildTable<M, StoreTable<M, Store>> build(tab);
            for_each_byte(tabled, (n_tab + 3) & ~3, build);
        }
        EachByte<Store> streamed(s, !table_b);
        v = measure_body<M>(measure, tab, streamed, la, lb, n_tab, table_b ? la : lb, out);
        {
            ClearTable<M, StoreTable<M, Store>> clear(tab);
            for_each_byte(tabled, (n_tab + 3) & ~3, clear);
        }
        if (measure == JARO_WINKLER && v > 0.7) {  // strsim.rs:260-267
            const uint32_t x = s.wa(0) ^ s.wb(0);
            int lim = la < lb ? la : lb;
            if (lim > 4) lim = 4;
            int l = 0;
            while (l < lim && ((x >> (8 * l)) & 0xFFu) == 0) l++;
            out.x2 = l;
            v = winkler_value(v, l);
        }
    } else {
        StoreWords<Store> wa(s, false), wb(s, true);
        const int la = count_chars(wa, na), lb = count_chars(wb, nb);
        out.la = la;
        out.lb = lb;
        if (is_jaro && la == 1 && lb == 1) {
            out.flag = F_SINGLE_CHAR;
            return 0.0;  // one character each and the bytes differ
        }
        const bool table_b = measure != LEVENSHTEIN || lb <= la;
        const int n_tab = table_b ? lb : la;
        HashTab<M, Store> tab(s);
        {
            HashBuild<M, HashTab<M, Store>> build(tab);
            for_each_char(table_b ? wb : wa, table_b ? nb : na, n_tab, LAST_WORD, build);
        }
        EachChar<Store> streamed(s, !table_b, table_b ? na : nb);
        v = measure_body<M>(measure, tab, streamed, la, lb, n_tab, table_b ? la : lb, out);
        tab.clear();
        if (measure == JARO_WINKLER && v > 0.7) {
            PrefixKeys pa, pb;
            for_each_char(wa, na, 4, LAST_WORD, pa);
            for_each_char(wb, nb, 4, LAST_WORD, pb);
            int lim = pa.n < pb.n ? pa.n : pb.n;
            int l = 0;
            while (l < lim && pa.k[l] == pb.k[l]) l++;
            out.x2 = l;
            v = winkler_value(v, l);
        }
    }
    return v;
}

}  // namespace strsim
