// row_short.cuh -- one (a, b) pair of short strings -> f64 similarity + integer intermediates.
//
// "Short" = both strings at most bits(M) BYTES (hence at most bits(M) codepoints), M = uint32_t or
// uint64_t.  The pair's bytes arrive zero-padded in 4-byte words through a `Store`, which also
// provides the scratch the algorithms need:
//     tab(c)  c in [0,128)   position-mask table for ASCII strings (all-zero between rows)
//     cp(s)   s in [0,2*bits(M))  codepoint keys of a (first half) and b (second half)
//     wa(k), wb(k)  k in [0,bits(M)/4)  the bytes of a and b, little-endian words
// short_kernel.cuh backs Store with per-thread shared-memory slabs laid out [slot][thread] so that a
// warp's accesses never bank-conflict; tests back it with plain arrays on the host.
//
// Two code paths, chosen per pair:
//   * ASCII (every byte < 0x80): characters are bytes, PM lookups are one table load.
//   * Unicode: each character is keyed by its packed UTF-8 bytes (injective for valid UTF-8, so
//     equality of keys == equality of Unicode scalar values, which is all the measures use;
//     strsim.rs:133,189,261,297 iterate `chars()`), PM lookups scan the stored keys.
//
// Row rules follow /root/reference/src/expressions/strsim.rs:128,182-186,197-199,260-270,288-292.
#pragma once
#include "pair_algos.cuh"

namespace strsim {

template <class Store>
struct ByteReader {
    const Store& s;
    bool second;  // false: a, true: b
    int j;
    uint32_t cur;
    SS_HD ByteReader(const Store& s_, bool second_) : s(s_), second(second_), j(0), cur(0) {}
    SS_HD uint32_t next() {
        if ((j & 3) == 0) cur = second ? s.wb(j >> 2) : s.wa(j >> 2);
        uint32_t c = cur & 0xFFu;
        cur >>= 8;
        j++;
        return c;
    }
};

template <class Store>
struct CpReader {
    const Store& s;
    int pos;
    SS_HD CpReader(const Store& s_, int base) : s(s_), pos(base) {}
    SS_HD uint32_t next() { return s.cp(pos++); }
};

template <class M, class Store>
struct TablePM {
    const Store& s;
    SS_HD explicit TablePM(const Store& s_) : s(s_) {}
    SS_HD M operator()(uint32_t c) const { return s.tab(c); }
};

template <class M, class Store>
struct ScanPM {
    const Store& s;
    int base, len;
    SS_HD ScanPM(const Store& s_, int base_, int len_) : s(s_), base(base_), len(len_) {}
    SS_HD M operator()(uint32_t c) const {
        M r = M(0);
        for (int k = 0; k < len; k++) r |= M(s.cp(base + k) == c) << k;
        return r;
    }
};

// adapters for the ASCII fast path of pair_algos.cuh
template <class Store>
struct StoreWords {
    const Store& s;
    bool second;
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

// UTF-8 bytes -> packed-byte character keys; returns the number of characters
template <class Store>
SS_HD int decode_keys(Store& s, bool second, int nbytes, int base) {
    ByteReader<Store> r(s, second);
    int k = 0, j = 0;
    while (j < nbytes) {
        uint32_t c = r.next();
        int len = c < 0xC0u ? 1 : c < 0xE0u ? 2 : c < 0xF0u ? 3 : 4;
        if (len > nbytes - j) len = nbytes - j;
        uint32_t key = c;
        for (int e = 1; e < len; e++) key = (key << 8) | r.next();
        s.cp(base + k) = key;
        k++;
        j += len;
    }
    return k;
}

// The measure-specific part once characters and a PM provider over the "tabled" string exist.
//   Levenshtein: tabled string = pattern (either side, the distance is symmetric), RT streams the
//                other one (n_text characters).
//   other measures: tabled string = b, RA/RA2 stream a.
template <class M, class PM, class RA>
SS_HD double measure_core(int measure, const PM& pm, RA& ra, RA& ra2, int la, int lb, int n_tabled,
                          int n_stream, PairInts& out) {
    switch (measure) {
        case LEVENSHTEIN: {
            int d = myers_single_word<M>(pm, n_tabled, ra, n_stream);
            out.x0 = d;
            return lev_value(d, la, lb);
        }
        case JARO:
        case JARO_WINKLER: {
            int m, t;
            jaro_match<M>(pm, ra, ra2, la, lb, m, t);
            out.x0 = m;
            out.x1 = t;
            return m == 0 ? 0.0 : jaro_value(m, t, la, lb);
        }
        case JACCARD: {
            int inter = multiset_intersection<M>(pm, ra, la);
            out.x0 = inter;
            out.x1 = la + lb - inter;  // sum_c max = la + lb - sum_c min
            return jaccard_value(inter, la + lb - inter);
        }
        default: {
            int inter = multiset_intersection<M>(pm, ra, la);
            out.x0 = inter;
            out.x1 = la + lb;
            return dice_value(inter, la + lb);
        }
    }
}

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
    constexpr int CAP = (int)sizeof(M) * 8;
    double v;
    if (ascii) {
        const int la = na, lb = nb;
        out.la = la;
        out.lb = lb;
        if ((measure == JARO || measure == JARO_WINKLER) && la == 1 && lb == 1) {  // strsim.rs:197
            out.flag = F_SINGLE_CHAR;
            return 0.0;  // bytes differ here
        }
        // Levenshtein tables the shorter string (fewer table writes, more text steps are cheap)
        const bool table_b = measure != LEVENSHTEIN || lb <= la;
        const int n_tab = table_b ? lb : la;
        const int n_str = table_b ? la : lb;
        StoreWords<Store> tabled(s, table_b), streamed(s, !table_b);
        StoreTable<M, Store> tab(s);
        {
            BuildTable<M, StoreTable<M, Store>> build(tab);
            for_each_byte(tabled, (n_tab + 3) & ~3, build);
        }
        switch (measure) {
            case LEVENSHTEIN: {
                int d = n_str;
                if (n_tab > 0) {
                    MyersStep<M, StoreTable<M, Store>> step(tab);
                    for_each_byte(streamed, n_str, step);
                    d = step.distance(n_tab, n_str);
                }
                out.x0 = d;
                v = lev_value(d, la, lb);
                break;
            }
            case JARO:
            case JARO_WINKLER: {
                const int mx = la > lb ? la : lb;
                const int bound = mx / 2 - 1;  // strsim.rs:200
                const int outer = la < lb + bound ? la : lb + bound;
                JaroMatchStep<M, StoreTable<M, Store>> match(tab, lb, bound);
                for_each_byte(streamed, outer, match);
                JaroTransStep<M, StoreTable<M, Store>> trans(tab, match.flag_a, match.flag_b);
                if (match.m > 0) for_each_byte(streamed, outer, trans);
                out.x0 = match.m;
                out.x1 = trans.t;
                v = match.m == 0 ? 0.0 : jaro_value(match.m, trans.t, la, lb);
                break;
            }
            default: {
                MultisetStep<M, StoreTable<M, Store>> ms(tab, lb);
                for_each_byte(streamed, la, ms);
                out.x0 = ms.inter;
                if (measure == JACCARD) {
                    out.x1 = la + lb - ms.inter;  // sum_c max = la + lb - sum_c min
                    v = jaccard_value(ms.inter, la + lb - ms.inter);
                } else {
                    out.x1 = la + lb;
                    v = dice_value(ms.inter, la + lb);
                }
                break;
            }
        }
        {
            ClearTable<M, StoreTable<M, Store>> clear(tab);
            for_each_byte(tabled, (n_tab + 3) & ~3, clear);
        }
        if (measure == JARO_WINKLER && v > 0.7) {  // strsim.rs:260-267
            uint32_t x = s.wa(0) ^ s.wb(0);
            int lim = la < lb ? la : lb;
            if (lim > 4) lim = 4;
            int l = 0;
            while (l < lim && ((x >> (8 * l)) & 0xFFu) == 0) l++;
            out.x2 = l;
            v = winkler_value(v, l);
        }
    } else {
        const int la = decode_keys(s, false, na, 0);
        const int lb = decode_keys(s, true, nb, CAP);
        out.la = la;
        out.lb = lb;
        if ((measure == JARO || measure == JARO_WINKLER) && la == 1 && lb == 1) {
            out.flag = F_SINGLE_CHAR;
            v = s.cp(0) == s.cp(CAP) ? 1.0 : 0.0;
        } else {
            const bool table_b = measure != LEVENSHTEIN || lb <= la;
            ScanPM<M, Store> pm(s, table_b ? CAP : 0, table_b ? lb : la);
            CpReader<Store> ra(s, table_b ? 0 : CAP), ra2(s, table_b ? 0 : CAP);
            v = measure_core<M>(measure, pm, ra, ra2, la, lb, table_b ? lb : la, table_b ? la : lb,
                                out);
            if (measure == JARO_WINKLER && v > 0.7) {
                int lim = la < lb ? la : lb;
                if (lim > 4) lim = 4;
                int l = 0;
                while (l < lim && s.cp(l) == s.cp(CAP + l)) l++;
                out.x2 = l;
                v = winkler_value(v, l);
            }
        }
        if (Store::CPS_ALIAS_TABLE) {  // the keys live in the table's memory: restore all-zero
            for (int k = 0; k < la; k++) s.cp(k) = 0u;
            for (int k = 0; k < lb; k++) s.cp(CAP + k) = 0u;
        }
    }
    return v;
}

}  // namespace strsim
