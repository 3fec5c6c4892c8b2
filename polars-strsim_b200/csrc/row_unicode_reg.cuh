// row_unicode_reg.cuh -- one pair of short strings of ANY script, position masks from register compares.
//
// The hash-table path of row_short.cuh keeps 512 B of slots per thread in shared memory, which caps
// the general kernel at 8 resident warps per SM.  Here the characters of the TABLED string are
// decoded once into registers (up to 32 keys; the character index is uniform across a warp, so the
// unrolled loops index registers statically) and the position mask of a streamed character c is
//     Eq(c) = sum_i (P[i] == c) << i          (one compare + one predicated OR per tabled character)
// -- 2m ALU operations for a pattern of m characters, no shared-memory table, no probing, no
// clearing.  The strings themselves stay in the thread's slab (their byte positions differ per lane)
// and are decoded on the fly with the branch-free window decoder of row_short.cuh.
// Keys are the characters' packed UTF-8 bytes (injective for valid UTF-8).  Serves strings of at most
// 32 bytes (M = uint32_t); the 33..64-byte overflow kernel keeps the hash path.
#pragma once
#include "row_short.cuh"

namespace strsim {

constexpr int UREG_MAX = 32;
constexpr uint32_t UREG_NONE = 0xFFFFFFFFu;  // no UTF-8 byte is 0xFF: never equals a key

// eq |= bit when x == c: one compare and one predicated OR (the compiler's SEL + add tree costs three)
SS_HD void or_if_equal(uint32_t& eq, uint32_t x, uint32_t c, uint32_t bit) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(eq) : "r"(x), "r"(c), "r"(bit));
#else
    if (x == c) eq |= bit;
#endif
}

struct CmpTab {
    uint32_t P[UREG_MAX];
    int bound;  // compares run over [0, bound): >= the pattern length, uniform across the warp
    SS_HD uint32_t operator()(uint32_t c) const {
        uint32_t eq = 0u;
        // groups of four positions; entries past the pattern hold UREG_NONE, so a group may overrun it
#pragma unroll
        for (int g = 0; g < UREG_MAX; g += 4) {
            if (g >= bound) break;
            or_if_equal(eq, P[g], c, 1u << g);
            or_if_equal(eq, P[g + 1], c, 2u << g);
            or_if_equal(eq, P[g + 2], c, 4u << g);
            or_if_equal(eq, P[g + 3], c, 8u << g);
        }
        return eq;
    }
};

// decodes up to UREG_MAX characters of a slab string into registers; returns the character count
template <class Src>
SS_HD int decode_to_regs(const Src& src, int nbytes, int last_word, CmpTab& tab) {
    int pos = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < UREG_MAX; i++) {
        tab.P[i] = UREG_NONE;
        if (pos < nbytes) {
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
            uint32_t extra = (0xE5000000u >> ((lead >> 4) * 2)) & 3u;
            if ((int)extra > nbytes - pos - 1) extra = (uint32_t)(nbytes - pos - 1);
            tab.P[i] = win & (0xFFFFFFFFu >> (24 - 8 * extra));
            pos += 1 + (int)extra;
            cnt++;
        }
    }
    return cnt;
}

// na, nb: byte lengths (<= 32).  bound_hint: 0 on the host (use the pattern length); on the device the
// caller passes a functor-free warp maximum through `warp_bound` (see short_kernel.cuh).
template <int MEASURE, class Store, class WarpMax>
SS_HD double row_unicode_reg(Store& s, int na, int nb, bool equal, const WarpMax& warp_max, PairInts& out) {
    out.flag = F_GENERAL;
    out.la = out.lb = out.x0 = out.x1 = out.x2 = 0;
    constexpr int LAST_WORD = 7;
    constexpr bool IS_JARO = MEASURE == JARO || MEASURE == JARO_WINKLER;
    StoreWords<Store> wa(s, false), wb(s, true);
    // codepoint counts first: they decide which string is tabled and the trivial cases
    int la = count_chars(wa, na), lb = count_chars(wb, nb);
    int kind = 0;  // 0 general, 1 equal, 2 one side empty, 3 single characters
    if (equal)
        kind = 1;
    else if (MEASURE != LEVENSHTEIN && (na == 0 || nb == 0))
        kind = 2;
    else if (IS_JARO && la == 1 && lb == 1)
        kind = 3;
    const bool table_b = MEASURE != LEVENSHTEIN || lb <= la;
    const int n_tab = kind ? 0 : (table_b ? lb : la);
    CmpTab tab;
    tab.bound = warp_max(n_tab);  // every lane of the warp reaches this point
    if (kind == 1) {  // strsim.rs:128,182,288,324
        out.flag = F_EQUAL;
        return 1.0;
    }
    if (kind == 2) {  // strsim.rs:184,290,326
        out.flag = F_ONE_EMPTY;
        return 0.0;
    }
    out.la = la;
    out.lb = lb;
    if (kind == 3) {  // strsim.rs:197; the bytes differ here
        out.flag = F_SINGLE_CHAR;
        return 0.0;
    }
    decode_to_regs(table_b ? wb : wa, table_b ? nb : na, LAST_WORD, tab);
    EachChar<Store> streamed(s, !table_b, table_b ? na : nb);
    double v = measure_body<uint32_t, true>(MEASURE, tab, streamed, la, lb, n_tab, table_b ? la : lb, out);
    if (MEASURE == JARO_WINKLER && v > 0.7) {  // strsim.rs:260-267
        PrefixKeys pa, pb;
        for_each_char(wa, na, 4, LAST_WORD, pa);
        for_each_char(wb, nb, 4, LAST_WORD, pb);
        const int lim = pa.n < pb.n ? pa.n : pb.n;
        int l = 0;
        while (l < lim && pa.k[l] == pb.k[l]) l++;
        out.x2 = l;
        v = winkler_value(v, l);
    }
    return v;
}

// ---- fused evaluation of several measures (row_short.cuh: multi_body) ---------------------------------
template <class Words>
struct PrefixChars {  // common prefix in characters, capped at 4 (strsim.rs:261-266)
    const Words &wa, &wb;
    int na, nb;
    SS_HD PrefixChars(const Words& a, const Words& b, int na_, int nb_) : wa(a), wb(b), na(na_), nb(nb_) {}
    SS_HD int operator()() const {
        PrefixKeys pa, pb;
        for_each_char(wa, na, 4, 7, pa);
        for_each_char(wb, nb, 4, 7, pb);
        const int lim = pa.n < pb.n ? pa.n : pb.n;
        int l = 0;
        while (l < lim && pa.k[l] == pb.k[l]) l++;
        return l;
    }
};

template <int GROUPS, class Store, class WarpMax, class Emit>
SS_HD void row_unicode_reg_multi(Store& s, int na, int nb, bool equal, const WarpMax& warp_max, Emit& emit) {
    StoreWords<Store> wa(s, false), wb(s, true);
    const int la = count_chars(wa, na), lb = count_chars(wb, nb);
    CmpTab tab;
    tab.bound = warp_max(equal ? 0 : lb);  // every lane of the warp reaches this point
    if (equal) {  // strsim.rs:128,182,288,324
        PairInts o;
        o.flag = F_EQUAL;
        o.la = o.lb = o.x0 = o.x1 = o.x2 = 0;
        emit_groups<GROUPS>(emit, 1.0, o);
        return;
    }
    decode_to_regs(wb, nb, 7, tab);
    EachChar<Store> each_a(s, false, na);
    PrefixChars<StoreWords<Store>> prefix(wa, wb, na, nb);
    multi_body<GROUPS, uint32_t>(tab, each_a, la, lb, na == 0 || nb == 0, prefix, TransByPass(), emit);
}

struct HostWarpMax {  // host tests: one "lane"
    SS_HD int operator()(int v) const { return v; }
};

}  // namespace strsim
