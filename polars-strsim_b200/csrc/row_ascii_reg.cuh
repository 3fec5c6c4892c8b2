// row_ascii_reg.cuh -- the PLANE PATH: one pair of strings of one-byte characters (ASCII, or Latin-1 after
// transcoding), position masks computed from bit planes, no table.
//
// The table-driven path of row_short.cuh keeps a 32..128-entry position-mask table and a copy of the
// pair per thread in shared memory; profiles showed that this shared memory limited the resident warps
// and that its load/store traffic kept the LSU half busy.  For strings of at most 32 characters:
//   * instead of table[c], the position mask of a character is computed from BIT PLANES of the
//     tabled string: plane k holds bit k of every character (bit i of B[k] = bit k of char i), and
//         Eq(c) = ~( OR_k ( B[k] ^ S_k(c) ) ),   S_k(c) = all-ones if bit k of c is set.
//     The S_k come from two multiplies that park bit k in the sign bit of some byte, and one
//     byte-permute with sign replication each (PRMT); with the column statistics proving a 32- or
//     64-code-point alphabet block only 5 or 6 planes are needed, 7 for any ASCII, 8 for Latin-1;
//   * the strings are read where they lie, through a "source" (second half of this file): the tabled
//     one becomes planes word by word, the streamed one is consumed byte by byte -- nothing is copied
//     into registers.
// Arithmetic and row rules are those of row_short.cuh (same step functors, same f64 formulas).
#pragma once
#include "row_short.cuh"
#include "wide_mask.cuh"

namespace strsim {

constexpr int REG_WORDS = 8;  // 32 bytes per string

SS_HD uint32_t sign_fill_byte(uint32_t x, int byte) {  // all-ones if bit 7 of byte `byte` of x is set
#if defined(__CUDA_ARCH__)
    // prmt with the msb of a selector nibble set replicates the SIGN of the selected byte over the
    // target byte (the __byte_perm intrinsic is documented to ignore that bit, so use PTX directly)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(0x8888u + 0x1111u * (uint32_t)byte));
    return r;
#else
    return 0u - ((x >> (8 * byte + 7)) & 1u);
#endif
}

// byte 1 or 2 of a word, zero-extended: one PRMT on the device where a shift and a mask are two operations
SS_HD uint32_t byte_of(uint32_t word, int byte) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(word, 0u, 0x4440u + (uint32_t)byte);
#else
    return (word >> (8 * byte)) & 0xFFu;
#endif
}

// a 32-bit pattern over every 32-bit word of M
template <class M>
struct Rep32 {
    SS_HD static M get(uint32_t s) { return sizeof(M) == 4 ? (M)s : (M)(((uint64_t)s << 32) | s); }
};
template <int N>
struct Rep32<Wide<N>> {
    SS_HD static Wide<N> get(uint32_t s) { return Wide<N>::fill(s); }
};
template <class M>
SS_HD M rep32(uint32_t s) {
    return Rep32<M>::get(s);
}

// M = uint32_t: strings of at most 32 characters; uint64_t: at most 64 (every plane is two registers)
template <int NBITS, class M = uint32_t>
struct PlaneTab {
    typedef M mask_type;
    M B[NBITS];
    // The mask is exact on the positions of the tabled string and ARBITRARY above them (the planes hold
    // whatever followed the string in its last word).  No consumer looks there: Myers' carries and shifts only
    // move upwards and its distance is read below m, Jaro's candidates are cut by `avail` and the multiset's
    // by `used`, both of which start from the length of b.  (A `valid` mask here was one more operation per
    // character of a, on the pipe that bounds the kernel.)
    SS_HD M operator()(uint32_t c) const {
        // u: bits 6,5,4,3 of c in the sign bits of bytes 0..3; v: bits 2,1,0 in bytes 0..2
        const uint32_t u = c * 0x10080402u, v = c * 0x00804020u;
        M X = B[0] ^ rep32<M>(sign_fill_byte(v, 2));
        if (NBITS > 1) X |= B[1] ^ rep32<M>(sign_fill_byte(v, 1));
        if (NBITS > 2) X |= B[2] ^ rep32<M>(sign_fill_byte(v, 0));
        if (NBITS > 3) X |= B[3] ^ rep32<M>(sign_fill_byte(u, 3));
        if (NBITS > 4) X |= B[4] ^ rep32<M>(sign_fill_byte(u, 2));
        if (NBITS > 5) X |= B[5] ^ rep32<M>(sign_fill_byte(u, 1));
        if (NBITS > 6) X |= B[6] ^ rep32<M>(sign_fill_byte(u, 0));
        if (NBITS > 7) X |= B[7] ^ rep32<M>(sign_fill_byte(c, 0));  // one byte per character up to U+00FF (Latin-1 rows)
        return ~X;
    }
};

// adds the four characters of word w (characters 4w..4w+3) to the planes
template <int NBITS, class M>
SS_HD void planes_add_word(PlaneTab<NBITS, M>& tab, uint32_t word, int w) {
#pragma unroll
    for (int k = 0; k < NBITS; k++) {
        // bit k of the four bytes -> four adjacent bits (character order) in the top nibble: source bit
        // 8j+k times 2^(28-7j-k) lands on 28+j; the sixteen partial products fall on distinct bits
        const uint32_t prod = (word & (0x01010101u << k)) * (0x10204080u >> k);
        if (sizeof(M) == 4)
            tab.B[k] |= (M)(w == 7 ? (prod & 0xF0000000u) : ((prod >> (28 - 4 * w)) & (0xFu << (4 * w))));
        else
            tab.B[k] |= (M)(prod >> 28) << (4 * w);
    }
}

// masks of several words (wide_mask.cuh): word w of the string fills nibble w % 8 of mask word w / 8
template <int NBITS, int N>
SS_HD void planes_add_word(PlaneTab<NBITS, Wide<N>>& tab, uint32_t word, int w) {
#pragma unroll
    for (int k = 0; k < NBITS; k++) {
        const uint32_t prod = (word & (0x01010101u << k)) * (0x10204080u >> k);
        tab.B[k].w[w >> 3] |= (prod >> 28) << (4 * (w & 7));
    }
}

// ---- Latin-1 rows of a mixed-script column ------------------------------------------------------------
// A string whose characters are all below U+0100 (no byte >= 0xC4: ASCII, or a C2/C3 lead with one
// continuation byte) has at most one byte of information per character: the continuation byte with
// the lead's low bit moved into bit 6 IS the code point (C2 xx -> xx, C3 xx -> xx | 0x40).  Transcoded
// that way, a pair of such strings runs the bit-plane path with 8 planes instead of the register-compare
// path, whose position masks cost two ALU operations per tabled character per streamed character.
//
// Core: rd(w) -> word w of the UTF-8 string (zero beyond nbytes), wr(w, word) receives the output words.
// `wide` collects the lead bytes >= 0xC4 met on the way (a character above U+00FF: the result is then
// meaningless and the caller hands the pair to the register-compare path).  A word holds at most two lead
// bytes (a lead is followed by a continuation byte); they are squeezed out highest first, branch-free --
// a missing lead turns its step into a no-op.  The output never overtakes the input, so rd and wr may name
// the same storage.  Returns the number of characters.
template <class Rd, class Wr>
SS_HD int transcode_latin1_stream(const Rd& rd, Wr& wr, int nbytes, uint32_t& wide) {
    uint64_t acc = 0;
    int acc_bytes = 0, out_w = 0, total = 0;
    uint32_t carry = 0;  // low bit of a lead byte that ended the previous word
    const int nw = (nbytes + 3) >> 2;
    for (int w = 0; w < nw; w++) {
        const uint32_t x = rd(w);
        const uint32_t lead = x & (x << 1) & 0x80808080u;   // bit 7 of every 11xxxxxx byte
        const uint32_t leadfull = (lead >> 7) * 0xFFu;      // those bytes, all bits
        wide |= x & leadfull & 0x3C3C3C3Cu;                 // a lead with any of bits 2..5 set is >= 0xC4
        const uint32_t lowbit = x & (lead >> 7);            // bit 0 of the lead bytes ...
        uint32_t y = x | (lowbit << 14) | (carry << 6);     // ... into bit 6 of the byte that follows
        carry = lowbit >> 24;
        const uint32_t l1 = lead & (0u - lead), l2 = lead ^ l1;  // the lower and the higher lead of the word (or 0)
        const uint32_t m2 = (l2 >> 7) - 1u, m1 = (l1 >> 7) - 1u;  // bytes below the lead; all ones when there is none
        y = (y & m2) | ((y >> 8) & ~m2);
        y = (y & m1) | ((y >> 8) & ~m1);
        const int in_word = nbytes - 4 * w < 4 ? nbytes - 4 * w : 4;
        const int cnt = in_word - popc(lead);
        acc |= (uint64_t)y << (8 * acc_bytes);
        acc_bytes += cnt;
        total += cnt;
        if (acc_bytes >= 4) {
            wr(out_w++, (uint32_t)acc);
            acc >>= 32;
            acc_bytes -= 4;
        }
    }
    if (acc_bytes > 0) wr(out_w++, (uint32_t)acc);
    return total;
}

// Slab: rd(w) / wr(w, word) over the string's words (zero padded), transcoded in place.
template <class Slab>
SS_HD int transcode_latin1(Slab& s, int nbytes) {
    struct Rd {
        const Slab& s;
        SS_HD uint32_t operator()(int w) const { return s.rd(w); }
    } rd{s};
    struct Wr {
        Slab& s;
        SS_HD void operator()(int w, uint32_t v) { s.wr(w, v); }
    } wr{s};
    uint32_t wide = 0;
    return transcode_latin1_stream(rd, wr, nbytes, wide);
}

// any byte >= 0xC4 (a character above U+00FF) in a zero-padded word?
SS_HD uint32_t wide_bytes(uint32_t x) {
    const uint32_t t = x & 0x3C3C3C3Cu;
    const uint32_t nz = (t + 0x7F7F7F7Fu) & 0x80808080u;  // bit 7 where bits 2..5 are not all zero
    return x & (x << 1) & nz & 0x80808080u;
}


// =========================================================================================================
// The plane path over string SOURCES.  A source is wherever a string of at most 32 one-byte characters
// lives -- the staged tile in shared memory (StagedSrc, short_kernel.cuh), a thread's slab after the
// Latin-1 transcoding (SlabSrc below), plain host memory in the tests -- and offers exactly what the
// measures need:
//     int len;                          characters
//     planes(tab)                       bit planes of the string (it is the tabled one)
//     each(n, f)                        f(c) for the first n characters (it is the streamed one)
//     first_word()                      characters 0..3, zero-masked to len (Winkler prefix)
//     byte_at() -> ByteAt               functor: character p (Jaro's transposition walk)
// Nothing is copied into registers: the tabled string becomes planes word by word, the streamed one
// is read once, byte by byte.  The pair is known NOT to be byte-equal (the sort key of the kernels
// settles equal pairs; strsim.rs:128,182,288,324).
// =========================================================================================================
SS_HD uint32_t low_bytes_mask(int nbytes) {  // low nbytes bytes set, nbytes >= 0
    return nbytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbytes)) - 1u);
}

// words at w[i * STRIDE] (STRIDE = threads per CTA for a slab in shared memory, 1 for plain memory),
// word-aligned, one byte per character; bytes past len are arbitrary
template <int STRIDE, class BA>
struct SlabSrc {
    const uint32_t* w;
    int len;
    BA ba;
    typedef BA ByteAt;
    SS_HD const BA& byte_at() const { return ba; }
    SS_HD uint32_t first_word() const { return w[0] & low_bytes_mask(len); }
    template <int NBITS, class M>
    SS_HD void planes(PlaneTab<NBITS, M>& tab) const {
#pragma unroll
        for (int k = 0; k < NBITS; k++) tab.B[k] = M(0);
#pragma unroll
        for (int i = 0; i < (int)sizeof(M) * 2; i++) {
            if (4 * i >= len) break;
            planes_add_word<NBITS, M>(tab, w[i * STRIDE], i);
        }
    }
    template <class F>
    SS_HD void each(int n, F& f) const {
        const uint32_t* p = w;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (; n >= 4; n -= 4, p += STRIDE) {
            const uint32_t word = p[0];
            f(word & 0xFFu);
            f(byte_of(word, 1));
            f(byte_of(word, 2));
            f(word >> 24);
        }
        if (n > 0) {
            uint32_t word = p[0];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (; n > 0; n--) {
                f(word & 0xFFu);
                word >>= 8;
            }
        }
    }
};

template <class Src>
struct EachOf {
    const Src& s;
    template <class F>
    SS_HD void operator()(int n, F& f) const {
        s.each(n, f);
    }
};

template <class Src>
struct PrefixOf {  // common prefix in characters, capped at 4 (strsim.rs:261-266)
    const Src &A, &B;
    SS_HD int operator()() const {
        const uint32_t x = A.first_word() ^ B.first_word();
        int lim = A.len < B.len ? A.len : B.len;
        if (lim > 4) lim = 4;
        // bytes past the shorter string may or may not differ (a NUL character equals the zero mask):
        // the count of equal leading bytes, capped by lim
        const int l = x == 0u ? 4 : ctz32(x) >> 3;
        return l < lim ? l : lim;
    }
};

// one measure; row rules and arithmetic of row_ascii_reg()
template <int MEASURE, int NBITS, class M = uint32_t, class Src>
SS_HD double row_planes(const Src& A, const Src& B, PairInts& out) {
    out.flag = F_GENERAL;
    out.la = out.lb = out.x0 = out.x1 = out.x2 = 0;
    const int la = A.len, lb = B.len;
    if (MEASURE != LEVENSHTEIN && (la == 0 || lb == 0)) {  // strsim.rs:184,290,326
        out.flag = F_ONE_EMPTY;
        return 0.0;
    }
    out.la = la;
    out.lb = lb;
    constexpr bool IS_JARO = MEASURE == JARO || MEASURE == JARO_WINKLER;
    if (IS_JARO && la == 1 && lb == 1) {  // strsim.rs:197
        out.flag = F_SINGLE_CHAR;
        return 0.0;
    }
    typedef PlaneTab<NBITS, M> Tab;
    Tab tab;
    double v;
    if (MEASURE == LEVENSHTEIN) {
        // the shorter string is tabled; the longer one is streamed (the distance is symmetric)
        const bool table_b = lb <= la;
        const Src& P = table_b ? B : A;
        const Src& X = table_b ? A : B;
        int d = X.len;
        if (P.len > 0) {
            P.planes(tab);
            MyersStep<M, Tab> step(tab);
            X.each(X.len, step);
            d = step.distance(P.len, X.len);
        }
        out.x0 = d;
        v = lev_value<(sizeof(M) <= 8)>(d, la, lb);  // (the quotient table ends at 64)
    } else {
        B.planes(tab);
        if (IS_JARO) {
            const int mx = la > lb ? la : lb;
            const int bound = mx / 2 - 1;  // strsim.rs:200
            const int outer = la < lb + bound ? la : lb + bound;
            JaroMatchStep<M, Tab> match(tab, lb, bound);
            A.each(outer, match);
            match.finish(outer);
            TransByBytes<typename Src::ByteAt> trans{A.byte_at(), B.byte_at()};
            EachOf<Src> each_a{A};
            const int t = match.m > 0 ? trans(tab, each_a, outer, match.flag_a, match.flag_b) : 0;
            out.x0 = match.m;
            out.x1 = t;
            v = match.m == 0 ? 0.0 : jaro_value<(sizeof(M) <= 8)>(match.m, t, la, lb);
            if (MEASURE == JARO_WINKLER && v > 0.7) {  // strsim.rs:260-267
                PrefixOf<Src> prefix{A, B};
                const int l = prefix();
                out.x2 = l;
                v = winkler_value(v, l);
            }
        } else {
            MultisetStep<M, Tab> ms(tab, lb);
            A.each(la, ms);
            ms.finish();
            out.x0 = ms.inter;
            if (MEASURE == JACCARD) {
                out.x1 = la + lb - ms.inter;
                v = jaccard_value<sizeof(M) == 4>(ms.inter, la + lb - ms.inter);
            } else {
                out.x1 = la + lb;
                v = dice_value<sizeof(M) == 4>(ms.inter, la + lb);
            }
        }
    }
    return v;
}

// several measures in one pass: b tabled, a streamed for every group (row_short.cuh: multi_body)
template <int GROUPS, int NBITS, class M = uint32_t, class Src, class Emit>
SS_HD void row_planes_multi(const Src& A, const Src& B, Emit& emit) {
    PlaneTab<NBITS, M> tab;
    B.planes(tab);
    TransByBytes<typename Src::ByteAt> trans{A.byte_at(), B.byte_at()};
    EachOf<Src> each_a{A};
    PrefixOf<Src> prefix{A, B};
    multi_body<GROUPS, M>(tab, each_a, A.len, B.len, A.len == 0 || B.len == 0, prefix, trans, emit);
}

}  // namespace strsim
