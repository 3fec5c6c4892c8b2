// pair_algos.cuh -- per-pair bit-parallel cores of the five measures.
//
// Every function here is `__host__ __device__` and free of memory-space assumptions: the position
// masks of one string ("PM": codepoint -> bitmask of the positions where it occurs) and the
// character streams are supplied by small functor/reader types.  On the device those are backed by
// per-thread shared-memory slabs (short_kernel.cuh); in tests/test_pair_algos.py the same templates
// are compiled for the host and checked against the oracle, so the arithmetic is validated without
// a GPU.
//
// Behavioural contract: /root/reference/src/expressions/strsim.rs (lines cited per function) as
// restated in SURVEY.md section 9.  All integer intermediates are exact; all f64 operations are
// single IEEE-754 round-to-nearest operations in the reference's order, never contracted to FMA.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SS_HD __host__ __device__ __forceinline__
#else
#define SS_HD inline
#endif

namespace strsim {

enum Measure : int {
    LEVENSHTEIN = 0,
    JARO = 1,
    JARO_WINKLER = 2,
    JACCARD = 3,
    SORENSEN_DICE = 4,
};

// flags reported in the debug record (same convention as oracle/strsim_oracle.c)
enum Flag : int { F_GENERAL = 0, F_EQUAL = 1, F_ONE_EMPTY = 2, F_SINGLE_CHAR = 3 };

struct PairInts {
    int flag, la, lb, x0, x1, x2;
};

// ---- f64 helpers: one rounding per written operation -------------------------------------------
SS_HD double f_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
SS_HD double f_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SS_HD double f_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
SS_HD double f_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

// ---- bit helpers --------------------------------------------------------------------------------
SS_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
SS_HD int popc(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
SS_HD int ctz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
SS_HD int ctz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// masks wider than 64 bits (wide_mask.cuh): N 32-bit words behind the same operators
template <int N>
struct Wide;

// bits < m of a mask (m >= 0; m >= bits(M) gives all ones)
template <class M>
struct LowMask {
    SS_HD static M get(int m) { return m >= (int)(sizeof(M) * 8) ? ~M(0) : ((M(1) << m) - M(1)); }
};

// bits lo..hi inclusive (0 <= lo, hi < bits(M)); empty when lo > hi
template <class M>
SS_HD M mask_range(int lo, int hi) {
    if (lo > hi) return M(0);
    M upto_hi = (M(2) << hi) - M(1);  // hi == bits-1 wraps to all ones
    M below_lo = (M(1) << lo) - M(1);
    return upto_hi & ~below_lo;
}

// ---- quotients of small integers ---------------------------------------------------------------------
// Every division of the five formulas but Jaro's final "/ 3.0" divides two small non-negative integers
// (counts and lengths, at most 64 in the short-string kernels).  __ddiv_rn is a ~15-instruction
// software routine, and a fused launch needs eight of them per pair, so the short-string kernels read
// the quotient from a 65 x 65 table instead.  The table is FILLED ON THE DEVICE WITH __ddiv_rn ITSELF
// (quotient_table_kernel, run once per device by host.cu), so a looked-up value is bit for bit the
// value the division would have produced.
constexpr int QUOT_N = 65;
#if defined(__CUDACC__)
__device__ double g_quotient[QUOT_N * QUOT_N];  // [y][x] = x / y for 0 <= x, 1 <= y <= 64
__global__ void quotient_table_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= QUOT_N * QUOT_N) return;
    const int y = i / QUOT_N, x = i % QUOT_N;
    g_quotient[i] = y == 0 ? 0.0 : __ddiv_rn((double)x, (double)y);
}
#endif
// x / y as f64.  SMALL: the caller guarantees 0 <= x <= 64 and 1 <= y <= 64 (device: table lookup)
template <bool SMALL>
SS_HD double int_quotient(int x, int y) {
#if defined(__CUDA_ARCH__)
    if (SMALL) return g_quotient[y * QUOT_N + x];
#endif
    return f_div((double)x, (double)y);
}

// ---- final f64 formulas (reference order) --------------------------------------------------------
// strsim.rs:160
template <bool SMALL = false>
SS_HD double lev_value(int d, int la, int lb) {
    int mx = la > lb ? la : lb;
    return f_sub(1.0, int_quotient<SMALL>(d, mx));
}
// strsim.rs:238-243 (m > 0); t/2 is the integer floor
template <bool SMALL = false>
SS_HD double jaro_value(int m, int t, int la, int lb) {
    double s = f_add(int_quotient<SMALL>(m, la), int_quotient<SMALL>(m, lb));
    s = f_add(s, int_quotient<SMALL>(m - t / 2, m));
    return f_div(s, 3.0);
}
// strsim.rs:260-270
SS_HD double winkler_value(double js, int l) {
    return f_add(js, f_mul(f_mul((double)l, 0.1), f_sub(1.0, js)));
}
// strsim.rs:306
template <bool SMALL = false>
SS_HD double jaccard_value(int inter, int uni) { return int_quotient<SMALL>(inter, uni); }
// strsim.rs:343: 2.0 * inter is exact, so the quotient of the integers 2*inter and la+lb is the same value
template <bool SMALL = false>
SS_HD double dice_value(int inter, int total) { return int_quotient<SMALL>(2 * inter, total); }

// =====================================================================================================
// Step functors.  `Tab` maps a character to the bitmask of the positions where it occurs in the
// tabled string: a direct table indexed by the byte on the ASCII path (characters are bytes packed
// four per word, loops run over whole words), a per-pair hash table on the Unicode path
// (row_short.cuh).  Each functor consumes one character of the streamed string per call.
// =====================================================================================================

// Applies f to the first n bytes of a string given as little-endian words (src(w) -> word w).
template <class Src, class F>
SS_HD void for_each_byte(const Src& src, int n, F& f) {
    int w = 0;
    for (; 4 * (w + 1) <= n; w++) {
        const uint32_t word = src(w);
        f(word & 0xFFu);
        f((word >> 8) & 0xFFu);
        f((word >> 16) & 0xFFu);
        f(word >> 24);
    }
    const int rem = n - 4 * w;
    if (rem > 0) {
        const uint32_t word = src(w);
        f(word & 0xFFu);
        if (rem > 1) f((word >> 8) & 0xFFu);
        if (rem > 2) f((word >> 16) & 0xFFu);
    }
}

// table[c] |= 1 << position, for the n characters of the tabled string.  Runs over whole words: the
// zero padding after the string sets bits >= n in table[0], which no consumer looks at (Myers reads
// bits < m only, Jaro masks with the window, the multiset marks them used).
template <class M, class Tab>
struct BuildTable {
    Tab& tab;
    M bit;
    SS_HD explicit BuildTable(Tab& t) : tab(t), bit(M(1)) {}
    SS_HD void operator()(uint32_t c) {
        tab(c) = tab(c) | bit;
        bit = bit << 1;
    }
};
template <class M, class Tab>
struct ClearTable {
    Tab& tab;
    SS_HD explicit ClearTable(Tab& t) : tab(t) {}
    SS_HD void operator()(uint32_t c) { tab(c) = M(0); }
};

// (x << 1) | (top bit of src): one funnel shift
SS_HD uint32_t shift_in_top(uint32_t x, uint32_t src) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(src, x, 1);
#else
    return (x << 1) | (src >> 31);
#endif
}
SS_HD uint64_t shift_in_top(uint64_t x, uint64_t src) { return (x << 1) | (src >> 63); }

// One Myers / Hyyro column per text character.  The distance is read off the final vertical delta
// vectors: D[m][n] = D[0][n] + sum_{i<m} (Pv_i - Mv_i) with D[0][n] = n, so no per-step score update.
template <class M, class Tab>
struct MyersStep {
    const Tab& tab;
    M Pv, Mv;
    SS_HD explicit MyersStep(const Tab& t) : tab(t), Pv(~M(0)), Mv(M(0)) {}
    SS_HD void operator()(uint32_t c) { step(tab(c)); }
    SS_HD void step(const M Eq) {
        const M Xv = Eq | Mv;
        const M Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        M Ph = Mv | ~(Xh | Pv);
        M Mh = Pv & Xh;
        Ph = shift_in_top(Ph, ~M(0));  // (Ph << 1) | 1: one funnel shift
        Mh = Mh + Mh;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    SS_HD int distance(int m, int n) const {
        const M mask = LowMask<M>::get(m);
        return n + popc(Pv & mask) - popc(Mv & mask);
    }
};

// bit reversal of a mask
SS_HD uint32_t brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
#endif
}
SS_HD uint64_t brev(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    return ((uint64_t)brev((uint32_t)x) << 32) | brev((uint32_t)(x >> 32));
#endif
}
SS_HD int flo(uint32_t x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
    int r;  // bfind IS the hardware's find-leading-one; 31 - __clz() costs a subtraction the compiler keeps
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
#else
    return 31 - __builtin_clz(x);
#endif
}
SS_HD int flo(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
}

// below = x - 1 and reg = (reg << 1) | (x != 0), the latter through the carry of the former:
// x + 0xFFFFFFFF carries out exactly when x >= 1 (add.cc / addc, two instructions)
SS_HD void dec_and_shift_in(uint32_t x, uint32_t& below, uint32_t& reg) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %2, 0xFFFFFFFF;\n\taddc.u32 %1, %1, %1;" : "=r"(below), "+r"(reg) : "r"(x));
#else
    below = x - 1u;
    reg = (reg << 1) | (x != 0u ? 1u : 0u);
#endif
}
SS_HD void dec_and_shift_in(uint64_t x, uint64_t& below, uint64_t& reg) {
    below = x - 1ull;
    reg = (reg << 1) | (x != 0ull ? 1ull : 0ull);
}

// The match window [i-bound, i+bound] of step i (strsim.rs:209-210) as a mask over the positions of b.
// Generic form: it grows at the top every step and starts dropping its lowest bit once i > bound.
template <class M>
struct JaroWindow {
    M win;
    int i, bound;
    SS_HD void init(int bound_) {
        bound = bound_;
        i = 0;
        win = (M(2) << bound_) - M(1);  // bits 0..bound
    }
    SS_HD M get() const { return win; }
    SS_HD void next() {
        i++;
        win = (win << 1) | M(i <= bound ? 1 : 0);
    }
};
// 32-bit masks: the window is the high word of a run of 2*bound+1 ones that moves up one bit per step
// through a 64-bit register (two shifts per step, no counter, no compare).  bound <= 15.
template <>
struct JaroWindow<uint32_t> {
    uint64_t t;
    SS_HD void init(int bound_) { t = ((2ull << (2 * bound_)) - 1ull) << (32 - bound_); }
    SS_HD uint32_t get() const { return (uint32_t)(t >> 32); }
    SS_HD void next() { t += t; }
};

// Jaro match pass (strsim.rs:208-219), branch-free: `avail` holds the positions of b that exist and are
// not flagged yet, the flag of a's character i enters `rev_a` at the bottom (so rev_a is flag_a in
// REVERSED order until finish()); the match count is read off the flags at the end.
template <class M, class Tab>
struct JaroMatchStep {
    const Tab& tab;
    JaroWindow<M> window;
    M lbmask, avail, rev_a;
    M flag_a, flag_b;  // valid after finish()
    int m;             // valid after finish()
    SS_HD JaroMatchStep(const Tab& t, int lb, int bound_) : tab(t), rev_a(M(0)), flag_a(M(0)), flag_b(M(0)), m(0) {
        window.init(bound_);
        lbmask = LowMask<M>::get(lb);
        avail = lbmask;
    }
    SS_HD void operator()(uint32_t c) { step(tab(c)); }
    SS_HD void step(const M Eq) {
        const M cand = Eq & window.get() & avail;
        M below;  // cand - 1; the borrow-free carry of that subtraction says "there is a candidate"
        dec_and_shift_in(cand, below, rev_a);
        avail = avail & ~(cand & ~below);  // the lowest candidate is taken (strsim.rs:211-217)
        window.next();
    }
    // steps = number of characters of a that went through step()
    SS_HD void finish(int steps) {
        flag_b = lbmask & ~avail;
        m = popc(flag_b);
        flag_a = steps > 0 ? (brev(rev_a) >> ((int)(sizeof(M) * 8) - steps)) : M(0);
    }
};

// Jaro transposition pass (strsim.rs:220-237)
template <class M, class Tab>
struct JaroTransStep {
    const Tab& tab;
    M flag_a, fb;
    int t;
    SS_HD JaroTransStep(const Tab& tb, M fa, M fbb) : tab(tb), flag_a(fa), fb(fbb), t(0) {}
    SS_HD void operator()(uint32_t c) {
        if (flag_a & M(1)) {
            const M low = fb & (M(0) - fb);
            fb ^= low;
            if (!(tab(c) & low)) t++;
        }
        flag_a = flag_a >> 1;
    }
};

// t += (x != y): one compare and one predicated add
SS_HD void count_if_differ(int& t, uint32_t x, uint32_t y) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(t) : "r"(x), "r"(y));
#else
    if (x != y) t++;
#endif
}

// How the transposition count t (strsim.rs:220-237) is obtained once the match pass has left its flags.
// TransByPass: a second pass over the characters of a (any Tab / any script).
struct TransByPass {
    template <class M, class Tab, class Each>
    SS_HD int operator()(const Tab& tab, const Each& each_a, int la, M flag_a, M flag_b) const {
        JaroTransStep<M, Tab> trans(tab, flag_a, flag_b);
        each_a(la, trans);
        return trans.t;
    }
};

// TransByBytes: ASCII strings whose bytes can be fetched at a dynamic position (A(p), B(p): byte p of
// a / b): walk the two flag sets lowest bit first and compare the k-th flagged characters directly --
// m iterations of about 16 instructions instead of la iterations that each rebuild a position mask.
template <class ByteAt>
struct TransByBytes {
    ByteAt A, B;
    template <class M, class Tab, class Each>
    SS_HD int operator()(const Tab&, const Each&, int, M flag_a, M flag_b) const {
        // the two flag sets have the same number of bits, so pairing them highest bit first pairs the same
        // ranks as lowest bit first; the highest set bit is one instruction (FLO), the lowest two
        int t = 0;
        while (flag_a) {
            const int ia = flo(flag_a), ib = flo(flag_b);
            flag_a ^= M(1) << ia;
            flag_b ^= M(1) << ib;
            count_if_differ(t, A(ia), B(ib));
        }
        return t;
    }
    // masks of several words: wide_mask.cuh
    template <int N, class Tab, class Each>
    SS_HD int operator()(const Tab&, const Each&, int, const Wide<N>& flag_a, const Wide<N>& flag_b) const {
        return trans_by_bytes_wide(flag_a, flag_b, A, B);
    }
};

// multiset intersection (strsim.rs:297-305); `used` starts with the padding positions >= lb set
template <class M, class Tab>
struct MultisetStep {
    const Tab& tab;
    M used;
    int pad;    // padding positions >= lb, set in `used` from the start
    int inter;  // valid after finish()
    SS_HD MultisetStep(const Tab& t, int lb)
        : tab(t), used(~LowMask<M>::get(lb)),
          pad(lb >= (int)(sizeof(M) * 8) ? 0 : (int)(sizeof(M) * 8) - lb), inter(0) {}
    SS_HD void operator()(uint32_t c) { step(tab(c)); }
    SS_HD void step(const M Eq) {  // branch-free: the lowest unused equal position of b is consumed
        const M cand = Eq & ~used;
        used |= cand & (M(0) - cand);
    }
    SS_HD void finish() { inter = popc(used) - pad; }
};

// ---- fused evaluation of several measures over ONE pass ----------------------------
// The README evaluates all five expressions over the same two columns (README.md:47-51).  With b
// tabled and a streamed -- the orientation Jaro and the multisets need anyway (strsim.rs:208,297) and
// one Levenshtein accepts because the distance is symmetric -- the position mask Eq(c) of every
// character of a is computed ONCE and feeds the Myers column, the Jaro match step and the multiset
// step together.  Groups: Jaro and Jaro-Winkler share the match/transposition counts, Jaccard and
// Sorensen-Dice share the intersection.
enum Group : int { G_LEV = 1, G_JARO = 2, G_SET = 4 };
constexpr int MULTI_BASE = 8;  // kernel template values MULTI_BASE + groups (9..15) = fused evaluation
SS_HD constexpr bool is_multi(int measure) { return measure >= MULTI_BASE; }
SS_HD constexpr int group_of(int measure) {
    return measure == LEVENSHTEIN ? G_LEV : (measure == JARO || measure == JARO_WINKLER) ? G_JARO : G_SET;
}

template <int GROUPS, class M, class Tab>
struct FusedStep {
    const Tab& tab;
    MyersStep<M, Tab> my;
    JaroMatchStep<M, Tab> jm;
    MultisetStep<M, Tab> ms;
    SS_HD FusedStep(const Tab& t, int lb, int bound) : tab(t), my(t), jm(t, lb, bound), ms(t, lb) {}
    SS_HD void operator()(uint32_t c) {
        const M Eq = tab(c);
        if (GROUPS & G_LEV) my.step(Eq);
        if (GROUPS & G_JARO) jm.step(Eq);
        if (GROUPS & G_SET) ms.step(Eq);
    }
};

// =====================================================================================================
// Multi-word Myers / Hyyro: one 64-cell block of one DP column.  (hp, hm) carry the horizontal delta
// (+1 / -1) entering the block from the block above on input and leaving it at the bottom on output;
// the top block of a column is entered with (1, 0) because D[0][j] = j.  long_lev_kernel.cuh runs the
// blocks of one pair as a wavefront across the lanes of a warp, handing (hp, hm) to the next lane
// with __shfl_up; the host tests chain the blocks sequentially.
// =====================================================================================================
SS_HD void myers_block(uint64_t& Pv, uint64_t& Mv, uint64_t Eq, uint32_t& hp, uint32_t& hm) {
    const uint64_t Xv = Eq | Mv;
    Eq |= (uint64_t)hm;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    const uint32_t hp_out = (uint32_t)(Ph >> 63), hm_out = (uint32_t)(Mh >> 63);
    Ph = (Ph << 1) | (uint64_t)hp;
    Mh = (Mh << 1) | (uint64_t)hm;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    hp = hp_out;
    hm = hm_out;
}

// contribution of one block's final vertical deltas to D[m][n] - n; `bits` = pattern cells in the block
SS_HD int myers_block_score(uint64_t Pv, uint64_t Mv, int bits) {
    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    return popc(Pv & mask) - popc(Mv & mask);
}

}  // namespace strsim
