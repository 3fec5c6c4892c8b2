// wide_mask.cuh -- position masks of more than 64 bits: the mask type M of the step functors in
// pair_algos.cuh as N 32-bit words, so that ONE THREAD evaluates a pair of ASCII strings of up to 32*N
// characters with the same Myers / Jaro / multiset steps (and the same f64 formulas) as the 32- and 64-bit
// kernels.
//
// Why: rows of 65..320 bytes (addresses, titles: workload T1) used to go to the warp-per-pair kernels
// (long_lev_kernel.cuh, long_pair_kernel.cuh), whose per-pair set-up -- hash, dense ids, Peq slab in HBM --
// and a third of the lanes idle on a 7-word pattern made a 200-character pair cost 20-40 x a 50-character
// one.  With the bit-plane position masks of row_ascii_reg.cuh nothing is set up per pair but the planes,
// and a warp keeps 32 pairs in flight.  strsim.rs:125-345 semantics, bit-exact (tests/test_pair_algos.py
// runs this header on the host against the oracle).
//
// Arithmetic: everything that propagates through the whole mask is an add with carry --
//     x + y                     Myers' (Eq & Pv) + Pv
//     (x << 1) | bit            x + x + bit            (Myers' Ph / Mh, Jaro's window and flag register)
//     x - 1, and "x != 0"       x + all-ones, and its carry out
//     -x                        ~x + 1                 (lowest set bit: x & -x)
// -- so the device code is ONE primitive, a chain of add.cc / addc.cc over five words at a time (the carry
// crosses asm statements as a 0/1 register).  Shifts by a variable amount (only outside the per-character
// loop) are a funnel shift per word plus a barrel over whole words, with static register indices.
#pragma once
#include "pair_algos.cuh"

namespace strsim {

template <int N>
struct Wide {
    uint32_t w[N];
    SS_HD Wide() {}
    SS_HD explicit Wide(int v) {
        w[0] = (uint32_t)v;
#pragma unroll
        for (int i = 1; i < N; i++) w[i] = 0u;
    }
    SS_HD explicit operator bool() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= w[i];
        return o != 0u;
    }
    SS_HD static Wide fill(uint32_t v) {
        Wide r;
#pragma unroll
        for (int i = 0; i < N; i++) r.w[i] = v;
        return r;
    }
};

#define SS_WIDE_BITWISE(OP)                                              \
    template <int N>                                                     \
    SS_HD Wide<N> operator OP(const Wide<N>& a, const Wide<N>& b) {     \
        Wide<N> r;                                                       \
        _Pragma("unroll") for (int i = 0; i < N; i++) r.w[i] = a.w[i] OP b.w[i]; \
        return r;                                                        \
    }                                                                    \
    template <int N>                                                     \
    SS_HD Wide<N>& operator OP##=(Wide<N>& a, const Wide<N>& b) {       \
        _Pragma("unroll") for (int i = 0; i < N; i++) a.w[i] OP## = b.w[i]; \
        return a;                                                        \
    }
SS_WIDE_BITWISE(&)
SS_WIDE_BITWISE(|)
SS_WIDE_BITWISE(^)
#undef SS_WIDE_BITWISE

template <int N>
SS_HD Wide<N> operator~(const Wide<N>& a) {
    Wide<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.w[i] = ~a.w[i];
    return r;
}

// r = r + b + carry (carry 0 or 1); returns the carry out of the top word
template <int N>
SS_HD uint32_t wide_add_in_place(Wide<N>& r, const Wide<N>& b, uint32_t carry) {
#if defined(__CUDA_ARCH__)
    static_assert(N % 5 == 0, "the carry chain is written for groups of five words");
#pragma unroll
    for (int i = 0; i < N; i += 5) {
        // the in/out words are early-clobber: none of them may share a register with a b word read later
        uint32_t out;
        asm("{\n\t.reg .u32 t;\n\t"
            "add.cc.u32 t, %6, 0xFFFFFFFF;\n\t"  // carry flag <- carry
            "addc.cc.u32 %0, %0, %7;\n\t"
            "addc.cc.u32 %1, %1, %8;\n\t"
            "addc.cc.u32 %2, %2, %9;\n\t"
            "addc.cc.u32 %3, %3, %10;\n\t"
            "addc.cc.u32 %4, %4, %11;\n\t"
            "addc.u32 %5, 0, 0;\n\t}"
            : "+&r"(r.w[i]), "+&r"(r.w[i + 1]), "+&r"(r.w[i + 2]), "+&r"(r.w[i + 3]), "+&r"(r.w[i + 4]), "=r"(out)
            : "r"(carry), "r"(b.w[i]), "r"(b.w[i + 1]), "r"(b.w[i + 2]), "r"(b.w[i + 3]), "r"(b.w[i + 4]));
        carry = out;
    }
    return carry;
#else
    for (int i = 0; i < N; i++) {
        const uint64_t s = (uint64_t)r.w[i] + b.w[i] + carry;
        r.w[i] = (uint32_t)s;
        carry = (uint32_t)(s >> 32);
    }
    return carry;
#endif
}

template <int N>
SS_HD Wide<N> operator+(const Wide<N>& a, const Wide<N>& b) {
    Wide<N> r = a;
    wide_add_in_place(r, b, 0u);
    return r;
}
template <int N>
SS_HD Wide<N> operator-(const Wide<N>& a, const Wide<N>& b) {  // a + ~b + 1
    Wide<N> r = a;
    wide_add_in_place(r, ~b, 1u);
    return r;
}

// (x << 1) | (top bit of src)
template <int N>
SS_HD Wide<N> shift_in_top(const Wide<N>& x, const Wide<N>& src) {
    Wide<N> r = x;
    wide_add_in_place(r, x, src.w[N - 1] >> 31);
    return r;
}

// below = x - 1; reg = (reg << 1) | (x != 0): the carry out of x + all-ones says "x is not zero"
template <int N>
SS_HD void dec_and_shift_in(const Wide<N>& x, Wide<N>& below, Wide<N>& reg) {
    below = x;
    const uint32_t nonzero = wide_add_in_place(below, Wide<N>::fill(0xFFFFFFFFu), 0u);
    const Wide<N> twice = reg;
    wide_add_in_place(reg, twice, nonzero);
}

template <int N>
SS_HD int popc(const Wide<N>& x) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) c += popc(x.w[i]);
    return c;
}

template <int N>
SS_HD Wide<N> brev(const Wide<N>& x) {
    Wide<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.w[i] = brev(x.w[N - 1 - i]);
    return r;
}

template <int N>
SS_HD int flo(const Wide<N>& x) {  // index of the highest set bit, x != 0
    int r = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
        if (x.w[i]) r = 32 * i + flo(x.w[i]);
    return r;
}

// shifts by a variable amount: bits by one funnel shift per word, whole words by a barrel (static indices,
// so the words stay in registers).  Outside the per-character loops only.
template <int N>
SS_HD Wide<N> operator<<(const Wide<N>& x, int s) {
    if (s >= 32 * N) return Wide<N>(0);
    Wide<N> r;
    const int bs = s & 31;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        const uint32_t lo = i > 0 ? x.w[i - 1] : 0u;
        r.w[i] = bs ? ((x.w[i] << bs) | (lo >> (32 - bs))) : x.w[i];
    }
    const int ws = s >> 5;
#pragma unroll
    for (int k = 1; k < N; k <<= 1) {
        if (ws & k) {
#pragma unroll
            for (int i = N - 1; i >= 0; i--) r.w[i] = i >= k ? r.w[i - k] : 0u;
        }
    }
    return r;
}
template <int N>
SS_HD Wide<N> operator>>(const Wide<N>& x, int s) {
    if (s >= 32 * N) return Wide<N>(0);
    Wide<N> r;
    const int bs = s & 31;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t hi = i + 1 < N ? x.w[i + 1] : 0u;
        r.w[i] = bs ? ((x.w[i] >> bs) | (hi << (32 - bs))) : x.w[i];
    }
    const int ws = s >> 5;
#pragma unroll
    for (int k = 1; k < N; k <<= 1) {
        if (ws & k) {
#pragma unroll
            for (int i = 0; i < N; i++) r.w[i] = i + k < N ? r.w[i + k] : 0u;
        }
    }
    return r;
}

template <int N>
struct LowMask<Wide<N>> {  // bits < m
    SS_HD static Wide<N> get(int m) {
        Wide<N> r;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int k = m - 32 * i;
            r.w[i] = k >= 32 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << k) - 1u));
        }
        return r;
    }
};

// Jaro's window: bits [i - bound, i + bound] of step i; it moves up one bit per step, one add with carry
template <int N>
struct JaroWindow<Wide<N>> {
    Wide<N> win;
    int i, bound;
    SS_HD void init(int bound_) {
        bound = bound_;
        i = 0;
        win = LowMask<Wide<N>>::get(bound_ + 1);
    }
    SS_HD Wide<N> get() const { return win; }
    SS_HD void next() {
        i++;
        const Wide<N> twice = win;
        wide_add_in_place(win, twice, i <= bound ? 1u : 0u);
    }
};

// TransByBytes (pair_algos.cuh) over wide flag sets: the two sets have the same number of bits; pair them
// highest first.  a's words are walked with static indices; b's cursor moves at its own pace, so b's flags
// sit in a small local array (one load per 32 flags).
template <int N, class ByteAt>
SS_HD int trans_by_bytes_wide(const Wide<N>& flag_a, const Wide<N>& flag_b, const ByteAt& A, const ByteAt& B) {
    uint32_t fb[N];
#pragma unroll
    for (int i = 0; i < N; i++) fb[i] = flag_b.w[i];
    int wb = N - 1;
    uint32_t cb = fb[wb];
    int t = 0;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        uint32_t ca = flag_a.w[i];
        while (ca) {
            while (!cb) cb = fb[--wb];
            const int ia = flo(ca), ib = flo(cb);
            ca ^= 1u << ia;
            cb ^= 1u << ib;
            count_if_differ(t, A(32 * i + ia), B(32 * wb + ib));
        }
    }
    return t;
}

}  // namespace strsim
