// generic_kernel.cuh -- any-length, any-content fallback (one pair per thread, scratch in HBM) and
// the validity-bitmap kernel.
//
// Rows that leave the fused short-string kernel (longer than 64 bytes, or no room in its stage
// area) end up on an overflow list.  This kernel finishes them with the textbook formulations of
// SURVEY.md section 9 (/root/reference/src/expressions/strsim.rs:125-345): two-row DP, greedy
// windowed Jaro over flag arrays, sorted-codepoint merge for the multisets.  It is the correctness
// net: simple, memory-bound, independent of the bit-parallel code (the GPU tests cross-check the
// two on rows both can handle).  Long-string Levenshtein has its own fast kernel
// (long_lev_kernel.cuh); this one keeps Jaro / Jaccard / Dice for long rows and everything when
// forced by STRSIM_B200_FORCE_GENERIC=1.
#pragma once
#include "short_kernel.cuh"

namespace strsim {

struct GenericArgs {
    DevCol a, b;
    double* out;
    int* dbg;
    const unsigned int* list;        // segment-relative rows
    const unsigned int* list_count;  // device-resident count
    uint32_t* scratch;               // n_slots slabs of slab_words u32
    long long slab_words;
    int cap_a, cap_b;  // max bytes (>= max codepoints) of a / b among the listed rows
    int n_slots;
};

__device__ __forceinline__ const unsigned char* view_ptr(const DevCol& c, long long row, int& len) {
    const uint4* pv = c.views + row * c.stride;
    const uint4 v = ld_view(pv);
    len = (int)v.x;
    if (len <= 12) return reinterpret_cast<const unsigned char*>(pv) + 4;
    return reinterpret_cast<const unsigned char*>(c.bufs[v.z]) + v.w;
}

// UTF-8 -> Unicode scalar values (same clamping rules as the oracle's decoder)
__device__ inline int decode_cps(const unsigned char* s, int n, uint32_t* out) {
    int i = 0, k = 0;
    while (i < n) {
        const uint32_t c = s[i];
        int len = c < 0xC0u ? 1 : c < 0xE0u ? 2 : c < 0xF0u ? 3 : 4;
        if (len > n - i) len = n - i;
        uint32_t cp = len == 1 ? c : (c & (0xFFu >> (len + 1)));
        for (int j = 1; j < len; j++) cp = (cp << 6) | (s[i + j] & 0x3Fu);
        out[k++] = cp;
        i += len;
    }
    return k;
}

__device__ inline void heap_sort(uint32_t* v, int n) {
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && v[child] < v[child + 1]) child++;
            if (v[root] >= v[child]) break;
            uint32_t t = v[root];
            v[root] = v[child];
            v[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        uint32_t t = v[0];
        v[0] = v[end];
        v[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && v[child] < v[child + 1]) child++;
            if (v[root] >= v[child]) break;
            uint32_t t2 = v[root];
            v[root] = v[child];
            v[child] = t2;
            root = child;
        }
    }
}

template <int MEASURE>
__global__ void __launch_bounds__(64) generic_kernel(const GenericArgs g) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= g.n_slots) return;
    const unsigned int count = *g.list_count;
    uint32_t* ua = g.scratch + (long long)slot * g.slab_words;
    uint32_t* ub = ua + g.cap_a;
    uint32_t* work = ub + g.cap_b;
    for (unsigned int e = slot; e < count; e += g.n_slots) {
        const long long row = g.list[e];
        int na, nb;
        const unsigned char* pa = view_ptr(g.a, row, na);
        const unsigned char* pb = view_ptr(g.b, row, nb);
        PairInts pi = {F_GENERAL, 0, 0, 0, 0, 0};
        double v;
        bool equal = na == nb;
        for (int i = 0; equal && i < na; i++) equal = pa[i] == pb[i];
        if (equal) {
            pi.flag = F_EQUAL;
            v = 1.0;
        } else if (MEASURE != LEVENSHTEIN && (na == 0 || nb == 0)) {
            pi.flag = F_ONE_EMPTY;
            v = 0.0;
        } else {
            const int la = decode_cps(pa, na, ua);
            const int lb = decode_cps(pb, nb, ub);
            pi.la = la;
            pi.lb = lb;
            if (MEASURE == LEVENSHTEIN) {
                uint32_t* prev = work;
                uint32_t* cur = work + (lb + 1);
                for (int j = 0; j <= lb; j++) prev[j] = (uint32_t)j;
                for (int i = 0; i < la; i++) {
                    cur[0] = (uint32_t)(i + 1);
                    const uint32_t ai = ua[i];
                    uint32_t left = cur[0], diag = prev[0];
                    for (int j = 0; j < lb; j++) {
                        const uint32_t up = prev[j + 1];
                        uint32_t x = diag + (ai != ub[j] ? 1u : 0u);
                        x = min(x, up + 1u);
                        x = min(x, left + 1u);
                        cur[j + 1] = x;
                        left = x;
                        diag = up;
                    }
                    uint32_t* t = prev;
                    prev = cur;
                    cur = t;
                }
                pi.x0 = (int)prev[lb];
                v = lev_value(pi.x0, la, lb);
            } else if (MEASURE == JARO || MEASURE == JARO_WINKLER) {
                if (la == 1 && lb == 1) {
                    pi.flag = F_SINGLE_CHAR;
                    v = ua[0] == ub[0] ? 1.0 : 0.0;
                } else {
                    unsigned char* fa = reinterpret_cast<unsigned char*>(work);
                    unsigned char* fb = fa + la;
                    for (int i = 0; i < la; i++) fa[i] = 0;
                    for (int j = 0; j < lb; j++) fb[j] = 0;
                    const int mx = la > lb ? la : lb;
                    const int bound = mx / 2 - 1;
                    const int outer = la < lb + bound ? la : lb + bound;
                    int m = 0;
                    for (int i = 0; i < outer; i++) {
                        const int lo = i > bound ? i - bound : 0;
                        const int hi = i + bound < lb - 1 ? i + bound : lb - 1;
                        for (int j = lo; j <= hi; j++) {
                            if (ua[i] == ub[j] && !fb[j]) {
                                fa[i] = 1;
                                fb[j] = 1;
                                m++;
                                break;
                            }
                        }
                    }
                    int t = 0, j = 0;
                    for (int i = 0; i < la; i++) {
                        if (!fa[i]) continue;
                        while (j < lb && !fb[j]) j++;
                        if (j >= lb) break;
                        if (ua[i] != ub[j]) t++;
                        j++;
                    }
                    pi.x0 = m;
                    pi.x1 = t;
                    v = m == 0 ? 0.0 : jaro_value(m, t, la, lb);
                    if (MEASURE == JARO_WINKLER && v > 0.7) {
                        int lim = la < lb ? la : lb;
                        if (lim > 4) lim = 4;
                        int l = 0;
                        while (l < lim && ua[l] == ub[l]) l++;
                        pi.x2 = l;
                        v = winkler_value(v, l);
                    }
                }
            } else {
                heap_sort(ua, la);
                heap_sort(ub, lb);
                int i = 0, j = 0, inter = 0;
                while (i < la && j < lb) {
                    if (ua[i] == ub[j]) {
                        inter++;
                        i++;
                        j++;
                    } else if (ua[i] < ub[j]) {
                        i++;
                    } else {
                        j++;
                    }
                }
                pi.x0 = inter;
                if (MEASURE == JACCARD) {
                    pi.x1 = la + lb - inter;
                    v = jaccard_value(inter, la + lb - inter);
                } else {
                    pi.x1 = la + lb;
                    v = dice_value(inter, la + lb);
                }
            }
        }
        g.out[row] = v;
        if (g.dbg) {
            int* d = g.dbg + row * 6;
            d[0] = pi.flag;
            d[1] = pi.la;
            d[2] = pi.lb;
            d[3] = pi.x0;
            d[4] = pi.x1;
            d[5] = pi.x2;
        }
    }
}

// test hook (STRSIM_B200_FORCE_GENERIC=1): list every valid row for the fallback kernel
__global__ void list_all_kernel(const SegArgs s) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= s.n) return;
    const bool valid = bit_valid(s.a.validity, s.a.vbit + row * s.a.stride) &&
                       bit_valid(s.b.validity, s.b.vbit + row * s.b.stride);
    if (!valid) {
        if (s.out) {
            s.out[row] = 0.0;
            if (s.dbg)
                for (int q = 0; q < 6; q++) s.dbg[row * 6 + q] = 0;
        } else {  // fused call: every wanted measure
            for (int m = 0; m < 5; m++) {
                if (!s.outs[m]) continue;
                s.outs[m][row] = 0.0;
                if (s.dbgs[m])
                    for (int q = 0; q < 6; q++) s.dbgs[m][row * 6 + q] = 0;
            }
        }
        return;
    }
    s.listlong[atomicAdd(&s.ovf->nlong, 1u)] = (unsigned int)row;
    atomicMax(&s.ovf->max_bytes_a, ld_view(s.a.views + row * s.a.stride).x);
    atomicMax(&s.ovf->max_bytes_b, ld_view(s.b.views + row * s.b.stride).x);
}

// ---- column statistics: OR and AND of every string byte (views' inline bytes + data buffers) --------
// Decides, per pair of columns, whether the whole alphabet is ASCII and whether it fits one aligned
// block of 32 / 64 code points (=> smaller position-mask tables, see DevStore).  Supersets are fine:
// unreferenced bytes in a data buffer can only make the answer more conservative.
struct ColumnStats {
    unsigned int or_bits;   // OR of all bytes, replicated over the four byte lanes
    unsigned int and_bits;  // AND of all bytes (padding treated as 0xFF)
    // views only: the longest string in bytes, and the largest out-of-line payload (lengths padded to 4) of
    // any aligned block of STATS_BLOCK consecutive rows of a chunk.  With them the host can PROVE that no row
    // of a launch will leave the short-string kernel (none longer than its masks, no tile larger than its
    // stage area) and skip the read-back of the overflow counters -- the device-resident call is then
    // asynchronous for real.
    unsigned int max_len;
    unsigned int max_block_pad;
};
constexpr int STATS_BLOCK = 256;

__device__ __forceinline__ void stats_commit(ColumnStats* out, uint32_t o, uint32_t a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        o |= __shfl_xor_sync(0xFFFFFFFFu, o, d);
        a &= __shfl_xor_sync(0xFFFFFFFFu, a, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&out->or_bits, o);
        atomicAnd(&out->and_bits, a);
    }
}

__device__ __forceinline__ void stats_of_view(const uint4& v, uint32_t& o, uint32_t& a) {
    const int len = (int)v.x;
    if (len > 12) {  // prefix; the rest is in the data buffer
        o |= v.y;
        a &= v.y;
    } else {
        const uint32_t m0 = byte_mask(len), m1 = byte_mask(len - 4 < 0 ? 0 : len - 4),
                       m2 = byte_mask(len - 8 < 0 ? 0 : len - 8);
        o |= (v.y & m0) | (v.z & m1) | (v.w & m2);
        a &= (v.y | ~m0) & (v.z | ~m1) & (v.w | ~m2);
    }
}

// Both statistics kernels are pure streaming reads: four independent 16-byte loads per thread and
// iteration keep enough bytes in flight to run at the HBM rate (one load per iteration: half of it).
// A CTA of STATS_BLOCK threads takes four aligned blocks of STATS_BLOCK rows per iteration, one row of each
// per thread, so that the per-block payload sums are CTA reductions.
__global__ void __launch_bounds__(STATS_BLOCK) stats_views_kernel(const uint4* views, long long n, ColumnStats* out) {
    __shared__ unsigned int block_sum[4];
    uint32_t o = 0, a = 0xFFFFFFFFu, mx = 0, best = 0;
    const long long n_blocks = (n + STATS_BLOCK - 1) / STATS_BLOCK;
    if (threadIdx.x < 4) block_sum[threadIdx.x] = 0u;
    __syncthreads();
    for (long long b0 = 4ll * blockIdx.x; b0 < n_blocks; b0 += 4ll * gridDim.x) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long row = (b0 + j) * STATS_BLOCK + threadIdx.x;
            v[j] = row < n ? ld_view(views + row) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((b0 + j) * STATS_BLOCK + threadIdx.x < n) stats_of_view(v[j], o, a);
            mx = max(mx, v[j].x);
            const unsigned int pad = v[j].x > 12u ? ((v[j].x + 3u) & ~3u) : 0u;
            const unsigned int s = __reduce_add_sync(0xFFFFFFFFu, pad);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(&block_sum[j], s);
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            best = max(best, block_sum[threadIdx.x]);
            block_sum[threadIdx.x] = 0u;
        }
        __syncthreads();
    }
    stats_commit(out, o, a);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&out->max_len, mx);
    if (threadIdx.x < 4 && best) atomicMax(&out->max_block_pad, best);
}

// data: 16-byte aligned device buffer, `bytes` valid bytes
__global__ void __launch_bounds__(256) stats_bytes_kernel(const unsigned char* data, long long bytes, ColumnStats* out) {
    uint32_t o = 0, a = 0xFFFFFFFFu;
    const long long n16 = bytes >> 4;
    const uint4* d4 = reinterpret_cast<const uint4*>(data);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 v0 = ld_view(d4 + i), v1 = ld_view(d4 + i + stride), v2 = ld_view(d4 + i + 2 * stride),
                    v3 = ld_view(d4 + i + 3 * stride);
        o |= (v0.x | v0.y | v0.z | v0.w) | (v1.x | v1.y | v1.z | v1.w) | (v2.x | v2.y | v2.z | v2.w) | (v3.x | v3.y | v3.z | v3.w);
        a &= (v0.x & v0.y & v0.z & v0.w) & (v1.x & v1.y & v1.z & v1.w) & (v2.x & v2.y & v2.z & v2.w) & (v3.x & v3.y & v3.z & v3.w);
    }
    for (; i < n16; i += stride) {
        const uint4 v = ld_view(d4 + i);
        o |= v.x | v.y | v.z | v.w;
        a &= v.x & v.y & v.z & v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (bytes & 15)) {
        const uint32_t c = data[(n16 << 4) + threadIdx.x];
        o |= c * 0x01010101u;
        a &= c * 0x01010101u;
    }
    stats_commit(out, o, a);
}

// ---- dictionary-encoded columns: materialise the views of the rows ----------------------------------------
// A dictionary-encoded (Polars Categorical / Arrow dictionary) chunk holds one small integer per row and the
// distinct strings once.  Only the indices cross PCIe (1-8 bytes per row instead of a 16-byte view plus
// payload); on the device row r's view is the dictionary's view of its index -- a 16-byte gather out of a
// table that sits in L2 -- and everything downstream runs on that view column unchanged (its out-of-line
// views point into the dictionary's data buffers).  Validity: the row's bit AND the dictionary value's
// bit AND index in range; one ballot per warp writes 32 rows' bits.
struct DictGatherArgs {
    const uint4* dict_views;
    const uint8_t* dict_validity;  // or nullptr
    long long dict_vbit;
    long long dict_len;
    const void* indices;  // row 0 of the chunk's logical range
    int index_bytes;      // 1, 2, 4, 8
    int index_signed;
    const uint8_t* validity;  // of the indices (device copy), or nullptr
    long long vbit;
    long long n;
    uint4* out_views;
    uint32_t* out_validity;  // ceil(n/32) words
};

__global__ void __launch_bounds__(256) dict_gather_kernel(const DictGatherArgs g) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    bool ok = false;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < g.n) {
        long long idx;
        switch (g.index_bytes) {
            case 1: idx = g.index_signed ? (long long)static_cast<const int8_t*>(g.indices)[row]
                                         : (long long)static_cast<const uint8_t*>(g.indices)[row]; break;
            case 2: idx = g.index_signed ? (long long)static_cast<const int16_t*>(g.indices)[row]
                                         : (long long)static_cast<const uint16_t*>(g.indices)[row]; break;
            case 4: idx = g.index_signed ? (long long)static_cast<const int32_t*>(g.indices)[row]
                                         : (long long)static_cast<const uint32_t*>(g.indices)[row]; break;
            default: idx = static_cast<const long long*>(g.indices)[row]; break;
        }
        ok = bit_valid(g.validity, g.vbit + row) && idx >= 0 && idx < g.dict_len &&
             bit_valid(g.dict_validity, g.dict_vbit + idx);
        if (ok) v = __ldg(g.dict_views + idx);
        g.out_views[row] = v;
    }
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, ok);
    if ((threadIdx.x & 31) == 0 && (row & ~31ll) < g.n) g.out_validity[row >> 5] = bits;
}

// ---- output validity: bit = both inputs valid (polars-core arity kernels; README.md:69-70) ---------
struct ValidityArgs {
    const uint8_t* va;
    long long abit;
    int astride;
    const uint8_t* vb;
    long long bbit;
    int bstride;
    long long n;         // rows of the segment
    long long out_row0;  // first output row of the segment
    uint32_t* out;       // whole output bitmap (zero-initialised), LSB-first words
    unsigned long long* null_count;
};

// 32 validity bits of rows [s, s+32) of one column (s >= 0): all ones without a bitmap, the broadcast bit
// for a scalar operand, else two aligned word loads and a funnel shift (the device copy of a bitmap is
// 256-byte aligned and padded, host.cu: upload_plan)
__device__ __forceinline__ uint32_t validity_word(const uint8_t* bm, long long bit0, int stride, long long s) {
    if (bm == nullptr) return 0xFFFFFFFFu;
    if (stride == 0) return ((bm[bit0 >> 3] >> (bit0 & 7)) & 1) ? 0xFFFFFFFFu : 0u;
    const long long b = bit0 + s;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(bm);
    const uint32_t lo = __ldg(words + (b >> 5)), hi = __ldg(words + (b >> 5) + 1);
    return __funnelshift_r(lo, hi, (int)(b & 31));
}

__global__ void validity_kernel(const ValidityArgs v) {
    const long long w_lo = v.out_row0 >> 5;
    const long long w = w_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long r_begin = max(w << 5, v.out_row0);
    const long long r_end = min((w + 1) << 5, v.out_row0 + v.n);
    const int rows = r_begin < r_end ? (int)(r_end - r_begin) : 0;  // (no early return: the reduction below names every lane)
    uint32_t bits = 0;
    if (rows == 32 && (reinterpret_cast<uintptr_t>(v.va) & 3) == 0 && (reinterpret_cast<uintptr_t>(v.vb) & 3) == 0) {
        const long long s = r_begin - v.out_row0;
        bits = validity_word(v.va, v.abit, v.astride, s) & validity_word(v.vb, v.bbit, v.bstride, s);
    } else {
        for (long long r = r_begin; r < r_end; r++) {
            const long long s = r - v.out_row0;
            const bool ok = bit_valid(v.va, v.abit + s * v.astride) && bit_valid(v.vb, v.bbit + s * v.bstride);
            bits |= (ok ? 1u : 0u) << (int)(r & 31);
        }
    }
    if (rows == 32)
        v.out[w] = bits;
    else if (rows > 0)
        atomicOr(&v.out[w], bits);
    // one atomic per warp: with nulls in most 32-row words, a million atomics per segment on ONE address
    // serialised in the L2 and made this kernel 12 % of a C3 step (658 us per 33 M rows; now a few us)
    const int nulls = rows - __popc(bits);
    const int total = __reduce_add_sync(0xFFFFFFFFu, nulls);
    if (total && (threadIdx.x & 31) == 0) atomicAdd(v.null_count, (unsigned long long)total);
}

}  // namespace strsim
