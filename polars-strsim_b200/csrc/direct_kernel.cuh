// direct_kernel.cuh -- fused short-string kernel, "direct" variant: the strings of a tile are NOT
// staged in shared memory.
//
// Why a second variant: profiles of short_kernel.cuh (profiles/r1_*) show that the kernel is bound by
// dependent-instruction latency with too few resident warps, and that shared memory is what limits
// the warps: per CTA 16 KB of views + ~14 KB of staged payload sit next to the per-thread tables.
// The DRAM traffic, on the other hand, is already exactly the algorithmic bytes.  This variant keeps
// the same tile structure (coalesced view loads -> per-tile length bucketing -> one pair per thread
// with row_short<M>()) but leaves the bytes where they are:
//   1. load    : coalesced 16-byte loads of both columns' views; only the LENGTHS are kept (sort key)
//   2. prefetch: the tile's out-of-line payload span is pulled into L2 by one TMA bulk prefetch per
//                column (cp.async.bulk.prefetch.L2) -- a hint, no shared memory, no barrier
//   3. bucket  : shared-memory counting sort on max byte length -> permutation
//   4. compute : each thread re-reads its (sorted) row's two views and payload words from L2 straight
//                into its private slab and runs row_short<M>()
// Shared memory per CTA shrinks to tables + slabs + 2 B/row, so about twice as many warps are
// resident.  ASCII vs Unicode is decided per WARP at compute time (the Unicode path is correct for
// ASCII pairs too), so no bytes are needed before the sort.
#pragma once
#include "short_kernel.cuh"

namespace strsim {

template <class M, int TPB, int RPT, int T>
struct DirectLayout {
    static constexpr int CAP = (int)sizeof(M) * 8;
    static constexpr int WORDS = CAP / 4;
    static constexpr int TILE = TPB * RPT;
    static constexpr int NB = CAP + 2;  // keys 0 .. CAP+1
    static constexpr int NWARP = TPB / 32;
    static constexpr size_t off_tab = 0;
    static constexpr size_t off_slab_a = off_tab + sizeof(M) * T * TPB;
    static constexpr size_t off_slab_b = off_slab_a + 4 * WORDS * TPB;
    static constexpr size_t off_hist = off_slab_b + 4 * WORDS * TPB;
    static constexpr size_t off_red = off_hist + 4 * ((NB + 3) & ~3);
    static constexpr size_t off_perm = off_red + 4 * 16 * NWARP;
    static constexpr size_t bytes = (off_perm + 2 * TILE + 15) & ~(size_t)15;
};

__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// bytes of one string (view v of column c) -> the thread's slab, zero-masked words; returns OR of the words
template <int TPB>
__device__ __forceinline__ uint32_t load_string_global(const uint4& v, const DevCol& c, uint32_t* slab) {
    const int len = (int)v.x;
    uint32_t acc = 0;
    if (len <= 12) {
        const uint32_t w0 = v.y & byte_mask(len);
        const uint32_t w1 = v.z & byte_mask(len - 4 < 0 ? 0 : len - 4);
        const uint32_t w2 = v.w & byte_mask(len - 8 < 0 ? 0 : len - 8);
        slab[0] = w0;
        slab[TPB] = w1;
        slab[2 * TPB] = w2;
        acc = w0 | w1 | w2;
    } else {
        const uintptr_t g = (uintptr_t)c.bufs[v.z] + v.w;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g & ~(uintptr_t)3);
        const int sh = (int)(g & 3) * 8;
        const int full = len >> 2;
        uint32_t lo = __ldg(src);
        int w = 0;
        for (; w < full; w++) {
            const uint32_t hi = __ldg(src + w + 1);  // at most 7 bytes past the string: padded buffer
            const uint32_t word = __funnelshift_r(lo, hi, sh);
            lo = hi;
            slab[w * TPB] = word;
            acc |= word;
        }
        if (len & 3) {
            const uint32_t word = __funnelshift_r(lo, __ldg(src + w + 1), sh) & byte_mask(len & 3);
            slab[w * TPB] = word;
            acc |= word;
        }
    }
    return acc;
}

template <class M, int MEASURE, int TPB, int RPT, bool GATHER, int T, bool ASCII_ONLY>
__device__ __forceinline__ void direct_body(const SegArgs& s) {
    static_assert(ASCII_ONLY || T >= DevStore<M, TPB, T>::HASH_ENTRIES,
                  "the Unicode path keeps its hash slots in the table memory");
    using L = DirectLayout<M, TPB, RPT, T>;
    constexpr int CAP = L::CAP;
    constexpr int TILE = L::TILE;
    constexpr int NB = L::NB;
    constexpr int NWARP = L::NWARP;

    extern __shared__ __align__(16) unsigned char smem[];
    M* tab = reinterpret_cast<M*>(smem + L::off_tab);
    uint32_t* slab_a = reinterpret_cast<uint32_t*>(smem + L::off_slab_a);
    uint32_t* slab_b = reinterpret_cast<uint32_t*>(smem + L::off_slab_b);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L::off_hist);
    uint32_t* red = reinterpret_cast<uint32_t*>(smem + L::off_red);
    uint16_t* perm = reinterpret_cast<uint16_t*>(smem + L::off_perm);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const long long n = GATHER ? (long long)*s.list_count : s.n;
    const long long n_tiles = (n + TILE - 1) / TILE;

    {
        uint4* t4 = reinterpret_cast<uint4*>(tab);
        constexpr int N4 = (int)(sizeof(M) * T * TPB / 16);
        for (int i = tid; i < N4; i += TPB) t4[i] = make_uint4(0, 0, 0, 0);
    }
    DevStore<M, TPB, T> store;
    store.tab_ = tab + tid;
    store.wa_ = slab_a + tid;
    store.wb_ = slab_b + tid;
    __syncthreads();

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long tile0 = tile * TILE;
        for (int i = tid; i < NB; i += TPB) hist[i] = 0;
        __syncthreads();

        // ---------------- 1. views -> lengths, validity, routing, sort key -----------------------------
        uint32_t key[RPT], rank[RPT];
        uint32_t mn_off[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_end[2] = {0, 0};
        uint32_t mn_buf[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_buf[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int i = k * TPB + tid;
            const long long idx = tile0 + i;
            key[k] = 0;
            rank[k] = 0;
            if (idx >= n) continue;
            const long long row = GATHER ? (long long)s.list[idx] : idx;
            const uint4 va = ld_view(s.a.views + row * s.a.stride);
            const uint4 vb = ld_view(s.b.views + row * s.b.stride);
            const bool valid = bit_valid(s.a.validity, s.a.vbit + row * s.a.stride) &&
                               bit_valid(s.b.validity, s.b.vbit + row * s.b.stride);
            const uint32_t mx = va.x > vb.x ? va.x : vb.x;
            if (!valid) {
                s.out[row] = 0.0;
                if (s.dbg) {
                    int* d = s.dbg + row * 6;
#pragma unroll
                    for (int q = 0; q < 6; q++) d[q] = 0;
                }
            } else if (mx > (uint32_t)CAP) {
                if (CAP == 32 && mx <= 64u) {
                    s.list64[atomicAdd(&s.ovf->n64, 1u)] = (unsigned int)row;
                } else {
                    s.listlong[atomicAdd(&s.ovf->nlong, 1u)] = (unsigned int)row;
                    atomicMax(&s.ovf->max_bytes_a, va.x);
                    atomicMax(&s.ovf->max_bytes_b, vb.x);
                }
            } else {
                key[k] = 1u + mx;
                rank[k] = atomicAdd(&hist[key[k]], 1u);
                if (!GATHER) {
                    if (va.x > 12u) {
                        mn_off[0] = min(mn_off[0], va.w);
                        mx_end[0] = max(mx_end[0], va.w + va.x);
                        mn_buf[0] = min(mn_buf[0], va.z);
                        mx_buf[0] = max(mx_buf[0], va.z);
                    }
                    if (vb.x > 12u) {
                        mn_off[1] = min(mn_off[1], vb.w);
                        mx_end[1] = max(mx_end[1], vb.w + vb.x);
                        mn_buf[1] = min(mn_buf[1], vb.z);
                        mx_buf[1] = max(mx_buf[1], vb.z);
                    }
                }
            }
        }

        // ---------------- 2. TMA bulk prefetch of the tile's payload spans into L2 -----------------------
        if (!GATHER) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn_off[c] = min(mn_off[c], __shfl_xor_sync(0xFFFFFFFFu, mn_off[c], o));
                    mx_end[c] = max(mx_end[c], __shfl_xor_sync(0xFFFFFFFFu, mx_end[c], o));
                    mn_buf[c] = min(mn_buf[c], __shfl_xor_sync(0xFFFFFFFFu, mn_buf[c], o));
                    mx_buf[c] = max(mx_buf[c], __shfl_xor_sync(0xFFFFFFFFu, mx_buf[c], o));
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    red[warp * 16 + c * 4 + 0] = mn_off[c];
                    red[warp * 16 + c * 4 + 1] = mx_end[c];
                    red[warp * 16 + c * 4 + 2] = mn_buf[c];
                    red[warp * 16 + c * 4 + 3] = mx_buf[c];
                }
            }
        }
        __syncthreads();
        if (!GATHER && tid < 2) {
            const int c = tid;
            uint32_t a0 = 0xFFFFFFFFu, a1 = 0, a2 = 0xFFFFFFFFu, a3 = 0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) {
                a0 = min(a0, red[w * 16 + c * 4 + 0]);
                a1 = max(a1, red[w * 16 + c * 4 + 1]);
                a2 = min(a2, red[w * 16 + c * 4 + 2]);
                a3 = max(a3, red[w * 16 + c * 4 + 3]);
            }
            // one buffer, a span no larger than a tile of maximum-length strings: contiguous payload
            if (a1 > a0 && a2 == a3 && a1 - a0 <= (uint32_t)(2 * CAP * TILE)) {
                const DevCol& col = c == 0 ? s.a : s.b;
                const uint32_t b16 = a0 & ~15u;
                tma_prefetch_l2(reinterpret_cast<const unsigned char*>(col.bufs[a2]) + b16,
                                ((a1 + 15u) & ~15u) - b16);
            }
        }

        // ---------------- 3. bucket: counting sort on the length key (descending) -------------------------
        if (warp == 0) {
            constexpr int CH = (NB + 31) / 32;
            uint32_t local[CH];
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < CH; q++) {
                const int bin = NB - 1 - (lane * CH + q);
                local[q] = bin >= 1 ? hist[bin] : 0u;
                sum += local[q];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int q = 0; q < CH; q++) {
                const int bin = NB - 1 - (lane * CH + q);
                if (bin >= 1) hist[bin] = run;
                run += local[q];
            }
            if (lane == 31) hist[0] = incl;
        }
        __syncthreads();
        const int n_active = (int)hist[0];
#pragma unroll
        for (int k = 0; k < RPT; k++)
            if (key[k]) perm[hist[key[k]] + rank[k]] = (uint16_t)(k * TPB + tid);
        __syncthreads();

        // ---------------- 4. compute: views and payload straight from L2 ----------------------------------
#pragma unroll 1
        for (int k = 0; k < RPT; k++) {
            const int p = k * TPB + ((k & 1) ? (TPB - 1 - tid) : tid);
            const bool has = p < n_active;
            long long row = 0;
            int na = 0, nb = 0;
            uint32_t or_bits = 0;
            if (has) {
                const long long idx = tile0 + perm[p];
                row = GATHER ? (long long)s.list[idx] : idx;
                const uint4 va = ld_view(s.a.views + row * s.a.stride);
                const uint4 vb = ld_view(s.b.views + row * s.b.stride);
                na = (int)va.x;
                nb = (int)vb.x;
                or_bits = load_string_global<TPB>(va, s.a, store.wa_) | load_string_global<TPB>(vb, s.b, store.wb_);
            }
            // one code path per warp: the Unicode path is also correct for ASCII pairs
            const bool ascii = ASCII_ONLY || __all_sync(0xFFFFFFFFu, (or_bits & 0x80808080u) == 0u);
            if (!has) continue;
            bool equal = na == nb;
            if (equal) {
                const int nw = (na + 3) >> 2;
                uint32_t diff = 0;
                for (int w = 0; w < nw; w++) diff |= store.wa(w) ^ store.wb(w);
                equal = diff == 0;
            }
            PairInts ints;
            const double v = row_short<M>(MEASURE, store, na, nb, equal, ascii, ints);
            s.out[row] = v;
            if (s.dbg) {
                int* d = s.dbg + row * 6;
                d[0] = ints.flag;
                d[1] = ints.la;
                d[2] = ints.lb;
                d[3] = ints.x0;
                d[4] = ints.x1;
                d[5] = ints.x2;
            }
        }
        __syncthreads();
    }
}

template <class M, int MEASURE, int TPB, int RPT, bool GATHER, int T, bool ASCII_ONLY>
__global__ void __launch_bounds__(TPB) direct_kernel(const SegArgs s) {
    direct_body<M, MEASURE, TPB, RPT, GATHER, T, ASCII_ONLY>(s);
}

// The 33..64-byte rows of a general column, every wanted measure in ONE launch: blockIdx.y names the
// measure (the five single-measure launches this replaces were a dozen small CTAs each and cost a launch
// latency apiece -- 150 us per segment for 0.05 % of C3's rows).  s.outs / s.dbgs by measure id.
template <int TPB, int T>
__global__ void __launch_bounds__(TPB) direct_multi64_kernel(const SegArgs s) {
    const int measure = (int)blockIdx.y;
    if (!s.outs[measure]) return;
    SegArgs m = s;
    m.out = s.outs[measure];
    m.dbg = s.dbgs[measure];
    switch (measure) {
        case 0: direct_body<uint64_t, 0, TPB, 1, true, T, false>(m); break;
        case 1: direct_body<uint64_t, 1, TPB, 1, true, T, false>(m); break;
        case 2: direct_body<uint64_t, 2, TPB, 1, true, T, false>(m); break;
        case 3: direct_body<uint64_t, 3, TPB, 1, true, T, false>(m); break;
        default: direct_body<uint64_t, 4, TPB, 1, true, T, false>(m); break;
    }
}

}  // namespace strsim
