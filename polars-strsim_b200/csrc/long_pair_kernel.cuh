// long_pair_kernel.cuh -- Jaro / Jaro-Winkler / Jaccard / Sorensen-Dice for rows that do not fit the
// 64-byte short-string kernels: one pair per WARP.
//
// The reference treats every length alike (strsim.rs:208-219 greedy windowed matching, :297-305 character
// multiset through a HashMap); the one-thread-per-pair fallback of generic_kernel.cuh did the same and
// took 3.9 ms for ONE 300-character Jaro pair (a quadratic scan by a single thread over HBM scratch).
// Here the 32 lanes of a warp share the pair:
//   * both strings are decoded once to code points (32 bytes per step, ballot-compacted) into the warp's
//     slab; the slab is private to the warp, a few KB for typical rows, and lives in L1/L2;
//   * Jaro match pass: for character i of a, the lanes compare 32 positions of b's window per step
//     (lane l looks at position 32k+l); `ballot & ~flag_b[k]` is the candidate set of that word, its
//     lowest bit the reference's first unflagged match (strsim.rs:211-217), found in at most
//     ceil(window/32) steps and usually in the first ones; flags are bit vectors, one word per 32
//     positions, read and written warp-uniformly;
//   * transpositions: the flagged characters of a and of b are compacted in place (ballot prefix sums),
//     then compared rank by rank, 32 ranks per step (strsim.rs:220-237); integer t/2, Winkler prefix on
//     code points (strsim.rs:241,260-267);
//   * multiset intersection: b's characters are counted in an open-addressing hash table of the slab
//     (atomicCAS on the key, atomicAdd on the count), every character of a then takes one unit of its
//     character's count if one is left (atomicSub, undone when it went below zero): sum of min(ca, cb)
//     without sorting (strsim.rs:297-305); union = la + lb - inter.
// One launch serves every wanted measure of the two groups (Jaro and Jaro-Winkler share m and t, Jaccard
// and Sorensen-Dice the intersection); Levenshtein has its own long kernel (long_lev_kernel.cuh).
// Arithmetic: the f64 formulas of pair_algos.cuh, same operation order as the reference.
#pragma once
#include "long_lev_kernel.cuh"

namespace strsim {

constexpr int LONGP_WPB = 4;  // warps per block

struct LongPairArgs {
    DevCol a, b;
    double* outs[5];  // by measure id; nullptr = not wanted (LEVENSHTEIN is never served here)
    int* dbgs[5];
    const unsigned int* list;        // segment-relative rows
    const unsigned int* list_count;  // device-resident count
    unsigned int* cursor;            // zero-initialised work counter
    unsigned char* scratch;
    long long slab_bytes;
    int n_warps;
    int cap_a, cap_b;  // max bytes of a / b over the listed rows (>= code points), multiples of 32
    int hash_size;     // power of two >= 2 * cap_b
};

__host__ __device__ inline long long long_pair_slab_bytes(int cap_a, int cap_b, int hash_size) {
    long long b = 4ll * cap_a + 4ll * cap_b;       // code points
    b += 4ll * (cap_a / 32 + 1) + 4ll * (cap_b / 32 + 1);  // flag words
    b = (b + 15) & ~15ll;
    b += 8ll * hash_size;                          // keys + counts
    return (b + 255) & ~255ll;
}

__device__ __forceinline__ void long_pair_store(const LongPairArgs& g, int measure, long long row, double v, int flag,
                                                int la, int lb, int x0, int x1, int x2) {
    if (!g.outs[measure]) return;
    g.outs[measure][row] = v;
    if (int* d = g.dbgs[measure]) {
        d += row * 6;
        d[0] = flag;
        d[1] = la;
        d[2] = lb;
        d[3] = x0;
        d[4] = x1;
        d[5] = x2;
    }
}

// flagged code points of `cp` (flag words `fw`, n positions) moved to the front, in order; returns their number
__device__ inline int long_pair_compact(uint32_t* cp, const uint32_t* fw, int n, int lane) {
    int count = 0;
    for (int w = 0; 32 * w < n; w++) {
        const uint32_t mask = fw[w];
        const uint32_t c = 32 * w + lane < n ? cp[32 * w + lane] : 0u;
        __syncwarp();  // every lane has read its character before any lane overwrites this word's range
        if ((mask >> lane) & 1u) cp[count + __popc(mask & ((1u << lane) - 1u))] = c;
        count += __popc(mask);
        __syncwarp();
    }
    return count;
}

__global__ void __launch_bounds__(32 * LONGP_WPB) long_pair_kernel(const LongPairArgs g) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * LONGP_WPB + (threadIdx.x >> 5);
    if (warp >= g.n_warps) return;
    unsigned char* slab = g.scratch + (long long)warp * g.slab_bytes;
    uint32_t* ca = reinterpret_cast<uint32_t*>(slab);
    uint32_t* cb = ca + g.cap_a;
    uint32_t* fa = cb + g.cap_b;
    uint32_t* fb = fa + (g.cap_a / 32 + 1);
    uint32_t* hkeys = reinterpret_cast<uint32_t*>(slab + ((4ll * g.cap_a + 4ll * g.cap_b + 4ll * (g.cap_a / 32 + 1) +
                                                           4ll * (g.cap_b / 32 + 1) + 15) & ~15ll));
    int* hcnt = reinterpret_cast<int*>(hkeys + g.hash_size);
    const unsigned int count = *g.list_count;
    const bool want_jaro = g.outs[JARO] || g.outs[JARO_WINKLER];
    const bool want_set = g.outs[JACCARD] || g.outs[SORENSEN_DICE];
    for (;;) {
        unsigned int e = 0;
        if (lane == 0) e = atomicAdd(g.cursor, 1u);
        e = __shfl_sync(0xFFFFFFFFu, e, 0);
        if (e >= count) break;
        const long long row = g.list[e];
        int na, nb;
        const unsigned char* pa = view_ptr(g.a, row, na);
        const unsigned char* pb = view_ptr(g.b, row, nb);
        // byte equality first (strsim.rs:182,288,324), then "one side empty"
        bool differ = na != nb;
        if (!differ)
            for (int i = lane; i < na; i += 32) differ = differ || pa[i] != pb[i];
        differ = __any_sync(0xFFFFFFFFu, differ);
        if (!differ || na == 0 || nb == 0) {
            if (lane == 0) {
                const double v = differ ? 0.0 : 1.0;
                const int flag = differ ? F_ONE_EMPTY : F_EQUAL;
                for (int m = JARO; m <= SORENSEN_DICE; m++) long_pair_store(g, m, row, v, flag, 0, 0, 0, 0, 0);
            }
            continue;
        }
        const int la = warp_decode(pa, na, ca, lane);
        const int lb = warp_decode(pb, nb, cb, lane);

        if (want_set) {
            // ---- character multiset intersection (strsim.rs:297-305) ------------------------------------
            int hs = 64;
            while (hs < 2 * lb) hs <<= 1;  // <= g.hash_size
            const uint32_t hmask = (uint32_t)hs - 1u;
            const int hshift = 32 - (31 - __clz(hs));
            for (int i = lane; i < hs; i += 32) {
                hkeys[i] = 0u;
                hcnt[i] = 0;
            }
            __syncwarp();
            for (int j = lane; j < lb; j += 32) {
                const uint32_t key = cb[j] + 1u;
                uint32_t slot = long_hash(cb[j], hshift);
                for (;;) {
                    const uint32_t k = atomicCAS(&hkeys[slot], 0u, key);
                    if (k == 0u || k == key) break;
                    slot = (slot + 1u) & hmask;
                }
                atomicAdd(&hcnt[slot], 1);
            }
            __syncwarp();
            int inter = 0;
            for (int i = lane; i < la; i += 32) {
                const uint32_t key = ca[i] + 1u;
                uint32_t slot = long_hash(ca[i], hshift);
                for (;;) {
                    const uint32_t k = __ldcg(&hkeys[slot]);  // written with atomicCAS (L2): do not trust L1
                    if (k == key) {
                        if (atomicSub(&hcnt[slot], 1) > 0) inter++;
                        else atomicAdd(&hcnt[slot], 1);
                        break;
                    }
                    if (k == 0u) break;
                    slot = (slot + 1u) & hmask;
                }
            }
            inter = __reduce_add_sync(0xFFFFFFFFu, inter);
            if (lane == 0) {
                long_pair_store(g, JACCARD, row, jaccard_value(inter, la + lb - inter), F_GENERAL, la, lb, inter,
                                la + lb - inter, 0);
                long_pair_store(g, SORENSEN_DICE, row, dice_value(inter, la + lb), F_GENERAL, la, lb, inter, la + lb, 0);
            }
            __syncwarp();
        }

        if (want_jaro) {
            // ---- Jaro match pass (strsim.rs:200-219) ----------------------------------------------------
            // (both strings longer than one character here: the pair left the 64-byte kernels)
            if (la == 1 && lb == 1) {  // strsim.rs:197; unreachable for rows above 64 bytes, kept for completeness
                if (lane == 0) {
                    const double v = ca[0] == cb[0] ? 1.0 : 0.0;
                    long_pair_store(g, JARO, row, v, F_SINGLE_CHAR, la, lb, 0, 0, 0);
                    long_pair_store(g, JARO_WINKLER, row, v, F_SINGLE_CHAR, la, lb, 0, 0, 0);
                }
                continue;
            }
            const int mx = la > lb ? la : lb;
            const int bound = mx / 2 - 1;
            const int outer = la < lb + bound ? la : lb + bound;
            for (int w = lane; 32 * w < la; w += 32) fa[w] = 0u;
            for (int w = lane; 32 * w < lb; w += 32) fb[w] = 0u;
            // Winkler prefix on code points, before the compaction below reorders the arrays
            int prefix = 0;
            {
                int lim = la < lb ? la : lb;
                if (lim > 4) lim = 4;
                while (prefix < lim && ca[prefix] == cb[prefix]) prefix++;
            }
            __syncwarp();
            int m = 0;
            for (int i = 0; i < outer; i++) {
                const uint32_t c = ca[i];
                const int lo = i > bound ? i - bound : 0;
                const int hi = i + bound < lb - 1 ? i + bound : lb - 1;
                for (int k = lo >> 5; k <= (hi >> 5); k++) {
                    const int j = 32 * k + lane;
                    const bool eq = j >= lo && j <= hi && cb[j] == c;
                    uint32_t cand = __ballot_sync(0xFFFFFFFFu, eq) & ~fb[k];
                    // lane 0 updates fb[k] below while other lanes may not have read it yet: every lane takes
                    // lane 0's view of the candidates, so the branch is warp-uniform by construction
                    cand = __shfl_sync(0xFFFFFFFFu, cand, 0);
                    if (cand) {
                        if (lane == 0) {
                            fb[k] |= cand & (0u - cand);  // the lowest unflagged match (strsim.rs:211-217)
                            fa[i >> 5] |= 1u << (i & 31);
                        }
                        m++;
                        __syncwarp();
                        break;
                    }
                }
            }
            __syncwarp();
            // ---- transpositions (strsim.rs:220-237): k-th flagged character of a vs k-th of b ------------
            int t = 0;
            if (m > 0) {
                long_pair_compact(ca, fa, la, lane);
                long_pair_compact(cb, fb, lb, lane);
                for (int k = lane; k < m; k += 32) t += ca[k] != cb[k] ? 1 : 0;
                t = __reduce_add_sync(0xFFFFFFFFu, t);
            }
            if (lane == 0) {
                const double js = m == 0 ? 0.0 : jaro_value(m, t, la, lb);
                long_pair_store(g, JARO, row, js, F_GENERAL, la, lb, m, t, 0);
                double jw = js;
                int l = 0;
                if (js > 0.7) {  // strsim.rs:260-267
                    l = prefix;
                    jw = winkler_value(js, l);
                }
                long_pair_store(g, JARO_WINKLER, row, jw, F_GENERAL, la, lb, m, t, l);
            }
            __syncwarp();
        }
    }
}

}  // namespace strsim
