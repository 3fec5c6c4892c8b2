// long_pair_kernel.cuh -- Jaro / Jaro-Winkler / Jaccard / Sorensen-Dice for rows that do not fit the
// 64-byte short-string kernels: one pair per WARP.
//
// The reference treats every length alike (strsim.rs:208-219 greedy windowed matching, :297-305 character
// multiset through a HashMap); the one-thread-per-pair fallback of generic_kernel.cuh did the same and
// took 3.9 ms for ONE 300-character Jaro pair (a quadratic scan by a single thread over HBM scratch).
// Here the 32 lanes of a warp share the pair:
//   * both strings are decoded once to code points (32 bytes per step, ballot-compacted) into the warp's
//     slab; the slab is private to the warp, a few KB for typical rows, and lives in L1/L2;
//   * the distinct characters of b get dense ids through a per-pair open-addressing hash table (any Unicode
//     scalar value); both strings become id sequences -- only equality of characters matters;
//   * Jaro match pass, b of at most 1024 characters: lane w owns word w (32 positions) of b; the position
//     masks Peq[id][w] are built once per pair, and for character i of a every lane tests its own word,
//     `Peq & window & ~flag_b`; one ballot names the lowest word with a candidate, whose lowest bit is the
//     reference's first unflagged match (strsim.rs:211-217) -- no loop over the window.  Longer b: the
//     lanes walk the window 32 positions per step (lane l at position 32k+l, `ballot & ~flag_b[k]`);
//   * transpositions: the flagged characters of a and of b are compacted in place (ballot prefix sums),
//     then compared rank by rank, 32 ranks per step (strsim.rs:220-237); integer t/2, Winkler prefix on
//     code points (strsim.rs:241,260-267);
//   * multiset intersection: b's characters are counted per id (atomicAdd), every character of a then takes
//     one unit of its character's count if one is left (atomicSub, undone when it went below zero): sum of
//     min(ca, cb) without sorting (strsim.rs:297-305); union = la + lb - inter.
// One launch serves every wanted measure of the two groups (Jaro and Jaro-Winkler share m and t, Jaccard
// and Sorensen-Dice the intersection); Levenshtein has its own long kernel (long_lev_kernel.cuh).
// Arithmetic: the f64 formulas of pair_algos.cuh, same operation order as the reference.
#pragma once
#include "long_lev_kernel.cuh"

namespace strsim {

constexpr int LONGP_WPB = 4;  // warps per block
constexpr int LONGP_PEQ_WORDS = 8192;  // position-mask words of one slab (32 KB): distinct characters x words of b
constexpr uint32_t LONGP_NONE = 0xFFFFFFFFu;  // id of a character of a that b does not contain

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

struct LongPairSlab {
    uint32_t *ca, *cb;      // code points, later character ids
    uint32_t *fa, *fb;      // Jaro flag words
    uint32_t *hkeys, *hvals;  // open-addressing hash: character + 1 -> dense id
    int* cnt;               // multiset: occurrences in b per id
    uint32_t* peq;          // Jaro fast path: position masks, [id][word of b]
};

__host__ __device__ inline long long long_pair_slab_bytes(int cap_a, int cap_b, int hash_size) {
    long long b = 4ll * cap_a + 4ll * cap_b;               // code points / ids
    b += 4ll * (cap_a / 32 + 1) + 4ll * (cap_b / 32 + 1);  // flag words
    b = (b + 15) & ~15ll;
    b += 8ll * hash_size;                                  // keys + ids
    b += 4ll * cap_b;                                      // counts per id
    b += 4ll * LONGP_PEQ_WORDS;
    return (b + 255) & ~255ll;
}

__device__ inline LongPairSlab long_pair_carve(unsigned char* base, const LongPairArgs& g) {
    LongPairSlab s;
    s.ca = reinterpret_cast<uint32_t*>(base);
    s.cb = s.ca + g.cap_a;
    s.fa = s.cb + g.cap_b;
    s.fb = s.fa + (g.cap_a / 32 + 1);
    const long long o = (4ll * g.cap_a + 4ll * g.cap_b + 4ll * (g.cap_a / 32 + 1) + 4ll * (g.cap_b / 32 + 1) + 15) & ~15ll;
    s.hkeys = reinterpret_cast<uint32_t*>(base + o);
    s.hvals = s.hkeys + g.hash_size;
    s.cnt = reinterpret_cast<int*>(s.hvals + g.hash_size);
    s.peq = reinterpret_cast<uint32_t*>(s.cnt + g.cap_b);
    return s;
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

// flagged entries of `cp` (flag words `fw`, n positions) moved to the front, in order; returns their number
__device__ inline int long_pair_compact(uint32_t* cp, const uint32_t* fw, int n, int lane) {
    int count = 0;
    for (int w = 0; 32 * w < n; w++) {
        const uint32_t mask = fw[w];
        const uint32_t c = 32 * w + lane < n ? cp[32 * w + lane] : 0u;
        __syncwarp();  // every lane has read its entry before any lane overwrites this word's range
        if ((mask >> lane) & 1u) cp[count + __popc(mask & ((1u << lane) - 1u))] = c;
        count += __popc(mask);
        __syncwarp();
    }
    return count;
}

// slot of character c in the pair's hash table, or LONGP_NONE
__device__ __forceinline__ uint32_t long_pair_find(const uint32_t* hkeys, uint32_t hmask, int hshift, uint32_t c) {
    const uint32_t key = c + 1u;
    uint32_t slot = long_hash(c, hshift);
    for (;;) {
        const uint32_t k = __ldcg(&hkeys[slot]);  // written with atomicCAS (L2): do not trust L1
        if (k == key) return slot;
        if (k == 0u) return LONGP_NONE;
        slot = (slot + 1u) & hmask;
    }
}

__global__ void __launch_bounds__(32 * LONGP_WPB) long_pair_kernel(const LongPairArgs g) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * LONGP_WPB + (threadIdx.x >> 5);
    if (warp >= g.n_warps) return;
    const LongPairSlab s = long_pair_carve(g.scratch + (long long)warp * g.slab_bytes, g);
    uint32_t *const ca = s.ca, *const cb = s.cb;
    const unsigned int count = *g.list_count;
    const bool want_jaro = g.outs[JARO] || g.outs[JARO_WINKLER];
    const bool want_set = g.outs[JACCARD] || g.outs[SORENSEN_DICE];
    const uint32_t lt_mask = (1u << lane) - 1u;
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
        if (la == 1 && lb == 1) {  // strsim.rs:197; unreachable for rows above 64 bytes, kept for completeness
            if (lane == 0) {
                const double v = ca[0] == cb[0] ? 1.0 : 0.0;
                long_pair_store(g, JARO, row, v, F_SINGLE_CHAR, la, lb, 0, 0, 0);
                long_pair_store(g, JARO_WINKLER, row, v, F_SINGLE_CHAR, la, lb, 0, 0, 0);
            }
            if (!want_set) continue;
        }
        // Winkler prefix on code points (strsim.rs:261-266), before the characters are replaced by ids
        int prefix = 0;
        {
            int lim = la < lb ? la : lb;
            if (lim > 4) lim = 4;
            while (prefix < lim && ca[prefix] == cb[prefix]) prefix++;
        }

        // ---- characters -> dense ids: the D distinct characters of b get ids 0..D-1 through the pair's hash
        // table; a character of a that b does not contain gets LONGP_NONE.  Only EQUALITY of characters
        // matters to every measure, so from here on both strings are id sequences.
        int hs = 64;
        while (hs < 2 * lb) hs <<= 1;  // <= g.hash_size
        const uint32_t hmask = (uint32_t)hs - 1u;
        const int hshift = 32 - (31 - __clz(hs));
        for (int i = lane; i < hs; i += 32) s.hkeys[i] = 0u;
        __syncwarp();
        for (int j = lane; j < lb; j += 32) {
            const uint32_t key = cb[j] + 1u;
            uint32_t slot = long_hash(cb[j], hshift);
            for (;;) {
                const uint32_t k = atomicCAS(&s.hkeys[slot], 0u, key);
                if (k == 0u || k == key) break;
                slot = (slot + 1u) & hmask;
            }
        }
        __syncwarp();
        int D = 0;
        for (int base = 0; base < hs; base += 32) {
            const bool has = __ldcg(&s.hkeys[base + lane]) != 0u;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
            if (has) s.hvals[base + lane] = (uint32_t)(D + __popc(bal & lt_mask));
            D += __popc(bal);
        }
        __syncwarp();
        for (int j = lane; j < lb; j += 32) cb[j] = s.hvals[long_pair_find(s.hkeys, hmask, hshift, cb[j])];
        for (int i = lane; i < la; i += 32) {
            const uint32_t slot = long_pair_find(s.hkeys, hmask, hshift, ca[i]);
            ca[i] = slot == LONGP_NONE ? LONGP_NONE : s.hvals[slot];
        }
        __syncwarp();

        if (want_set) {
            // ---- character multiset intersection (strsim.rs:297-305): sum over characters of min(ca, cb) ----
            for (int i = lane; i < D; i += 32) s.cnt[i] = 0;
            __syncwarp();
            for (int j = lane; j < lb; j += 32) atomicAdd(&s.cnt[cb[j]], 1);
            __syncwarp();
            int inter = 0;
            for (int i = lane; i < la; i += 32) {
                const uint32_t id = ca[i];
                if (id == LONGP_NONE) continue;
                // one unit of that character's count in b, if one is left
                if (atomicSub(&s.cnt[id], 1) > 0) inter++;
                else atomicAdd(&s.cnt[id], 1);
            }
            inter = __reduce_add_sync(0xFFFFFFFFu, inter);
            if (lane == 0) {
                long_pair_store(g, JACCARD, row, jaccard_value(inter, la + lb - inter), F_GENERAL, la, lb, inter,
                                la + lb - inter, 0);
                long_pair_store(g, SORENSEN_DICE, row, dice_value(inter, la + lb), F_GENERAL, la, lb, inter, la + lb, 0);
            }
            __syncwarp();
        }

        if (want_jaro && !(la == 1 && lb == 1)) {
            // ---- Jaro match pass (strsim.rs:200-219) ----------------------------------------------------
            const int mx = la > lb ? la : lb;
            const int bound = mx / 2 - 1;
            const int outer = la < lb + bound ? la : lb + bound;
            const int W = (lb + 31) >> 5;
            int m = 0;
            if (W <= 32 && la <= 1024 && D * W <= LONGP_PEQ_WORDS) {
                // Fast path (b of at most 1024 characters): lane w owns word w of b's positions.  The
                // position masks of b's characters are built once per pair (one __match_any per word: the
                // lanes that hold the same character ARE that character's bits in this word); per
                // character of a every lane then tests its own word -- `Eq & window & ~flag_b` -- and one
                // ballot finds the lowest word with a candidate: no loop over the window.
                for (int i = lane; i < D * W; i += 32) s.peq[i] = 0u;
                __syncwarp();
                for (int k = 0; k < W; k++) {
                    const int j = 32 * k + lane;
                    const uint32_t active = __ballot_sync(0xFFFFFFFFu, j < lb);
                    if (j < lb) {
                        const uint32_t id = cb[j];
                        const uint32_t same = __match_any_sync(active, id);
                        if (lane == __ffs((int)same) - 1) s.peq[id * (uint32_t)W + (uint32_t)k] = same;
                    }
                }
                __syncwarp();
                uint32_t fbreg = 0u, fareg = 0u;
                const int bit0 = 32 * lane;  // first position of this lane's word
                // lanes beyond b's last word read word 0 (their window mask is empty: no predicated load, and
                // the row address is one multiply-add on a pointer that stays in registers)
                const uint32_t* const peq_lane = s.peq + (lane < W ? lane : 0);
                const uint32_t* const ida = ca;
                uint32_t id = ida[0];
                for (int i = 0; i < outer; i++) {
                    const uint32_t cur = id;  // warp-uniform
                    id = ida[i + 1];          // next character's id, requested one iteration ahead (slab has slack)
                    if (cur == LONGP_NONE) continue;
                    const uint32_t eq = peq_lane[cur * (uint32_t)W];
                    // window [i - bound, i + bound] cut to this lane's 32 positions (strsim.rs:209-210)
                    const int l = i - bound - bit0, h = i + bound - bit0;  // first / last window position, word-relative
                    const int hc = min(h, lb - 1 - bit0);
                    uint32_t win = l > 0 ? (l > 31 ? 0u : 0xFFFFFFFFu << l) : 0xFFFFFFFFu;
                    win = hc < 31 ? (hc < 0 ? 0u : win & (0xFFFFFFFFu >> (31 - hc))) : win;
                    const uint32_t cand = eq & win & ~fbreg;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, cand != 0u);
                    if (bal) {
                        if (lane == __ffs((int)bal) - 1) fbreg |= cand & (0u - cand);  // the lowest unflagged match (strsim.rs:211-217)
                        if (lane == (i >> 5)) fareg |= 1u << (i & 31);
                        m++;
                    }
                }
                if (32 * lane < la) s.fa[lane] = fareg;
                if (lane < W) s.fb[lane] = fbreg;
            } else {
                // General path (any length): 32 positions of the window per step, lane l at position 32k+l
                for (int w = lane; 32 * w < la; w += 32) s.fa[w] = 0u;
                for (int w = lane; 32 * w < lb; w += 32) s.fb[w] = 0u;
                __syncwarp();
                for (int i = 0; i < outer; i++) {
                    const uint32_t c = ca[i];
                    if (c == LONGP_NONE) continue;
                    const int lo = i > bound ? i - bound : 0;
                    const int hi = i + bound < lb - 1 ? i + bound : lb - 1;
                    for (int k = lo >> 5; k <= (hi >> 5); k++) {
                        const int j = 32 * k + lane;
                        const bool eq = j >= lo && j <= hi && cb[j] == c;
                        uint32_t cand = __ballot_sync(0xFFFFFFFFu, eq) & ~s.fb[k];
                        // lane 0 updates fb[k] below while other lanes may not have read it yet: every lane takes
                        // lane 0's view of the candidates, so the branch is warp-uniform by construction
                        cand = __shfl_sync(0xFFFFFFFFu, cand, 0);
                        if (cand) {
                            if (lane == 0) {
                                s.fb[k] |= cand & (0u - cand);  // the lowest unflagged match (strsim.rs:211-217)
                                s.fa[i >> 5] |= 1u << (i & 31);
                            }
                            m++;
                            __syncwarp();
                            break;
                        }
                    }
                }
            }
            __syncwarp();
            // ---- transpositions (strsim.rs:220-237): k-th flagged character of a vs k-th of b ------------
            int t = 0;
            if (m > 0) {
                long_pair_compact(ca, s.fa, la, lane);
                long_pair_compact(cb, s.fb, lb, lane);
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
