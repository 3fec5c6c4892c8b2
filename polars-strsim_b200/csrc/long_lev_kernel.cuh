// long_lev_kernel.cuh -- Levenshtein for long strings: warp-cooperative multi-word Myers / Hyyro.
//
// One warp per pair (pairs are pulled from the overflow list through an atomic cursor, so long and
// short pairs balance dynamically).  The shorter string is the pattern (m codepoints, W = ceil(m/64)
// 64-cell blocks), the longer one the text (n codepoints).  Lane l owns K = ceil(W/32) consecutive
// blocks, with their vertical delta vectors Pv/Mv in registers.  The blocks of one DP column depend
// on each other top to bottom through the horizontal delta (hp, hm) that leaves a block at its last
// row, so the warp runs a WAVEFRONT: at step s lane l works on text column s - l and receives the
// carry of the lane above -- produced one step earlier for the same column -- with __shfl_up.  A pair
// takes n + L - 1 steps (L = lanes in use) instead of n * W sequential block updates.
//
// Pattern-match vectors: per pair, the pattern's distinct codepoints get dense ids through an
// open-addressing hash table (any Unicode scalar value), the text is translated to ids once, and
// Peq[id][block] (64-bit words) is built with atomicOr.  All of it lives in a per-warp scratch slab in
// HBM that stays L2-resident while the pair is processed; the Eq words of the next column are
// prefetched into registers while the current column is computed.
//
// Result: D[m][n] = n + sum over blocks of popc(Pv) - popc(Mv) (the final vertical deltas), then
// 1 - d / max(la, lb) exactly as /root/reference/src/expressions/strsim.rs:160.
// Patterns longer than LONG_PAT_MAX codepoints go to the generic kernel (their Peq would not fit the
// slab budget); nothing is approximated.
#pragma once
#include "generic_kernel.cuh"

namespace strsim {

constexpr int LONG_PAT_MAX = 8192;  // codepoints; 128 blocks = 4 per lane
constexpr int LONG_WPB = 4;         // warps per block
// The wavefront loop runs without bounds checks: the text's id array carries LONG_TID_PAD entries of the
// all-zero Peq row before its first and after its last column (lanes that have not started or have
// finished -- or wait for the longer pair of the warp -- prefetch those), and the Peq area LONG_PEQ_PAD spare words (the last lane's blocks past W).
constexpr int LONG_TID_PAD = 304;   // 48 for one pair's own lanes + LONG_PAIR_SLACK
constexpr int LONG_PAIR_SLACK = 256;  // two pairs share a warp only if their step counts differ by less
constexpr int LONG_PEQ_PAD = 8;

struct LongLevArgs {
    DevCol a, b;
    double* out;
    int* dbg;
    const unsigned int* list;
    const unsigned int* list_count;
    unsigned int* cursor;      // zero-initialised work counter
    unsigned int* huge_list;   // rows whose pattern exceeds LONG_PAT_MAX
    unsigned int* huge_count;
    unsigned char* scratch;
    long long slab_bytes;
    int n_warps;
    int cap_a, cap_b;  // max bytes of a / b over the listed rows (>= codepoints)
    int cap_pat;       // min(cap_a, cap_b, LONG_PAT_MAX)
    int hash_size;     // power of two >= 2 * cap_pat (slab capacity; each pair uses what it needs)
    int w_max;         // ceil(cap_pat / 64)
    long long peq_words;  // capacity of one slab's Peq area in 64-bit words; a pair that needs more is
                          // deferred to `huge_list` (the host relaunches it with worst-case slabs)
};

struct LongLevSlab {
    uint32_t* cps_a;  // the PATTERN's code points (cap_pat entries); the text is never stored decoded
    uint16_t* tid;
    uint32_t* hkeys;
    uint32_t* hvals;  // per hash slot: occurrence count while the pattern is inserted, then the character's code
    unsigned long long* peq;
};

__host__ __device__ inline long long long_lev_slab_bytes(int cap_a, int cap_b, int cap_pat, int hash_size, long long peq_words) {
    long long b = 0;
    b += 4ll * cap_pat;
    b += 2ll * ((cap_a > cap_b ? cap_a : cap_b) + 2 * LONG_TID_PAD);
    b = (b + 15) & ~15ll;
    b += 4ll * hash_size;
    b += 4ll * hash_size;
    b = (b + 15) & ~15ll;
    b += 8ll * (peq_words + LONG_PEQ_PAD);
    return (b + 255) & ~255ll;
}

__device__ inline LongLevSlab long_lev_carve(unsigned char* base, const LongLevArgs& g) {
    LongLevSlab s;
    long long o = 0;
    s.cps_a = reinterpret_cast<uint32_t*>(base + o);
    o += 4ll * g.cap_pat;
    s.tid = reinterpret_cast<uint16_t*>(base + o);
    o += 2ll * ((g.cap_a > g.cap_b ? g.cap_a : g.cap_b) + 2 * LONG_TID_PAD);
    o = (o + 15) & ~15ll;
    s.hkeys = reinterpret_cast<uint32_t*>(base + o);
    o += 4ll * g.hash_size;
    s.hvals = reinterpret_cast<uint32_t*>(base + o);
    o += 4ll * g.hash_size;
    o = (o + 15) & ~15ll;
    s.peq = reinterpret_cast<unsigned long long*>(base + o);
    return s;
}

// UTF-8 -> Unicode scalar values, 32 bytes per iteration (valid UTF-8, as Polars guarantees)
__device__ inline int warp_decode(const unsigned char* p, int nbytes, uint32_t* out, int lane) {
    int count = 0;
    for (int base = 0; base < nbytes; base += 32) {
        const int i = base + lane;
        const uint32_t c = i < nbytes ? p[i] : 0x80u;
        const bool lead = i < nbytes && (c & 0xC0u) != 0x80u;
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, lead);
        if (lead) {
            int len = c < 0x80u ? 1 : c < 0xE0u ? 2 : c < 0xF0u ? 3 : 4;
            if (len > nbytes - i) len = nbytes - i;
            uint32_t cp = len == 1 ? c : (c & (0xFFu >> (len + 1)));
            for (int e = 1; e < len; e++) cp = (cp << 6) | (p[i + e] & 0x3Fu);
            out[count + __popc(mask & ((1u << lane) - 1u))] = cp;
        }
        count += __popc(mask);
    }
    __syncwarp();
    return count;
}

// number of characters (bytes that are not UTF-8 continuation bytes) of p[0..nbytes), by the whole warp:
// aligned 4-byte loads, 128 bytes per step
__device__ inline int warp_count_chars(const unsigned char* p, int nbytes, int lane) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const int head = (int)(addr & 3);  // bytes of the first aligned word that precede the string
    const uint32_t* w = reinterpret_cast<const uint32_t*>(addr - (uintptr_t)head);
    const int total = head + nbytes, nwords = (total + 3) >> 2;
    int cont = 0;
    for (int i = lane; i < nwords; i += 32) {
        uint32_t x = w[i];  // reads at most 3 bytes before / after the string: inside the padded device buffer
        if (i == 0) x &= ~byte_mask(head);
        if (i == nwords - 1) x &= byte_mask(total - 4 * i);
        cont += __popc(((x >> 7) & ~(x >> 6)) & 0x01010101u);
    }
    cont = __reduce_add_sync(0xFFFFFFFFu, cont);
    return nbytes - cont;
}

__device__ __forceinline__ uint32_t long_hash(uint32_t cp, int shift) { return (cp * 2654435761u) >> shift; }

// Character codes (16 bit, what the text is translated to): a character that occurs at least twice in the
// pattern owns a Peq row, code = row id (< 0x8000); a character that occurs ONCE has no row -- its position
// mask is a single bit, code = LONG_SINGLE | position, and the lane that owns that block sets the bit in
// registers; a character the pattern does not contain gets the id of the all-zero row.  On C4 (a quarter
// of the characters are spaces and letters, the rest rare 2- and 3-byte code points) this shrinks the Peq
// of a pair from ~230 rows to ~40: the matrix used to stream from HBM and now stays in L2.
constexpr uint32_t LONG_SINGLE = 0x8000u;
constexpr uint32_t LONG_PENDING = 0xFFFFFFFFu;  // a single character whose position is not recorded yet

__device__ __forceinline__ uint32_t long_find_slot(const LongLevSlab& s, uint32_t hmask, int hshift, uint32_t cp) {
    const uint32_t key = cp + 1u;  // returns the slot holding cp, or ~0u when the pattern does not contain it
    uint32_t slot = long_hash(cp, hshift);
    for (;;) {
        const uint32_t k = __ldcg(&s.hkeys[slot]);  // written with atomicCAS (L2): do not trust L1
        if (k == key) return slot;
        if (k == 0u) return ~0u;
        slot = (slot + 1u) & hmask;
    }
}

// The K Eq words of one lane: consecutive 64-bit words of a Peq row, starting at the lane's first block.
// Every lane reads another row (its own text column), so a load instruction touches up to 32 sectors
// whatever its width -- the round-1 kernel issued K separate 64-bit loads and kept the L1 data pipe 82 %
// busy.  Rows are 16-byte aligned (even stride), so K = 2 and K = 4 are one / two 128-bit loads; K = 3
// starts at an odd word for odd lanes: two 128-bit loads from the aligned word below, and the three
// wanted words are picked by the lane's parity.  (Reads one word outside the lane's blocks: inside the
// row, or inside LONG_PEQ_PAD for the last lane.)
template <int K>
__device__ __forceinline__ void long_load_eq(const unsigned long long* row, bool odd, uint64_t (&eq)[K]) {
    if (K == 1) {
        eq[0] = row[0];
    } else if (K == 2) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row);
        eq[0] = v.x;
        eq[1] = v.y;
    } else if (K == 4) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row), w = *reinterpret_cast<const ulonglong2*>(row + 2);
        eq[0] = v.x;
        eq[1] = v.y;
        eq[2] = w.x;
        eq[3] = w.y;
    } else {  // K == 3
        const unsigned long long* base = row - (odd ? 1 : 0);
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(base), w = *reinterpret_cast<const ulonglong2*>(base + 2);
        eq[0] = odd ? v.y : v.x;
        eq[1] = odd ? w.x : v.y;
        eq[2] = odd ? w.y : w.x;
    }
}

// tid = the text's ids with LONG_TID_PAD padding entries in front (tid[0] is column 0).
// Per step a lane does: one shuffle for the carry, one 16-bit load of the id three columns ahead, one
// IMAD.WIDE + K loads for the Eq words two columns ahead, and -- between its first and last column --
// K Myers blocks.  The step loop is unrolled three times so that the three Eq buffers and the three
// ids rotate by renaming, not by moves.
// WIDTH = lanes per pair: 32 (one pair per warp) or 16 (two pairs per warp, each half with its own slab,
// lengths and lane count; `lane` is the lane inside the group, `steps` the step count of the whole warp).
template <int K, int WIDTH>
__device__ inline int long_wavefront(const LongLevSlab& s, const uint16_t* tid, uint32_t zero_row, int m, int n, int W,
                                     int L, int lane, int steps) {
    uint64_t Pv[K], Mv[K], eq[3][K];
    uint32_t tq[3];
#pragma unroll
    for (int k = 0; k < K; k++) {
        Pv[k] = ~0ull;
        Mv[k] = 0ull;
    }
    const bool lane_on = lane < L;
    const int blk0 = lane_on ? lane * K : 0;  // idle lanes prefetch block 0 (harmless) and never compute
    const unsigned n_on = lane_on ? (unsigned)n : 0u;
    const unsigned long long* rowbase = s.peq + blk0;
    const bool odd = (blk0 & 1) != 0;  // K == 3 only: the lane's first block sits at an odd word of its (16-byte aligned) row
    const unsigned W8 = (unsigned)((W + 1) & ~1);  // row stride of Peq: even, so that rows start 16-byte aligned
    const uint16_t* tp = tid - lane;  // tp[st + c] = id of column (st - lane) + c
    // codes of this lane's columns j, j+1, j+2 sit in tq[1], tq[2], tq[0] at step 0 and rotate from there
    tq[1] = tp[0];
    tq[2] = tp[1];
    tq[0] = tp[2];
    {
        const unsigned long long* r0 = rowbase + (size_t)((tq[1] & LONG_SINGLE) ? zero_row : tq[1]) * W8;
        const unsigned long long* r1 = rowbase + (size_t)((tq[2] & LONG_SINGLE) ? zero_row : tq[2]) * W8;
        long_load_eq<K>(r0, odd, eq[0]);
        long_load_eq<K>(r1, odd, eq[1]);
#pragma unroll
        for (int k = 0; k < K; k++) eq[2][k] = 0ull;
    }
    uint32_t carry_prev = 0;  // bit 0: hp, bit 1: hm of this lane's last block in the previous step
#pragma unroll 1
    const uint16_t* q = tp + 3;  // q[u] = id of the column three ahead of step st0 + u
    int j0 = -lane;              // this lane's column at step st0
    for (int st0 = 0; st0 < steps; st0 += 3, q += 3, j0 += 3) {
#pragma unroll
        for (int u = 0; u < 3; u++) {
            uint32_t carry = __shfl_up_sync(0xFFFFFFFFu, carry_prev, 1, WIDTH);
            if (lane == 0) carry = 1u;  // D[0][j] - D[0][j-1] = +1
            const uint32_t tcur = tq[(u + 1) % 3];  // code of the column computed in this step
            tq[(u + 1) % 3] = q[u];
            {
                const uint32_t t = tq[u];
                const unsigned long long* row = rowbase + (size_t)((t & LONG_SINGLE) ? zero_row : t) * W8;
                long_load_eq<K>(row, odd, eq[(u + 2) % 3]);
            }
            if ((unsigned)(j0 + u) < n_on) {
                // a character that occurs once in the pattern: its bit, if it falls into this lane's blocks
                const unsigned rel = ((tcur & (LONG_SINGLE - 1u)) >> 6) - (unsigned)blk0;
                if ((tcur & LONG_SINGLE) && rel < (unsigned)K) {
                    const uint64_t bit = 1ull << (tcur & 63u);
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (rel == (unsigned)k) eq[u][k] |= bit;
                }
                uint32_t hp = carry & 1u, hm = carry >> 1;
#pragma unroll
                for (int k = 0; k < K; k++) myers_block(Pv[k], Mv[k], eq[u][k], hp, hm);
                carry_prev = hp | (hm << 1);
            }
        }
    }
    int score = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
        if (lane_on && blk0 + k < W) score += myers_block_score(Pv[k], Mv[k], m - 64 * (blk0 + k));
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, o);
    return n + score;
}

// What the prelude of one pair leaves behind.
struct LongPair {
    int status;  // 0: run the wavefront, 1: settled (v, pi), 2: deferred to the huge list (nothing is written)
    long long row;
    int la, lb, m, n, W;
    const uint16_t* tid;
    uint32_t zero_row;
    double v;
    PairInts pi;
};

// Equality test, decoding, pattern hash, text ids and Peq of one pair, by the whole warp, into slab s.
__device__ __noinline__ void long_prepare(const LongLevArgs& g, const LongLevSlab& s, long long row, int lane, LongPair& o) {
    o.row = row;
    o.status = 1;
    o.pi = {F_GENERAL, 0, 0, 0, 0, 0};
    o.v = 0.0;
    o.m = o.n = o.W = o.la = o.lb = 0;
    o.tid = s.tid + LONG_TID_PAD;
    int na, nb;
    const unsigned char* pa = view_ptr(g.a, row, na);
    const unsigned char* pb = view_ptr(g.b, row, nb);
    bool differ = na != nb;
    if (!differ) {
        bool d = false;
        for (int i = lane; i < na; i += 32) d = d || pa[i] != pb[i];
        differ = __any_sync(0xFFFFFFFFu, d);
    }
    if (!differ) {
        o.pi.flag = F_EQUAL;
        o.v = 1.0;
        return;
    }
    // Character counts first: they decide which string is the pattern (the shorter one), and only the
    // PATTERN is kept as code points -- the text is decoded and translated to codes in one pass (step 4).
    // (The first version stored both strings as code points, 34 KB written and read again per C4 pair; with
    // thousands of pairs in flight the slabs outgrew the L2 and that traffic went to HBM.)
    const int la = warp_count_chars(pa, na, lane);
    const int lb = warp_count_chars(pb, nb, lane);
    o.la = o.pi.la = la;
    o.lb = o.pi.lb = lb;
    const bool pat_is_b = lb <= la;
    const unsigned char* const pat_bytes = pat_is_b ? pb : pa;
    const unsigned char* const txt_bytes = pat_is_b ? pa : pb;
    const int pat_nbytes = pat_is_b ? nb : na, txt_nbytes = pat_is_b ? na : nb;
    const int m = pat_is_b ? lb : la, n = pat_is_b ? la : lb;
    o.m = m;
    o.n = n;
    if (m > g.cap_pat) {  // Peq would not fit the slab: the generic kernel finishes this row
        if (lane == 0) g.huge_list[atomicAdd(g.huge_count, 1u)] = (unsigned int)row;
        o.status = 2;
        return;
    }
    if (m == 0) {
        o.pi.x0 = n;
        o.v = lev_value(n, la, lb);
        return;
    }
    uint32_t* const P = s.cps_a;
    warp_decode(pat_bytes, pat_nbytes, P, lane);
    const int W = (m + 63) >> 6;
    o.W = W;
    // 1. hash set of the pattern's codepoints.  Sized for the DISTINCT characters, which are few in text (230
    // of 2100 on C4): start with 1024 slots and grow only when the table turns out more than half full, so a
    // typical pair zeroes and probes 8 KB instead of the 32 KB a table for 2m keys would take.
    int hbits = 6;
    while ((1 << hbits) < 2 * m && hbits < 10) hbits++;
    int hsize, hshift;
    uint32_t hmask;
    for (;;) {
        hsize = 1 << hbits;
        hshift = 32 - hbits;
        hmask = (uint32_t)hsize - 1u;
        for (int i = lane; i < hsize; i += 32) {
            s.hkeys[i] = 0u;
            s.hvals[i] = 0u;
        }
        __syncwarp();
        int fresh = 0;
        bool full = false;
        for (int i = lane; i < m; i += 32) {
            const uint32_t key = P[i] + 1u;
            uint32_t slot = long_hash(P[i], hshift);
            int probes = 0;
            for (;;) {
                const uint32_t old = atomicCAS(&s.hkeys[slot], 0u, key);
                if (old == 0u) fresh++;
                if (old == 0u || old == key) break;
                slot = (slot + 1u) & hmask;
                if (++probes > hsize) {
                    full = true;
                    break;
                }
            }
            if (full) break;
            atomicAdd(&s.hvals[slot], 1u);  // occurrences of this character in the pattern
        }
        fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
        full = __any_sync(0xFFFFFFFFu, full);
        if (!full && 2 * fresh <= hsize) break;
        hbits += 2;  // (1 << hbits) stays within the slab: it ends at the first power of two >= 2m
        while ((1 << (hbits - 1)) >= 2 * m && hbits > 6) hbits--;
        __syncwarp();
    }
    __syncwarp();
    // 2. dense row ids, in slot order, for the characters that occur at least twice
    uint32_t rows = 0;
    for (int base = 0; base < hsize; base += 32) {
        const uint32_t cnt = __ldcg(&s.hvals[base + lane]);
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, cnt >= 2u);
        if (cnt >= 2u)
            s.hvals[base + lane] = rows + __popc(mask & ((1u << lane) - 1u));
        else if (cnt == 1u)
            s.hvals[base + lane] = LONG_PENDING;
        rows += __popc(mask);
    }
    __syncwarp();
    o.zero_row = rows;  // row `rows` stays all zero: characters the pattern does not contain
    // 3. Peq[row][block]; single characters record their position instead
    const int Wp = (W + 1) & ~1;  // row stride (see long_load_eq)
    const size_t words = (size_t)(rows + 1u) * Wp;
    if ((long long)words > g.peq_words) {  // needs a bigger slab: second launch
        if (lane == 0) g.huge_list[atomicAdd(g.huge_count, 1u)] = (unsigned int)row;
        o.status = 2;
        return;
    }
    for (size_t i = lane; i < words; i += 32) s.peq[i] = 0ull;
    __syncwarp();
    for (int i = lane; i < m; i += 32) {
        const uint32_t slot = long_find_slot(s, hmask, hshift, P[i]);
        const uint32_t code = __ldcg(&s.hvals[slot]);
        if (code == LONG_PENDING)
            s.hvals[slot] = LONG_SINGLE | (uint32_t)i;  // the only occurrence: nobody else writes this slot
        else
            atomicOr(&s.peq[(size_t)code * Wp + (i >> 6)], 1ull << (i & 63));
    }
    __syncwarp();
    __threadfence();  // the atomics landed in L2; drop possibly stale L1 lines before reading Peq
    // 4. text -> codes, padded on both sides with the all-zero row: decoded 32 bytes per step and looked up
    // at once (valid UTF-8, as Polars guarantees)
    uint16_t* tid = s.tid + LONG_TID_PAD;
    {
        int done = 0;
        for (int base = 0; base < txt_nbytes; base += 32) {
            const int i = base + lane;
            const uint32_t c = i < txt_nbytes ? txt_bytes[i] : 0x80u;
            const bool lead = i < txt_nbytes && (c & 0xC0u) != 0x80u;
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, lead);
            if (lead) {
                int len = c < 0x80u ? 1 : c < 0xE0u ? 2 : c < 0xF0u ? 3 : 4;
                if (len > txt_nbytes - i) len = txt_nbytes - i;
                uint32_t cp = len == 1 ? c : (c & (0xFFu >> (len + 1)));
                for (int e = 1; e < len; e++) cp = (cp << 6) | (txt_bytes[i + e] & 0x3Fu);
                const uint32_t slot = long_find_slot(s, hmask, hshift, cp);
                tid[done + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)(slot == ~0u ? rows : __ldcg(&s.hvals[slot]));
            }
            done += __popc(mask);
        }
    }
    for (int j = lane; j < LONG_TID_PAD; j += 32) {
        s.tid[j] = (uint16_t)rows;
        tid[n + j] = (uint16_t)rows;
    }
    __syncwarp();
    o.status = 0;
}

__device__ inline void long_store(const LongLevArgs& g, const LongPair& o) {
    g.out[o.row] = o.v;
    if (g.dbg) {
        int* d = g.dbg + o.row * 6;
        d[0] = o.pi.flag;
        d[1] = o.pi.la;
        d[2] = o.pi.lb;
        d[3] = o.pi.x0;
        d[4] = o.pi.x1;
        d[5] = o.pi.x2;
    }
}

// one pair on all 32 lanes
__device__ inline int long_run_single(const LongLevSlab& s, const LongPair& o, int lane) {
    const int K = (o.W + 31) >> 5;
    const int L = (o.W + K - 1) / K;
    const int steps = o.n + L - 1;
    switch (K) {
        case 1: return long_wavefront<1, 32>(s, o.tid, o.zero_row, o.m, o.n, o.W, L, lane, steps);
        case 2: return long_wavefront<2, 32>(s, o.tid, o.zero_row, o.m, o.n, o.W, L, lane, steps);
        case 3: return long_wavefront<3, 32>(s, o.tid, o.zero_row, o.m, o.n, o.W, L, lane, steps);
        default: return long_wavefront<4, 32>(s, o.tid, o.zero_row, o.m, o.n, o.W, L, lane, steps);
    }
}

// Each warp owns TWO slabs and takes two list entries at a time.  The host sorts the list by (text
// blocks, pattern blocks), so neighbours cost nearly the same number of steps and blocks: when both fit
// 16 lanes (W <= 64 blocks at up to four per lane) the two wavefronts run side by side, lanes 0-15 on
// the first pair and lanes 16-31 on the second -- a pattern of W blocks keeps ceil(W/K) lanes busy, so
// one pair per warp left a third of the lanes idle on C4, and every step's fixed cost (carry shuffle,
// id and Eq loads, bounds test) is now shared by up to four blocks per lane instead of two.
__global__ void __launch_bounds__(32 * LONG_WPB, 8) long_lev_kernel(const LongLevArgs g) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * LONG_WPB + (threadIdx.x >> 5);
    if (warp >= g.n_warps) return;
    const LongLevSlab s0 = long_lev_carve(g.scratch + (long long)(2 * warp) * g.slab_bytes, g);
    const LongLevSlab s1 = long_lev_carve(g.scratch + (long long)(2 * warp + 1) * g.slab_bytes, g);
    const unsigned int count = *g.list_count;
    for (;;) {
        unsigned int e = 0;
        if (lane == 0) e = atomicAdd(g.cursor, 2u);
        e = __shfl_sync(0xFFFFFFFFu, e, 0);
        if (e >= count) break;
        LongPair p0, p1;
        long_prepare(g, s0, g.list[e], lane, p0);
        p1.status = 2;
        if (e + 1 < count) long_prepare(g, s1, g.list[e + 1], lane, p1);
        const int K0 = (p0.W + 15) >> 4, K1 = (p1.W + 15) >> 4;
        const int K = K0 > K1 ? K0 : K1;
        const int L0 = K ? (p0.W + K - 1) / K : 0, L1 = K ? (p1.W + K - 1) / K : 0;
        const int st0 = p0.n + L0 - 1, st1 = p1.n + L1 - 1;
        const int steps = st0 > st1 ? st0 : st1;
        if (p0.status == 0 && p1.status == 0 && K <= 4 && steps - (st0 < st1 ? st0 : st1) < LONG_PAIR_SLACK) {
            const bool hi = lane >= 16;
            const LongLevSlab& s = hi ? s1 : s0;
            const LongPair& p = hi ? p1 : p0;
            const int L = hi ? L1 : L0;
            int d;
            switch (K) {
                case 1: d = long_wavefront<1, 16>(s, p.tid, p.zero_row, p.m, p.n, p.W, L, lane & 15, steps); break;
                case 2: d = long_wavefront<2, 16>(s, p.tid, p.zero_row, p.m, p.n, p.W, L, lane & 15, steps); break;
                case 3: d = long_wavefront<3, 16>(s, p.tid, p.zero_row, p.m, p.n, p.W, L, lane & 15, steps); break;
                default: d = long_wavefront<4, 16>(s, p.tid, p.zero_row, p.m, p.n, p.W, L, lane & 15, steps); break;
            }
            const int d0 = __shfl_sync(0xFFFFFFFFu, d, 0), d1 = __shfl_sync(0xFFFFFFFFu, d, 16);
            p0.pi.x0 = d0;
            p0.v = lev_value(d0, p0.la, p0.lb);
            p1.pi.x0 = d1;
            p1.v = lev_value(d1, p1.la, p1.lb);
            p0.status = p1.status = 1;
        } else {
            if (p0.status == 0) {
                const int d = long_run_single(s0, p0, lane);
                p0.pi.x0 = d;
                p0.v = lev_value(d, p0.la, p0.lb);
                p0.status = 1;
            }
            if (p1.status == 0) {
                const int d = long_run_single(s1, p1, lane);
                p1.pi.x0 = d;
                p1.v = lev_value(d, p1.la, p1.lb);
                p1.status = 1;
            }
        }
        if (lane == 0) {
            if (p0.status == 1) long_store(g, p0);
            if (p1.status == 1) long_store(g, p1);
        }
        __syncwarp();
    }
}

// ---- sorting the long list by cost ----------------------------------------------------------------------
// key = (64-byte blocks of the longer string, 64-byte blocks of the shorter one), both clamped to 127:
// bytes bound codepoints from above and are what the views hold.  Longest first, so that the tail of the
// launch is made of short pairs.  Counting sort in three small launches over the list.
constexpr int LONG_KEY_BITS = 7;
constexpr int LONG_KEYS = 1 << (2 * LONG_KEY_BITS);

__device__ __forceinline__ unsigned int long_sort_key(const DevCol& a, const DevCol& b, long long row) {
    const unsigned int na = a.views[row * a.stride].x, nb = b.views[row * b.stride].x;
    unsigned int hi = (na > nb ? na : nb) >> 6, lo = (na > nb ? nb : na) >> 6;
    const unsigned int top = (1u << LONG_KEY_BITS) - 1u;
    if (hi > top) hi = top;
    if (lo > top) lo = top;
    return (unsigned int)LONG_KEYS - 1u - ((hi << LONG_KEY_BITS) | lo);  // descending cost
}

__global__ void long_sort_hist_kernel(DevCol a, DevCol b, const unsigned int* list, const unsigned int* list_count,
                                      unsigned int* hist) {
    const unsigned int n = *list_count;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&hist[long_sort_key(a, b, list[i])], 1u);
}

// exclusive scan of LONG_KEYS bins by one CTA of 1024 threads (16 bins per thread)
__global__ void long_sort_scan_kernel(unsigned int* hist) {
    __shared__ unsigned int part[1024];
    constexpr int PER = LONG_KEYS / 1024;
    unsigned int local[PER], sum = 0;
    for (int q = 0; q < PER; q++) {
        local[q] = hist[threadIdx.x * PER + q];
        sum += local[q];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned int t = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned int run = part[threadIdx.x] - sum;
    for (int q = 0; q < PER; q++) {
        hist[threadIdx.x * PER + q] = run;
        run += local[q];
    }
}

__global__ void long_sort_scatter_kernel(DevCol a, DevCol b, const unsigned int* list, const unsigned int* list_count,
                                         unsigned int* start, unsigned int* sorted) {
    const unsigned int n = *list_count;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned int row = list[i];
        sorted[atomicAdd(&start[long_sort_key(a, b, row)], 1u)] = row;
    }
}

}  // namespace strsim
