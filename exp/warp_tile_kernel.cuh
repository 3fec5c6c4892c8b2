// exp/warp_tile_kernel.cuh -- EXPERIMENT (round 2, measured and not adopted; not part of the build).
// The plane path of short_kernel.cuh with WARP-PRIVATE tiles: no CTA barrier.
//
// Result on C2 (10 M pairs x 5 measures, one B200; profiles/r2_warp_tile_experiment.md): bit-exact, barrier
// stalls gone, issue slots busy 61 % -> 65 % -- but a warp that sorts only 96 rows keeps 21.3 of 32 lanes
// busy instead of 26.5, so the same thread instructions take 25 % more warp instructions: 0.748 ms against
// 0.641 ms.  Larger warp tiles restore the sort (RPT 4 / 6 / 8) but halve the resident warps: 0.739 / 0.832 /
// 1.094 ms.  To build it again: include it from host.cu and dispatch launch_warp_tile (git history, the
// commit that added this file).
//
//
// short_kernel.cuh walks tiles of 768 / 1024 rows with all 8 warps of a CTA in lock step: load ->
// __syncthreads -> TMA wait -> sort -> __syncthreads x3 -> compute -> __syncthreads.  The round-1 profile
// (profiles/r1_ncu_regions_C2_fused.txt) charged those fixed phases 25 % of the instructions but 43 % of
// the stall samples: while a CTA loads or sorts, its 8 warps issue next to nothing, and only the other
// three CTAs of the SM fill the gap -- 60 % of the issue slots were used.
//
// Here every warp owns its tiles: 32 x RPT consecutive rows, its own slice of shared memory (views, stage
// area, histogram, permutation, mbarrier), its own TMA bulk copies, its own counting sort, and only
// __syncwarp between the steps.  The 32 warps of an SM drift apart within a few tiles, so at any time
// they are spread over all phases and the latency of one warp's loads / TMA / sort is covered by the
// compute of the others -- the pipelining a producer warp would buy, without a second buffer.  Price: a
// warp sorts 96 rows instead of a CTA 768, so the 32 pairs of a round span more lengths (about 7 values
// of 4..24 instead of 2): a few more idle lanes in the byte loops.
//
// Serves ASCII-only columns (the plane path, REG) in their natural row order; general columns, gather
// launches over overflow lists and the table / hash path stay with short_kernel.cuh.  Per-pair functions,
// routing of rows that do not fit, and results are those of short_kernel.cuh.
#pragma once
#include "short_kernel.cuh"

namespace strsim {

template <int RPT, int CAP>
struct WarpTileLayout {
    static constexpr int TILE = 32 * RPT;
    static constexpr int NB = 2 * CAP + 3;  // keys 0 .. 2*CAP+2
    static constexpr size_t off_sva = 0;
    static constexpr size_t off_svb = off_sva + 16 * TILE;
    static constexpr size_t off_hist = off_svb + 16 * TILE;
    static constexpr size_t off_mbar = off_hist + 4 * ((NB + 3) & ~3);
    static constexpr size_t off_perm = off_mbar + 16;
    static constexpr size_t off_stage = (off_perm + 2 * TILE + 15) & ~(size_t)15;
    // bytes of one warp's slice; stage_bytes = capacity of each column's stage area (multiple of 16)
    static size_t warp_bytes(int stage_bytes) { return (off_stage + 2 * ((size_t)stage_bytes + 16) + 127) & ~(size_t)127; }
};

template <class M, int MEASURE, int WPB, int RPT, int T>
__global__ void __launch_bounds__(32 * WPB) warp_tile_kernel(const SegArgs s, const int warp_bytes) {
    static_assert(T == 32 || T == 64 || T == 128, "plane path: 5, 6 or 7 bit planes");
    constexpr int GROUPS = is_multi(MEASURE) ? MEASURE - MULTI_BASE : 0;
    constexpr int NBITS = T == 32 ? 5 : T == 64 ? 6 : 7;
    constexpr int CAP = (int)sizeof(M) * 8;
    using L = WarpTileLayout<RPT, CAP>;
    constexpr int TILE = L::TILE;
    constexpr int NB = L::NB;

    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* const wbase = smem + (size_t)warp * (size_t)warp_bytes;
    const uint32_t woff = (uint32_t)warp * (uint32_t)warp_bytes;  // this warp's slice, as an offset into the CTA's shared memory
    uint4* const sva = reinterpret_cast<uint4*>(wbase + L::off_sva);
    uint4* const svb = reinterpret_cast<uint4*>(wbase + L::off_svb);
    uint32_t* const hist = reinterpret_cast<uint32_t*>(wbase + L::off_hist);
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(wbase + L::off_mbar);
    uint16_t* const perm = reinterpret_cast<uint16_t*>(wbase + L::off_perm);
    unsigned char* const stage_a = wbase + L::off_stage;
    unsigned char* const stage_b = stage_a + s.stage_bytes + 16;
    const uint32_t off_sva = woff + (uint32_t)L::off_sva, off_svb = woff + (uint32_t)L::off_svb;
    const uint32_t off_stage_a = woff + (uint32_t)L::off_stage, off_stage_b = off_stage_a + (uint32_t)s.stage_bytes + 16u;

    if (lane == 0) mbar_init(mbar, 1);
    uint32_t mbar_phase = 0;
    __syncwarp();

    const long long n = s.n;
    const long long n_tiles = (n + TILE - 1) / TILE;
    const long long n_warps = (long long)gridDim.x * WPB;
    for (long long tile = (long long)blockIdx.x * WPB + warp; tile < n_tiles; tile += n_warps) {
        const long long tile0 = tile * TILE;
        for (int i = lane; i < NB; i += 32) hist[i] = 0;

        // ---------------- 1. load views, validity; route rows that do not fit --------------------
        unsigned active = 0;  // bit k: row k*32+lane goes through the sort
        uint32_t mn_off[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_end[2] = {0, 0};
        uint32_t mn_buf[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_buf[2] = {0, 0};
        uint32_t cnt[2] = {0, 0}, pad_bytes[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int i = k * 32 + lane;
            const long long row = tile0 + i;
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (row < n) {
                va = ld_view(s.a.views + row * s.a.stride);
                vb = ld_view(s.b.views + row * s.b.stride);
                // pull this warp's NEXT tile of views into L2 while the current one is processed
                const long long nxt = row + n_warps * TILE;
                if (nxt < n) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.a.views + nxt * s.a.stride));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.b.views + nxt * s.b.stride));
                }
                const bool valid = bit_valid(s.a.validity, s.a.vbit + row * s.a.stride) &&
                                   bit_valid(s.b.validity, s.b.vbit + row * s.b.stride);
                const uint32_t mx = va.x > vb.x ? va.x : vb.x;
                if (!valid) {
                    store_settled<MEASURE>(s, row, 0.0, 0);
                } else if ((va.x > 12u && !payload_resident(va, s.a)) || (vb.x > 12u && !payload_resident(vb, s.b))) {
                    atomicAdd(&s.ovf->ndefer, 1u);  // nothing is written for this row in this pass
                } else if (mx > (uint32_t)CAP) {
                    if (CAP == 32 && mx <= 64u) {
                        s.list64[atomicAdd(&s.ovf->n64, 1u)] = (unsigned int)row;
                    } else {
                        s.listlong[atomicAdd(&s.ovf->nlong, 1u)] = (unsigned int)row;
                        atomicMax(&s.ovf->max_bytes_a, va.x);
                        atomicMax(&s.ovf->max_bytes_b, vb.x);
                    }
                } else {
                    active |= 1u << k;
                    if (va.x > 12u) {
                        mn_off[0] = min(mn_off[0], va.w);
                        mx_end[0] = max(mx_end[0], va.w + va.x);
                        mn_buf[0] = min(mn_buf[0], va.z);
                        mx_buf[0] = max(mx_buf[0], va.z);
                        cnt[0]++;
                        pad_bytes[0] += (va.x + 3u) & ~3u;
                    }
                    if (vb.x > 12u) {
                        mn_off[1] = min(mn_off[1], vb.w);
                        mx_end[1] = max(mx_end[1], vb.w + vb.x);
                        mn_buf[1] = min(mn_buf[1], vb.z);
                        mx_buf[1] = max(mx_buf[1], vb.z);
                        cnt[1]++;
                        pad_bytes[1] += (vb.x + 3u) & ~3u;
                    }
                }
            }
            sva[i] = va;
            svb[i] = vb;
        }

        // ---------------- 2. stage the out-of-line payload -----------------------------------------
        int mode[2];          // 0 nothing out of line, 1 TMA bulk span, 2 cooperative gather copy
        uint32_t base16[2];   // TMA mode: 16-aligned start offset of the span in the data buffer
        uint32_t span[2], bufidx[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {  // one REDUX each: the span descriptors of the warp's tile
            const uint32_t a0 = __reduce_min_sync(0xFFFFFFFFu, mn_off[c]);
            const uint32_t a1 = __reduce_max_sync(0xFFFFFFFFu, mx_end[c]);
            const uint32_t a2 = __reduce_min_sync(0xFFFFFFFFu, mn_buf[c]);
            const uint32_t a3 = __reduce_max_sync(0xFFFFFFFFu, mx_buf[c]);
            const uint32_t a4 = __reduce_add_sync(0xFFFFFFFFu, cnt[c]);
            base16[c] = a0 & ~15u;
            span[c] = ((a1 + 15u) & ~15u) - base16[c];
            bufidx[c] = a2;
            mode[c] = a4 == 0 ? 0 : (a2 == a3 && span[c] <= (uint32_t)s.stage_bytes) ? 1 : 2;
        }
        const bool used_tma = mode[0] == 1 || mode[1] == 1;
        __syncwarp();  // every lane's views are in shared memory; the stage areas are free (previous tile done)
        if (used_tma && lane == 0) {
            // one arrival per phase: announce the bytes of both copies, then issue them
            mbar_expect_tx(mbar, (mode[0] == 1 ? span[0] : 0u) + (mode[1] == 1 ? span[1] : 0u));
            if (mode[0] == 1)
                tma_bulk_g2s(stage_a, reinterpret_cast<const unsigned char*>(s.a.bufs[bufidx[0]]) + base16[0], span[0], mbar);
            if (mode[1] == 1)
                tma_bulk_g2s(stage_b, reinterpret_cast<const unsigned char*>(s.b.bufs[bufidx[1]]) + base16[1], span[1], mbar);
        }
        // gather-copy fallback (views scattered over the data buffer, or a span larger than the stage area):
        // exclusive scan of the padded byte counts over the warp's rows, then every lane copies its rows
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (mode[c] != 2) continue;  // warp-uniform
            uint32_t incl = pad_bytes[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t pos = incl - pad_bytes[c];
            uint4* sv = c == 0 ? sva : svb;
            unsigned char* stage = c == 0 ? stage_a : stage_b;
            const DevCol& col = c == 0 ? s.a : s.b;
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int i = k * 32 + lane;
                if (!((active >> k) & 1u)) continue;
                const uint4 v = sv[i];
                if (v.x <= 12u) continue;
                const uint32_t padded = (v.x + 3u) & ~3u;
                if (pos + padded > (uint32_t)s.stage_bytes) {
                    // stage full: finish this row in the long kernels
                    s.listlong[atomicAdd(&s.ovf->nlong, 1u)] = (unsigned int)(tile0 + i);
                    atomicMax(&s.ovf->max_bytes_a, sva[i].x);
                    atomicMax(&s.ovf->max_bytes_b, svb[i].x);
                    active &= ~(1u << k);
                    continue;
                }
                const unsigned char* gp = reinterpret_cast<const unsigned char*>(col.bufs[v.z]) + v.w;
                const uint32_t* gw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(gp) & ~(uintptr_t)3);
                const int sh = (int)(reinterpret_cast<uintptr_t>(gp) & 3) * 8;
                uint32_t* dst = reinterpret_cast<uint32_t*>(stage + pos);
                uint32_t lo = __ldg(gw);
                const int nw = (int)(padded >> 2);
                for (int w = 0; w < nw; w++) {
                    // reads at most 7 bytes past the string: inside the padded device buffer
                    const uint32_t hi = __ldg(gw + w + 1);
                    dst[w] = __funnelshift_r(lo, hi, sh);
                    lo = hi;
                }
                sv[i].y = pos;
                pos += padded;
            }
        }
        // TMA mode: rewrite the views so that .y is the offset inside the stage area
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (mode[c] != 1) continue;
            uint4* sv = c == 0 ? sva : svb;
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int i = k * 32 + lane;
                if (((active >> k) & 1u) && sv[i].x > 12u) sv[i].y = sv[i].w - base16[c];
            }
        }
        if (used_tma) {
            mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
        }
        __syncwarp();

        // ---------------- 3. bucket: cost key per row, counting sort (descending) -----------------
        uint32_t key[RPT], rank[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int i = k * 32 + lane;
            const uint4 va = sva[i], vb = svb[i];
            const bool on = ((active >> k) & 1u) != 0u;
            // byte-equal pairs (strsim.rs:128,182,288,324), found by all lanes in lock step
            const bool equal = staged_equal_conv(smem, staged_str(va, i, off_sva, off_stage_a),
                                                 staged_str(vb, i, off_svb, off_stage_b), on);
            key[k] = 0;
            rank[k] = 0;
            if (!on) continue;
            // the loops of the row functions run over a (fused evaluation, Jaro, multisets); Levenshtein
            // streams the longer string
            uint32_t cost = va.x > vb.x ? va.x : vb.x;
            if (is_multi(MEASURE) || MEASURE != LEVENSHTEIN) cost = va.x;
            key[k] = equal ? 1u : 2u + cost;  // key 1: byte-equal pairs, the cheapest bucket of their own
            rank[k] = atomicAdd(&hist[key[k]], 1u);
        }
        __syncwarp();
        {
            // start[key] = number of rows with a larger key; lane l owns bins [l*CH, (l+1)*CH)
            constexpr int CH = (NB + 31) / 32;
            uint32_t local[CH];
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < CH; q++) {
                const int bin = NB - 1 - (lane * CH + q);  // descending
                local[q] = bin >= 1 ? hist[bin] : 0u;      // bin 0 = inactive rows
                sum += local[q];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t run = incl - sum;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < CH; q++) {
                const int bin = NB - 1 - (lane * CH + q);
                if (bin >= 1) hist[bin] = run;
                run += local[q];
            }
            if (lane == 31) hist[0] = incl;  // total active rows
        }
        __syncwarp();
        const int n_active = (int)hist[0];
        const int n_differ = (int)hist[1];  // rows before this position differ; the byte-equal ones (key 1) come last
#pragma unroll
        for (int k = 0; k < RPT; k++)
            if (key[k]) perm[hist[key[k]] + rank[k]] = (uint16_t)(k * 32 + lane);
        __syncwarp();

        // ---------------- 4. compute -----------------------------------------------------------------
#pragma unroll 1
        for (int k = 0; k < RPT; k++) {
            const int p = k * 32 + lane;
            if (p >= n_active) continue;
            const int i = perm[p];
            const long long row = tile0 + i;
            if (p >= n_differ) {  // whole warps of byte-equal pairs leave here
                store_settled<MEASURE>(s, row, 1.0, F_EQUAL);
                continue;
            }
            const uint4 va = sva[i], vb = svb[i];
            const StagedStr A = staged_str(va, i, off_sva, off_stage_a), B = staged_str(vb, i, off_svb, off_stage_b);
            if constexpr (is_multi(MEASURE)) {
                // b is tabled as bit planes, a is streamed -- both straight from the staged tile
                RowEmit emit{s, row, true};
                row_planes_multi<GROUPS, NBITS, M>(StagedSrc{smem, A.off, A.len}, StagedSrc{smem, B.off, B.len}, emit);
            } else {
                PairInts ints;
                const double v = row_planes<is_multi(MEASURE) ? 0 : MEASURE, NBITS, M>(StagedSrc{smem, A.off, A.len},
                                                                                      StagedSrc{smem, B.off, B.len}, ints);
                s.out[row] = v;
                if (s.dbg) store_dbg(s.dbg + row * 6, ints);
            }
        }
        __syncwarp();  // the slice is reused by the next tile
    }
}

}  // namespace strsim
