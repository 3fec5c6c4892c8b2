// short_kernel.cuh -- the short-string kernel: one pair per thread, strings of at most bits(M) bytes
// (M = u32, u64, or Wide<10> = 320 bits), one measure or a fused set of measures per launch.
//
// Persistent CTAs walk tiles of TILE = TPB*RPT consecutive rows (gather launches: consecutive entries of a
// row list).  Tiles are handed out through a counter (Overflow::tile_ctr), not by a fixed stride: a CTA starts
// on tile blockIdx.x and asks for its next tile while it works on the current one.  Per tile:
//   1. load   : coalesced 16-byte loads of both columns' Arrow views (+ validity bits) -> shared memory;
//               null rows are settled, rows that cannot be served here go to the overflow lists
//   2. stage  : the tile's out-of-line payload (byte length > 12) is one contiguous span of the data
//               buffer when the column was built sequentially, so ONE TMA bulk copy
//               (cp.async.bulk.shared.global, SASS UBLKCP) per column brings it into shared memory,
//               completion on an mbarrier; columns whose views are scattered (gather/filter results,
//               dictionary columns) fall back to a cooperative word copy into the same stage area
//   3. bucket : byte-equal pairs are found by all lanes in lock step; every other row gets a cost key (the
//               length of the streamed string / its character count); a shared-memory counting sort yields
//               a permutation so that the 32 lanes of a warp work on pairs of (nearly) equal cost -- the
//               length-bucketing pre-pass, done per tile on chip, no extra HBM traffic, output order untouched
//   4. compute: each thread takes one sorted pair and runs one of the register paths on it --
//               REG  (ASCII-only columns): bit planes of the tabled string, the streamed one read byte by byte,
//                    both straight from the staged tile (row_ascii_reg.cuh);
//               ULAT (general columns, first launch): the pairs without a character above U+00FF, transcoded to
//                    one byte per character into the thread's slab, then the plane path with 8 planes; the
//                    other pairs are listed for
//               UREG (general columns, second launch, gather mode): the tabled string decoded into registers,
//                    position masks by compares (row_unicode_reg.cuh)
//               -- and writes the f64 result(s).
// Rows that do not fit (longer than bits(M) bytes, or the stage area is full) are appended to an overflow list
// and finished by the wider instantiations of this kernel in gather mode (64-bit and ten-word masks, ASCII
// columns), by direct_kernel.cuh (33..64-byte rows of general columns) or by the warp-per-pair kernels
// (long_lev_kernel.cuh, long_pair_kernel.cuh).  Null rows (either input null, README.md:69-70) get 0.0 and
// are never dereferenced beyond their view.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "row_ascii_reg.cuh"
#include "row_short.cuh"
#include "row_unicode_reg.cuh"

namespace strsim {

struct DevCol {
    const uint4* views;       // row 0 of the segment
    const uint8_t* validity;  // or nullptr
    long long vbit;           // bit index of row 0 inside validity
    const unsigned long long* bufs;  // device table of data-buffer base addresses
    int stride;               // 1, or 0 for a broadcast scalar (strsim.rs:61-66)
    // residency frontier of the data buffers while a host call is still uploading them (host.cu:
    // progressive upload): buffers with index < res_buf are complete, buffer res_buf holds its first
    // res_off bytes, later buffers nothing yet.  0xFFFFFFFF = everything is resident.
    unsigned int res_buf, res_off;
    // lower frontier (row shards of a sharded call upload only the stretch of the data their rows
    // reference): buffers with index < lo_buf and the first lo_off bytes of buffer lo_buf have no storage
    // on this device.  (0, 0) = everything from the start.
    unsigned int lo_buf, lo_off;
};

// is the out-of-line payload of view v (length > 12) on the device (yet)?
__device__ __forceinline__ bool payload_resident(const uint4& v, const DevCol& c) {
    return (v.z < c.res_buf || (v.z == c.res_buf && v.w + v.x <= c.res_off)) &&
           (v.z > c.lo_buf || (v.z == c.lo_buf && v.w >= c.lo_off));
}

struct Overflow {
    unsigned int n64;    // rows for the 64-bit short kernel
    unsigned int nlong;  // rows for the long kernel
    unsigned int max_bytes_a, max_bytes_b;  // over the long rows
    unsigned int ndefer;  // rows whose payload had not been uploaded yet: the host recomputes the slice
    unsigned int nwide;   // ULAT launch: pairs with a character above U+00FF, left to the UREG launch
    // tile scheduling (short_kernel): tiles handed out beyond the first one of every CTA / CTAs that are done,
    // one pair of counters per launch kind (SegArgs::ctr); the last CTA of a launch leaves both at zero
    unsigned int tile_ctr[3], done_ctr[3];
};

struct SegArgs {
    DevCol a, b;
    long long n;   // rows of the segment, or entries of `list` in gather mode
    double* out;   // row 0 of the segment
    int* dbg;      // row 0 of the segment (STRSIM_DBG_INTS per row) or nullptr
    // fused evaluation (MEASURE >= MULTI_BASE): one output per measure, row 0 of the segment; nullptr =
    // this measure is not wanted (its group may still be computed for the group's other member)
    double* outs[5];
    int* dbgs[5];
    const unsigned int* list;  // gather mode: segment-relative row numbers
    const unsigned int* list_count;  // gather mode: device-resident number of entries
    Overflow* ovf;
    unsigned int* list64;
    unsigned int* listlong;
    unsigned int* listwide;  // ULAT launch: the pairs it leaves to the register-compare launch (count: ovf->nwide)
    int stage_bytes;  // capacity of each column's stage area (multiple of 16)
    // general columns are served by two launches (ULAT, then UREG): the second one skips the Latin-1
    // pairs and leaves null rows and the overflow lists alone
    int skip_latin;
    // which of ovf's tile counters this launch uses: launches that may be in flight behind each other over the
    // same overflow record take different ones (0: the launch over every row, 1: the second launch over a
    // general column, 2: the 64-bit follow-up)
    int ctr;
};

// ---- small device helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_view(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ bool bit_valid(const uint8_t* validity, long long bit) {
    return validity == nullptr || ((__ldg(validity + (bit >> 3)) >> (bit & 7)) & 1);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// TMA bulk copy global -> shared (1-D, no tensor map): dst, src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// T = entries of the per-thread position-mask table.  128 is exact for any ASCII string; when the
// column statistics (host.cu: column_stats) prove that every byte of both columns lies in one aligned
// block of 32 or 64 code points (e.g. lower-case names: 0x60..0x7F), c & (T-1) is injective on the
// alphabet and the smaller table quadruples / doubles the resident warps per SM.
template <class M, int TPB, int T>
struct DevStore {
    using mask_type = M;
    static constexpr bool CPS_ALIAS_TABLE = true;
    M* tab_;          // &table[tid]
    uint32_t* wa_;    // &slab_a[tid]
    uint32_t* wb_;    // &slab_b[tid]
    static constexpr uint32_t CMASK = (T >= 128 ? 128u : (uint32_t)T) - 1u;  // ASCII byte -> table entry
    __device__ __forceinline__ M& tab(uint32_t c) { return tab_[(c & CMASK) * TPB]; }
    __device__ __forceinline__ const M& tab(uint32_t c) const { return tab_[(c & CMASK) * TPB]; }
    // Unicode path: 2*bits(M) hash slots alias the thread's OWN table entries -- the keys (32 bit)
    // sit in entries [0, 2*bits(M)/R), R keys per entry, the masks in the 2*bits(M) entries after them
    static constexpr int R = (int)(sizeof(M) / 4);
    static constexpr int SLOTS = 2 * 8 * (int)sizeof(M);
    static constexpr int HASH_ENTRIES = SLOTS / R + SLOTS;  // table entries the hash needs
    __device__ __forceinline__ uint32_t& hkey(int s) {
        return reinterpret_cast<uint32_t*>(&tab_[(s / R) * TPB])[s % R];
    }
    __device__ __forceinline__ const uint32_t& hkey(int s) const {
        return reinterpret_cast<const uint32_t*>(&tab_[(s / R) * TPB])[s % R];
    }
    __device__ __forceinline__ M& hmask(int s) { return tab_[(SLOTS / R + s) * TPB]; }
    __device__ __forceinline__ const M& hmask(int s) const { return tab_[(SLOTS / R + s) * TPB]; }
    __device__ __forceinline__ uint32_t wa(int k) const { return wa_[k * TPB]; }
    __device__ __forceinline__ uint32_t wb(int k) const { return wb_[k * TPB]; }
};

// REG: plane path for ASCII-only columns (row_ascii_reg.cuh), strings read from the staged tile -- no table,
// no slabs in shared memory
// UREG: register-compare path for any script (row_unicode_reg.cuh) -- no table, slabs only
template <class M, int TPB, int RPT, int T, bool REG = false, bool UREG = false>
struct ShortLayout {
    static constexpr int CAP = (int)sizeof(M) * 8;
    static constexpr int WORDS = CAP / 4;
    static constexpr int TILE = TPB * RPT;
    static constexpr int NB = 2 * CAP + 3;  // keys 0 .. 2*CAP+2
    static constexpr int NWARP = TPB / 32;
    static constexpr size_t off_sva = 0;
    static constexpr size_t off_svb = off_sva + sizeof(uint4) * TILE;
    // 16 spare bytes behind the views: reading an inline string by whole words looks one word past the
    // last view, which must not be a word another thread writes (its slab) -- racecheck flagged exactly that
    static constexpr size_t off_slab_a = off_svb + sizeof(uint4) * TILE + 16;
    static constexpr size_t off_slab_b = off_slab_a + (REG ? 0 : 4 * WORDS * TPB);
    static constexpr size_t off_hist = off_slab_b + (REG ? 0 : 4 * WORDS * TPB);
    static constexpr size_t off_red = off_hist + 4 * ((NB + 3) & ~3);
    static constexpr size_t off_mbar = off_red + 4 * 16 * NWARP;
    static constexpr size_t off_perm = off_mbar + 16;
    static constexpr size_t off_stage = (off_perm + 2 * TILE + 15) & ~(size_t)15;
    static size_t bytes(int stage_bytes) { return off_stage + 2 * ((size_t)stage_bytes + 16); }
};

__device__ __forceinline__ uint32_t byte_mask(int nbytes) {  // low nbytes bytes set, nbytes in 0..4
    return nbytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbytes)) - 1u);
}

// Loads the bytes of one string (view v; out-of-line bytes already staged at stage + v.y) into the
// thread's slab as zero-masked little-endian words.  Returns OR of all words (for the ASCII test).
template <int WORDS, int TPB>
__device__ __forceinline__ uint32_t load_string(const uint4& v, const unsigned char* stage,
                                                uint32_t* slab /* &slab[tid] */) {
    const int len = (int)v.x;
    uint32_t acc = 0;
    if (len <= 12) {
        uint32_t w0 = v.y & byte_mask(len);
        uint32_t w1 = v.z & byte_mask(len - 4 < 0 ? 0 : len - 4);
        uint32_t w2 = v.w & byte_mask(len - 8 < 0 ? 0 : len - 8);
        slab[0] = w0;
        slab[TPB] = w1;
        slab[2 * TPB] = w2;
        acc = w0 | w1 | w2;
    } else {
        const uint32_t soff = v.y;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(stage) + (soff >> 2);
        const int sh = (int)(soff & 3) * 8;
        const int full = len >> 2;  // whole words
        uint32_t lo = src[0];
        int w = 0;
        for (; w < full; w++) {
            const uint32_t hi = src[w + 1];
            const uint32_t word = __funnelshift_r(lo, hi, sh);
            lo = hi;
            slab[w * TPB] = word;
            acc |= word;
        }
        if (len & 3) {
            const uint32_t word = __funnelshift_r(lo, src[w + 1], sh) & byte_mask(len & 3);
            slab[w * TPB] = word;
            acc |= word;
        }
    }
    return acc;
}

// number of characters (bytes that are not UTF-8 continuation bytes) of a staged string; `wide` collects
// the bytes >= 0xC4 (characters above U+00FF, see row_ascii_reg.cuh: wide_bytes)
__device__ __forceinline__ int staged_char_count(const uint4& v, const unsigned char* stage, uint32_t& wide) {
    const int len = (int)v.x;
    int cont = 0;
    if (len <= 12) {
        const uint32_t w0 = v.y & byte_mask(len), w1 = v.z & byte_mask(len - 4 < 0 ? 0 : len - 4),
                       w2 = v.w & byte_mask(len - 8 < 0 ? 0 : len - 8);
        cont = __popc(((w0 >> 7) & ~(w0 >> 6)) & 0x01010101u) + __popc(((w1 >> 7) & ~(w1 >> 6)) & 0x01010101u) +
               __popc(((w2 >> 7) & ~(w2 >> 6)) & 0x01010101u);
        wide |= wide_bytes(w0) | wide_bytes(w1) | wide_bytes(w2);
    } else {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(stage) + (v.y >> 2);
        const int head = (int)(v.y & 3u);               // bytes of the first word before the string
        const int nwords = (head + len + 3) >> 2;
        for (int w = 0; w < nwords; w++) {
            uint32_t x = p[w];
            if (w == 0) x &= ~byte_mask(head);
            if (w == nwords - 1) x &= byte_mask(head + len - 4 * w);
            cont += __popc(((x >> 7) & ~(x >> 6)) & 0x01010101u);
            wide |= wide_bytes(x);
        }
    }
    return len - cont;
}

// byte p of a string that sits in shared memory (inline in its staged view, or in the stage area)
struct SmemByteAt {
    uint32_t base;  // shared-space address of the first byte
    __device__ __forceinline__ uint32_t operator()(int p) const {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)p));
        return v;
    }
};

// A string of the staged tile: both kinds of string sit in shared memory after step 2 -- inline in
// their staged view (<= 12 bytes, bytes 4..15 of the view) or in the stage area (view.y = offset) -- so
// one code path reads either: `off` is the byte offset of the first byte from the start of the CTA's
// shared memory.  Reading whole aligned words overruns the string by at most 7 bytes: the next view,
// or the 16-byte pad of the stage area.
struct StagedStr {
    uint32_t off;
    int len;
};
__device__ __forceinline__ StagedStr staged_str(const uint4& v, int i, uint32_t off_views, uint32_t off_stage) {
    StagedStr s;
    s.len = (int)v.x;
    s.off = v.x <= 12u ? off_views + 16u * (uint32_t)i + 4u : off_stage + v.y;
    return s;
}

// Exact byte equality of two staged strings, for the sort key.  EVERY lane of the warp walks the loop
// (pairs of unequal length with a trip count of zero), so the warp pays max(words) iterations of ten
// instructions -- the former version (prefix test, then separate inline / out-of-line compares behind
// branches) ran 160 warp instructions per 32 rows with four or five lanes active.
__device__ __forceinline__ bool staged_equal_conv(const unsigned char* smem, const StagedStr& A, const StagedStr& B,
                                                  bool candidate) {
    const bool eq_len = candidate && A.len == B.len;
    // rows that are not candidates (null, routed to an overflow list, beyond n) hold no stage offset in
    // their view: they read word 0 of the shared memory instead
    const uint32_t oa = candidate ? A.off : 0u, ob = candidate ? B.off : 0u;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(smem + (oa & ~3u));
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(smem + (ob & ~3u));
    const int sa = (int)(oa & 3u) * 8, sb = (int)(ob & 3u) * 8;
    const int full = eq_len ? A.len >> 2 : 0;
    uint32_t la = pa[0], lb = pb[0], diff = 0;
    int w = 0;
#pragma unroll 1
    for (; w + 2 <= full; w += 2) {  // two words per trip: half the loop control, no register moves
        const uint32_t ma = pa[w + 1], mb = pb[w + 1], ha = pa[w + 2], hb = pb[w + 2];
        diff |= __funnelshift_r(la, ma, sa) ^ __funnelshift_r(lb, mb, sb);
        diff |= __funnelshift_r(ma, ha, sa) ^ __funnelshift_r(mb, hb, sb);
        la = ha;
        lb = hb;
    }
    if (w < full) {
        const uint32_t ha = pa[w + 1], hb = pb[w + 1];
        diff |= __funnelshift_r(la, ha, sa) ^ __funnelshift_r(lb, hb, sb);
        la = ha;
        lb = hb;
    }
    if (eq_len && (A.len & 3))
        diff |= (__funnelshift_r(la, pa[full + 1], sa) ^ __funnelshift_r(lb, pb[full + 1], sb)) & byte_mask(A.len & 3);
    return eq_len && diff == 0u;
}

// streams the bytes of a staged string through f, straight from shared memory (no register copy):
// whole words first, then the 1..3 bytes of the tail one at a time
struct EachByteStaged {
    const unsigned char* smem;
    uint32_t off;
    template <class F>
    __device__ __forceinline__ void operator()(int n, F& f) const {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(smem + (off & ~3u));
        const int sh = (int)(off & 3u) * 8;
        uint32_t lo = p[0];
#pragma unroll 1
        for (; n >= 4; n -= 4) {
            p++;
            const uint32_t hi = p[0];
            const uint32_t word = __funnelshift_r(lo, hi, sh);
            lo = hi;
            f(word & 0xFFu);
            f(byte_of(word, 1));
            f(byte_of(word, 2));
            f(word >> 24);
        }
        if (n > 0) {
            uint32_t word = __funnelshift_r(lo, p[1], sh);
#pragma unroll 1
            for (; n > 0; n--) {
                f(word & 0xFFu);
                word >>= 8;
            }
        }
    }
};

// bit planes of a staged string of m <= 32 characters, word by word from shared memory.  The bytes after
// the string need no masking: no consumer of a position mask looks above the string (PlaneTab).
template <int NBITS, class M>
__device__ __forceinline__ void build_planes_staged(const unsigned char* smem, const StagedStr& S,
                                                    PlaneTab<NBITS, M>& tab) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(smem + (S.off & ~3u));
    const int sh = (int)(S.off & 3u) * 8;
    const int m = S.len;
#pragma unroll
    for (int k = 0; k < NBITS; k++) tab.B[k] = M(0);
    uint32_t lo = p[0];
#pragma unroll
    for (int w = 0; w < (int)sizeof(M) * 2; w++) {
        if (4 * w >= m) break;
        const uint32_t hi = p[w + 1];
        const uint32_t word = __funnelshift_r(lo, hi, sh);
        lo = hi;
        planes_add_word<NBITS, M>(tab, word, w);
    }
}

// first word (zero-masked to the string's length) of a staged string
__device__ __forceinline__ uint32_t staged_first_word(const unsigned char* smem, const StagedStr& S) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(smem + (S.off & ~3u));
    return __funnelshift_r(p[0], p[1], (int)(S.off & 3u) * 8) & byte_mask(S.len);
}

// a string of the staged tile as a source of the plane path (row_ascii_reg.cuh: row_planes)
struct StagedSrc {
    const unsigned char* smem;
    uint32_t off;
    int len;
    typedef SmemByteAt ByteAt;
    __device__ __forceinline__ SmemByteAt byte_at() const { return SmemByteAt{smem_u32(smem) + off}; }
    __device__ __forceinline__ uint32_t first_word() const { return staged_first_word(smem, StagedStr{off, len}); }
    template <int NBITS, class M>
    __device__ __forceinline__ void planes(PlaneTab<NBITS, M>& tab) const {
        build_planes_staged<NBITS, M>(smem, StagedStr{off, len}, tab);
    }
    template <class F>
    __device__ __forceinline__ void each(int n, F& f) const {
        EachByteStaged e{smem, off};
        e(n, f);
    }
};

// ULAT launch: appends the rows a thread flagged (bit k of `flags` = its k-th row of the tile) to the list the
// register-compare launch gathers, with one atomic per warp.  EVERY lane of the warp calls it from converged
// code (the ballots name the full mask).  It replaces the opportunistic form -- __activemask() inside the
// divergent branch, then __shfl_sync over that mask -- in which nothing ties the point where the compiler reads
// the active mask to the lanes that take the branch.
template <int RPT, typename RowOf>
__device__ __forceinline__ void list_wide_rows(const SegArgs& s, uint32_t flags, int lane, RowOf row_of) {
    if (!__any_sync(0xFFFFFFFFu, flags != 0u)) return;
    unsigned total = 0;
#pragma unroll
    for (int k = 0; k < RPT; k++) total += (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, (flags >> k) & 1u));
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&s.ovf->nwide, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
    for (int k = 0; k < RPT; k++) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, (flags >> k) & 1u);
        if ((flags >> k) & 1u) s.listwide[base + (unsigned)__popc(b & ((1u << lane) - 1u))] = row_of(k);
        base += (unsigned)__popc(b);
    }
}

__device__ __forceinline__ void store_dbg(int* d, const PairInts& o) {
    d[0] = o.flag;
    d[1] = o.la;
    d[2] = o.lb;
    d[3] = o.x0;
    d[4] = o.x1;
    d[5] = o.x2;
}

// result of a row settled without looking at its characters (null row: 0.0 / zero record; byte-equal
// row: 1.0 / F_EQUAL), written to the single output or to every wanted output of a fused launch
template <int MEASURE>
__device__ __forceinline__ void store_settled(const SegArgs& s, long long row, double v, int flag) {
    PairInts o;
    o.flag = flag;
    o.la = o.lb = o.x0 = o.x1 = o.x2 = 0;
    if (is_multi(MEASURE)) {
#pragma unroll
        for (int m = 0; m < 5; m++) {
            if (!s.outs[m]) continue;
            s.outs[m][row] = v;
            if (s.dbgs[m]) store_dbg(s.dbgs[m] + row * 6, o);
        }
    } else {
        s.out[row] = v;
        if (s.dbg) store_dbg(s.dbg + row * 6, o);
    }
}

// sink of the fused row functions: keeps the measures whose output pointer is set
struct RowEmit {
    const SegArgs& s;
    long long row;
    bool has;
    __device__ __forceinline__ void operator()(int measure, double v, const PairInts& o) const {
        double* out = s.outs[measure];
        if (has && out) {
            out[row] = v;
            int* d = s.dbgs[measure];
            if (d) store_dbg(d + row * 6, o);
        }
    }
};

// the thread's slab of shared memory ([word][thread] layout) as transcode_latin1's Slab / as bytes
template <int TPB>
struct SlabWords {
    uint32_t* w;  // &slab[tid]
    __device__ __forceinline__ uint32_t rd(int i) const { return w[i * TPB]; }
    __device__ __forceinline__ void wr(int i, uint32_t v) { w[i * TPB] = v; }
};
template <int TPB>
struct SlabByteAt {
    uint32_t base;  // shared-space address of &slab[tid]
    __device__ __forceinline__ uint32_t operator()(int p) const {
        uint32_t v;
        // "memory": the slab was rewritten by transcode_latin1 just before
        asm volatile("ld.shared.u8 %0, [%1];"
                     : "=r"(v)
                     : "r"(base + (uint32_t)(p >> 2) * (uint32_t)(4 * TPB) + (uint32_t)(p & 3))
                     : "memory");
        return v;
    }
};

// A staged string (any byte offset inside the tile) as the word source of transcode_latin1_stream: aligned
// shared-memory words, funnel-shifted to the string's start, zero beyond its length.
struct StagedWords {
    const uint32_t* p;  // aligned word that holds the first byte
    int sh;             // bit offset of the first byte inside it
    int len;
    __device__ __forceinline__ uint32_t operator()(int w) const {
        const uint32_t word = __funnelshift_r(p[w], p[w + 1], sh);
        const int left = len - 4 * w;
        return left >= 4 ? word : word & byte_mask(left);
    }
};
template <int TPB>
struct SlabWriter {
    uint32_t* w;  // &slab[tid]
    __device__ __forceinline__ void operator()(int i, uint32_t v) { w[i * TPB] = v; }
};

struct WarpMaxDev {  // maximum over the 32 lanes of the warp (all lanes must call it)
    __device__ __forceinline__ int operator()(int v) const { return __reduce_max_sync(0xFFFFFFFFu, v); }
};

// ULAT (with UREG): the Latin-1 half of a general column -- this instantiation computes only the pairs
// without a character above U+00FF, by the 8-plane path; the plain UREG launch that follows (with
// SegArgs::skip_latin) takes the others by the register-compare path.  One kernel holding both paths
// measured SLOWER than the compare path alone: 121 KB of SASS, a third of the stall samples
// instruction-cache misses, another third barrier waits of warps with unequal work.
template <class M, int MEASURE, int TPB, int RPT, bool GATHER, int T, bool ASCII_ONLY, bool REG = false,
          bool UREG = false, bool ULAT = false>
__global__ void __launch_bounds__(TPB) short_kernel(const SegArgs s) {
    static_assert(!ULAT || UREG, "the Latin-1 instantiation uses the general kernel's slabs");
    static_assert(REG || UREG, "register paths only (the table / hash path lives in direct_kernel.cuh)");
    static_assert(!UREG || (!ASCII_ONLY && !REG && sizeof(M) == 4), "register-compare path: general u32 kernel");
    static_assert(!REG || (ASCII_ONLY && (T == 32 || T == 64 || T == 128)),
                  "the plane path serves ASCII-only columns (strings of at most 32 bytes for M = u32, 64 for u64)");
    static_assert(!is_multi(MEASURE) || REG || UREG, "fused evaluation: register paths only");
    constexpr int GROUPS = is_multi(MEASURE) ? MEASURE - MULTI_BASE : 0;
    constexpr int NBITS = T == 32 ? 5 : T == 64 ? 6 : 7;
    using L = ShortLayout<M, TPB, RPT, T, REG, UREG>;
    constexpr int CAP = L::CAP;
    constexpr int WORDS = L::WORDS;
    constexpr int TILE = L::TILE;
    constexpr int NB = L::NB;
    constexpr int NWARP = L::NWARP;

    extern __shared__ __align__(16) unsigned char smem[];
    uint4* sva = reinterpret_cast<uint4*>(smem + L::off_sva);
    uint4* svb = reinterpret_cast<uint4*>(smem + L::off_svb);
    uint32_t* slab_a = reinterpret_cast<uint32_t*>(smem + L::off_slab_a);
    uint32_t* slab_b = reinterpret_cast<uint32_t*>(smem + L::off_slab_b);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L::off_hist);
    uint32_t* red = reinterpret_cast<uint32_t*>(smem + L::off_red);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L::off_mbar);
    uint16_t* perm = reinterpret_cast<uint16_t*>(smem + L::off_perm);
    unsigned char* stage_a = smem + L::off_stage;
    unsigned char* stage_b = stage_a + s.stage_bytes + 16;
    const uint32_t off_stage_a = (uint32_t)L::off_stage, off_stage_b = off_stage_a + (uint32_t)s.stage_bytes + 16u;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const long long n = GATHER ? (long long)*s.list_count : s.n;
    const long long n_tiles = (n + TILE - 1) / TILE;

    if (tid == 0) mbar_init(mbar, 1);
    uint32_t mbar_phase = 0;
    DevStore<M, TPB, T> store;
    store.tab_ = nullptr;  // (the register paths keep no table; the slabs serve ULAT / UREG)
    store.wa_ = slab_a + tid;
    store.wb_ = slab_b + tid;
    __syncthreads();

    // Gather launches (overflow lists, the wide pairs of a general column) reach a row through its list entry:
    // two dependent trips to memory per row, and the profile of the wide-pair launch of C3 showed the CTAs
    // waiting on exactly that (40 % of the issue slots used, long-scoreboard and barrier stalls on top).  The
    // entries of this CTA's NEXT tile are therefore fetched into registers while the current tile is loaded,
    // and the views they name are pulled into L2 at the start of the compute phase.
    unsigned int next_row[RPT];
    if (GATHER) {
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const long long idx = (long long)blockIdx.x * TILE + k * TPB + tid;
            next_row[k] = idx < n ? s.list[idx] : 0u;
        }
    }
    // Tiles are handed out through a counter: with a fixed share of tiles per CTA, the CTAs of an SM finished at
    // different times (a tile's cost follows its rows) and the SM ran under-occupied until its last CTA was done
    // -- C2 fused 0.622 -> 0.565 ms.  A CTA starts on tile blockIdx.x; thread 0 asks for the next tile at the
    // start of the current one (the round trip to L2 hides behind the loads of the views) and the answer crosses
    // the CTA through shared memory behind the barrier that opens the stage phase.
    uint32_t* next_slot = reinterpret_cast<uint32_t*>(smem + L::off_mbar + 8);
    // both columns complete on the device (no upload in flight, no row shard): no view needs the residency test
    const bool all_resident = s.a.res_buf == 0xFFFFFFFFu && s.b.res_buf == 0xFFFFFFFFu &&
                              (s.a.lo_buf | s.a.lo_off | s.b.lo_buf | s.b.lo_off) == 0u;
    long long tile_after = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile = tile_after) {
        const long long tile0 = tile * TILE;
        for (int i = tid; i < NB; i += TPB) hist[i] = 0;
        uint32_t ticket = 0;
        if (tid == 0) ticket = atomicAdd(&s.ovf->tile_ctr[s.ctr], 1u);

        // ---------------- 1. load views, validity; route rows that do not fit --------------------
        unsigned active = 0;  // bit k: row k*TPB+tid goes through the sort
        uint32_t mn_off[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_end[2] = {0, 0};
        uint32_t mn_buf[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, mx_buf[2] = {0, 0};
        uint32_t cnt[2] = {0, 0}, pad_bytes[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int i = k * TPB + tid;
            const long long idx = tile0 + i;
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (idx < n) {
                const long long row = GATHER ? (long long)next_row[k] : idx;
                va = ld_view(s.a.views + row * s.a.stride);
                vb = ld_view(s.b.views + row * s.b.stride);
                const bool valid = bit_valid(s.a.validity, s.a.vbit + row * s.a.stride) &&
                                   bit_valid(s.b.validity, s.b.vbit + row * s.b.stride);
                const uint32_t mx = va.x > vb.x ? va.x : vb.x;
                const bool second = UREG && !ULAT && s.skip_latin;  // the first launch settled these rows
                if (!valid) {
                    if (!second) store_settled<MEASURE>(s, row, 0.0, 0);
                } else if (!all_resident &&
                           ((va.x > 12u && !payload_resident(va, s.a)) || (vb.x > 12u && !payload_resident(vb, s.b)))) {
                    if (!second) atomicAdd(&s.ovf->ndefer, 1u);  // nothing is written for this row in this pass
                } else if (mx > (uint32_t)CAP) {
                    if (second) {
                        // already on the lists
                    } else
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
        // block reduction of the span descriptors (warp shuffle, then warp 0 over the partials)
#pragma unroll
        for (int c = 0; c < 2; c++) {  // one REDUX each
            mn_off[c] = __reduce_min_sync(0xFFFFFFFFu, mn_off[c]);
            mx_end[c] = __reduce_max_sync(0xFFFFFFFFu, mx_end[c]);
            mn_buf[c] = __reduce_min_sync(0xFFFFFFFFu, mn_buf[c]);
            mx_buf[c] = __reduce_max_sync(0xFFFFFFFFu, mx_buf[c]);
            cnt[c] = __reduce_add_sync(0xFFFFFFFFu, cnt[c]);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                red[warp * 16 + c * 5 + 0] = mn_off[c];
                red[warp * 16 + c * 5 + 1] = mx_end[c];
                red[warp * 16 + c * 5 + 2] = mn_buf[c];
                red[warp * 16 + c * 5 + 3] = mx_buf[c];
                red[warp * 16 + c * 5 + 4] = cnt[c];
            }
        }
        if (tid == 0) *next_slot = gridDim.x + ticket;
        __syncthreads();
        int mode[2];          // 0 nothing out of line, 1 TMA bulk span, 2 cooperative gather copy
        uint32_t base16[2];   // TMA mode: 16-aligned start offset of the span in the data buffer
        uint32_t span[2], bufidx[2];
        // every warp folds the NWARP partials again, lane l taking warp l % NWARP's (one load and one REDUX per
        // field; each thread looping over all of them was 160 instructions per thread and tile)
        const uint32_t* my_red = red + (lane & (NWARP - 1)) * 16;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint32_t a0 = __reduce_min_sync(0xFFFFFFFFu, my_red[c * 5 + 0]);
            const uint32_t a1 = __reduce_max_sync(0xFFFFFFFFu, my_red[c * 5 + 1]);
            const uint32_t a2 = __reduce_min_sync(0xFFFFFFFFu, my_red[c * 5 + 2]);
            const uint32_t a3 = __reduce_max_sync(0xFFFFFFFFu, my_red[c * 5 + 3]);
            const uint32_t a4 = __reduce_max_sync(0xFFFFFFFFu, my_red[c * 5 + 4]);  // only "any at all" matters
            base16[c] = a0 & ~15u;
            span[c] = ((a1 + 15u) & ~15u) - base16[c];
            bufidx[c] = a2;
            mode[c] = a4 == 0 ? 0 : (a2 == a3 && span[c] <= (uint32_t)s.stage_bytes) ? 1 : 2;
        }
        const bool used_tma = mode[0] == 1 || mode[1] == 1;
        if (used_tma && tid == 0) {
            // one arrival per phase: announce the bytes of both copies, then issue them
            mbar_expect_tx(mbar, (mode[0] == 1 ? span[0] : 0u) + (mode[1] == 1 ? span[1] : 0u));
            if (mode[0] == 1)
                tma_bulk_g2s(stage_a,
                             reinterpret_cast<const unsigned char*>(s.a.bufs[bufidx[0]]) + base16[0],
                             span[0], mbar);
            if (mode[1] == 1)
                tma_bulk_g2s(stage_b,
                             reinterpret_cast<const unsigned char*>(s.b.bufs[bufidx[1]]) + base16[1],
                             span[1], mbar);
        }

        // gather-copy fallback: exclusive scan of padded byte counts, then each thread copies its
        // rows' bytes global -> stage (aligned words, funnel shift on the source)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (mode[c] != 2) continue;  // CTA-uniform
            __syncthreads();             // red[] reuse
            uint32_t incl = pad_bytes[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) red[warp] = incl;
            __syncthreads();
            uint32_t warp_base = 0;
            for (int w = 0; w < warp; w++) warp_base += red[w];
            uint32_t pos = warp_base + incl - pad_bytes[c];
            uint4* sv = c == 0 ? sva : svb;
            unsigned char* stage = c == 0 ? stage_a : stage_b;
            const DevCol& col = c == 0 ? s.a : s.b;
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int i = k * TPB + tid;
                if (!((active >> k) & 1u)) continue;
                uint4 v = sv[i];
                if (v.x <= 12u) continue;
                const uint32_t padded = (v.x + 3u) & ~3u;
                if (pos + padded > (uint32_t)s.stage_bytes) {
                    // stage full: finish this row in the long kernel.  (Also in the second launch over a general
                    // column: its tiles hold wide pairs only, which can be longer than the column's mean -- the
                    // first launch staged this row and listed it as wide, so it is on no other list yet.)
                    const long long idx = tile0 + i;
                    const long long row = GATHER ? (long long)s.list[idx] : idx;
                    s.listlong[atomicAdd(&s.ovf->nlong, 1u)] = (unsigned int)row;
                    atomicMax(&s.ovf->max_bytes_a, sva[i].x);
                    atomicMax(&s.ovf->max_bytes_b, svb[i].x);
                    active &= ~(1u << k);
                    continue;
                }
                const unsigned char* g =
                    reinterpret_cast<const unsigned char*>(col.bufs[v.z]) + v.w;
                const uint32_t* gw = reinterpret_cast<const uint32_t*>(
                    reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)3);
                const int sh = (int)(reinterpret_cast<uintptr_t>(g) & 3) * 8;
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
                const int i = k * TPB + tid;
                if (((active >> k) & 1u) && sv[i].x > 12u) sv[i].y = sv[i].w - base16[c];
            }
        }
        // the CTA's next tile is known now (written before the barrier that opened this phase): pull its views
        // into L2 while this tile is processed; gather launches fetch its list entries (the views they name are
        // prefetched in step 4)
        tile_after = (long long)*next_slot;
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const long long nxt = tile_after * TILE + k * TPB + tid;
            if (nxt < n) {
                if (GATHER) {
                    next_row[k] = s.list[nxt];
                } else {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.a.views + nxt * s.a.stride));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.b.views + nxt * s.b.stride));
                }
            }
        }
        if (used_tma) {
            mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
        }
        __syncthreads();

        // ---------------- 3. bucket: cost key per row, counting sort (descending) -----------------
        uint32_t key[RPT], rank[RPT];
        bool row_equal[RPT];
        if constexpr (REG || ULAT || is_multi(MEASURE)) {
            // byte-equal pairs (strsim.rs:128,182,288,324), found by all lanes in lock step
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int i = k * TPB + tid;
                const uint4 va = sva[i], vb = svb[i];
                row_equal[k] = staged_equal_conv(smem, staged_str(va, i, (uint32_t)L::off_sva, off_stage_a),
                                                 staged_str(vb, i, (uint32_t)L::off_svb, off_stage_b),
                                                 ((active >> k) & 1u) != 0u);
            }
        }
        uint32_t listed = 0;  // ULAT: bit k = this thread's k-th row goes to the register-compare launch
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            const int i = k * TPB + tid;
            key[k] = 0;
            rank[k] = 0;
            if (!((active >> k) & 1u)) continue;
            const uint4 va = sva[i], vb = svb[i];
            // byte-equal pairs score 1.0 (strsim.rs:128,182,288,324).  The plane launches and every fused launch
            // found them above (all lanes in lock step) and give them the cheapest key; the single-measure
            // register-compare launch finds them in step 4
            bool settle_equal = false;
            if constexpr (REG || ULAT || is_multi(MEASURE)) settle_equal = row_equal[k];
            uint32_t mx = va.x > vb.x ? va.x : vb.x;
            // fused evaluation streams a against the tabled b for every group, and so do the single Jaro and
            // multiset kernels: the loops run la times (Levenshtein streams the longer string)
            if (REG && (is_multi(MEASURE) || MEASURE != LEVENSHTEIN)) mx = va.x;
            uint32_t wide = 0;  // UREG: a character above U+00FF somewhere in the pair
            if (ULAT) {
                // The Latin-1 launch looks at the first word of each string only: a name in a script above
                // U+00FF starts with such a character (C3: every CJK row), and a pair whose first wide
                // character comes later is caught when it is transcoded in step 4.  The sort key is the
                // BYTE length of the streamed string -- a few diacritics more or less do not change the cost
                // class.  (The first version counted the characters of both strings here, word by word with a
                // third of the lanes active: a fifth of the launch's instructions.)
                wide = wide_bytes(staged_first_word(smem, staged_str(va, i, (uint32_t)L::off_sva, off_stage_a))) |
                       wide_bytes(staged_first_word(smem, staged_str(vb, i, (uint32_t)L::off_svb, off_stage_b)));
                if (is_multi(MEASURE) || MEASURE != LEVENSHTEIN) mx = va.x;  // the plane path streams a
            } else if (UREG) {
                // the register-compare path costs (streamed characters) x (tabled characters): bucket by
                // CHARACTER counts so that e.g. 6-character CJK rows do not share a warp with 18-character
                // Latin rows of the same byte length.
                const int ca = staged_char_count(va, stage_a, wide), cb = staged_char_count(vb, stage_b, wide);
                mx = (uint32_t)(ca > cb ? ca : cb);
            }
            if (UREG) {
                // each of the two launches over a general column takes one class; the first one counts what
                // it leaves to the second (none in a Latin-1 column: the host then skips that launch)
                if (ULAT && wide != 0u) {
                    listed |= 1u << k;  // the second launch gathers these rows only (list_wide_rows below)
                    continue;
                }
                if (!ULAT && s.skip_latin && wide == 0u) continue;
            }
            // fused kernel: the equal pairs get the cheapest bucket of their own (key 1 otherwise holds only
            // empty/empty pairs, which are equal too) -- whole warps of them leave the row function at its
            // first test, and the stores stay dense
            // (the register kernels read "key 1" as "byte-equal" in step 4: there a pair with an empty a and a
            // non-empty b must not share that key, so their other keys start at 2)
            key[k] = settle_equal ? 1u : (REG || ULAT ? 2u : 1u) + mx;
            rank[k] = atomicAdd(&hist[key[k]], 1u);
        }
        if constexpr (ULAT)
            list_wide_rows<RPT>(s, listed, lane, [&](int k) {
                const long long idx = tile0 + k * TPB + tid;
                return (unsigned int)(GATHER ? (long long)s.list[idx] : idx);
            });
        __syncthreads();
        if (warp == 0) {
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
            if (lane == 31) hist[0] = incl;  // total active rows
        }
        __syncthreads();
        const int n_active = (int)hist[0];
#pragma unroll
        for (int k = 0; k < RPT; k++)
            if (key[k]) perm[hist[key[k]] + rank[k]] = (uint16_t)(k * TPB + tid);
        __syncthreads();

        // ---------------- 4. compute -----------------------------------------------------------------
        if (GATHER) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const long long nxt = tile_after * TILE + k * TPB + tid;
                if (nxt < n) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.a.views + (long long)next_row[k] * s.a.stride));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(s.b.views + (long long)next_row[k] * s.b.stride));
                }
            }
        }
        uint32_t late = 0;  // ULAT: bit k = the pair of round k turned out wide while it was transcoded
#pragma unroll 1
        for (int k = 0; k < RPT; k++) {
            const int p = k * TPB + ((k & 1) ? (TPB - 1 - tid) : tid);  // snake order balances warps
            if constexpr (ULAT) {
                // the Latin-1 half of a general column: no character above U+00FF in these pairs.  The sort put
                // the byte-equal pairs (key 1) last; the others are copied into the thread's slab, rewritten
                // there to one byte per character (row_ascii_reg.cuh: transcode_latin1) and run the plane
                // path with 8 planes straight from the slab
                if (p >= n_active) continue;
                const int i = perm[p];
                const long long idx = tile0 + i;
                const long long row = GATHER ? (long long)s.list[idx] : idx;
                if (p >= (int)hist[1]) {
                    store_settled<MEASURE>(s, row, 1.0, F_EQUAL);
                    continue;
                }
                const uint4 va = sva[i], vb = svb[i];
                // straight from the staged tile into the thread's slab, one byte per character
                const StagedStr SA = staged_str(va, i, (uint32_t)L::off_sva, off_stage_a),
                                SB = staged_str(vb, i, (uint32_t)L::off_svb, off_stage_b);
                uint32_t wide = 0;
                SlabWriter<TPB> wa{store.wa_}, wb{store.wb_};
                const int ca = transcode_latin1_stream(
                    StagedWords{reinterpret_cast<const uint32_t*>(smem + (SA.off & ~3u)), (int)(SA.off & 3u) * 8, SA.len}, wa,
                    SA.len, wide);
                const int cb = transcode_latin1_stream(
                    StagedWords{reinterpret_cast<const uint32_t*>(smem + (SB.off & ~3u)), (int)(SB.off & 3u) * 8, SB.len}, wb,
                    SB.len, wide);
                if (wide != 0u) {
                    // a character above U+00FF behind the first word: the register-compare launch takes the pair
                    late |= 1u << k;
                    continue;
                }
                typedef SlabSrc<TPB, SlabByteAt<TPB>> Src;
                const Src A{store.wa_, ca, SlabByteAt<TPB>{smem_u32(store.wa_)}},
                          B{store.wb_, cb, SlabByteAt<TPB>{smem_u32(store.wb_)}};
                if constexpr (is_multi(MEASURE)) {
                    RowEmit emit{s, row, true};
                    row_planes_multi<GROUPS, 8>(A, B, emit);
                } else {
                    PairInts ints;
                    s.out[row] = row_planes<is_multi(MEASURE) ? 0 : MEASURE, 8>(A, B, ints);
                    if (s.dbg) store_dbg(s.dbg + row * 6, ints);
                }
                continue;
            }
            if (UREG) {
                // every lane of the warp runs the row function (idle lanes on an empty, "equal" pair):
                // the compare loops are bounded by a warp-wide maximum
                const bool has = p < n_active;
                const int i = has ? (int)perm[p] : 0;
                int na = 0, nb = 0;
                bool equal = true;
                if (has) {
                    const uint4 va = sva[i], vb = svb[i];
                    na = (int)va.x;
                    nb = (int)vb.x;
                    load_string<WORDS, TPB>(va, stage_a, store.wa_);
                    load_string<WORDS, TPB>(vb, stage_b, store.wb_);
                    equal = na == nb;
                    if (equal) {
                        const int nw = (na + 3) >> 2;
                        uint32_t diff = 0;
                        for (int w = 0; w < nw; w++) diff |= store.wa(w) ^ store.wb(w);
                        equal = diff == 0;
                    }
                }
                WarpMaxDev wm;
                if constexpr (is_multi(MEASURE)) {
                    const long long idx = tile0 + i;
                    RowEmit emit{s, has ? (GATHER ? (long long)s.list[idx] : idx) : 0ll, has};
                    row_unicode_reg_multi<GROUPS>(store, na, nb, equal, wm, emit);
                    continue;
                }
                PairInts ints;
                const double v = row_unicode_reg<is_multi(MEASURE) ? 0 : MEASURE>(store, na, nb, equal, wm, ints);
                if (has) {
                    const long long idx = tile0 + i;
                    const long long row = GATHER ? (long long)s.list[idx] : idx;
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
                continue;
            }
            if (p >= n_active) continue;
            const int i = perm[p];
            const uint4 va = sva[i], vb = svb[i];
            const int na = (int)va.x, nb = (int)vb.x;
            PairInts ints;
            double v;
            if constexpr (REG && is_multi(MEASURE)) {
                // the sort put the byte-equal pairs (key 1) last: whole warps of them leave here
                const long long idx = tile0 + i;
                RowEmit emit{s, GATHER ? (long long)s.list[idx] : idx, true};
                if (p >= (int)hist[1]) {
                    PairInts o;
                    o.flag = F_EQUAL;
                    o.la = o.lb = o.x0 = o.x1 = o.x2 = 0;
                    emit_groups<GROUPS>(emit, 1.0, o);
                    continue;
                }
                // b is tabled as bit planes, a is streamed -- both straight from the staged tile
                const StagedStr A = staged_str(va, i, (uint32_t)L::off_sva, off_stage_a),
                                B = staged_str(vb, i, (uint32_t)L::off_svb, off_stage_b);
                row_planes_multi<GROUPS, NBITS, M>(StagedSrc{smem, A.off, A.len}, StagedSrc{smem, B.off, B.len}, emit);
                continue;
            }
            // REG, one measure.  The sort put the byte-equal pairs (key 1) last: whole warps of them leave here
            if (p >= (int)hist[1]) {
                const long long idx = tile0 + i;
                store_settled<MEASURE>(s, GATHER ? (long long)s.list[idx] : idx, 1.0, F_EQUAL);
                continue;
            }
            {
                const StagedStr A = staged_str(va, i, (uint32_t)L::off_sva, off_stage_a),
                                B = staged_str(vb, i, (uint32_t)L::off_svb, off_stage_b);
                v = row_planes<is_multi(MEASURE) ? 0 : MEASURE, NBITS, M>(StagedSrc{smem, A.off, A.len},
                                                                         StagedSrc{smem, B.off, B.len}, ints);
            }
            const long long idx = tile0 + i;
            const long long row = GATHER ? (long long)s.list[idx] : idx;
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
        if constexpr (ULAT)
            list_wide_rows<RPT>(s, late, lane, [&](int k) {
                const long long idx = tile0 + perm[k * TPB + ((k & 1) ? (TPB - 1 - tid) : tid)];
                return (unsigned int)(GATHER ? (long long)s.list[idx] : idx);
            });
        __syncthreads();  // smem is reused by the next tile
    }
    // the last CTA to leave puts the counters back to zero for the next launch over this overflow record
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&s.ovf->done_ctr[s.ctr], 1u) == gridDim.x - 1u) {
            s.ovf->tile_ctr[s.ctr] = 0u;
            s.ovf->done_ctr[s.ctr] = 0u;
        }
    }
}

}  // namespace strsim
