// host.cu -- host driver of libpolars_strsim_b200.so: column upload, segment alignment, kernel
// launches, result download.  Replaces parallel_apply() (/root/reference/src/expressions/strsim.rs:
// 41-107) and the polars-core arity kernels it calls (chunk alignment, validity AND).
//
// No CPU fallback: every path ends in the CUDA kernels of short_kernel.cuh / generic_kernel.cuh /
// long_lev_kernel.cuh, and fails with STRSIM_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>
#include <ctype.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sched.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <algorithm>
#include <mutex>
#include <new>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "../../include/strsim_b200.h"
#include "direct_kernel.cuh"
#include "generic_kernel.cuh"
#include "long_lev_kernel.cuh"
#include "long_pair_kernel.cuh"
#include "short_kernel.cuh"

using namespace strsim;

// ---- error plumbing --------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static thread_local int64_t g_last_overflow[2] = {0, 0};
static thread_local int g_last_redo_slices = 0;  // slices of the last host call that had to be recomputed
static thread_local int64_t g_deferred = 0;  // rows left out because their payload was still in flight
static std::atomic<uint64_t> g_launches{0};

extern "C" void strsim_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            strsim_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                             __LINE__, #expr);                                                \
            return STRSIM_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

// Nothing unwinds across the C ABI: an exception on the host side (std::bad_alloc from the bookkeeping
// vectors) becomes a status code and a last-error message, like every other failure.
template <class F>
static int guarded(const char* what, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        strsim_set_error("%s: out of host memory", what);
        return STRSIM_ERR_NOMEM;
    } catch (const std::exception& e) {
        strsim_set_error("%s: %s", what, e.what());
        return STRSIM_ERR_ARGUMENT;
    } catch (...) {
        strsim_set_error("%s: unknown exception", what);
        return STRSIM_ERR_ARGUMENT;
    }
}

constexpr int STRSIM_RETRY_UNTRIMMED = 100;  // internal status of host_call_impl(trim = true), never returned to callers
constexpr int SHARD_DEVICE_BASE = 1000;      // strsim_b200_column_device() of a column kept by a sharded call

// row slices of one host call (compute_host_multi): H2D of slice s+1 overlaps compute / D2H of slice s
constexpr int MAX_SLICES = 16;
constexpr long long SLICE_ROWS = 1ll << 21;

// ---- per-thread device context ---------------------------------------------------------------------
// Small device -> host readbacks (overflow counters, column statistics) go through a tiny kernel that
// stores into pinned host memory instead of cudaMemcpyAsync: a 20-byte copy queued on the D2H copy
// engine waits behind the 80 MB result download of the previous row slice, which stalled the upload
// and compute streams for 0.2-0.35 ms per slice (timeline in profiles/r1_e2e_timeline.md).
__global__ void publish_words_kernel(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst_host, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst_host[i] = src[i];
    __threadfence_system();
}
static cudaError_t publish_to_host(void* h_pinned, const void* d_src, size_t bytes, cudaStream_t st) {
    void* d_alias = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&d_alias, h_pinned, 0);
    if (e != cudaSuccess) return e;
    publish_words_kernel<<<1, 64, 0, st>>>(static_cast<const unsigned int*>(d_src), static_cast<unsigned int*>(d_alias),
                                          (int)(bytes / 4));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

struct Workspace {
    void* ptr = nullptr;
    size_t cap = 0;
};

// ---- downloads into pageable host memory --------------------------------------------------------------
// A Polars plugin call returns its results in freshly allocated, pageable memory.  cudaMemcpyAsync into
// such memory is staged by the driver on the calling thread and takes every first-touch page fault of
// the new buffer there: measured 2.1 GB/s (37 ms for the 80 MB of one measure over 10 M rows, against
// 0.5 ms of kernels).  Instead the results are DMA'd into a ring of pinned slots at the full PCIe rate
// and a small pool of host threads copies the slots out -- the page faults and the copy spread over
// the pool's threads.  Pinned destinations (bench.py, callers that pin) are still DMA'd directly.
constexpr int STAGE_SLOTS = 16;
constexpr size_t STAGE_SLOT_BYTES = 4u << 20;

struct StageTask {
    int device;
    cudaEvent_t ev;  // wait for it before copying (downloads); nullptr for uploads (host -> ring)
    const void* src;
    void* dst;
    size_t bytes;
    std::atomic<int>* flag;  // set to flag_value once the bytes are copied
    int flag_value;
    std::atomic<long long>* pending;  // decremented afterwards (may be nullptr)
    bool streaming;  // destination is the write-combined upload ring: non-temporal stores
};

// memcpy with non-temporal stores: the destination (a slot of the pinned upload ring) is next read by the
// DMA engine, not by a CPU, so the lines need neither be fetched for ownership nor stay in a cache.
// Measured on the pool's 16-core host (exp/stage_bw.cu, 8 threads feeding a ring that is DMA'd
// concurrently): memcpy 39.6 GB/s, this into a write-combined ring 49.2 GB/s.
static void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__SSE2__)
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        for (; i + 64 <= bytes; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
        }
    }
    if (i < bytes) memcpy(d + i, s + i, bytes - i);
    _mm_sfence();
#else
    memcpy(dst, src, bytes);
#endif
}

static const std::vector<int>& shard_devices();  // the GPUs one host call is sharded over (STRSIM_B200_DEVICES)

class StagePool {
   public:
    static StagePool& get() {
        static StagePool* p = new StagePool();  // never destroyed: workers outlive static destructors
        return *p;
    }
    void push(const StageTask& t) {
        {
            std::lock_guard<std::mutex> lock(m_);
            q_.push_back(t);
        }
        cv_.notify_one();
    }

   private:
    StagePool() {
        // half the cores, at most 8 (16 when one call is sharded over four or more GPUs) -- shared with the
        // other processes of a one-process-per-GPU launch (torchrun sets LOCAL_WORLD_SIZE), which would
        // otherwise oversubscribe the host with 8 copy threads each
        unsigned n = std::thread::hardware_concurrency() / 2;
        if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {
            const int ranks = atoi(lw);  // measured at 2 ranks on 16 cores: 8 threads per rank beat 4 (30 vs 42 ms per C2 step)
            if (ranks > 1) n = std::thread::hardware_concurrency() / (unsigned)ranks;
        }
        const unsigned cap = shard_devices().size() >= 4 ? 16u : 8u;
        if (n < 2) n = 2;
        if (n > cap) n = cap;
        if (const char* e = getenv("STRSIM_B200_COPY_THREADS")) {
            const int v = atoi(e);
            if (v > 0 && v <= 64) n = (unsigned)v;
        }
        for (unsigned i = 0; i < n; i++) std::thread([this] { run(); }).detach();
    }
    void run() {
        int device = -1;
        for (;;) {
            StageTask t;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [this] { return !q_.empty(); });
                t = q_.front();
                q_.pop_front();
            }
            if (t.ev) {
                if (t.device != device) {
                    cudaSetDevice(t.device);
                    device = t.device;
                }
                cudaEventSynchronize(t.ev);
            }
            if (t.streaming)
                stream_copy(t.dst, t.src, t.bytes);
            else
                memcpy(t.dst, t.src, t.bytes);
            t.flag->store(t.flag_value, std::memory_order_release);
            if (t.pending) t.pending->fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<StageTask> q_;
};

struct Stager {  // one per host thread (ThreadCtx)
    void* ring = nullptr;  // STAGE_SLOTS x STAGE_SLOT_BYTES, pinned
    cudaEvent_t ev[STAGE_SLOTS] = {};
    std::atomic<int> busy[STAGE_SLOTS];
    std::atomic<long long> pending{0};
    int next = 0;
};

struct ThreadCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // D2H of finished measures overlaps the next kernel
    cudaStream_t upload_stream = nullptr;  // H2D of later row slices overlaps compute of earlier ones
    cudaEvent_t done_event[8] = {};
    cudaEvent_t slice_event[MAX_SLICES] = {};
    cudaEvent_t up_event[MAX_SLICES] = {};  // the H2D copies of a slice have landed
    cudaStream_t stats_stream = nullptr;    // column statistics of uploaded slices (off the copy queue)
    ColumnStats* d_slice_stats = nullptr;  // [0] data of a, [1] data of b, [2+s] views of slice s
    ColumnStats* h_slice_stats = nullptr;  // pinned
    int sm_count = 0;
    Overflow* d_ovf = nullptr;      // the counters the current launches fill
    Overflow* d_ovf_alt = nullptr;  // a second set: a follow-up launch that reads one list and fills another (finish_wide)
    Overflow* d_ovf_base = nullptr; // the allocation behind the two (they swap roles)
    Overflow* h_ovf = nullptr;  // pinned
    unsigned long long* d_nulls = nullptr;
    unsigned long long* h_nulls = nullptr;  // pinned
    ColumnStats* d_stats = nullptr;
    unsigned int* d_counters = nullptr;  // [0] long-kernel cursor, [1] huge-row count
    unsigned int* h_counters = nullptr;  // pinned
    ColumnStats* h_stats = nullptr;  // pinned
    Workspace lists, scratch;
    Stager* stager = nullptr;     // created by the first download into pageable memory
    struct UpStager* up_stager = nullptr;  // created by the first upload from pageable memory
};

static thread_local ThreadCtx g_ctx;
static thread_local int g_requested_device = -2;  // -2: not chosen yet

static int default_device() {
    const char* e = getenv("STRSIM_B200_DEVICE");
    if (e && *e) return atoi(e);
    e = getenv("LOCAL_RANK");  // one process per GPU under torchrun
    if (e && *e) {
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) return atoi(e) % n;
    }
    return 0;
}

static void destroy_up_stager(struct UpStager* sg);
static const std::vector<int>& shard_devices();

static void destroy_ctx(ThreadCtx& c) {
    // best effort: the context may be half built (failed initialisation) or belong to a device that is
    // still busy with other threads' work; every handle is checked before it is released
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    for (cudaStream_t st : {c.stream, c.copy_stream, c.upload_stream, c.stats_stream})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (auto& ev : c.done_event) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.slice_event) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.up_event) if (ev) cudaEventDestroy(ev);
    for (void* p : {(void*)c.d_slice_stats, (void*)c.d_ovf_base, (void*)c.d_nulls, (void*)c.d_stats, (void*)c.d_counters,
                    c.lists.ptr, c.scratch.ptr})
        if (p) cudaFree(p);
    for (void* p : {(void*)c.h_slice_stats, (void*)c.h_ovf, (void*)c.h_nulls, (void*)c.h_counters, (void*)c.h_stats})
        if (p) cudaFreeHost(p);
    destroy_up_stager(c.up_stager);
    for (Stager* sg : {c.stager})
        if (sg) {
            while (sg->pending.load(std::memory_order_acquire) > 0) std::this_thread::yield();
            for (auto& ev : sg->ev) if (ev) cudaEventDestroy(ev);
            if (sg->ring) cudaFreeHost(sg->ring);
            delete sg;
        }
    cudaGetLastError();
    c = ThreadCtx();
}

static int init_ctx(ThreadCtx& c, int device) {
    c.device = device;
    CUDA_TRY(cudaSetDevice(c.device));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.upload_stream, cudaStreamNonBlocking));
    for (auto& ev : c.slice_event) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c.up_event) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.stats_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&c.d_slice_stats, sizeof(ColumnStats) * (2 + MAX_SLICES)));
    CUDA_TRY(cudaMallocHost(&c.h_slice_stats, sizeof(ColumnStats) * (2 + MAX_SLICES)));
    for (auto& ev : c.done_event) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, c.device));
    CUDA_TRY(cudaMalloc(&c.d_ovf_base, 2 * sizeof(Overflow)));
    c.d_ovf = c.d_ovf_base;
    c.d_ovf_alt = c.d_ovf_base + 1;
    CUDA_TRY(cudaMallocHost(&c.h_ovf, sizeof(Overflow)));
    CUDA_TRY(cudaMalloc(&c.d_nulls, sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&c.h_nulls, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&c.d_stats, sizeof(ColumnStats)));
    CUDA_TRY(cudaMalloc(&c.d_counters, 4 * sizeof(unsigned int)));
    CUDA_TRY(cudaMallocHost(&c.h_counters, 4 * sizeof(unsigned int)));
    CUDA_TRY(cudaMallocHost(&c.h_stats, sizeof(ColumnStats)));
    // the quotient table of pair_algos.cuh, once per device (every host thread has its own context)
    static std::mutex quot_mutex;
    static bool quot_ready[64] = {false};
    std::lock_guard<std::mutex> lock(quot_mutex);
    if (c.device >= 64 || !quot_ready[c.device]) {
        quotient_table_kernel<<<(QUOT_N * QUOT_N + 255) / 256, 256, 0, c.stream>>>();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        if (c.device < 64) quot_ready[c.device] = true;
    }
    return STRSIM_OK;
}

static int ensure_ctx(ThreadCtx** out) {
    if (g_requested_device == -2) g_requested_device = default_device();
    ThreadCtx& c = g_ctx;
    if (c.device != g_requested_device) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            strsim_set_error(
                "no usable CUDA device (%s): polars-strsim_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
            return STRSIM_ERR_CUDA;
        }
        if (g_requested_device < 0 || g_requested_device >= n) {
            strsim_set_error("device %d out of range (0..%d)", g_requested_device, n - 1);
            return STRSIM_ERR_ARGUMENT;
        }
        // the context is built aside and committed only when every stream, event and buffer exists: a
        // failed initialisation (out of memory) must not leave a half-built context that the next call
        // would take for a usable one, and a thread that switches device gives the old resources back
        destroy_ctx(c);
        ThreadCtx fresh;
        const int rc = init_ctx(fresh, g_requested_device);
        if (rc != STRSIM_OK) {
            const std::string msg = g_last_error;
            destroy_ctx(fresh);
            g_last_error = msg;
            return rc;
        }
        c = fresh;
    }
    CUDA_TRY(cudaSetDevice(c.device));
    *out = &c;
    return STRSIM_OK;
}

static int ws_reserve(Workspace& w, size_t bytes) {
    if (bytes <= w.cap) return STRSIM_OK;
    if (w.ptr) CUDA_TRY(cudaFree(w.ptr));
    w.ptr = nullptr;
    w.cap = 0;
    size_t cap = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMalloc(&w.ptr, cap);
    if (e != cudaSuccess) {
        strsim_set_error("cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
        return STRSIM_ERR_NOMEM;
    }
    w.cap = cap;
    return STRSIM_OK;
}

// ---- device block pool: cudaMalloc / cudaFree cost milliseconds (and cudaFree synchronises), so the
//      per-call blocks (uploaded columns, results) are recycled ------------------------------------------
struct PoolBlock {
    int device;
    void* ptr;
    size_t bytes;
    std::chrono::steady_clock::time_point freed_at;
};
// (never destroyed: the plugin's reaper thread may still be trimming when the process runs its static
// destructors -- a vector freed under it shows up as heap corruption at exit)
static std::mutex& g_pool_mutex = *new std::mutex();
static std::vector<PoolBlock>& g_pool = *new std::vector<PoolBlock>();
static size_t pool_limit() {
    static const size_t lim = [] {
        const char* e = getenv("STRSIM_B200_POOL_BYTES");
        return e && *e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)16 << 30);
    }();
    return lim;
}

static int pool_alloc(int device, size_t bytes, void** out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        size_t best = g_pool.size();
        for (size_t i = 0; i < g_pool.size(); i++) {
            const PoolBlock& b = g_pool[i];
            if (b.device == device && b.bytes >= bytes && b.bytes <= bytes + bytes / 2 + (1 << 20) &&
                (best == g_pool.size() || b.bytes < g_pool[best].bytes))
                best = i;
        }
        if (best != g_pool.size()) {
            *out = g_pool[best].ptr;
            g_pool.erase(g_pool.begin() + (long)best);
            return STRSIM_OK;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
        // make room: drop every cached block of this device and retry once
        {
            std::lock_guard<std::mutex> lock(g_pool_mutex);
            for (size_t i = g_pool.size(); i-- > 0;)
                if (g_pool[i].device == device) {
                    cudaFree(g_pool[i].ptr);
                    g_pool.erase(g_pool.begin() + (long)i);
                }
        }
        cudaGetLastError();
        e = cudaMalloc(out, bytes);
    }
    if (e != cudaSuccess) {
        strsim_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return STRSIM_ERR_NOMEM;
    }
    return STRSIM_OK;
}

// the block must no longer be in use by any stream (callers synchronise first)
static void pool_free(int device, void* ptr, size_t bytes) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    size_t total = bytes;
    for (const PoolBlock& b : g_pool) total += b.bytes;
    if (total > pool_limit() || g_pool.size() >= 16) {
        cudaFree(ptr);
        return;
    }
    g_pool.push_back({device, ptr, bytes, std::chrono::steady_clock::now()});
}

// hands the blocks that sat unused for at least `idle_seconds` back to the driver (the plugin's cache reaper
// calls this once the cache has emptied, so an idle process does not sit on gigabytes of HBM)
extern "C" void strsim_pool_trim(int idle_seconds) {
    std::vector<PoolBlock> drop;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        const auto now = std::chrono::steady_clock::now();
        for (size_t i = g_pool.size(); i-- > 0;)
            if (std::chrono::duration_cast<std::chrono::seconds>(now - g_pool[i].freed_at).count() >= idle_seconds) {
                drop.push_back(g_pool[i]);
                g_pool.erase(g_pool.begin() + (long)i);
            }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (const PoolBlock& b : drop) {
        cudaSetDevice(b.device);
        cudaFree(b.ptr);
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

// ---- pinned result buffers ------------------------------------------------------------------------------
// A plugin call returns its Float64 column in a buffer the plugin allocates.  When that buffer is PINNED the
// results are DMA'd straight into it (1.5 ms per 80 MB); a pageable buffer costs a second pass through the
// pinned ring plus every first-touch page fault of fresh memory (3+ ms).  Pinning is expensive (tens of
// milliseconds per 80 MB), so buffers are pooled: a call takes a free block when there is one -- otherwise
// it falls back to pageable memory AND asks a background thread to pin a block of that size for the calls
// to come; the release callback of the Arrow array hands the block back.  At most
// STRSIM_B200_PINNED_RESULT_BYTES (default 4 GiB) are pinned at any time; blocks idle for the cache's
// time-to-live are unpinned by the plugin's reaper (strsim_result_pool_trim).
struct PinnedBlock {
    void* ptr;
    size_t bytes;
    bool in_use;
    std::chrono::steady_clock::time_point freed_at;
};
static std::mutex& g_pinned_mutex = *new std::mutex();  // never destroyed, like the device pool above
static std::condition_variable& g_pinned_cv = *new std::condition_variable();
static std::vector<PinnedBlock>& g_pinned = *new std::vector<PinnedBlock>();
static std::deque<size_t>& g_pinned_requests = *new std::deque<size_t>();
static bool g_pinned_grower_running = false;

static size_t pinned_limit() {
    static const size_t lim = [] {
        const char* e = getenv("STRSIM_B200_PINNED_RESULT_BYTES");
        return e && *e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)4 << 30);
    }();
    return lim;
}

static void pinned_grower_main() {
    for (;;) {
        size_t want = 0;
        {
            std::unique_lock<std::mutex> lock(g_pinned_mutex);
            if (!g_pinned_cv.wait_for(lock, std::chrono::seconds(5), [] { return !g_pinned_requests.empty(); })) {
                g_pinned_grower_running = false;  // idle: the next request starts a new thread
                return;
            }
            want = g_pinned_requests.front();
            g_pinned_requests.pop_front();
            size_t total = want;
            for (const PinnedBlock& b : g_pinned) total += b.bytes;
            // (a free block of this size does not cancel the request: requests are counted against the results
            // alive at once, and the burst that raised them has usually been released by now)
            if (total > pinned_limit()) continue;  // over budget
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        g_pinned.push_back({p, want, false, std::chrono::steady_clock::now()});
    }
}

// results of (roughly) this size that live in pageable memory right now because the pool had no block for them
static std::vector<std::pair<size_t, int>>& g_unpinned_live = *new std::vector<std::pair<size_t, int>>();  // guarded by g_pinned_mutex

static bool same_class(size_t block, size_t want) { return block >= want && block <= want + want / 2 + (1u << 20); }

extern "C" void* strsim_result_alloc(size_t bytes) {
    if (bytes < (1u << 20) || pinned_limit() == 0) return nullptr;
    const size_t want = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    size_t best = g_pinned.size();
    for (size_t i = 0; i < g_pinned.size(); i++) {
        const PinnedBlock& b = g_pinned[i];
        if (!b.in_use && same_class(b.bytes, want) && (best == g_pinned.size() || b.bytes < g_pinned[best].bytes)) best = i;
    }
    if (best != g_pinned.size()) {
        g_pinned[best].in_use = true;
        return g_pinned[best].ptr;
    }
    // No block: this result goes to pageable memory.  The pool should hold as many blocks of this size as
    // there are results of this size alive at once (held by the engine, or computed ahead and waiting): ask
    // the background thread for one more block only while blocks + pending requests fall short of that --
    // every miss used to queue a request, and a dozen 80 MB cudaHostAlloc calls (35 ms each, serialised
    // with the streams by the driver) then ran into the calls that followed.
    int live = 1;
    bool found = false;
    for (auto& u : g_unpinned_live)
        if (u.first == want) {
            live = ++u.second;
            found = true;
        }
    if (!found) g_unpinned_live.emplace_back(want, 1);
    int have = 0;
    for (const PinnedBlock& b : g_pinned)
        if (same_class(b.bytes, want)) have += b.in_use ? 0 : 1;  // (free blocks of the class: none, or we would not be here)
    int pending = 0;
    for (size_t r : g_pinned_requests) pending += r == want ? 1 : 0;
    if (pending + have < live && g_pinned_requests.size() < 16) {
        g_pinned_requests.push_back(want);
        if (!g_pinned_grower_running) {
            try {
                std::thread(pinned_grower_main).detach();
                g_pinned_grower_running = true;
            } catch (...) {
                g_pinned_requests.clear();
            }
        }
        g_pinned_cv.notify_one();
    }
    return nullptr;
}

// a result that strsim_result_alloc() turned away (pageable memory, `bytes` as asked for then) has been freed
extern "C" void strsim_result_unpinned_released(size_t bytes) {
    if (bytes < (1u << 20)) return;
    const size_t want = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    for (auto& u : g_unpinned_live)
        if (u.first == want && u.second > 0) u.second--;
}

// true when `p` was a block of the pool (now free again); callable from any thread, makes no CUDA call
extern "C" int strsim_result_free(void* p) {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    for (PinnedBlock& b : g_pinned)
        if (b.ptr == p) {
            b.in_use = false;
            b.freed_at = std::chrono::steady_clock::now();
            return 1;
        }
    return 0;
}

// returns the number of blocks the pool still owns
extern "C" int strsim_result_pool_trim(int idle_seconds) {
    std::vector<void*> drop;
    int left = 0;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        const auto now = std::chrono::steady_clock::now();
        for (size_t i = g_pinned.size(); i-- > 0;)
            if (!g_pinned[i].in_use &&
                std::chrono::duration_cast<std::chrono::seconds>(now - g_pinned[i].freed_at).count() >= idle_seconds) {
                drop.push_back(g_pinned[i].ptr);
                g_pinned.erase(g_pinned.begin() + (long)i);
            }
        left = (int)g_pinned.size();
    }
    for (void* p : drop) cudaFreeHost(p);
    cudaGetLastError();
    return left;
}

// ---- device-resident column --------------------------------------------------------------------------
struct DevChunk {
    const uint4* views;  // row 0 of the chunk's logical range
    const uint8_t* validity;
    long long vbit;
    const unsigned long long* bufs;
    int64_t length;
    int64_t data_bytes;  // sum of data buffer sizes
};

struct strsim_b200_column {
    int device = 0;
    // A column a SHARDED host call kept (STRSIM_B200_DEVICES): one resident column per device, shard g
    // holding rows [shard_lo[g], shard_lo[g+1]); `device` is then the marker SHARD_DEVICE_BASE + number of
    // shards and no other field but `length` is used.
    std::vector<strsim_b200_column*> shards;
    std::vector<int64_t> shard_lo;
    // a column materialised from dictionary-encoded chunks: the dictionaries (resident columns whose data
    // buffers this column's views point into), freed with it
    std::vector<strsim_b200_column*> dictionaries;
    void* block = nullptr;  // one allocation holds everything
    size_t block_bytes = 0;
    std::vector<DevChunk> chunks;
    int64_t length = 0;
    int64_t data_bytes = 0;
    int64_t alg_bytes = -1;
    bool has_validity = false;
    bool scalar_null = false;  // the column has exactly one row and that row is null
    unsigned or_byte = 0, and_byte = 0xFF;  // OR / AND over every string byte of the column
    // longest string (bytes) and largest padded out-of-line payload of an aligned STATS_BLOCK-row block of a
    // chunk; 0xFFFFFFFF = not known (columns assembled by host calls, dictionary columns)
    unsigned max_len = 0xFFFFFFFFu, max_block_pad = 0xFFFFFFFFu;
    // distinct data buffers in upload order (chunks made by slicing share buffers) and how much of each
    // is on the device right now: equal to the size except while a host call uploads progressively
    std::vector<int64_t> buf_size, buf_resident;
    // the part of each distinct buffer this column's rows reference, [buf_lo, buf_hi) (multiples of 256 unless
    // they are 0 / the size): everything for a whole column; a row shard of a sequentially built column
    // (one GPU's slice of a sharded call) needs -- and uploads -- only its own stretch of the data
    std::vector<int64_t> buf_lo, buf_hi;
    // per distinct data buffer: offset of the buffer's (virtual) byte 0 inside `block`; bytes below buf_lo
    // have no storage, so the offset may be negative
    std::vector<int64_t> buf_dev_off;
    std::vector<std::vector<int>> chunk_buf_ids;  // [chunk][buffer index] -> distinct buffer id
    // general columns: share of pairs with a character above U+00FF seen by the last Latin-1 launch over
    // this column (-1: unknown); a hint only, read and written without synchronisation
    mutable float wide_share = -1.0f;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
extern "C" void strsim_b200_column_free(strsim_b200_column* col);

static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// device -> host copy of `bytes` on stream `st`: direct DMA into pinned memory, ring + copy threads otherwise
static cudaError_t download(ThreadCtx& ctx, void* dst, const void* d_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    static const bool no_stage = getenv("STRSIM_B200_STAGED_D2H") != nullptr && !strcmp(getenv("STRSIM_B200_STAGED_D2H"), "0");
    if (no_stage || bytes < (1u << 20) || is_pinned(dst)) return cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
    if (!ctx.stager) {
        Stager* sg = new Stager();
        cudaError_t e = cudaMallocHost(&sg->ring, (size_t)STAGE_SLOTS * STAGE_SLOT_BYTES);
        if (e != cudaSuccess) {
            delete sg;
            cudaGetLastError();
            return cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
        }
        for (int i = 0; i < STAGE_SLOTS; i++) {
            cudaEventCreateWithFlags(&sg->ev[i], cudaEventDisableTiming);
            sg->busy[i].store(0);
        }
        ctx.stager = sg;
    }
    Stager& sg = *ctx.stager;
    StagePool& pool = StagePool::get();
    for (size_t off = 0; off < bytes; off += STAGE_SLOT_BYTES) {
        const size_t len = bytes - off < STAGE_SLOT_BYTES ? bytes - off : STAGE_SLOT_BYTES;
        const int slot = sg.next;
        sg.next = (sg.next + 1) % STAGE_SLOTS;
        while (sg.busy[slot].load(std::memory_order_acquire)) std::this_thread::yield();
        sg.busy[slot].store(1, std::memory_order_relaxed);
        sg.pending.fetch_add(1, std::memory_order_acq_rel);
        char* pinned = static_cast<char*>(sg.ring) + (size_t)slot * STAGE_SLOT_BYTES;
        cudaError_t e = cudaMemcpyAsync(pinned, static_cast<const char*>(d_src) + off, len, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(sg.ev[slot], st);
        if (e != cudaSuccess) {
            sg.busy[slot].store(0);
            sg.pending.fetch_sub(1);
            return e;
        }
        pool.push(StageTask{ctx.device, sg.ev[slot], pinned, static_cast<char*>(dst) + off, len, &sg.busy[slot], 0, &sg.pending, false});
    }
    return cudaSuccess;
}

// every staged download of this thread has reached its destination
static void download_wait(ThreadCtx& ctx) {
    if (!ctx.stager) return;
    while (ctx.stager->pending.load(std::memory_order_acquire) > 0) std::this_thread::yield();
}

// ---- uploads from pageable host memory --------------------------------------------------------------------
// Pinned sources are DMA'd directly (asynchronous).  Pageable sources -- what Polars hands a plugin -- would
// be staged by the driver on the calling thread at ~12 GB/s; here the pool's threads copy them (non-temporal
// stores) into a ring of write-combined pinned slots and every filled slot is DMA'd at once.  The ring is a
// pipeline ACROSS upload_copy() calls: a call only hands its chunks to the copy threads and sends whatever
// slots are full by then; upload_flush() sends the rest.  (The first version drained the ring at the end of
// every call -- four calls per row slice, each with a start-up and a tail of one slot's copy time during
// which the other threads idled: 29 GB/s where the same threads reach 49 GB/s in exp/stage_bw.cu.)
constexpr int UP_SLOTS = 32;
constexpr size_t UP_SLOT_BYTES = 2u << 20;

struct UpStager {  // one per host thread (ThreadCtx)
    void* ring = nullptr;  // UP_SLOTS x UP_SLOT_BYTES, pinned, write-combined
    cudaEvent_t ev[UP_SLOTS] = {};  // the DMA that last read the slot
    std::atomic<int> state[UP_SLOTS];  // 0 free, 1 being filled, 2 filled
    struct Pending {
        void* d_dst;
        size_t len;
        cudaStream_t st;
    } pend[UP_SLOTS];
    unsigned long long head = 0, tail = 0;  // chunks handed to the copy threads / chunks whose DMA is queued
};

static void destroy_up_stager(UpStager* sg) {
    if (!sg) return;
    while (sg->tail < sg->head) {  // copy threads may still be writing into the ring
        if (sg->state[sg->tail % UP_SLOTS].load(std::memory_order_acquire) == 2) sg->tail++;
        else std::this_thread::yield();
    }
    for (auto& ev : sg->ev)
        if (ev) cudaEventDestroy(ev);
    if (sg->ring) cudaFreeHost(sg->ring);
    delete sg;
}

// sends the filled slots at the tail of the ring, in order; `all`: waits for the copy threads and sends
// everything that was handed out
static cudaError_t upload_pump(UpStager& sg, bool all) {
    cudaError_t e = cudaSuccess;
    while (sg.tail < sg.head) {
        const int slot = (int)(sg.tail % UP_SLOTS);
        if (sg.state[slot].load(std::memory_order_acquire) != 2) {
            if (!all) break;
            std::this_thread::yield();
            continue;
        }
        const UpStager::Pending& p = sg.pend[slot];
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(p.d_dst, static_cast<char*>(sg.ring) + (size_t)slot * UP_SLOT_BYTES, p.len, cudaMemcpyHostToDevice, p.st);
        if (e == cudaSuccess) e = cudaEventRecord(sg.ev[slot], p.st);
        sg.state[slot].store(0, std::memory_order_relaxed);
        sg.tail++;
    }
    return e;
}

// every chunk handed to upload_copy() so far has been copied out of its (pageable) source and its DMA is
// queued: call before recording an event that stands for "the upload has landed", and before the sources
// may go away
static cudaError_t upload_flush(ThreadCtx& ctx) {
    return ctx.up_stager ? upload_pump(*ctx.up_stager, true) : cudaSuccess;
}

static cudaError_t upload_copy(ThreadCtx& ctx, void* d_dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    static const bool no_stage = getenv("STRSIM_B200_STAGED_H2D") != nullptr && !strcmp(getenv("STRSIM_B200_STAGED_H2D"), "0");
    if (no_stage || bytes < (1u << 20) || is_pinned(src)) return cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st);
    if (!ctx.up_stager) {
        UpStager* sg = new UpStager();
        cudaError_t e = cudaHostAlloc(&sg->ring, (size_t)UP_SLOTS * UP_SLOT_BYTES, cudaHostAllocWriteCombined);
        if (e != cudaSuccess) {
            delete sg;
            cudaGetLastError();
            return cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st);
        }
        for (int i = 0; i < UP_SLOTS; i++) {
            cudaEventCreateWithFlags(&sg->ev[i], cudaEventDisableTiming);
            sg->state[i].store(0);
        }
        ctx.up_stager = sg;
    }
    UpStager& sg = *ctx.up_stager;
    StagePool& pool = StagePool::get();
    cudaError_t e = cudaSuccess;
    for (size_t off = 0; off < bytes && e == cudaSuccess; off += UP_SLOT_BYTES) {
        const size_t len = bytes - off < UP_SLOT_BYTES ? bytes - off : UP_SLOT_BYTES;
        while (sg.head - sg.tail >= (unsigned long long)UP_SLOTS && e == cudaSuccess) {  // ring full: send the oldest slot
            if (sg.state[sg.tail % UP_SLOTS].load(std::memory_order_acquire) != 2) std::this_thread::yield();
            e = upload_pump(sg, false);
        }
        if (e != cudaSuccess) break;
        const int slot = (int)(sg.head % UP_SLOTS);
        if (sg.head >= (unsigned long long)UP_SLOTS) e = cudaEventSynchronize(sg.ev[slot]);  // the DMA that last read this slot is done
        if (e != cudaSuccess) break;
        sg.pend[slot] = UpStager::Pending{static_cast<char*>(d_dst) + off, len, st};
        sg.state[slot].store(1, std::memory_order_relaxed);
        pool.push(StageTask{ctx.device, nullptr, static_cast<const char*>(src) + off,
                            static_cast<char*>(sg.ring) + (size_t)slot * UP_SLOT_BYTES, len, &sg.state[slot], 2, nullptr, true});
        sg.head++;
        e = upload_pump(sg, false);
    }
    if (e != cudaSuccess) upload_pump(sg, true);  // never leave copy threads writing into a ring nobody drains
    return e;
}

// H2D copy that never blocks on pageable memory longer than needed: pinned sources are DMA'd
// directly, pageable sources are staged by the driver (cudaMemcpyAsync semantics).
static int h2d(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return STRSIM_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return STRSIM_OK;
}

// Upload of one column, split into steps so that a host call can pipeline row slices:
//   upload_plan  : one device block for everything, DevChunks with their final device addresses
//   upload_data  : buffer tables + every distinct data buffer (+ byte statistics of the buffers)
//   upload_rows  : views + validity of a row range (+ byte statistics of the inline strings)
struct ChunkPlan {
    size_t views_off, validity_off, table_off;
    size_t validity_bytes;
    int64_t first_byte;
    int64_t row0;  // first row of the chunk in the column
};

struct Uploader {
    strsim_b200_column* col = nullptr;
    const strsim_view_chunk* chunks = nullptr;
    size_t n_chunks = 0;
    std::vector<ChunkPlan> plans;
    std::vector<unsigned long long> tables;  // all chunks' buffer tables, alive until the copies ran
    std::vector<size_t> table_pos;
    std::vector<const void*> buf_src;   // per distinct data buffer: host address
    int64_t uploaded = 0;               // progressive upload: linear position over the distinct buffers
};

static void stats_init_value(ColumnStats* s) {
    s->or_bits = 0u;
    s->and_bits = 0xFFFFFFFFu;
    s->max_len = 0u;
    s->max_block_pad = 0u;
}
static void stats_fold(const ColumnStats& s, unsigned* or_byte, unsigned* and_byte) {
    const unsigned o = s.or_bits, a = s.and_bits;
    *or_byte = (o | (o >> 8) | (o >> 16) | (o >> 24)) & 0xFFu;
    *and_byte = (a & (a >> 8) & (a >> 16) & (a >> 24)) & 0xFFu;
}

// is the single row of a length-1 column null?
static bool host_scalar_is_null(const strsim_view_chunk* chunks, size_t n_chunks) {
    for (size_t i = 0; i < n_chunks; i++)
        if (chunks[i].length == 1)
            return chunks[i].validity && !((chunks[i].validity[chunks[i].offset >> 3] >> (chunks[i].offset & 7)) & 1);
    return false;
}

// The null-literal rule.  A length-1 operand against a longer column is a literal (strsim.rs:61-66,85-92).
// The reference unwraps it (`b.get(0).unwrap()`, strsim.rs:62,65,87,90) and PANICS on a null literal; the
// derive wrapper turns that into a failed query.  Here the query fails too, as an ordinary error with a
// message that says why.  A one-row frame (both operands of length 1) is not a literal: nulls propagate
// as for any other row (the reference panics there as well when b is null -- a side effect of the same
// unwrap that is not reproduced).
static int null_literal_error() {
    strsim_set_error("the literal operand is null: a String literal compared with a column must not be null "
                     "(polars-strsim unwraps it and panics, strsim.rs:62,65)");
    return STRSIM_ERR_ARGUMENT;
}

static int64_t view_end_position(const Uploader& up, size_t c, const int32_t* v);
static bool looks_sequential(const Uploader& up);

// `trim`: the chunks are a row shard cut out of a longer column (sharded host call).  When the views look
// sequential, only the stretch of the data buffers between the shard's first and last out-of-line string
// gets device storage and is uploaded; the kernels check every view against that stretch (DevCol::lo_* /
// res_*) and the host falls back to whole buffers should a view point outside it.
static int upload_plan(ThreadCtx& ctx, const strsim_view_chunk* chunks, size_t n_chunks, Uploader* up, bool trim = false) {
    auto* col = new strsim_b200_column();
    col->device = ctx.device;
    up->col = col;
    up->chunks = chunks;
    up->n_chunks = n_chunks;
    up->plans.assign(n_chunks, ChunkPlan());
    struct Seen {
        const void* ptr;
        int64_t size;
        int id;
    };
    std::vector<Seen> seen;
    // pass 1: rows, distinct data buffers (chunks produced by slicing share theirs: each is uploaded once)
    for (size_t i = 0; i < n_chunks; i++) {
        const strsim_view_chunk& ch = chunks[i];
        if (ch.length < 0 || ch.offset < 0 || (ch.length > 0 && !ch.views) || ch.n_data_buffers < 0) {
            strsim_set_error("chunk %zu: bad length/offset/views", i);
            delete col;
            up->col = nullptr;
            return STRSIM_ERR_ARGUMENT;
        }
        ChunkPlan& p = up->plans[i];
        p.row0 = col->length;
        p.first_byte = ch.offset >> 3;
        p.validity_bytes =
            ch.validity && ch.length > 0 ? (size_t)(((ch.offset + ch.length + 7) >> 3) - p.first_byte) : 0;
        col->chunk_buf_ids.emplace_back((size_t)ch.n_data_buffers, 0);
        for (int64_t b = 0; b < ch.n_data_buffers; b++) {
            bool dup = false;
            for (const Seen& sn : seen)
                if (sn.ptr == ch.data_buffers[b] && sn.size == ch.data_buffer_sizes[b]) {
                    col->chunk_buf_ids.back()[(size_t)b] = sn.id;
                    dup = true;
                    break;
                }
            if (dup) continue;
            const int id = (int)col->buf_size.size();
            col->chunk_buf_ids.back()[(size_t)b] = id;
            col->buf_size.push_back(ch.data_buffer_sizes[b] > 0 ? ch.data_buffer_sizes[b] : 0);
            up->buf_src.push_back(ch.data_buffers[b]);
            if (seen.size() < 4096) seen.push_back({ch.data_buffers[b], ch.data_buffer_sizes[b], id});
        }
        col->length += ch.length;
        if (p.validity_bytes) col->has_validity = true;
    }
    col->scalar_null = col->length == 1 && host_scalar_is_null(chunks, n_chunks);
    col->buf_lo.assign(col->buf_size.size(), 0);
    col->buf_hi = col->buf_size;
    col->buf_resident = col->buf_size;
    // pass 2 (row shards only): the stretch of the data the shard references
    if (trim && col->length > 0 && !col->buf_size.empty() && looks_sequential(*up)) {
        // linear positions (over the distinct buffers in upload order) of the first out-of-line string's
        // start and of the last one's end, each found within 4096 rows of the shard's ends
        int64_t first = -1, last = -1, left = 4096;
        for (size_t c = 0; c < n_chunks && left > 0 && first < 0; c++) {
            const int32_t* v = static_cast<const int32_t*>(chunks[c].views) + 4 * chunks[c].offset;
            for (int64_t r = 0; r < chunks[c].length && left > 0; r++, left--) {
                const int64_t e = view_end_position(*up, c, v + 4 * r);
                if (e < 0) continue;
                // start of that string: the end position is rounded up to 256, so recompute from the view
                const auto& ids = col->chunk_buf_ids[c];
                int64_t start = 0;
                for (int id = 0; id < ids[(size_t)v[4 * r + 2]]; id++) start += col->buf_size[(size_t)id];
                first = start + ((int64_t)v[4 * r + 3] & ~255ll);
                break;
            }
        }
        left = 4096;
        for (size_t c = n_chunks; c-- > 0 && left > 0 && last < 0;) {
            const int32_t* v = static_cast<const int32_t*>(chunks[c].views) + 4 * chunks[c].offset;
            for (int64_t r = chunks[c].length; r-- > 0 && left > 0; left--) {
                const int64_t e = view_end_position(*up, c, v + 4 * r);
                if (e >= 0) {
                    last = e;
                    break;
                }
            }
        }
        if (first >= 0 && last >= first) {
            int64_t start = 0;
            for (size_t id = 0; id < col->buf_size.size(); id++) {
                const int64_t size = col->buf_size[id];
                int64_t lo = first - start, hi = last - start;
                lo = lo < 0 ? 0 : (lo > size ? size : lo);
                hi = hi < 0 ? 0 : (hi > size ? size : hi);
                if (hi < lo) hi = lo;
                col->buf_lo[id] = lo;
                col->buf_hi[id] = hi;
                start += size;
            }
        }
    }
    // pass 3: one device block for everything
    size_t total = 0;
    for (size_t i = 0; i < n_chunks; i++) {
        const strsim_view_chunk& ch = chunks[i];
        ChunkPlan& p = up->plans[i];
        p.views_off = total;
        total = align_up(total + 16 * (size_t)ch.length, 256);
        p.validity_off = total;
        total = align_up(total + p.validity_bytes, 256);
        p.table_off = total;
        total = align_up(total + 8 * (size_t)(ch.n_data_buffers > 0 ? ch.n_data_buffers : 1), 256);
    }
    col->buf_dev_off.assign(col->buf_size.size(), 0);
    for (size_t id = 0; id < col->buf_size.size(); id++) {
        const int64_t lo = col->buf_lo[id], hi = col->buf_hi[id];  // lo is a multiple of 256
        col->buf_dev_off[id] = (int64_t)total - lo;
        // 64 spare bytes: TMA spans are rounded to 16 B and word copies read a few bytes past
        total = align_up(total + (size_t)(hi - lo) + 64, 256);
        col->data_bytes += hi - lo;
    }
    total += 256;
    int rc = pool_alloc(ctx.device, total, &col->block);
    if (rc) {
        delete col;
        up->col = nullptr;
        return rc;
    }
    col->block_bytes = total;
    char* base = static_cast<char*>(col->block);
    up->table_pos.assign(n_chunks, 0);
    for (size_t i = 0; i < n_chunks; i++) {
        const strsim_view_chunk& ch = chunks[i];
        const ChunkPlan& p = up->plans[i];
        DevChunk dc;
        dc.length = ch.length;
        dc.views = reinterpret_cast<const uint4*>(base + p.views_off);
        dc.validity = p.validity_bytes ? reinterpret_cast<const uint8_t*>(base + p.validity_off) : nullptr;
        dc.vbit = p.validity_bytes ? (ch.offset & 7) : 0;
        dc.bufs = reinterpret_cast<const unsigned long long*>(base + p.table_off);
        dc.data_bytes = 0;
        up->table_pos[i] = up->tables.size();
        if (ch.n_data_buffers == 0) up->tables.push_back(0ull);
        for (int64_t b = 0; b < ch.n_data_buffers; b++)
            up->tables.push_back(reinterpret_cast<unsigned long long>(base) +
                                 (unsigned long long)col->buf_dev_off[(size_t)col->chunk_buf_ids[i][(size_t)b]]);
        col->chunks.push_back(dc);
    }
    return STRSIM_OK;
}

static int upload_tables(Uploader& up, cudaStream_t st) {
    char* base = static_cast<char*>(up.col->block);
    for (size_t i = 0; i < up.n_chunks; i++) {
        const strsim_view_chunk& ch = up.chunks[i];
        const ChunkPlan& p = up.plans[i];
        const size_t n_tab = ch.n_data_buffers > 0 ? (size_t)ch.n_data_buffers : 1;
        CUDA_TRY(cudaMemcpyAsync(base + p.table_off, up.tables.data() + up.table_pos[i], 8 * n_tab,
                                 cudaMemcpyHostToDevice, st));
    }
    return STRSIM_OK;
}

static int64_t total_data_bytes(const strsim_b200_column* col) {
    int64_t t = 0;
    for (int64_t b : col->buf_size) t += b;
    return t;
}

// Uploads the data bytes of the linear range [up.uploaded, to) -- positions run over the column's
// distinct data buffers in upload order -- with their byte statistics, and advances up.uploaded.
// Range ends inside a buffer are multiples of 256 (see frontier_after), so every piece starts aligned.
static int upload_data_range(ThreadCtx& ctx, Uploader& up, int64_t from, int64_t to, cudaStream_t st,
                             ColumnStats* d_stats, bool do_copy, bool do_stats) {
    char* base = static_cast<char*>(up.col->block);
    int64_t start = 0;
    for (size_t id = 0; id < up.col->buf_size.size(); id++) {
        const int64_t size = up.col->buf_size[id];
        int64_t lo = from > start ? from - start : 0;
        int64_t hi = to - start < size ? to - start : size;
        if (lo < up.col->buf_lo[id]) lo = up.col->buf_lo[id];  // a row shard stores only [buf_lo, buf_hi)
        if (hi > up.col->buf_hi[id]) hi = up.col->buf_hi[id];
        if (hi > lo) {
            if (do_copy)
                CUDA_TRY(upload_copy(ctx, base + up.col->buf_dev_off[id] + lo, static_cast<const char*>(up.buf_src[id]) + lo,
                                     (size_t)(hi - lo), st));
            if (do_stats) {
                if (do_copy) CUDA_TRY(upload_flush(ctx));  // the kernel below must be queued behind every DMA of the copy
                long long blocks = (((hi - lo) >> 4) + 255) / 256;
                if (blocks > 148 * 16) blocks = 148 * 16;
                if (blocks < 1) blocks = 1;
                stats_bytes_kernel<<<(unsigned)blocks, 256, 0, st>>>(
                    reinterpret_cast<const unsigned char*>(base + up.col->buf_dev_off[id] + lo), hi - lo, d_stats);
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
        }
        start += size;
    }
    CUDA_TRY(cudaGetLastError());
    return STRSIM_OK;
}
static int upload_data_until(ThreadCtx& ctx, Uploader& up, int64_t to, cudaStream_t st, ColumnStats* d_stats) {
    if (to <= up.uploaded) return STRSIM_OK;
    const int rc = upload_data_range(ctx, up, up.uploaded, to, st, d_stats, true, true);
    up.uploaded = to;
    return rc;
}

// buffer tables + every distinct data buffer (+ byte statistics of the buffers)
static int upload_data(ThreadCtx& ctx, Uploader& up, cudaStream_t st, ColumnStats* d_stats) {
    int rc = upload_tables(up, st);
    if (rc) return rc;
    return upload_data_until(ctx, up, total_data_bytes(up.col), st, d_stats);
}

// ---- progressive upload (host calls) -----------------------------------------------------------------
// A freshly built column references its data buffers in increasing order, so the rows of slice s need
// only a PREFIX of the data.  Instead of sending all data before the first kernel can start, the host
// call sends, slice by slice, the data prefix that slice needs followed by its views; the kernels
// check every out-of-line view against the residency frontier (DevCol::res_buf / res_off) and count
// the rows they had to leave out, and a slice with such rows is recomputed once everything has
// arrived -- so any layout stays correct, and sequential layouts start computing (and downloading
// results) after the first slice instead of after the whole data upload.

// linear position of the end of view v of chunk c (-1: not an out-of-line view / malformed)
static int64_t view_end_position(const Uploader& up, size_t c, const int32_t* v) {
    const int32_t len = v[0];
    if (len <= 12) return -1;
    const auto& ids = up.col->chunk_buf_ids[c];
    const int32_t bi = v[2], off = v[3];
    if (bi < 0 || (size_t)bi >= ids.size() || off < 0) return -1;
    int64_t start = 0;
    for (int id = 0; id < ids[(size_t)bi]; id++) start += up.col->buf_size[(size_t)id];
    int64_t end = (int64_t)off + len;
    const int64_t size = up.col->buf_size[(size_t)ids[(size_t)bi]];
    end = (end + 255) & ~255ll;
    if (end > size) end = size;
    return start + end;
}

// data prefix needed by the rows before `hi`: end of the last out-of-line string, looking back at most
// 4096 rows from row hi-1; -1 when there is none that close (the caller then sends everything left:
// a column with so few out-of-line strings has little data to send)
static int64_t frontier_after(const Uploader& up, int64_t hi) {
    int64_t left = 4096;
    for (size_t c = up.n_chunks; c-- > 0 && left > 0;) {
        const strsim_view_chunk& ch = up.chunks[c];
        const int64_t row0 = up.plans[c].row0;
        if (row0 >= hi || ch.length == 0) continue;
        const int64_t last = hi - row0 < ch.length ? hi - row0 : ch.length;  // exclusive, chunk-relative
        const int32_t* v = static_cast<const int32_t*>(ch.views) + 4 * ch.offset;
        for (int64_t r = last; r-- > 0 && left > 0; left--) {
            const int64_t e = view_end_position(up, c, v + 4 * r);
            if (e >= 0) return e;
        }
    }
    return -1;
}

// sampled host-side test: do the views reference the data in increasing order?
static bool looks_sequential(const Uploader& up) {
    int64_t prev = -1;
    const int64_t n = up.col->length;
    for (int k = 0; k < 64; k++) {
        const int64_t row = n * k / 64;
        // the chunk holding `row`
        size_t c = 0;
        while (c + 1 < up.n_chunks && up.plans[c + 1].row0 <= row) c++;
        const strsim_view_chunk& ch = up.chunks[c];
        const int32_t* v = static_cast<const int32_t*>(ch.views) + 4 * ch.offset;
        int64_t r = row - up.plans[c].row0;
        const int64_t stop = r + 64 < ch.length ? r + 64 : ch.length;
        for (; r < stop; r++) {
            const int64_t e = view_end_position(up, c, v + 4 * r);
            if (e < 0) continue;
            if (e + 256 < prev) return false;
            prev = e;
            break;
        }
    }
    return true;
}

static void set_resident(strsim_b200_column* col, int64_t uploaded) {
    int64_t start = 0;
    for (size_t id = 0; id < col->buf_size.size(); id++) {
        const int64_t size = col->buf_size[id];
        int64_t r = uploaded - start;
        r = r < 0 ? 0 : (r > size ? size : r);
        col->buf_resident[id] = r > col->buf_hi[id] ? col->buf_hi[id] : r;  // bytes above buf_hi never arrive
        start += size;
    }
}

// residency frontiers of one chunk for the kernels (DevCol::res_* upper, DevCol::lo_* lower)
static void chunk_frontier(const strsim_b200_column* col, size_t chunk, DevCol* dc) {
    dc->res_buf = 0xFFFFFFFFu;
    dc->res_off = 0;
    dc->lo_buf = 0;
    dc->lo_off = 0;
    if (chunk >= col->chunk_buf_ids.size()) return;
    const auto& ids = col->chunk_buf_ids[chunk];
    for (size_t b = 0; b < ids.size(); b++) {
        const size_t id = (size_t)ids[b];
        if (col->buf_resident[id] < col->buf_size[id]) {
            dc->res_buf = (unsigned)b;
            dc->res_off = (unsigned)col->buf_resident[id];
            break;
        }
    }
    for (size_t b = ids.size(); b-- > 0;) {
        const size_t id = (size_t)ids[b];
        if (col->buf_lo[id] > 0) {
            dc->lo_buf = (unsigned)b;
            dc->lo_off = (unsigned)col->buf_lo[id];
            break;
        }
    }
}

// rows [lo, hi) of the column (a scalar column of length 1 is copied whenever lo == 0)
static int upload_rows(ThreadCtx& ctx, Uploader& up, int64_t lo, int64_t hi, cudaStream_t st, ColumnStats* d_stats,
                       bool do_copy = true, bool do_stats = true) {
    (void)ctx;
    char* base = static_cast<char*>(up.col->block);
    for (size_t i = 0; i < up.n_chunks; i++) {
        const strsim_view_chunk& ch = up.chunks[i];
        const ChunkPlan& p = up.plans[i];
        const int64_t c_lo = lo > p.row0 ? lo - p.row0 : 0;
        const int64_t c_hi = hi - p.row0 < ch.length ? hi - p.row0 : ch.length;
        if (c_lo >= c_hi) continue;
        if (do_copy)
            CUDA_TRY(upload_copy(ctx, base + p.views_off + 16 * c_lo,
                                 static_cast<const char*>(ch.views) + 16 * (ch.offset + c_lo), 16 * (size_t)(c_hi - c_lo), st));
        if (do_copy && p.validity_bytes) {
            const int64_t b_lo = ((ch.offset + c_lo) >> 3) - p.first_byte;
            const int64_t b_hi = ((ch.offset + c_hi + 7) >> 3) - p.first_byte;
            CUDA_TRY(cudaMemcpyAsync(base + p.validity_off + b_lo, ch.validity + p.first_byte + b_lo,
                                     (size_t)(b_hi - b_lo), cudaMemcpyHostToDevice, st));
        }
        if (!do_stats) continue;
        if (do_copy) CUDA_TRY(upload_flush(ctx));  // the kernel below must be queued behind every DMA of the copy
        long long blocks = (c_hi - c_lo + 4 * STATS_BLOCK - 1) / (4 * STATS_BLOCK);
        if (blocks > 148 * 8) blocks = 148 * 8;
        stats_views_kernel<<<(unsigned)blocks, STATS_BLOCK, 0, st>>>(
            reinterpret_cast<const uint4*>(base + p.views_off) + c_lo, c_hi - c_lo, d_stats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    CUDA_TRY(cudaGetLastError());
    return STRSIM_OK;
}

// whole column in one go (device-resident API): synchronises and folds the statistics
static int upload_column(ThreadCtx& ctx, const strsim_view_chunk* chunks, size_t n_chunks,
                         bool want_alg_bytes, strsim_b200_column** out) {
    Uploader up;
    int rc = upload_plan(ctx, chunks, n_chunks, &up);
    if (rc) return rc;
    stats_init_value(ctx.h_stats);
    cudaError_t e = cudaMemcpyAsync(ctx.d_stats, ctx.h_stats, sizeof(ColumnStats), cudaMemcpyHostToDevice, ctx.stream);
    if (e == cudaSuccess) {
        rc = upload_data(ctx, up, ctx.stream, ctx.d_stats);
        if (rc == STRSIM_OK) rc = upload_rows(ctx, up, 0, up.col->length > 0 ? up.col->length : 1, ctx.stream, ctx.d_stats);
    }
    if (e == cudaSuccess && rc == STRSIM_OK)
        e = publish_to_host(ctx.h_stats, ctx.d_stats, sizeof(ColumnStats), ctx.stream);
    if (e == cudaSuccess && rc == STRSIM_OK) e = cudaStreamSynchronize(ctx.stream);
    if (e != cudaSuccess || rc != STRSIM_OK) {
        if (e != cudaSuccess) {
            strsim_set_error("CUDA error during column upload: %s", cudaGetErrorString(e));
            rc = STRSIM_ERR_CUDA;
        }
        upload_flush(ctx);
        cudaStreamSynchronize(ctx.stream);
        strsim_b200_column_free(up.col);
        return rc;
    }
    strsim_b200_column* col = up.col;
    stats_fold(*ctx.h_stats, &col->or_byte, &col->and_byte);
    col->max_len = ctx.h_stats->max_len;
    col->max_block_pad = ctx.h_stats->max_block_pad;
    if (want_alg_bytes) {
        // 16 B of view per row + out-of-line payload (byte length > 12) + validity bits
        int64_t bytes = 0;
        for (size_t i = 0; i < n_chunks; i++) {
            const strsim_view_chunk& ch = chunks[i];
            const int32_t* v = static_cast<const int32_t*>(ch.views) + 4 * ch.offset;
            for (int64_t r = 0; r < ch.length; r++) {
                int32_t len = v[4 * r];
                bytes += 16 + (len > 12 ? len : 0);
            }
            if (ch.validity) bytes += (ch.length + 7) / 8;
        }
        col->alg_bytes = bytes;
    }
    *out = col;
    return STRSIM_OK;
}

// ---- kernel launch helpers -----------------------------------------------------------------------------
// raises the dynamic shared memory limit of `kern` on `device` to `bytes` (capped at what a CTA may opt
// into) the first time the pair (device, kernel) is seen; process-wide, safe from any host thread
static cudaError_t configure_smem(const void* kern, int device, size_t bytes) {
    static std::mutex& m = *new std::mutex();
    static std::vector<std::pair<const void*, int>>& done = *new std::vector<std::pair<const void*, int>>();
    std::lock_guard<std::mutex> lock(m);
    for (const auto& d : done)
        if (d.first == kern && d.second == device) return cudaSuccess;
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    if (bytes > (size_t)optin) bytes = (size_t)optin;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done.emplace_back(kern, device);
    return e;
}

template <class M, int MEASURE, int TPB, int RPT, bool GATHER, int T, bool ASCII_ONLY, bool REG = false,
          bool UREG = false, bool ULAT = false>
static int launch_short(ThreadCtx& ctx, SegArgs args, long long n_upper, cudaStream_t st) {
    using L = ShortLayout<M, TPB, RPT, T, REG, UREG>;
    auto kern = short_kernel<M, MEASURE, TPB, RPT, GATHER, T, ASCII_ONLY, REG, UREG, ULAT>;
    if (args.stage_bytes < 0) {
        // -stage_bytes = mean out-of-line bytes per row (x16) of the heavier column: size the stage
        // area for this tile shape with 25 % headroom (rows that still do not fit take the long path)
        long long stage = (long long)(-args.stage_bytes) * L::TILE * 5 / (16 * 4) + 256;
        if (stage < 1024) stage = 1024;
        if (stage > (long long)L::CAP * L::TILE) stage = (long long)L::CAP * L::TILE;
        args.stage_bytes = (int)((stage + 15) & ~15ll);
    }
    if (args.stage_bytes > L::CAP * L::TILE) args.stage_bytes = L::CAP * L::TILE;  // nothing longer than CAP bytes is staged
    const size_t smem = L::bytes(args.stage_bytes);
    // cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (device, kernel) and is shared by every host
    // thread (Polars calls the plugin from several): it is raised ONCE per device to the most this
    // instantiation can ever ask for -- a per-thread cache of the last size set would let one thread lower
    // the limit under another thread's larger launch
    CUDA_TRY(configure_smem(reinterpret_cast<const void*>(kern), ctx.device, L::bytes(L::CAP * L::TILE)));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
    if (per_sm < 1) {
        strsim_set_error("short kernel does not fit an SM (smem %zu)", smem);
        return STRSIM_ERR_CUDA;
    }
    const long long tiles = (n_upper + L::TILE - 1) / L::TILE;
    long long grid = (long long)per_sm * ctx.sm_count;
    if (tiles < grid) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TPB, smem, st>>>(args);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return STRSIM_OK;
}

template <class M, int MEASURE, int TPB, int RPT, bool GATHER, int T, bool ASCII_ONLY>
static int launch_direct(ThreadCtx& ctx, const SegArgs& args, long long n_upper, cudaStream_t st) {
    using L = DirectLayout<M, TPB, RPT, T>;
    auto kern = direct_kernel<M, MEASURE, TPB, RPT, GATHER, T, ASCII_ONLY>;
    const size_t smem = L::bytes;
    CUDA_TRY(configure_smem(reinterpret_cast<const void*>(kern), ctx.device, smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
    if (per_sm < 1) {
        strsim_set_error("direct kernel does not fit an SM (smem %zu)", smem);
        return STRSIM_ERR_CUDA;
    }
    const long long tiles = (n_upper + L::TILE - 1) / L::TILE;
    long long grid = (long long)per_sm * ctx.sm_count;
    if (tiles < grid) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TPB, smem, st>>>(args);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return STRSIM_OK;
}

// A general column (some byte >= 0x80) is served by two launches over the same rows: the pairs without
// a character above U+00FF by the 8-plane path on transcoded bytes (ULAT), the others by the
// register-compare path; the first launch also settles null rows and fills the overflow lists.
static thread_local float* g_wide_share = nullptr;  // the current column pair's hint (compute_on_device)

template <int MEASURE, int RPT_LATIN, int RPT_WIDE>
static int launch_general(ThreadCtx& ctx, const SegArgs& args, long long rows, cudaStream_t st) {
    SegArgs a = args;
    a.skip_latin = 0;
    // The first launch (ULAT) reads every row, computes the Latin-1 pairs and LISTS the others; the second
    // one (UREG) gathers the listed rows only.  Measured per 50 M rows x 5 measures on C3 (30 % of the rows
    // CJK): 8.9 ms against 10.2 ms for the register-compare kernel alone (round 1); L1 (no wide pairs): 1.2 ms
    // vs 2.2 ms per 10 M rows.  When most pairs are wide the first launch is a wasted pass: once a segment of
    // this column pair has shown more than half of its pairs wide, the following segments / slices / calls
    // go to the register-compare kernel alone.
    const bool mostly_latin = !(g_wide_share && *g_wide_share > 0.5f);
    if (!mostly_latin)
        return launch_short<uint32_t, MEASURE, 256, RPT_WIDE, false, 128, false, false, true>(ctx, a, rows, st);
    int rc = launch_short<uint32_t, MEASURE, 256, RPT_LATIN, false, 128, false, false, true, true>(ctx, a, rows, st);
    if (rc) return rc;
    // the pairs the first launch listed: gather mode over that list only (null rows and the overflow lists
    // were settled by the first launch: skip_latin).  The list's length stays on the device -- the second
    // launch reads it there and its CTAs leave at once when the list is empty -- so the two launches go out
    // back to back, without the host round trip that used to sit between them; the share of wide pairs
    // (the hint above) is read with the segment's overflow counters (note_wide_share).
    a.skip_latin = 1;
    a.ctr = 1;
    a.list = args.listwide;
    a.list_count = &ctx.d_ovf->nwide;
    a.n = rows;
    return launch_short<uint32_t, MEASURE, 256, RPT_WIDE, true, 128, false, false, true>(ctx, a, rows, st);
}

static void note_wide_share(const Overflow& ov, long long rows) {
    if (g_wide_share && rows > 0 && ov.nwide > 0) *g_wide_share = (float)ov.nwide / (float)rows;
}

// Tile shapes (256 threads x rows per thread).  Since the tiles are handed out through a counter, smaller tiles
// cost no balance, and their smaller stage areas let one more CTA share an SM: measured on one box, 3 / 4 / 3
// rows per thread -> 2 / 3 / 2: C2 fused 0.535 -> 0.515 ms, single measures 0.353 -> 0.336 (Levenshtein),
// L1 1.07 -> 0.96 ms, C3 6.30 -> 5.95 ms per 50 M rows, M1 2.96 -> 2.87 ms.
constexpr int FUSED_RPT = 2;   // fused ASCII launches
constexpr int SINGLE_RPT = 3;  // single-measure ASCII launches
constexpr int LATIN_RPT = 2;   // the fused Latin-1 launch over a general column
constexpr int M64_TPB = 128, M64_RPT = 2;  // the 64-bit plane launch over list64 (M1: 128 x 3 -> 128 x 2: 2.87 -> 2.54 ms;
                                           // 128 x 1: 2.67, 256 x 1: 2.57, 256 x 2: 2.74, 64 x 2: 2.72, 128 x 4: 3.41)

// Which instantiation of the fused kernel serves a segment (see DevStore): decided from the union of
// the two columns' byte statistics.
enum Alphabet { ALPHA_GENERAL = 0, ALPHA_ASCII128 = 1, ALPHA_ASCII64 = 2, ALPHA_ASCII32 = 3 };

static Alphabet classify_alphabet(unsigned o, unsigned n) {  // OR / AND over every byte of both columns
    Alphabet al;
    if (o & 0x80u)
        al = ALPHA_GENERAL;
    else if (((o ^ n) & 0x60u) == 0)
        al = ALPHA_ASCII32;  // every byte in one aligned block of 32 code points
    else if (((o ^ n) & 0x40u) == 0)
        al = ALPHA_ASCII64;
    else
        al = ALPHA_ASCII128;
    return al;
}

// One measure over a segment: the plane path for ASCII-only columns (tiles of 256 x SINGLE_RPT rows), the two
// launches of launch_general otherwise.  Variants that were measured and dropped -- a kernel without
// shared-memory staging (0.70 vs 0.58 ms: the sorted threads' scattered global loads cost 32 L1TEX
// wavefronts per warp instruction), the per-thread table path for ASCII columns (0.58 vs 0.48 ms), other
// tile shapes -- are in the git history of round 1, not in the build.
template <int MEASURE>
static int launch_fused(ThreadCtx& ctx, Alphabet al, const SegArgs& args, long long rows, cudaStream_t st) {
    switch (al) {
        case ALPHA_ASCII32:
            return launch_short<uint32_t, MEASURE, 256, SINGLE_RPT, false, 32, true, true>(ctx, args, rows, st);
        case ALPHA_ASCII64:
            return launch_short<uint32_t, MEASURE, 256, SINGLE_RPT, false, 64, true, true>(ctx, args, rows, st);
        case ALPHA_ASCII128:
            return launch_short<uint32_t, MEASURE, 256, SINGLE_RPT, false, 128, true, true>(ctx, args, rows, st);
        default:
            // any script: Latin-1 pairs by the plane path, the others by the register-compare path
            // (row_unicode_reg.cuh); no table in shared memory
            return launch_general<MEASURE, 4, 4>(ctx, args, rows, st);
    }
}

// fused evaluation of several measures in one pass (pair_algos.cuh: FusedStep): register paths only
template <int GROUPS>
static int launch_multi(ThreadCtx& ctx, Alphabet al, const SegArgs& args, long long rows, cudaStream_t st) {
    constexpr int ME = MULTI_BASE + GROUPS;
    // tiles of 256 x FUSED_RPT rows: 33 KB of shared memory per CTA, five CTAs (40 warps) share an SM (see the
    // tile shapes above; with a fixed stride of tiles per CTA, in round 1, 256 x 3 was the best shape)
    switch (al) {
        case ALPHA_ASCII32:
            return launch_short<uint32_t, ME, 256, FUSED_RPT, false, 32, true, true>(ctx, args, rows, st);
        case ALPHA_ASCII64:
            return launch_short<uint32_t, ME, 256, FUSED_RPT, false, 64, true, true>(ctx, args, rows, st);
        case ALPHA_ASCII128:
            return launch_short<uint32_t, ME, 256, FUSED_RPT, false, 128, true, true>(ctx, args, rows, st);
        default:
            // register-compare path: 256 x 2 measured best on C3 (4.82 ms vs 5.11 ms per 10M rows x 5 measures)
            return launch_general<ME, LATIN_RPT, 2>(ctx, args, rows, st);
    }
}

// rows on the long list -> fallback kernel, scratch slabs sized from the device-side maxima
template <int MEASURE>
static int run_generic(ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, cudaStream_t st,
                       const unsigned int* list = nullptr, const unsigned int* list_count = nullptr,
                       unsigned int n_rows = 0) {
    GenericArgs g{};
    g.a = args.a;
    g.b = args.b;
    g.out = args.out;
    g.dbg = args.dbg;
    g.list = list ? list : args.listlong;
    g.list_count = list_count ? list_count : &ctx.d_ovf->nlong;
    if (!list) n_rows = ov.nlong;
    g.cap_a = (int)ov.max_bytes_a + 1;
    g.cap_b = (int)ov.max_bytes_b + 1;
    long long work = 2ll * (g.cap_b + 1);
    const long long flags = ((long long)g.cap_a + g.cap_b) / 4 + 2;
    if (flags > work) work = flags;
    g.slab_words = (long long)g.cap_a + g.cap_b + work;
    long long slots = (long long)ctx.sm_count * 512;
    if ((long long)n_rows < slots) slots = n_rows;
    const long long budget = 2ll << 30;  // bytes of scratch at most
    if (slots * g.slab_words * 4 > budget) slots = budget / (g.slab_words * 4);
    if (slots < 1) {
        strsim_set_error("a row of %u/%u bytes needs more than the 2 GiB fallback scratch",
                         ov.max_bytes_a, ov.max_bytes_b);
        return STRSIM_ERR_NOMEM;
    }
    int rc = ws_reserve(ctx.scratch, (size_t)(slots * g.slab_words * 4));
    if (rc) return rc;
    g.scratch = static_cast<uint32_t*>(ctx.scratch.ptr);
    g.n_slots = (int)slots;
    generic_kernel<MEASURE><<<(unsigned)((slots + 63) / 64), 64, 0, st>>>(g);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return STRSIM_OK;
}

// long Levenshtein rows -> warp-cooperative multi-word Myers.  Two tiers: first many warps with
// modest Peq slabs (typical pairs need (distinct+1) x W words, far below the worst case), then the
// pairs that did not fit with worst-case slabs and fewer warps.  Rows whose pattern exceeds
// LONG_PAT_MAX come back on `huge_list` and are finished by the generic kernel.
struct Leftover {  // rows the long kernel could not take: finished by the generic kernel
    const unsigned int* list = nullptr;
    const unsigned int* count = nullptr;  // device-resident
    unsigned int n = 0;
};

static int run_long_lev(ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, unsigned int* spare_list,
                        Leftover* left, cudaStream_t st) {
    LongLevArgs g{};
    g.a = args.a;
    g.b = args.b;
    g.out = args.out;
    g.dbg = args.dbg;
    g.cursor = ctx.d_counters;
    g.cap_a = (int)ov.max_bytes_a + 4;
    g.cap_b = (int)ov.max_bytes_b + 4;
    int cap_pat = g.cap_a < g.cap_b ? g.cap_a : g.cap_b;
    if (cap_pat > LONG_PAT_MAX) cap_pat = LONG_PAT_MAX;
    g.cap_pat = cap_pat;
    int hs = 64;
    while (hs < 2 * cap_pat) hs <<= 1;
    g.hash_size = hs;
    g.w_max = (cap_pat + 63) / 64;
    const long long worst_words = (long long)(cap_pat + 1) * ((g.w_max + 1) & ~1);  // rows have an even stride (long_load_eq)
    const long long budget = 8ll << 30;
    // tier 0 reads listlong and defers into spare_list (count d_counters[1]);
    // tier 1 reads spare_list and defers into listlong, free again by then (count d_counters[2])
    Leftover src;
    src.list = args.listlong;
    src.count = &ctx.d_ovf->nlong;
    src.n = ov.nlong;
    CUDA_TRY(cudaMemsetAsync(ctx.d_counters, 0, 4 * sizeof(unsigned int), st));
    for (int tier = 0; tier < 2 && src.n > 0; tier++) {
        long long peq_words = tier == 0 ? (64ll << 10) : worst_words;  // tier 0: 512 KiB of Peq per slab
        if (peq_words > worst_words) peq_words = worst_words;
        g.peq_words = peq_words;
        g.slab_bytes = long_lev_slab_bytes(g.cap_a, g.cap_b, g.cap_pat, g.hash_size, peq_words);
        constexpr int tier0_warps = 32;  // resident warps per SM (24 / 16 measured 4 % / 20 % slower on C4)
        // a warp works on two pairs at a time (long_lev_kernel.cuh): two slabs per warp
        long long warps = (long long)ctx.sm_count * (tier == 0 ? tier0_warps : 8);
        if (((long long)src.n + 1) / 2 < warps) warps = ((long long)src.n + 1) / 2;
        if (2 * warps * g.slab_bytes > budget) warps = budget / (2 * g.slab_bytes);
        if (warps < 1) break;  // not even two slabs fit the budget: the generic kernel takes `src`
        // the list sorted by cost (longest first, neighbours alike) lives in front of the slabs
        const size_t hist_bytes = ((size_t)LONG_KEYS * sizeof(unsigned int) + 255) & ~(size_t)255;
        const size_t sort_bytes = hist_bytes + (((size_t)src.n * sizeof(unsigned int) + 255) & ~(size_t)255);
        int rc = ws_reserve(ctx.scratch, sort_bytes + (size_t)(2 * warps * g.slab_bytes));
        if (rc) return rc;
        unsigned char* base = static_cast<unsigned char*>(ctx.scratch.ptr);
        unsigned int* hist = reinterpret_cast<unsigned int*>(base);
        unsigned int* sorted = reinterpret_cast<unsigned int*>(base + hist_bytes);
        {
            CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)LONG_KEYS * sizeof(unsigned int), st));
            unsigned int blocks = (unsigned int)((src.n + 255) / 256);
            if (blocks > 4u * (unsigned int)ctx.sm_count) blocks = 4u * (unsigned int)ctx.sm_count;
            long_sort_hist_kernel<<<blocks, 256, 0, st>>>(g.a, g.b, src.list, src.count, hist);
            long_sort_scan_kernel<<<1, 1024, 0, st>>>(hist);
            long_sort_scatter_kernel<<<blocks, 256, 0, st>>>(g.a, g.b, src.list, src.count, hist, sorted);
            g_launches.fetch_add(3, std::memory_order_relaxed);
            CUDA_TRY(cudaGetLastError());
        }
        g.scratch = base + sort_bytes;
        g.n_warps = (int)warps;
        g.list = sorted;
        g.list_count = src.count;
        g.huge_list = tier == 0 ? spare_list : args.listlong;
        g.huge_count = ctx.d_counters + 1 + tier;
        CUDA_TRY(cudaMemsetAsync(ctx.d_counters, 0, sizeof(unsigned int), st));  // the cursor
        long_lev_kernel<<<(unsigned)((warps + LONG_WPB - 1) / LONG_WPB), 32 * LONG_WPB, 0, st>>>(g);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(publish_to_host(ctx.h_counters, ctx.d_counters, 4 * sizeof(unsigned int), st));
        CUDA_TRY(cudaStreamSynchronize(st));
        src.list = g.huge_list;
        src.count = g.huge_count;
        src.n = ctx.h_counters[1 + tier];
    }
    *left = src;
    return STRSIM_OK;
}

// long rows of Jaro / Jaro-Winkler / Jaccard / Sorensen-Dice -> one pair per warp (long_pair_kernel.cuh), every
// wanted measure in ONE launch.  `outs` / `dbgs` by measure id (nullptr = not wanted).  Returns
// STRSIM_ERR_NOMEM-free: *served = false when a slab for the longest row does not fit the scratch budget
// (the caller then falls back to the one-thread-per-pair kernel).
static int run_long_pair(ThreadCtx& ctx, const SegArgs& args, double* const outs[5], int* const dbgs[5], const Overflow& ov,
                         cudaStream_t st, bool* served) {
    *served = false;
    LongPairArgs g{};
    g.a = args.a;
    g.b = args.b;
    bool any = false;
    for (int m = 1; m < 5; m++) {
        g.outs[m] = outs[m];
        g.dbgs[m] = dbgs[m];
        any = any || outs[m] != nullptr;
    }
    if (!any) {
        *served = true;
        return STRSIM_OK;
    }
    g.list = args.listlong;
    g.list_count = &ctx.d_ovf->nlong;
    g.cursor = ctx.d_counters + 3;
    g.cap_a = (int)((ov.max_bytes_a + 32u) & ~31u);
    g.cap_b = (int)((ov.max_bytes_b + 32u) & ~31u);
    int hs = 64;
    while (hs < 2 * g.cap_b) hs <<= 1;
    g.hash_size = hs;
    g.slab_bytes = long_pair_slab_bytes(g.cap_a, g.cap_b, g.hash_size);
    long long warps = (long long)ctx.sm_count * 32;
    if ((long long)ov.nlong < warps) warps = ov.nlong;
    const long long budget = 2ll << 30;
    if (warps * g.slab_bytes > budget) warps = budget / g.slab_bytes;
    if (warps < 1) return STRSIM_OK;  // not served
    int rc = ws_reserve(ctx.scratch, (size_t)(warps * g.slab_bytes));
    if (rc) return rc;
    g.scratch = static_cast<unsigned char*>(ctx.scratch.ptr);
    g.n_warps = (int)warps;
    CUDA_TRY(cudaMemsetAsync(ctx.d_counters + 3, 0, sizeof(unsigned int), st));
    long_pair_kernel<<<(unsigned)((warps + LONGP_WPB - 1) / LONGP_WPB), 32 * LONGP_WPB, 0, st>>>(g);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    *served = true;
    return STRSIM_OK;
}

// test hook: STRSIM_B200_WIDE_ROWS=0 sends the 65..320-byte rows of ASCII columns to the warp-per-pair kernels,
// as before finish_wide existed (the parity tests run both ways)
static bool wide_rows_enabled() {
    static const bool v = !(getenv("STRSIM_B200_WIDE_ROWS") != nullptr && atoi(getenv("STRSIM_B200_WIDE_ROWS")) == 0);
    return v;
}

static bool force_generic_rows() {
    static const bool v = getenv("STRSIM_B200_FORCE_GENERIC") != nullptr && atoi(getenv("STRSIM_B200_FORCE_GENERIC")) != 0;
    return v;
}

// overflow lists (worst case every row) and counters for one segment
static int prepare_segment(ThreadCtx& ctx, SegArgs& args, int stage32, int64_t seg_rows, cudaStream_t st) {
    int rc = ws_reserve(ctx.lists, 3 * sizeof(unsigned int) * (size_t)seg_rows + 64);
    if (rc) return rc;
    args.list64 = static_cast<unsigned int*>(ctx.lists.ptr);
    args.listlong = args.list64 + seg_rows;
    args.listwide = args.listlong + seg_rows;
    args.ovf = ctx.d_ovf;
    CUDA_TRY(cudaMemsetAsync(ctx.d_ovf, 0, sizeof(Overflow), st));
    args.stage_bytes = stage32;
    args.list = nullptr;
    args.list_count = nullptr;
    return STRSIM_OK;
}

static int read_overflow(ThreadCtx& ctx, Overflow* ov, cudaStream_t st, bool first = true) {
    CUDA_TRY(publish_to_host(ctx.h_ovf, ctx.d_ovf, sizeof(Overflow), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *ov = *ctx.h_ovf;
    if (first) g_deferred += ov->ndefer;  // counted once, however often the counters are re-read
    return STRSIM_OK;
}

// rows of 33..64 bytes of ASCII-only columns: the plane path with 64-bit masks (two registers per plane),
// gather mode over list64; tiles of 128 x 2 rows with a worst-case stage area (every listed string is
// out of line).  MEASURE may be a fused set (MULTI_BASE + groups).  Medium ASCII strings (M1 workload,
// 20-60 characters, 73 % of the rows on this list): 22.8 -> see profiles/r1_bench_M1.json ms per 10 M rows.
template <int MEASURE>
static int finish_64_planes(ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, cudaStream_t st) {
    SegArgs a64 = args;
    a64.list = args.list64;
    a64.list_count = &ctx.d_ovf->n64;
    a64.n = ov.n64;
    a64.stage_bytes = 64 * M64_TPB * M64_RPT;
    a64.ctr = 2;
    return launch_short<uint64_t, MEASURE, M64_TPB, M64_RPT, true, 128, true, true>(ctx, a64, ov.n64, st);
}

// Rows of 65..320 bytes of ASCII-only columns: the plane path again, with masks of ten 32-bit words
// (wide_mask.cuh) -- one pair per thread, gather mode over listlong, tiles of 128 rows.  The launch reads
// listlong (count: the current counters) and appends what it cannot take -- a string above 320 bytes, a tile
// whose payload does not fit the stage area -- to the OTHER list buffer under the second set of counters;
// then the two sets and the two buffers swap roles, so that the warp-per-pair kernels behind it find "their"
// list where they always do.  `ov` is re-read: it describes what is left.
// Measured on T1 (10 M address-like rows, one in ten of 100-300 characters, all five measures): see
// DESIGN.md 3.3b.
constexpr int WIDE_WORDS = 10;
template <int MEASURE>
static int finish_wide(ThreadCtx& ctx, SegArgs& args, Overflow* ov, cudaStream_t st) {
    SegArgs aw = args;
    aw.list = args.listlong;
    aw.list_count = &ctx.d_ovf->nlong;
    aw.n = ov->nlong;
    aw.ovf = ctx.d_ovf_alt;
    aw.listlong = args.list64;  // free: the 64-bit launch has consumed it (same stream)
    aw.list64 = nullptr;        // never written by this instantiation (CAP != 32)
    CUDA_TRY(cudaMemsetAsync(ctx.d_ovf_alt, 0, sizeof(Overflow), st));
    // stage area: the listed strings are all out of line; 224 bytes a row on average fit 100-300-character
    // rows, longer tiles send their last rows on to the warp kernels
    aw.stage_bytes = 224 * 128;
    int rc = launch_short<Wide<WIDE_WORDS>, MEASURE, 128, 1, true, 128, true, true>(ctx, aw, ov->nlong, st);
    if (rc) return rc;
    std::swap(ctx.d_ovf, ctx.d_ovf_alt);
    std::swap(args.listlong, args.list64);
    args.ovf = ctx.d_ovf;
    return read_overflow(ctx, ov, st, false);
}

// rows of 33..64 bytes -> 64-bit instantiation in gather mode
template <int MEASURE>
static int finish_64(ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, cudaStream_t st, bool ascii = false) {
    if (ascii) return finish_64_planes<MEASURE>(ctx, args, ov, st);
    SegArgs a64 = args;
    a64.list = args.list64;
    a64.list_count = &ctx.d_ovf->n64;
    a64.n = ov.n64;
    // 64-row tiles: the list is short (C3: 10k of 6.7M rows), so many small CTAs finish it in one wave;
    // with 256-row tiles a dozen CTAs took 0.1 ms per measure, a quarter of the segment's time
    return launch_direct<uint64_t, MEASURE, 64, 1, true, 192, false>(ctx, a64, ov.n64, st);
}

// rows on the long list -> multi-word Myers (Levenshtein; consumes list64 and listlong) or the generic kernel
template <int MEASURE>
static int finish_long(ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, cudaStream_t st) {
    int rc;
    if (MEASURE == LEVENSHTEIN && !force_generic_rows()) {
        Leftover left;
        rc = run_long_lev(ctx, args, ov, args.list64, &left, st);  // list64 is free again here
        if (rc) return rc;
        if (left.n > 0) rc = run_generic<MEASURE>(ctx, args, ov, st, left.list, left.count, left.n);
    } else {
        bool served = false;
        if (!force_generic_rows()) {
            double* outs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            int* dbgs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            outs[MEASURE] = args.out;
            dbgs[MEASURE] = args.dbg;
            rc = run_long_pair(ctx, args, outs, dbgs, ov, st, &served);
            if (rc) return rc;
        }
        if (!served) rc = run_generic<MEASURE>(ctx, args, ov, st);
    }
    return rc;
}

template <int MEASURE>
static int run_segment(ThreadCtx& ctx, SegArgs args, Alphabet al, int stage32, int64_t seg_rows, cudaStream_t st,
                       bool proven_clean) {
    int rc = prepare_segment(ctx, args, stage32, seg_rows, st);
    if (rc) return rc;
    if (force_generic_rows()) {
        // test hook: every valid row goes straight to the fallback kernel
        list_all_kernel<<<(unsigned)((seg_rows + 255) / 256), 256, 0, st>>>(args);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    } else {
        rc = launch_fused<MEASURE>(ctx, al, args, seg_rows, st);
        if (rc) return rc;
        // the column statistics prove that no row leaves the launch (see compute_on_device): nothing to read back
        if (proven_clean) return STRSIM_OK;
    }
    Overflow ov;
    rc = read_overflow(ctx, &ov, st);
    if (rc) return rc;
    if (al == ALPHA_GENERAL) note_wide_share(ov, seg_rows);
    g_last_overflow[0] += ov.n64;
    if (ov.n64 > 0) {
        rc = finish_64<MEASURE>(ctx, args, ov, st, al != ALPHA_GENERAL);
        if (rc) return rc;
        // rows the 64-bit kernel could not take are appended to listlong; re-read the counters
        rc = read_overflow(ctx, &ov, st, false);
        if (rc) return rc;
    }
    g_last_overflow[1] += ov.nlong;
    if (ov.nlong > 0 && al != ALPHA_GENERAL && !force_generic_rows() && wide_rows_enabled()) {
        rc = finish_wide<MEASURE>(ctx, args, &ov, st);
        if (rc) return rc;
    }
    if (ov.nlong > 0) rc = finish_long<MEASURE>(ctx, args, ov, st);
    return rc;
}

static int finish_long_any(int measure, ThreadCtx& ctx, const SegArgs& args, const Overflow& ov, cudaStream_t st) {
    switch (measure) {
        case 0: return finish_long<0>(ctx, args, ov, st);
        case 1: return finish_long<1>(ctx, args, ov, st);
        case 2: return finish_long<2>(ctx, args, ov, st);
        case 3: return finish_long<3>(ctx, args, ov, st);
        default: return finish_long<4>(ctx, args, ov, st);
    }
}

// Several measures over one segment: ONE fused launch settles every row of at most 32 bytes for all
// wanted measures (args.outs / args.dbgs); the overflow lists it leaves are then finished measure by
// measure with the single-measure follow-up kernels.  Levenshtein goes last: its long-row kernel
// reuses both lists as scratch.
static int run_segment_multi(ThreadCtx& ctx, SegArgs args, int groups, Alphabet al, int stage32, int64_t seg_rows,
                             cudaStream_t st, bool proven_clean) {
    int rc = prepare_segment(ctx, args, stage32, seg_rows, st);
    if (rc) return rc;
    args.out = nullptr;
    args.dbg = nullptr;
    if (force_generic_rows()) {
        list_all_kernel<<<(unsigned)((seg_rows + 255) / 256), 256, 0, st>>>(args);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    } else {
        switch (groups) {
            case 2: rc = launch_multi<2>(ctx, al, args, seg_rows, st); break;
            case 3: rc = launch_multi<3>(ctx, al, args, seg_rows, st); break;
            case 4: rc = launch_multi<4>(ctx, al, args, seg_rows, st); break;
            case 5: rc = launch_multi<5>(ctx, al, args, seg_rows, st); break;
            case 6: rc = launch_multi<6>(ctx, al, args, seg_rows, st); break;
            case 7: rc = launch_multi<7>(ctx, al, args, seg_rows, st); break;
            default:
                strsim_set_error("fused evaluation: bad group mask %d", groups);
                rc = STRSIM_ERR_ARGUMENT;
        }
        if (rc) return rc;
        if (proven_clean) return STRSIM_OK;  // no row can have left the launch: nothing to read back
    }
    Overflow ov;
    rc = read_overflow(ctx, &ov, st);
    if (rc) return rc;
    if (al == ALPHA_GENERAL) note_wide_share(ov, seg_rows);
    g_last_overflow[0] += ov.n64;
    if (ov.n64 > 0 && al != ALPHA_GENERAL && groups >= 2) {
        // ASCII-only columns: ONE fused launch of the 64-bit plane path over list64
        switch (groups) {
            case 2: rc = finish_64_planes<MULTI_BASE + 2>(ctx, args, ov, st); break;
            case 3: rc = finish_64_planes<MULTI_BASE + 3>(ctx, args, ov, st); break;
            case 4: rc = finish_64_planes<MULTI_BASE + 4>(ctx, args, ov, st); break;
            case 5: rc = finish_64_planes<MULTI_BASE + 5>(ctx, args, ov, st); break;
            case 6: rc = finish_64_planes<MULTI_BASE + 6>(ctx, args, ov, st); break;
            default: rc = finish_64_planes<MULTI_BASE + 7>(ctx, args, ov, st); break;
        }
        if (rc) return rc;
    } else if (ov.n64 > 0) {
        // general columns: the table / hash path with 64-bit masks, every wanted measure in one launch
        // (blockIdx.y = measure)
        using L = DirectLayout<uint64_t, 64, 1, 192>;
        auto kern = direct_multi64_kernel<64, 192>;
        CUDA_TRY(configure_smem(reinterpret_cast<const void*>(kern), ctx.device, L::bytes));
        SegArgs a64 = args;
        a64.list = args.list64;
        a64.list_count = &ctx.d_ovf->n64;
        a64.n = ov.n64;
        long long gx = ((long long)ov.n64 + L::TILE - 1) / L::TILE;
        if (gx > 8ll * ctx.sm_count) gx = 8ll * ctx.sm_count;
        if (gx < 1) gx = 1;
        kern<<<dim3((unsigned)gx, 5), 64, L::bytes, st>>>(a64);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    }
    if (ov.n64 > 0) {  // rows the 64-bit kernels could not take are appended to listlong: re-read the counters
        rc = read_overflow(ctx, &ov, st, false);
        if (rc) return rc;
    }
    g_last_overflow[1] += ov.nlong;
    if (ov.nlong > 0 && al != ALPHA_GENERAL && groups >= 2 && !force_generic_rows() && wide_rows_enabled()) {
        switch (groups) {
            case 2: rc = finish_wide<MULTI_BASE + 2>(ctx, args, &ov, st); break;
            case 3: rc = finish_wide<MULTI_BASE + 3>(ctx, args, &ov, st); break;
            case 4: rc = finish_wide<MULTI_BASE + 4>(ctx, args, &ov, st); break;
            case 5: rc = finish_wide<MULTI_BASE + 5>(ctx, args, &ov, st); break;
            case 6: rc = finish_wide<MULTI_BASE + 6>(ctx, args, &ov, st); break;
            default: rc = finish_wide<MULTI_BASE + 7>(ctx, args, &ov, st); break;
        }
        if (rc) return rc;
    }
    if (ov.nlong > 0) {
        bool served = false;  // Jaro / Jaro-Winkler / Jaccard / Sorensen-Dice of the long rows: ONE warp-per-pair launch
        if (!force_generic_rows()) {
            rc = run_long_pair(ctx, args, args.outs, args.dbgs, ov, st, &served);
            if (rc) return rc;
        }
        for (int k = 0; k < 5; k++) {
            const int m = (k + 1) % 5;  // 1, 2, 3, 4, then Levenshtein
            if (!args.outs[m] || (served && m != 0)) continue;
            SegArgs am = args;
            am.out = args.outs[m];
            am.dbg = args.dbgs[m];
            rc = finish_long_any(m, ctx, am, ov, st);
            if (rc) return rc;
        }
    }
    return STRSIM_OK;
}

// Rows [row_lo, row_lo + n_rows) of measure(a, b) -> d_out[row] / validity bits / debug records, all
// indexed by the absolute row.  `reset`: zero the validity words of the range and (first slice only)
// the null counter.
//
// Several distinct measures (n_measures > 1) are evaluated by ONE fused launch per segment
// (run_segment_multi); d_outs[k] / d_dbgs[k] belong to measures[k].
static int compute_on_device(ThreadCtx& ctx, const int* measures, size_t n_measures, const strsim_b200_column* a,
                             const strsim_b200_column* b, int64_t row_lo, int64_t n_rows, Alphabet al,
                             double* const* d_outs, uint32_t* d_validity, int32_t* const* d_dbgs, cudaStream_t st) {
    int groups = 0;
    unsigned seen = 0;
    for (size_t k = 0; k < n_measures; k++) {
        const int m = measures[k];
        if (m < 0 || m > 4) {
            strsim_set_error("unknown measure %d", m);
            return STRSIM_ERR_ARGUMENT;
        }
        if (seen & (1u << m)) {
            strsim_set_error("measure %d requested twice in one fused call", m);
            return STRSIM_ERR_ARGUMENT;
        }
        seen |= 1u << m;
        groups |= group_of(m);
    }
    if (n_measures == 0) return STRSIM_OK;
    const int measure = measures[0];
    double* const d_out = d_outs[0];
    int32_t* const d_dbg = d_dbgs ? d_dbgs[0] : nullptr;
    const int64_t la = a->length, lb = b->length;
    if (la != lb && la != 1 && lb != 1) {
        strsim_set_error("Inputs must have the same length, or one of them must be a Utf8 literal.");
        return STRSIM_ERR_SHAPE;
    }
    const int64_t n_total = (la == 1) ? lb : la;
    if ((la == 1 && lb != 1 && a->scalar_null) || (lb == 1 && la != 1 && b->scalar_null)) return null_literal_error();
    if (n_total == 0 || n_rows <= 0) return STRSIM_OK;
    const int64_t n = row_lo + n_rows;  // exclusive end row
    const bool bc_a = la == 1 && n_total != 1, bc_b = lb == 1 && n_total != 1;
    const bool any_validity = a->has_validity || b->has_validity;
    if (d_validity) {
        // row_lo is a multiple of 32 for every slice but the caller's first (0)
        uint32_t* w0 = d_validity + (row_lo >> 5);
        const size_t words = (size_t)(((n + 31) >> 5) - (row_lo >> 5));
        CUDA_TRY(cudaMemsetAsync(w0, any_validity ? 0 : 0xFF, 4 * words, st));
        if (row_lo == 0) CUDA_TRY(cudaMemsetAsync(ctx.d_nulls, 0, sizeof(unsigned long long), st));
    }
    g_wide_share = &a->wide_share;
    // walk both chunk lists in lock step (polars-core align_chunks equivalent)
    size_t ia = 0, ib = 0;
    int64_t oa = 0, ob = 0, row = row_lo;
    if (!bc_a) {
        oa = row_lo;
        while (ia < a->chunks.size() && oa >= a->chunks[ia].length) oa -= a->chunks[ia++].length;
    }
    if (!bc_b) {
        ob = row_lo;
        while (ib < b->chunks.size() && ob >= b->chunks[ib].length) ob -= b->chunks[ib++].length;
    }
    while (row < n) {
        while (!bc_a && ia < a->chunks.size() && oa >= a->chunks[ia].length) {
            ia++;
            oa = 0;
        }
        while (!bc_b && ib < b->chunks.size() && ob >= b->chunks[ib].length) {
            ib++;
            ob = 0;
        }
        const DevChunk& ca = a->chunks[bc_a ? 0 : ia];
        const DevChunk& cb = b->chunks[bc_b ? 0 : ib];
        int64_t len = n - row;
        if (!bc_a && ca.length - oa < len) len = ca.length - oa;
        if (!bc_b && cb.length - ob < len) len = cb.length - ob;
        if (len > 0x7FFFFFF0ll) len = 0x7FFFFFF0ll;  // list entries are 32-bit
        SegArgs s{};
        s.a.views = ca.views + (bc_a ? 0 : oa);
        s.a.validity = ca.validity;
        s.a.vbit = ca.vbit + (bc_a ? 0 : oa);
        s.a.bufs = ca.bufs;
        s.a.stride = bc_a ? 0 : 1;
        chunk_frontier(a, bc_a ? 0 : ia, &s.a);
        chunk_frontier(b, bc_b ? 0 : ib, &s.b);
        s.b.views = cb.views + (bc_b ? 0 : ob);
        s.b.validity = cb.validity;
        s.b.vbit = cb.vbit + (bc_b ? 0 : ob);
        s.b.bufs = cb.bufs;
        s.b.stride = bc_b ? 0 : 1;
        s.n = len;
        s.out = d_out + row;
        s.dbg = d_dbg ? d_dbg + 6 * row : nullptr;
        // stage capacity is derived per tile shape in launch_short() from the mean out-of-line bytes
        // per row of the heavier column (passed as a negative fixed-point number)
        const double avg_a = bc_a ? 0.0 : (double)a->data_bytes / (double)(a->length > 0 ? a->length : 1);
        const double avg_b = bc_b ? 0.0 : (double)b->data_bytes / (double)(b->length > 0 ? b->length : 1);
        const double avg = avg_a > avg_b ? avg_a : avg_b;
        long long stage = -(long long)(avg * 16.0 + 1.0);
        if (stage < -64 * 16) stage = -64 * 16;
        // Can a row leave the short-string launch at all?  Not when (a) every byte is ASCII (one launch, the
        // plane path), (b) both columns are completely resident and their statistics carry the longest
        // string and the largest payload of an aligned 256-row block, (c) no string exceeds the 32-byte
        // masks, and (d) the stage area holds the payload of any tile -- a tile of 256 x RPT rows covers RPT
        // aligned blocks, one more when the segment does not start on a block boundary.  Then the stage
        // area is sized from that bound and the host neither reads the overflow counters back nor waits.
        bool proven_clean = false;
        static const bool no_proof = getenv("STRSIM_B200_READBACK") != nullptr && !strcmp(getenv("STRSIM_B200_READBACK"), "1");
        if (!no_proof && al != ALPHA_GENERAL && !force_generic_rows() && a->max_len <= 32u && b->max_len <= 32u &&
            a->max_block_pad != 0xFFFFFFFFu && b->max_block_pad != 0xFFFFFFFFu && s.a.res_buf == 0xFFFFFFFFu &&
            s.b.res_buf == 0xFFFFFFFFu && s.a.lo_off == 0u && s.b.lo_off == 0u) {
            const long long tile_blocks = n_measures > 1 ? FUSED_RPT : SINGLE_RPT;
            const long long blocks_a = tile_blocks + ((bc_a || oa % STATS_BLOCK == 0) ? 0 : 1);
            const long long blocks_b = tile_blocks + ((bc_b || ob % STATS_BLOCK == 0) ? 0 : 1);
            long long need = blocks_a * (long long)a->max_block_pad;
            if (blocks_b * (long long)b->max_block_pad > need) need = blocks_b * (long long)b->max_block_pad;
            need += 48;  // the TMA span is rounded to 16 bytes at both ends
            // what launch_short would size from the mean (25 % headroom); the proof may ask for a little more,
            // but not for so much more that a CTA less would fit an SM
            const long long by_mean = (long long)((avg < 64.0 ? avg : 64.0) * 16.0 + 1.0) * (tile_blocks * 256) * 5 / (16 * 4) + 256;
            if (need <= by_mean + by_mean / 8 && need <= 32ll * tile_blocks * 256) {
                // the bound itself, not the mean-based size: a column uploaded from a SLICE of a longer array
                // brings whole data buffers along, and their bytes per row say nothing about its rows
                stage = need < 1024 ? 1024 : need;
                stage = (stage + 15) & ~15ll;
                proven_clean = true;
            }
        }
        int rc;
        if (n_measures > 1) {
            for (size_t k = 0; k < n_measures; k++) {
                s.outs[measures[k]] = d_outs[k] + row;
                s.dbgs[measures[k]] = (d_dbgs && d_dbgs[k]) ? d_dbgs[k] + 6 * row : nullptr;
            }
            rc = run_segment_multi(ctx, s, groups, al, (int)stage, len, st, proven_clean);
        } else {
            switch (measure) {
                case 0: rc = run_segment<0>(ctx, s, al, (int)stage, len, st, proven_clean); break;
                case 1: rc = run_segment<1>(ctx, s, al, (int)stage, len, st, proven_clean); break;
                case 2: rc = run_segment<2>(ctx, s, al, (int)stage, len, st, proven_clean); break;
                case 3: rc = run_segment<3>(ctx, s, al, (int)stage, len, st, proven_clean); break;
                default: rc = run_segment<4>(ctx, s, al, (int)stage, len, st, proven_clean); break;
            }
        }
        if (rc) return rc;
        if (d_validity && any_validity) {
            ValidityArgs v;
            v.va = s.a.validity;
            v.abit = s.a.vbit;
            v.astride = s.a.stride;
            v.vb = s.b.validity;
            v.bbit = s.b.vbit;
            v.bstride = s.b.stride;
            v.n = len;
            v.out_row0 = row;
            v.out = d_validity;
            v.null_count = ctx.d_nulls;
            const long long words = ((row + len - 1) >> 5) - (row >> 5) + 1;
            validity_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(v);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            CUDA_TRY(cudaGetLastError());
        }
        row += len;
        oa += len;
        ob += len;
    }
    return STRSIM_OK;
}

// ---- exported C ABI ----------------------------------------------------------------------------------------
extern "C" {

int strsim_b200_set_device(int device) {
    g_requested_device = device;
    ThreadCtx* c;
    return ensure_ctx(&c);
}

int strsim_b200_bind_thread_near_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* p = bus; *p; p++) *p = (char)tolower((unsigned char)*p);
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE* f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node < 0) return -1;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    cpu_set_t have, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof have, &have) != 0) {
        fclose(f);
        return -1;
    }
    int lo, hi, any = 0;  // "0-15,32-47"
    while (fscanf(f, "%d", &lo) == 1) {
        hi = lo;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &hi) != 1) break;
            c = fgetc(f);
        }
        for (int cpu = lo; cpu <= hi && cpu < CPU_SETSIZE; cpu++)
            if (CPU_ISSET(cpu, &have)) {
                CPU_SET(cpu, &want);
                any++;
            }
        if (c != ',') break;
    }
    fclose(f);
    if (!any || sched_setaffinity(0, sizeof want, &want) != 0) return -1;
    return node;
}

int strsim_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* strsim_b200_last_error(void) { return g_last_error.c_str(); }
uint64_t strsim_b200_kernel_launches(void) { return g_launches.load(); }
void strsim_b200_last_overflow(int64_t out[2]) {
    out[0] = g_last_overflow[0];
    out[1] = g_last_overflow[1];
}
int strsim_b200_last_redo_slices(void) { return g_last_redo_slices; }
const char* strsim_b200_version(void) { return "polars-strsim_b200 0.1.0 (sm_100a)"; }

int strsim_b200_column_upload(const strsim_view_chunk* chunks, size_t n_chunks, strsim_b200_column** out) {
    if (!out || (n_chunks && !chunks)) {
        strsim_set_error("column_upload: NULL argument");
        return STRSIM_ERR_ARGUMENT;
    }
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    rc = guarded("column_upload", [&] { return upload_column(*ctx, chunks, n_chunks, true, out); });
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return STRSIM_OK;
}

int strsim_b200_column_upload_dictionary(const strsim_dict_chunk* chunks, size_t n_chunks, strsim_b200_column** out) {
    if (!out || (n_chunks && !chunks)) {
        strsim_set_error("column_upload_dictionary: NULL argument");
        return STRSIM_ERR_ARGUMENT;
    }
    *out = nullptr;
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    return guarded("column_upload_dictionary", [&]() -> int {
        std::unique_ptr<strsim_b200_column, void (*)(strsim_b200_column*)> col(new strsim_b200_column(), strsim_b200_column_free);
        col->device = ctx->device;
        // device storage of the materialised column: per chunk views, validity words and a staging area for
        // the indices (+ their validity bytes)
        size_t total = 0;
        std::vector<size_t> off_views(n_chunks), off_valid(n_chunks), off_idx(n_chunks), off_ival(n_chunks);
        for (size_t i = 0; i < n_chunks; i++) {
            const strsim_dict_chunk& ch = chunks[i];
            if (ch.length < 0 || ch.offset < 0 || (ch.length > 0 && !ch.indices) ||
                (ch.index_bytes != 1 && ch.index_bytes != 2 && ch.index_bytes != 4 && ch.index_bytes != 8)) {
                strsim_set_error("dictionary chunk %zu: bad length / offset / index width", i);
                return STRSIM_ERR_ARGUMENT;
            }
            off_views[i] = total;
            total = align_up(total + 16 * (size_t)ch.length, 256);
            off_valid[i] = total;
            total = align_up(total + 4 * (size_t)((ch.length + 31) / 32) + 8, 256);
            off_idx[i] = total;
            total = align_up(total + (size_t)ch.index_bytes * (size_t)ch.length, 256);
            off_ival[i] = total;
            total = align_up(total + (ch.validity ? (size_t)((ch.offset + ch.length + 7) / 8 - ch.offset / 8) : 0) + 8, 256);
        }
        total += 256;
        int rc2 = pool_alloc(ctx->device, total, &col->block);
        if (rc2) return rc2;
        col->block_bytes = total;
        col->has_validity = true;
        char* base = static_cast<char*>(col->block);
        // each distinct dictionary is uploaded once (the chunks of a Polars Categorical share theirs)
        std::vector<std::pair<const void*, strsim_b200_column*>> seen;
        unsigned or_byte = 0, and_byte = 0xFF;
        double out_of_line = 0.0;
        for (size_t i = 0; i < n_chunks; i++) {
            const strsim_dict_chunk& ch = chunks[i];
            strsim_b200_column* dict = nullptr;
            for (auto& s2 : seen)
                if (s2.first == ch.values.views && s2.second->length == ch.values.length) dict = s2.second;
            if (!dict) {
                rc2 = upload_column(*ctx, &ch.values, 1, false, &dict);
                if (rc2) return rc2;
                col->dictionaries.push_back(dict);
                seen.emplace_back(ch.values.views, dict);
                or_byte |= dict->or_byte;
                and_byte &= dict->and_byte;
                col->block_bytes += dict->block_bytes;
            }
            if (dict->length > 0) out_of_line = std::max(out_of_line, (double)dict->data_bytes / (double)dict->length);
            if (ch.length == 0) {
                DevChunk dc{};
                dc.length = 0;
                col->chunks.push_back(dc);
                col->chunk_buf_ids.emplace_back();
                continue;
            }
            const size_t idx_bytes = (size_t)ch.index_bytes * (size_t)ch.length;
            CUDA_TRY(upload_copy(*ctx, base + off_idx[i], static_cast<const char*>(ch.indices) + (size_t)ch.index_bytes * (size_t)ch.offset,
                                 idx_bytes, ctx->stream));
            const int64_t first_byte = ch.offset >> 3;
            if (ch.validity)
                CUDA_TRY(cudaMemcpyAsync(base + off_ival[i], ch.validity + first_byte,
                                         (size_t)(((ch.offset + ch.length + 7) >> 3) - first_byte), cudaMemcpyHostToDevice,
                                         ctx->stream));
            CUDA_TRY(upload_flush(*ctx));
            const DevChunk& dv = dict->chunks[0];
            DictGatherArgs g{};
            g.dict_views = dv.views;
            g.dict_validity = dv.validity;
            g.dict_vbit = dv.vbit;
            g.dict_len = dict->length;
            g.indices = base + off_idx[i];
            g.index_bytes = ch.index_bytes;
            g.index_signed = ch.index_signed;
            g.validity = ch.validity ? reinterpret_cast<const uint8_t*>(base + off_ival[i]) : nullptr;
            g.vbit = ch.offset & 7;
            g.n = ch.length;
            g.out_views = reinterpret_cast<uint4*>(base + off_views[i]);
            g.out_validity = reinterpret_cast<uint32_t*>(base + off_valid[i]);
            dict_gather_kernel<<<(unsigned)((ch.length + 255) / 256), 256, 0, ctx->stream>>>(g);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            CUDA_TRY(cudaGetLastError());
            DevChunk dc;
            dc.views = g.out_views;
            dc.validity = reinterpret_cast<const uint8_t*>(g.out_validity);
            dc.vbit = 0;
            dc.bufs = dv.bufs;
            dc.length = ch.length;
            dc.data_bytes = 0;
            col->chunks.push_back(dc);
            col->chunk_buf_ids.emplace_back();  // no residency bookkeeping: the dictionary is complete
            col->length += ch.length;
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        col->or_byte = or_byte;
        col->and_byte = and_byte;
        // rows repeat dictionary entries in any order: the stage area is sized for the dictionary's mean
        // out-of-line bytes per entry
        col->data_bytes = (int64_t)(out_of_line * (double)col->length);
        if (col->length == 1) {
            // a literal must not be null (see null_literal_error): read the one validity bit back
            uint32_t word = 1;
            for (size_t i = 0; i < n_chunks; i++)
                if (chunks[i].length == 1)
                    CUDA_TRY(cudaMemcpy(&word, base + off_valid[i], 4, cudaMemcpyDeviceToHost));
            col->scalar_null = !(word & 1u);
        }
        *out = col.release();
        return STRSIM_OK;
    });
}

void strsim_b200_column_free(strsim_b200_column* col) {
    if (!col) return;
    for (strsim_b200_column* sh : col->shards) strsim_b200_column_free(sh);
    for (strsim_b200_column* d : col->dictionaries) strsim_b200_column_free(d);
    if (col->block) {
        cudaSetDevice(col->device);
        pool_free(col->device, col->block, col->block_bytes);
    }
    delete col;
}

int64_t strsim_b200_column_length(const strsim_b200_column* col) { return col ? col->length : -1; }
int64_t strsim_b200_column_device_bytes(const strsim_b200_column* col) { return col ? (int64_t)col->block_bytes : -1; }
int strsim_b200_column_device(const strsim_b200_column* col) { return col ? col->device : -1; }
int strsim_b200_get_device(void) {
    ThreadCtx* c;
    if (ensure_ctx(&c) != STRSIM_OK) return -1;
    // sharded mode: what columns kept by host calls report as their "device" (strsim_b200_column_device)
    const size_t g = shard_devices().size();
    return g >= 2 ? SHARD_DEVICE_BASE + (int)g : c->device;
}
int64_t strsim_b200_column_algorithmic_bytes(const strsim_b200_column* col) {
    return col ? col->alg_bytes : -1;
}

int strsim_b200_column_restat(strsim_b200_column* col, void* stream) {
    if (!col) {
        strsim_set_error("column_restat: NULL column");
        return STRSIM_ERR_ARGUMENT;
    }
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    if (col->device != ctx->device) {
        strsim_set_error("column lives on device %d, calling thread uses device %d", col->device, ctx->device);
        return STRSIM_ERR_ARGUMENT;
    }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    stats_init_value(ctx->h_stats);
    CUDA_TRY(cudaMemcpyAsync(ctx->d_stats, ctx->h_stats, sizeof(ColumnStats), cudaMemcpyHostToDevice, st));
    char* base = static_cast<char*>(col->block);
    for (size_t id = 0; id < col->buf_size.size(); id++) {
        const int64_t lo = col->buf_lo[id], size = col->buf_hi[id] - lo;
        if (size <= 0) continue;
        long long blocks = ((size >> 4) + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        stats_bytes_kernel<<<(unsigned)blocks, 256, 0, st>>>(
            reinterpret_cast<const unsigned char*>(base + col->buf_dev_off[id] + lo), size, ctx->d_stats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    for (const DevChunk& dc : col->chunks) {
        if (dc.length <= 0) continue;
        long long blocks = (dc.length + 4 * STATS_BLOCK - 1) / (4 * STATS_BLOCK);
        if (blocks > 148 * 8) blocks = 148 * 8;
        stats_views_kernel<<<(unsigned)blocks, STATS_BLOCK, 0, st>>>(dc.views, dc.length, ctx->d_stats);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(publish_to_host(ctx->h_stats, ctx->d_stats, sizeof(ColumnStats), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    stats_fold(*ctx->h_stats, &col->or_byte, &col->and_byte);
    if (col->dictionaries.empty()) {  // (a materialised dictionary column's rows are not scanned here)
        col->max_len = ctx->h_stats->max_len;
        col->max_block_pad = ctx->h_stats->max_block_pad;
    }
    return STRSIM_OK;
}

int strsim_b200_compute_device(int measure, const strsim_b200_column* a, const strsim_b200_column* b,
                               double* d_out_values, uint32_t* d_out_validity, int32_t* d_dbg_ints,
                               void* stream) {
    if (!a || !b || !d_out_values) {
        strsim_set_error("compute_device: NULL argument");
        return STRSIM_ERR_ARGUMENT;
    }
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    if (a->device != ctx->device || b->device != ctx->device) {
        strsim_set_error("columns live on device %d/%d, calling thread uses device %d", a->device,
                         b->device, ctx->device);
        return STRSIM_ERR_ARGUMENT;
    }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    g_last_overflow[0] = g_last_overflow[1] = 0;
    const int64_t n = a->length == 1 ? b->length : a->length;
    double* outs[1] = {d_out_values};
    int32_t* dbgs[1] = {d_dbg_ints};
    return compute_on_device(*ctx, &measure, 1, a, b, 0, n, classify_alphabet(a->or_byte | b->or_byte, a->and_byte & b->and_byte),
                             outs, d_out_validity, dbgs, st);
}

int strsim_b200_compute_device_multi(const int* measures, size_t n_measures, const strsim_b200_column* a,
                                     const strsim_b200_column* b, double* const* d_out_values,
                                     uint32_t* d_out_validity, int32_t* const* d_dbg_ints, void* stream) {
    if (!a || !b || !measures || !d_out_values || n_measures == 0 || n_measures > 5) {
        strsim_set_error("compute_device_multi: NULL / out-of-range argument");
        return STRSIM_ERR_ARGUMENT;
    }
    for (size_t k = 0; k < n_measures; k++)
        if (!d_out_values[k]) {
            strsim_set_error("compute_device_multi: NULL d_out_values[%zu]", k);
            return STRSIM_ERR_ARGUMENT;
        }
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    if (a->device != ctx->device || b->device != ctx->device) {
        strsim_set_error("columns live on device %d/%d, calling thread uses device %d", a->device,
                         b->device, ctx->device);
        return STRSIM_ERR_ARGUMENT;
    }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    g_last_overflow[0] = g_last_overflow[1] = 0;
    const int64_t n = a->length == 1 ? b->length : a->length;
    return compute_on_device(*ctx, measures, n_measures, a, b, 0, n,
                             classify_alphabet(a->or_byte | b->or_byte, a->and_byte & b->and_byte), d_out_values,
                             d_out_validity, d_dbg_ints, st);
}

}  // extern "C"

// The host call.  A column is either given as host chunks (uploaded in pipelined row slices) or as
// `res_x`, a column that is already resident in HBM (then x / n_x are ignored).  keep_x != nullptr: the
// uploaded column is handed to the caller instead of being freed (strsim_b200_compute_host_keep).
static int host_call_impl(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                          const strsim_b200_column* res_a, strsim_b200_column** keep_a, const strsim_view_chunk* b,
                          size_t n_b, const strsim_b200_column* res_b, strsim_b200_column** keep_b,
                          double* const* out_values, uint8_t* out_validity, int64_t* out_null_count,
                          int32_t* const* dbg_ints, bool trim);
static bool shard_this_call(int64_t la, int64_t lb, const strsim_b200_column* res_a, const strsim_b200_column* res_b);
static int host_call_sharded(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                             const strsim_b200_column* res_a, strsim_b200_column** keep_a, const strsim_view_chunk* b,
                             size_t n_b, const strsim_b200_column* res_b, strsim_b200_column** keep_b,
                             double* const* out_values, uint8_t* out_validity, int64_t* out_null_count,
                             int32_t* const* dbg_ints);

static int host_call(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                     const strsim_b200_column* res_a, strsim_b200_column** keep_a, const strsim_view_chunk* b,
                     size_t n_b, const strsim_b200_column* res_b, strsim_b200_column** keep_b,
                     double* const* out_values, uint8_t* out_validity, int64_t* out_null_count,
                     int32_t* const* dbg_ints) {
    return guarded("compute_host", [&] {
        int64_t la = 0, lb = 0;
        if (res_a) la = res_a->length;
        else for (size_t i = 0; a && i < n_a; i++) la += a[i].length;
        if (res_b) lb = res_b->length;
        else for (size_t i = 0; b && i < n_b; i++) lb += b[i].length;
        if (shard_this_call(la, lb, res_a, res_b))
            return host_call_sharded(measures, n_measures, a, n_a, res_a, keep_a, b, n_b, res_b, keep_b, out_values,
                                     out_validity, out_null_count, dbg_ints);
        if ((res_a && !res_a->shards.empty()) || (res_b && !res_b->shards.empty())) {
            strsim_set_error("a column kept by a sharded call can only be used by sharded calls of the same shape");
            return (int)STRSIM_ERR_ARGUMENT;
        }
        return host_call_impl(measures, n_measures, a, n_a, res_a, keep_a, b, n_b, res_b, keep_b, out_values, out_validity,
                              out_null_count, dbg_ints, false);
    });
}

static int host_call_impl(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                          const strsim_b200_column* res_a, strsim_b200_column** keep_a, const strsim_view_chunk* b,
                          size_t n_b, const strsim_b200_column* res_b, strsim_b200_column** keep_b,
                          double* const* out_values, uint8_t* out_validity, int64_t* out_null_count,
                          int32_t* const* dbg_ints, bool trim) {
    if (keep_a) *keep_a = nullptr;
    if (keep_b) *keep_b = nullptr;
    if ((!res_a && n_a && !a) || (!res_b && n_b && !b) || !measures || n_measures == 0 || n_measures > 8 || !out_values) {
        strsim_set_error("compute_host: NULL / out-of-range argument");
        return STRSIM_ERR_ARGUMENT;
    }
    for (size_t m = 0; m < n_measures; m++)
        if (measures[m] < 0 || measures[m] > 4) {
            strsim_set_error("unknown measure %d", measures[m]);
            return STRSIM_ERR_ARGUMENT;
        }
    int64_t la = 0, lb = 0;
    if (res_a) la = res_a->length;
    else for (size_t i = 0; i < n_a; i++) la += a[i].length;
    if (res_b) lb = res_b->length;
    else for (size_t i = 0; i < n_b; i++) lb += b[i].length;
    if (la != lb && la != 1 && lb != 1) {
        strsim_set_error("Inputs must have the same length, or one of them must be a Utf8 literal.");
        return STRSIM_ERR_SHAPE;
    }
    const int64_t n = la == 1 ? lb : la;
    if (la != lb) {
        const bool null_a = la == 1 && (res_a ? res_a->scalar_null : host_scalar_is_null(a, n_a));
        const bool null_b = lb == 1 && (res_b ? res_b->scalar_null : host_scalar_is_null(b, n_b));
        if (null_a || null_b) return null_literal_error();
    }
    if (out_null_count) *out_null_count = 0;
    for (size_t m = 0; n > 0 && m < n_measures; m++)
        if (!out_values[m]) {
            strsim_set_error("compute_host: NULL out_values[%zu]", m);
            return STRSIM_ERR_ARGUMENT;
        }
    ThreadCtx* ctx;
    int rc = ensure_ctx(&ctx);
    if (rc) return rc;
    if ((res_a && res_a->device != ctx->device) || (res_b && res_b->device != ctx->device)) {
        strsim_set_error("resident column lives on another device than the calling thread uses (%d)", ctx->device);
        return STRSIM_ERR_ARGUMENT;
    }
    if (n == 0) return STRSIM_OK;
    unsigned seen_measures = 0;
    bool distinct = true;
    for (size_t m = 0; m < n_measures; m++) {
        distinct = distinct && !(seen_measures & (1u << measures[m]));
        seen_measures |= 1u << measures[m];
    }
    static const bool no_fuse = getenv("STRSIM_B200_NO_FUSE") != nullptr && atoi(getenv("STRSIM_B200_NO_FUSE")) != 0;
    const bool fuse = n_measures > 1 && n_measures <= 5 && distinct && !no_fuse;

    // One upload of the two columns serves every requested measure.  The rows are cut into slices:
    // all H2D copies are queued up front on the upload stream (data buffers first, then the views
    // slice by slice), slice s is computed as soon as its views have landed, and its results go back
    // on the download stream -- so H2D, kernels and D2H of different slices overlap (PCIe is full
    // duplex) instead of running back to back.
    Uploader ua, ub;
    if (!res_a) {
        rc = upload_plan(*ctx, a, n_a, &ua, trim && la > 1);
        if (rc) return rc;
    }
    if (!res_b) {
        rc = upload_plan(*ctx, b, n_b, &ub, trim && lb > 1);
        if (rc) {
            if (!res_a) strsim_b200_column_free(ua.col);
            return rc;
        }
    }
    // resident columns are shared (possibly with other host threads): read-only here
    strsim_b200_column* ca = res_a ? const_cast<strsim_b200_column*>(res_a) : ua.col;
    strsim_b200_column* cb = res_b ? const_cast<strsim_b200_column*>(res_b) : ub.col;
    auto free_uploaded = [&]() {
        if (!res_a) strsim_b200_column_free(ca);
        if (!res_b) strsim_b200_column_free(cb);
    };
    static const long long slice_target = [] {
        const char* e = getenv("STRSIM_B200_SLICE_ROWS");
        const long long v = e && *e ? atoll(e) : 0;
        return v > 0 ? v : SLICE_ROWS;
    }();
    int n_slices = (int)((n + slice_target - 1) / slice_target);
    if (n_slices > MAX_SLICES) n_slices = MAX_SLICES;
    if (n_slices < 1) n_slices = 1;
    int64_t slice_rows = ((n + n_slices - 1) / n_slices + 63) & ~63ll;

    const size_t val_words = (size_t)((n + 31) / 32);
    const size_t out_stride = align_up(8 * (size_t)n, 256);
    const size_t dbg_stride = align_up(24 * (size_t)n, 256);
    bool any_dbg = false;
    for (size_t m = 0; dbg_ints && m < n_measures; m++) any_dbg = any_dbg || dbg_ints[m] != nullptr;
    const size_t bytes = n_measures * out_stride + align_up(4 * val_words, 256) + (any_dbg ? n_measures * dbg_stride : 0);
    void* d_block = nullptr;
    rc = pool_alloc(ctx->device, bytes, &d_block);
    if (rc) {
        free_uploaded();
        return rc;
    }
    char* base = static_cast<char*>(d_block);
    uint32_t* d_val = reinterpret_cast<uint32_t*>(base + n_measures * out_stride);
    char* d_dbg_base = reinterpret_cast<char*>(d_val) + align_up(4 * val_words, 256);
    const bool want_validity = out_validity != nullptr || out_null_count != nullptr;

    // ---- queue every upload; statistics slots: [0] data of a, [1] data of b, [2+s] views of slice s
    // Progressive order (see "progressive upload" above): per slice the data prefix it needs, then its
    // views.  Otherwise (scalar operand, one slice, views that do not look sequential, test hooks that
    // bypass the residency check) all data first, as before.
    static const bool no_progressive = [] {
        const char* e = getenv("STRSIM_B200_PROGRESSIVE");
        return e && !strcmp(e, "0");
    }();
    n_slices = (int)((n + slice_rows - 1) / slice_rows);
    const bool progressive = !no_progressive && !force_generic_rows() && la == lb && n_slices > 1 &&
                             (res_a || looks_sequential(ua)) && (res_b || looks_sequential(ub));
    const int64_t total_a = total_data_bytes(ca), total_b = total_data_bytes(cb);
    if (res_a) ua.uploaded = total_a;  // nothing to send
    if (res_b) ub.uploaded = total_b;
    int64_t front_a[MAX_SLICES], front_b[MAX_SLICES];
    cudaError_t ce = cudaSuccess;
    // STRSIM_B200_TRACE=1: per-slice timeline (upload landed, kernels start/end, download done) on stderr
    static const bool trace = getenv("STRSIM_B200_TRACE") != nullptr && atoi(getenv("STRSIM_B200_TRACE")) != 0;
    static thread_local cudaEvent_t tr_t0 = nullptr, tr_u[MAX_SLICES], tr_c0[MAX_SLICES], tr_c1[MAX_SLICES], tr_d1[MAX_SLICES];
    if (trace) {
        if (!tr_t0) {
            cudaEventCreate(&tr_t0);
            for (int i = 0; i < MAX_SLICES; i++) {
                cudaEventCreate(&tr_u[i]);
                cudaEventCreate(&tr_c0[i]);
                cudaEventCreate(&tr_c1[i]);
                cudaEventCreate(&tr_d1[i]);
            }
        }
        cudaEventRecord(tr_t0, ctx->upload_stream);
    }
    for (int i = 0; i < 2 + MAX_SLICES; i++) stats_init_value(&ctx->h_slice_stats[i]);
    ce = cudaMemcpyAsync(ctx->d_slice_stats, ctx->h_slice_stats, sizeof(ColumnStats) * (2 + MAX_SLICES),
                         cudaMemcpyHostToDevice, ctx->upload_stream);
    if (ce == cudaSuccess && !res_a) rc = upload_tables(ua, ctx->upload_stream);
    if (ce == cudaSuccess && rc == STRSIM_OK && !res_b) rc = upload_tables(ub, ctx->upload_stream);
    // Pinned sources: every copy of the call is queued up front (asynchronous DMA).  Pageable sources are
    // staged through the pinned ring by upload_copy(), which blocks this thread while it feeds the DMA
    // engine, so their slices are queued one ahead of the slice being computed.
    auto source_pinned = [](const Uploader& up) {
        for (size_t i = 0; i < up.n_chunks; i++) {
            if (up.chunks[i].length > 65536 && !is_pinned(up.chunks[i].views)) return false;
            for (size_t b = 0; b < up.buf_src.size(); b++)
                if (up.col->buf_size[b] > (1 << 20) && !is_pinned(up.buf_src[b])) return false;
        }
        return true;
    };
    const bool queue_all_upfront = (res_a || source_pinned(ua)) && (res_b || source_pinned(ub));
    int queued_slices = 0;
    auto queue_slice_upload = [&](int sidx) {
        if (sidx != queued_slices || sidx >= n_slices || ce != cudaSuccess || rc != STRSIM_OK) return;
        queued_slices++;
        const int64_t lo = sidx * slice_rows, hi = lo + slice_rows < n ? lo + slice_rows : n;
        int64_t fa = total_a, fb = total_b;
        if (progressive && sidx + 1 < n_slices) {
            fa = res_a ? total_a : frontier_after(ua, hi);
            fb = res_b ? total_b : frontier_after(ub, hi);
            if (fa < 0) fa = total_a;
            if (fb < 0) fb = total_b;
            if (fa < ua.uploaded) fa = ua.uploaded;
            if (fb < ub.uploaded) fb = ub.uploaded;
        }
        // copies back to back on the upload stream; the statistics kernels follow on their own stream (a
        // copy -> kernel -> copy chain on one stream costs a copy-engine / compute hand-over per link,
        // measured at 0.15-0.2 ms per slice)
        const int64_t from_a = ua.uploaded < fa ? ua.uploaded : fa, from_b = ub.uploaded < fb ? ub.uploaded : fb;
        const int64_t r_lo_a = la == 1 && n != 1 ? 0 : lo, r_hi_a = la == 1 && n != 1 ? (sidx == 0 ? 1 : 0) : hi;
        const int64_t r_lo_b = lb == 1 && n != 1 ? 0 : lo, r_hi_b = lb == 1 && n != 1 ? (sidx == 0 ? 1 : 0) : hi;
        for (int phase = 0; phase < 2 && rc == STRSIM_OK && ce == cudaSuccess; phase++) {
            const bool cp = phase == 0, stt = phase == 1;
            cudaStream_t st = cp ? ctx->upload_stream : ctx->stats_stream;
            if (stt) {
                ce = cudaEventRecord(ctx->up_event[sidx], ctx->upload_stream);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stats_stream, ctx->up_event[sidx], 0);
                if (ce != cudaSuccess) break;
            }
            if (!res_a) rc = upload_data_range(*ctx, ua, from_a, fa, st, ctx->d_slice_stats + 0, cp, stt);
            if (rc == STRSIM_OK && !res_b) rc = upload_data_range(*ctx, ub, from_b, fb, st, ctx->d_slice_stats + 1, cp, stt);
            // a scalar (length-1) column is uploaded with the first slice
            if (rc == STRSIM_OK && !res_a)
                rc = upload_rows(*ctx, ua, r_lo_a, r_hi_a, st, ctx->d_slice_stats + 2 + sidx, cp, stt);
            if (rc == STRSIM_OK && !res_b)
                rc = upload_rows(*ctx, ub, r_lo_b, r_hi_b, st, ctx->d_slice_stats + 2 + sidx, cp, stt);
            // pageable sources: the four copies above shared one pipeline through the pinned ring; what is
            // still in it is sent now, before the event that stands for "this slice has landed"
            if (cp && ce == cudaSuccess) ce = upload_flush(*ctx);
        }
        if (fa > ua.uploaded) ua.uploaded = fa;
        if (fb > ub.uploaded) ub.uploaded = fb;
        front_a[sidx] = ua.uploaded;
        front_b[sidx] = ub.uploaded;
        if (rc == STRSIM_OK && ce == cudaSuccess)
            ce = publish_to_host(ctx->h_slice_stats, ctx->d_slice_stats, sizeof(ColumnStats) * (3 + sidx),
                                 ctx->stats_stream);
        if (rc == STRSIM_OK && ce == cudaSuccess) ce = cudaEventRecord(ctx->slice_event[sidx], ctx->stats_stream);
        if (trace) cudaEventRecord(tr_u[sidx], ctx->upload_stream);
    };
    for (int sidx = 0; sidx < (queue_all_upfront ? n_slices : 1); sidx++) queue_slice_upload(sidx);

    // ---- compute + download slice by slice
    // redo = false: as soon as the slice's views (and data prefix) have landed; returns in *deferred the
    // rows whose payload was not there yet -- such a slice is not downloaded but computed again (redo =
    // true) after the whole upload, without touching the validity bitmap / null count a second time
    auto run_slice = [&](int sidx, bool redo, int64_t* deferred) {
        const int64_t lo = sidx * slice_rows, hi = lo + slice_rows < n ? lo + slice_rows : n;
        const int last = redo ? n_slices - 1 : sidx;  // statistics slots that are valid by now
        ColumnStats acc = ctx->h_slice_stats[0];
        acc.or_bits |= ctx->h_slice_stats[1].or_bits | ctx->h_slice_stats[2 + sidx].or_bits;
        acc.and_bits &= ctx->h_slice_stats[1].and_bits & ctx->h_slice_stats[2 + sidx].and_bits;
        if (la == 1 || lb == 1) {  // the scalar's inline bytes were counted with slice 0
            acc.or_bits |= ctx->h_slice_stats[2].or_bits;
            acc.and_bits &= ctx->h_slice_stats[2].and_bits;
        }
        unsigned ob, nb_;
        stats_fold(acc, &ob, &nb_);
        if (res_a) {  // a resident column brings the statistics of its own upload
            ob |= res_a->or_byte;
            nb_ &= res_a->and_byte;
        }
        if (res_b) {
            ob |= res_b->or_byte;
            nb_ &= res_b->and_byte;
        }
        const Alphabet al = classify_alphabet(ob, nb_);
        if (!res_a) set_resident(ca, redo ? total_a : front_a[sidx]);
        if (!res_b) set_resident(cb, redo ? total_b : front_b[sidx]);
        ce = cudaStreamWaitEvent(ctx->stream, ctx->slice_event[last], 0);
        g_deferred = 0;
        // fused: ONE launch per segment evaluates every requested measure (run_segment_multi); otherwise
        // (a single measure, or the same measure requested twice) one pass per measure
        const size_t n_pass = fuse ? 1 : n_measures;
        double* d_outs[8];
        int32_t* d_dbgs[8];
        for (size_t m = 0; m < n_measures; m++) {
            d_outs[m] = reinterpret_cast<double*>(base + m * out_stride);
            d_dbgs[m] = (any_dbg && dbg_ints[m]) ? reinterpret_cast<int32_t*>(d_dbg_base + m * dbg_stride) : nullptr;
        }
        for (size_t p = 0; ce == cudaSuccess && rc == STRSIM_OK && p < n_pass; p++) {
            const size_t m0 = p, m1 = fuse ? n_measures : p + 1;
            rc = compute_on_device(*ctx, measures + m0, m1 - m0, ca, cb, lo, hi - lo, al, d_outs + m0,
                                   (want_validity && p == 0 && !redo) ? d_val : nullptr, d_dbgs + m0, ctx->stream);
        }
        *deferred = g_deferred;
        if (rc != STRSIM_OK || ce != cudaSuccess || g_deferred > 0) return;
        // download the finished slice on the copy stream while the next kernel runs
        cudaEvent_t ev = ctx->done_event[0];
        ce = cudaEventRecord(ev, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copy_stream, ev, 0);
        for (size_t m = 0; ce == cudaSuccess && m < n_measures; m++) {
            ce = download(*ctx, out_values[m] + lo, d_outs[m] + lo, 8 * (size_t)(hi - lo), ctx->copy_stream);
            if (ce == cudaSuccess && d_dbgs[m])
                ce = download(*ctx, dbg_ints[m] + 6 * lo, d_dbgs[m] + 6 * lo, 24 * (size_t)(hi - lo), ctx->copy_stream);
        }
    };
    int redo_list[MAX_SLICES], n_redo = 0;
    for (int sidx = 0; ce == cudaSuccess && rc == STRSIM_OK && sidx < n_slices; sidx++) {
        queue_slice_upload(sidx + 1);  // no-op when everything was queued up front
        if (ce != cudaSuccess || rc != STRSIM_OK) break;
        ce = cudaEventSynchronize(ctx->slice_event[sidx]);  // views + statistics of this slice are here
        if (ce != cudaSuccess) break;
        if (sidx == 0) g_last_overflow[0] = g_last_overflow[1] = 0;
        int64_t deferred = 0;
        if (trace) cudaEventRecord(tr_c0[sidx], ctx->stream);
        run_slice(sidx, false, &deferred);
        if (trace) {
            cudaEventRecord(tr_c1[sidx], ctx->stream);
            cudaEventRecord(tr_d1[sidx], ctx->copy_stream);
        }
        if (deferred > 0) redo_list[n_redo++] = sidx;
    }
    if (n_redo > 0 && ce == cudaSuccess && rc == STRSIM_OK) {
        ce = cudaEventSynchronize(ctx->slice_event[n_slices - 1]);  // every byte has arrived
        for (int k = 0; ce == cudaSuccess && rc == STRSIM_OK && k < n_redo; k++) {
            int64_t deferred = 0;
            run_slice(redo_list[k], true, &deferred);
            if (rc == STRSIM_OK && deferred > 0) {
                // a row shard uploaded only the stretch of the data its first and last rows delimit and some
                // view points outside it: the caller repeats the shard with whole buffers
                strsim_set_error("internal: rows still deferred after the whole upload");
                rc = trim ? STRSIM_RETRY_UNTRIMMED : (int)STRSIM_ERR_CUDA;
            }
        }
    }
    g_last_redo_slices = n_redo;
    if (!res_a) set_resident(ca, total_a);
    if (!res_b) set_resident(cb, total_b);
    if (ce == cudaSuccess && rc == STRSIM_OK && out_validity) {
        // the validity bitmap does not depend on the data: one copy once every slice's kernel has run
        ce = cudaEventRecord(ctx->done_event[1], ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->done_event[1], 0);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(out_validity, d_val, (size_t)((n + 7) / 8), cudaMemcpyDeviceToHost, ctx->copy_stream);
    }
    if (ce == cudaSuccess && rc == STRSIM_OK && want_validity) {
        ce = cudaEventRecord(ctx->done_event[0], ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->done_event[0], 0);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(ctx->h_nulls, ctx->d_nulls, 8, cudaMemcpyDeviceToHost, ctx->copy_stream);
    }
    upload_flush(*ctx);  // error paths: nothing may stay in the ring (it would be sent into freed memory later)
    cudaError_t s0 = cudaStreamSynchronize(ctx->upload_stream);
    if (s0 == cudaSuccess) s0 = cudaStreamSynchronize(ctx->stats_stream);
    cudaError_t s1 = cudaStreamSynchronize(ctx->stream);
    cudaError_t s2 = cudaStreamSynchronize(ctx->copy_stream);
    download_wait(*ctx);
    if (rc == STRSIM_OK && (ce != cudaSuccess || s0 != cudaSuccess || s1 != cudaSuccess || s2 != cudaSuccess)) {
        const cudaError_t first = ce != cudaSuccess ? ce : s0 != cudaSuccess ? s0 : s1 != cudaSuccess ? s1 : s2;
        strsim_set_error("CUDA error while uploading / computing / downloading: %s", cudaGetErrorString(first));
        rc = STRSIM_ERR_CUDA;
    }
    if (rc == STRSIM_OK && out_null_count) *out_null_count = (int64_t)*ctx->h_nulls;
    if (trace && rc == STRSIM_OK) {
        fprintf(stderr, "[strsim trace] %d slices of %lld rows, progressive=%d, redo=%d\n", n_slices, (long long)slice_rows,
                (int)progressive, n_redo);
        for (int i = 0; i < n_slices; i++) {
            float u = 0, c0 = 0, c1 = 0, d1 = 0;
            cudaEventElapsedTime(&u, tr_t0, tr_u[i]);
            cudaEventElapsedTime(&c0, tr_t0, tr_c0[i]);
            cudaEventElapsedTime(&c1, tr_t0, tr_c1[i]);
            cudaEventElapsedTime(&d1, tr_t0, tr_d1[i]);
            fprintf(stderr, "[strsim trace] slice %2d: uploaded %7.3f  kernels %7.3f .. %7.3f  downloaded %7.3f ms\n", i, u, c0,
                    c1, d1);
        }
    }
    pool_free(ctx->device, d_block, bytes);
    if (rc == STRSIM_OK && (keep_a || keep_b)) {
        // byte statistics of the kept columns: the union over everything this call uploaded (a superset
        // of each column's own bytes -- the statistics only ever pick a more general kernel)
        ColumnStats acc = ctx->h_slice_stats[0];
        for (int i = 1; i < 2 + n_slices; i++) {
            acc.or_bits |= ctx->h_slice_stats[i].or_bits;
            acc.and_bits &= ctx->h_slice_stats[i].and_bits;
        }
        unsigned ob, nb_;
        stats_fold(acc, &ob, &nb_);
        for (strsim_b200_column* c : {res_a ? nullptr : ca, res_b ? nullptr : cb})
            if (c) {
                c->or_byte = ob | (res_a ? res_a->or_byte : 0u) | (res_b ? res_b->or_byte : 0u);
                c->and_byte = nb_ & (res_a ? res_a->and_byte : 0xFFu) & (res_b ? res_b->and_byte : 0xFFu);
            }
    }
    if (rc == STRSIM_OK && keep_a && !res_a) {
        *keep_a = ca;
        ca = nullptr;
    }
    if (rc == STRSIM_OK && keep_b && !res_b) {
        *keep_b = cb;
        cb = nullptr;
    }
    if (!res_a && ca) strsim_b200_column_free(ca);
    if (!res_b && cb) strsim_b200_column_free(cb);
    return rc;
}

// ---- one host call over several GPUs (north_star 4) ---------------------------------------
// The reference fans one call out over the workers of Polars' pool and re-assembles the chunks in the same
// call (strsim.rs:72-104: split_offsets, one task per range, from_chunk_iter).  Here the workers are the
// GPUs named by STRSIM_B200_DEVICES ("0,1,2,3", "0-7" or "all"): the rows are cut into one contiguous range
// per device with the reference's rule (split_offsets, strsim.rs:21-39: equal ranges, the last one takes the
// remainder; ranges start at multiples of 64 rows so that no validity byte is shared), every range runs
// the ordinary host call -- upload of ITS views and of the stretch of the data buffers its rows reference,
// kernels, download -- on its device's worker thread, and the results land in disjoint ranges of the
// caller's buffers: the "concatenation" is the layout itself.  No collective, no peer traffic.
static const std::vector<int>& shard_devices() {
    static const std::vector<int>& devs = *new std::vector<int>([] {
        std::vector<int> v;
        const char* e = getenv("STRSIM_B200_DEVICES");
        if (!e || !*e) return v;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            cudaGetLastError();
            return v;
        }
        if (!strcmp(e, "all")) {
            for (int d = 0; d < n; d++) v.push_back(d);
            return v;
        }
        for (const char* p = e; *p;) {
            char* end = nullptr;
            const long lo = strtol(p, &end, 10);
            if (end == p) break;
            long hi = lo;
            p = end;
            if (*p == '-') {
                hi = strtol(p + 1, &end, 10);
                if (end == p + 1) break;
                p = end;
            }
            for (long d = lo; d <= hi && d < n; d++) {
                bool dup = false;
                for (int x : v) dup = dup || x == (int)d;
                if (d >= 0 && !dup) v.push_back((int)d);
            }
            if (*p == ',') p++;
            else break;
        }
        return v;
    }());
    return devs;
}

constexpr int64_t SHARD_MIN_ROWS = 65536;  // below this one device is quicker than the fan-out

static bool shard_this_call(int64_t la, int64_t lb, const strsim_b200_column* res_a, const strsim_b200_column* res_b) {
    const size_t g = shard_devices().size();
    if (g < 2) return false;
    if (la != lb && la != 1 && lb != 1) return false;  // the ordinary path reports the shape error
    const int64_t n = la == 1 ? lb : la;
    if (n < SHARD_MIN_ROWS) return false;
    // resident operands must be columns a sharded call of the same fan-out kept
    for (const strsim_b200_column* r : {res_a, res_b})
        if (r && r->length > 1 && r->shards.size() != g) return false;
    return true;
}

// split_offsets (strsim.rs:21-39) on units of 64 rows: range g starts at g * (units / G) * 64, the last one
// takes the remainder.  Exported so that the rule can be checked without a GPU (tests/test_sharding.py).
extern "C" void strsim_b200_shard_cuts(int64_t n_rows, int n_shards, int64_t* cuts) {
    if (n_shards < 1 || !cuts) return;
    const int64_t units = (n_rows + 63) / 64, per = units / n_shards;
    for (int g = 0; g < n_shards; g++) cuts[g] = (int64_t)g * per * 64;
    cuts[n_shards] = n_rows;
}

// one worker thread per device: its thread-local context (streams, pinned buffers, staging rings) lives as
// long as the process, so a sharded call costs no set-up
class DeviceWorker {
   public:
    static DeviceWorker& of(int device) {
        static std::mutex& m = *new std::mutex();
        static std::vector<DeviceWorker*>& all = *new std::vector<DeviceWorker*>();
        std::lock_guard<std::mutex> lock(m);
        for (DeviceWorker* w : all)
            if (w->device_ == device) return *w;
        all.push_back(new DeviceWorker(device));  // never destroyed: the threads outlive static destructors
        return *all.back();
    }
    struct Job {
        std::function<int()> fn;
        int rc = 0;
        std::string error;
        int64_t overflow[2] = {0, 0};
        int redo = 0;
        bool done = false;
    };
    void submit(Job* j) {
        {
            std::lock_guard<std::mutex> lock(m_);
            q_.push_back(j);
        }
        cv_.notify_all();
    }
    void wait(Job* j) {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [j] { return j->done; });
    }

   private:
    explicit DeviceWorker(int device) : device_(device) {
        std::thread([this] { run(); }).detach();
    }
    void run() {
        g_requested_device = device_;
        for (;;) {
            Job* j;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [this] { return !q_.empty(); });
                j = q_.front();
                q_.pop_front();
            }
            g_last_error.clear();
            const int rc = guarded("sharded call", j->fn);
            j->error = g_last_error;
            j->overflow[0] = g_last_overflow[0];
            j->overflow[1] = g_last_overflow[1];
            j->redo = g_last_redo_slices;
            {
                std::lock_guard<std::mutex> lock(m_);
                j->rc = rc;
                j->done = true;
            }
            cv_.notify_all();
        }
    }
    int device_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job*> q_;
};

// rows [lo, hi) of a chunk list as a chunk list over the same buffers
static std::vector<strsim_view_chunk> slice_chunks(const strsim_view_chunk* ch, size_t n, int64_t lo, int64_t hi) {
    std::vector<strsim_view_chunk> out;
    int64_t row0 = 0;
    for (size_t i = 0; i < n; i++) {
        const int64_t c_lo = lo > row0 ? lo - row0 : 0;
        const int64_t c_hi = hi - row0 < ch[i].length ? hi - row0 : ch[i].length;
        if (c_hi > c_lo) {
            strsim_view_chunk c = ch[i];
            c.offset += c_lo;
            c.length = c_hi - c_lo;
            out.push_back(c);
        }
        row0 += ch[i].length;
    }
    return out;
}

static int host_call_sharded(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                             const strsim_b200_column* res_a, strsim_b200_column** keep_a, const strsim_view_chunk* b,
                             size_t n_b, const strsim_b200_column* res_b, strsim_b200_column** keep_b,
                             double* const* out_values, uint8_t* out_validity, int64_t* out_null_count,
                             int32_t* const* dbg_ints) {
    if (keep_a) *keep_a = nullptr;
    if (keep_b) *keep_b = nullptr;
    if ((!res_a && n_a && !a) || (!res_b && n_b && !b) || !measures || n_measures == 0 || n_measures > 8 || !out_values) {
        strsim_set_error("compute_host: NULL / out-of-range argument");
        return STRSIM_ERR_ARGUMENT;
    }
    const std::vector<int>& devs = shard_devices();
    const size_t G = devs.size();
    int64_t la = 0, lb = 0;
    if (res_a) la = res_a->length;
    else for (size_t i = 0; i < n_a; i++) la += a[i].length;
    if (res_b) lb = res_b->length;
    else for (size_t i = 0; i < n_b; i++) lb += b[i].length;
    const int64_t n = la == 1 ? lb : la;
    const bool bc_a = la == 1 && n != 1, bc_b = lb == 1 && n != 1;
    // the row ranges: those of the resident column(s), else split_offsets (strsim.rs:21-39) on 64-row units
    std::vector<int64_t> cut(G + 1, n);
    const strsim_b200_column* shaped = (res_a && !res_a->shards.empty()) ? res_a : (res_b && !res_b->shards.empty()) ? res_b : nullptr;
    if (shaped) {
        for (size_t g = 0; g <= G; g++) cut[g] = shaped->shard_lo[g];
        if (res_a && res_b && !res_a->shards.empty() && !res_b->shards.empty() && res_a->shard_lo != res_b->shard_lo) {
            strsim_set_error("the two resident columns were sharded differently");
            return STRSIM_ERR_ARGUMENT;
        }
    } else {
        strsim_b200_shard_cuts(n, (int)G, cut.data());
    }
    static const bool no_trim = [] {
        const char* t = getenv("STRSIM_B200_SHARD_TRIM");
        return t && !strcmp(t, "0");
    }();
    struct Shard {
        std::vector<strsim_view_chunk> a, b;
        std::vector<double*> outs;
        std::vector<int32_t*> dbgs;
        int64_t nulls = 0;
        strsim_b200_column *kept_a = nullptr, *kept_b = nullptr;
        DeviceWorker::Job job;
        bool used = false;
    };
    std::vector<Shard> shards(G);
    for (size_t g = 0; g < G; g++) {
        Shard& sh = shards[g];
        const int64_t lo = cut[g], hi = cut[g + 1];
        if (hi <= lo) continue;
        sh.used = true;
        if (!res_a) sh.a = bc_a ? std::vector<strsim_view_chunk>(a, a + n_a) : slice_chunks(a, n_a, lo, hi);
        if (!res_b) sh.b = bc_b ? std::vector<strsim_view_chunk>(b, b + n_b) : slice_chunks(b, n_b, lo, hi);
        for (size_t m = 0; m < n_measures; m++) {
            sh.outs.push_back(out_values[m] ? out_values[m] + lo : nullptr);
            sh.dbgs.push_back(dbg_ints && dbg_ints[m] ? dbg_ints[m] + 6 * lo : nullptr);
        }
        const strsim_b200_column* ra = res_a ? (res_a->shards.empty() ? res_a : res_a->shards[g]) : nullptr;
        const strsim_b200_column* rb = res_b ? (res_b->shards.empty() ? res_b : res_b->shards[g]) : nullptr;
        uint8_t* val = out_validity ? out_validity + (lo >> 3) : nullptr;
        const bool want_dbg = dbg_ints != nullptr;
        sh.job.fn = [&sh, measures, n_measures, ra, rb, keep_a, keep_b, val, want_dbg, force = force_generic_rows()]() {
            const bool trim = !no_trim && !force;
            int rc = host_call_impl(measures, n_measures, sh.a.data(), sh.a.size(), ra, keep_a && !ra ? &sh.kept_a : nullptr,
                                    sh.b.data(), sh.b.size(), rb, keep_b && !rb ? &sh.kept_b : nullptr, sh.outs.data(), val,
                                    &sh.nulls, want_dbg ? sh.dbgs.data() : nullptr, trim);
            if (rc == STRSIM_RETRY_UNTRIMMED)
                rc = host_call_impl(measures, n_measures, sh.a.data(), sh.a.size(), ra, keep_a && !ra ? &sh.kept_a : nullptr,
                                    sh.b.data(), sh.b.size(), rb, keep_b && !rb ? &sh.kept_b : nullptr, sh.outs.data(), val,
                                    &sh.nulls, want_dbg ? sh.dbgs.data() : nullptr, false);
            return rc;
        };
        DeviceWorker::of(devs[g]).submit(&sh.job);
    }
    int rc = STRSIM_OK;
    int64_t nulls = 0;
    g_last_overflow[0] = g_last_overflow[1] = 0;
    g_last_redo_slices = 0;
    for (size_t g = 0; g < G; g++) {
        Shard& sh = shards[g];
        if (!sh.used) continue;
        DeviceWorker::of(devs[g]).wait(&sh.job);
        if (sh.job.rc != STRSIM_OK && rc == STRSIM_OK) {
            rc = sh.job.rc;
            strsim_set_error("device %d: %s", devs[g], sh.job.error.c_str());
        }
        nulls += sh.nulls;
        g_last_overflow[0] += sh.job.overflow[0];
        g_last_overflow[1] += sh.job.overflow[1];
        g_last_redo_slices += sh.job.redo;
    }
    if (out_null_count) *out_null_count = nulls;
    // kept columns: one composite per operand (a literal operand is never kept as a composite)
    auto assemble = [&](bool is_a, strsim_b200_column** keep, int64_t len) {
        bool all = keep != nullptr && rc == STRSIM_OK && len > 1;
        for (size_t g = 0; g < G && all; g++)
            if (shards[g].used && !(is_a ? shards[g].kept_a : shards[g].kept_b)) all = false;
        if (all) {
            auto* comp = new strsim_b200_column();
            comp->device = SHARD_DEVICE_BASE + (int)G;
            comp->length = len;
            comp->shard_lo = cut;
            for (size_t g = 0; g < G; g++) {
                strsim_b200_column* k = is_a ? shards[g].kept_a : shards[g].kept_b;
                if (!k) {  // an empty range: a placeholder keeps the indices aligned
                    k = new strsim_b200_column();
                    k->device = devs[g];
                }
                comp->shards.push_back(k);
                comp->block_bytes += k->block_bytes;
            }
            *keep = comp;
        } else {
            for (size_t g = 0; g < G; g++) strsim_b200_column_free(is_a ? shards[g].kept_a : shards[g].kept_b);
        }
    };
    assemble(true, keep_a, la);
    assemble(false, keep_b, lb);
    return rc;
}

extern "C" {

int strsim_b200_compute_host_multi(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                                   const strsim_view_chunk* b, size_t n_b, double* const* out_values,
                                   uint8_t* out_validity, int64_t* out_null_count, int32_t* const* dbg_ints) {
    return host_call(measures, n_measures, a, n_a, nullptr, nullptr, b, n_b, nullptr, nullptr, out_values, out_validity,
                     out_null_count, dbg_ints);
}

int strsim_b200_compute_host_keep(const int* measures, size_t n_measures, const strsim_view_chunk* a, size_t n_a,
                                  const strsim_b200_column* resident_a, strsim_b200_column** keep_a,
                                  const strsim_view_chunk* b, size_t n_b, const strsim_b200_column* resident_b,
                                  strsim_b200_column** keep_b, double* const* out_values, uint8_t* out_validity,
                                  int64_t* out_null_count, int32_t* const* dbg_ints) {
    return host_call(measures, n_measures, a, n_a, resident_a, keep_a, b, n_b, resident_b, keep_b, out_values,
                     out_validity, out_null_count, dbg_ints);
}

int strsim_b200_compute_host(int measure, const strsim_view_chunk* a, size_t n_a, const strsim_view_chunk* b,
                             size_t n_b, double* out_values, uint8_t* out_validity, int64_t* out_null_count,
                             int32_t* dbg_ints) {
    double* outs[1] = {out_values};
    int32_t* dbgs[1] = {dbg_ints};
    return strsim_b200_compute_host_multi(&measure, 1, a, n_a, b, n_b, outs, out_validity, out_null_count, dbgs);
}

}  // extern "C"
