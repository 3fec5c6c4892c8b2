// exp/stage_bw.cu -- how fast can host threads move PAGEABLE memory into a pinned staging ring, and does
// the DMA out of the ring keep up?  (round 2: the plugin's pageable upload path runs at 29 GB/s against
// 55 GB/s of PCIe; this separates the CPU copy from the DMA.)
//   nvcc -O3 -o /tmp/stage_bw exp/stage_bw.cu -lpthread && /tmp/stage_bw
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void nt_copy(void* dst, const void* src, size_t bytes) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    size_t i = 0;
    for (; i + 64 <= bytes; i += 64) {
        __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    if (i < bytes) memcpy(d + i, s + i, bytes - i);
    _mm_sfence();
}

int main() {
    const size_t total = 512u << 20, slot = 4u << 20, ring_slots = 16;
    char* src = static_cast<char*>(malloc(total));
    memset(src, 1, total);
    char* dev;
    cudaMalloc(&dev, total);
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int wc = 0; wc < 2; wc++) {
        char* ring;
        cudaHostAlloc(&ring, slot * ring_slots, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
        for (int nt = 0; nt < 2; nt++)
            for (int threads : {2, 4, 6, 8, 12}) {
                for (int dma = 0; dma < 2; dma++) {
                    // each thread takes slots t, t+T, ...; with dma the main thread sends every filled slot
                    std::vector<std::atomic<int>> filled(total / slot);
                    for (auto& f : filled) f.store(0);
                    cudaEvent_t ev[16];
                    for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
                    std::atomic<long> freed{(long)ring_slots};  // chunks whose ring slot may be overwritten
                    const auto t0 = std::chrono::steady_clock::now();
                    std::vector<std::thread> pool;
                    for (int t = 0; t < threads; t++)
                        pool.emplace_back([&, t] {
                            for (size_t c = (size_t)t; c < total / slot; c += (size_t)threads) {
                                while ((long)c >= freed.load(std::memory_order_acquire)) std::this_thread::yield();
                                char* d = ring + (c % ring_slots) * slot;
                                if (nt) nt_copy(d, src + c * slot, slot);
                                else memcpy(d, src + c * slot, slot);
                                filled[c].store(1, std::memory_order_release);
                            }
                        });
                    size_t done = 0, issued = 0;
                    auto poll_done = [&] {
                        while (done < issued && cudaEventQuery(ev[done % ring_slots]) == cudaSuccess) {
                            done++;
                            freed.store((long)(done + ring_slots), std::memory_order_release);
                        }
                    };
                    for (size_t c = 0; c < total / slot; c++) {
                        while (!filled[c].load(std::memory_order_acquire)) {
                            if (dma) poll_done();
                            std::this_thread::yield();
                        }
                        if (dma) {
                            cudaMemcpyAsync(dev + c * slot, ring + (c % ring_slots) * slot, slot, cudaMemcpyHostToDevice, st);
                            cudaEventRecord(ev[c % ring_slots], st);
                            issued = c + 1;
                            poll_done();
                        } else {
                            freed.store((long)(c + 1 + ring_slots), std::memory_order_release);
                        }
                    }
                    for (auto& th : pool) th.join();
                    cudaStreamSynchronize(st);
                    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    printf("ring %s  copy %s  threads %2d  dma %d : %6.1f GB/s\n", wc ? "WC " : "std", nt ? "nt    " : "memcpy",
                           threads, dma, total / s / 1e9);
                    for (auto& e : ev) cudaEventDestroy(e);
                }
            }
        cudaFreeHost(ring);
    }
    return 0;
}
