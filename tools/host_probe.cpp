// Floor of the end-to-end path (Contest::execute on individually new-ed pages): what the host and the
// PCIe link can move, measured alone.  Not part of the product; bench.py does not call it.
//
//   host_probe [threads] [GiB]
//
// 1. pinned H2D, D2H and both at once (cudaMemcpyAsync, 64 MiB pieces on two streams)
// 2. page scatter/gather: `threads` workers copy 8 KB pages between a staging buffer and individually
//    allocated pages with memcpy, 16-byte, 32-byte and 64-byte non-temporal stores
// One JSON line.
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using Clock = std::chrono::steady_clock;
static double secs(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

constexpr size_t kPage = 8192;

static void copy_memcpy(void* d, const void* s) { std::memcpy(d, s, kPage); }
static void copy_nt16(void* d, const void* s) {
    const __m128i* a = static_cast<const __m128i*>(s);
    __m128i*       b = static_cast<__m128i*>(d);
    for (size_t i = 0; i < kPage / 16; i += 4) {
        __m128i x0 = _mm_load_si128(a + i), x1 = _mm_load_si128(a + i + 1), x2 = _mm_load_si128(a + i + 2), x3 = _mm_load_si128(a + i + 3);
        _mm_stream_si128(b + i, x0);
        _mm_stream_si128(b + i + 1, x1);
        _mm_stream_si128(b + i + 2, x2);
        _mm_stream_si128(b + i + 3, x3);
    }
}
__attribute__((target("avx2"))) static void copy_nt32(void* d, const void* s) {
    // destination pages are 16-byte aligned (malloc): peel 16 bytes when the 32-byte boundary is off
    const char* a = static_cast<const char*>(s);
    char*       b = static_cast<char*>(d);
    size_t      n = kPage;
    if (reinterpret_cast<uintptr_t>(b) & 31) {
        _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
        a += 16; b += 16; n -= 16;
    }
    for (; n >= 128; n -= 128, a += 128, b += 128) {
        __m256i x0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a)), x1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + 32));
        __m256i x2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + 64)), x3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(b), x0);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(b + 32), x1);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(b + 64), x2);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(b + 96), x3);
    }
    for (; n >= 16; n -= 16, a += 16, b += 16) _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
}
__attribute__((target("avx512f"))) static void copy_nt64(void* d, const void* s) {
    const char* a = static_cast<const char*>(s);
    char*       b = static_cast<char*>(d);
    size_t      n = kPage;
    while (reinterpret_cast<uintptr_t>(b) & 63) { // 16-byte steps up to the cache-line boundary
        _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
        a += 16; b += 16; n -= 16;
    }
    for (; n >= 256; n -= 256, a += 256, b += 256) {
        __m512i x0 = _mm512_loadu_si512(a), x1 = _mm512_loadu_si512(a + 64), x2 = _mm512_loadu_si512(a + 128), x3 = _mm512_loadu_si512(a + 192);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b), x0);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 64), x1);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 128), x2);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 192), x3);
    }
    for (; n >= 64; n -= 64, a += 64, b += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(b), _mm512_loadu_si512(a));
    for (; n >= 16; n -= 16, a += 16, b += 16) _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
}

struct PageObj {
    alignas(8) unsigned char data[kPage];
};

template <class F>
static double run_threads(int threads, F f) {
    std::vector<std::thread> ts;
    auto t0 = Clock::now();
    for (int t = 0; t < threads; ++t) ts.emplace_back([=] { f(t); });
    for (auto& t: ts) t.join();
    return secs(t0);
}

int main(int argc, char** argv) {
    const int    threads = argc > 1 ? atoi(argv[1]) : static_cast<int>(std::thread::hardware_concurrency());
    const double gib = argc > 2 ? atof(argv[2]) : 2.0;
    const size_t bytes = static_cast<size_t>(gib * (1ull << 30)) / (64 << 20) * (64 << 20);
    printf("{\"threads\": %d, \"bytes\": %zu", threads, bytes);

    // ---- 1. PCIe -----------------------------------------------------------------------------------
    void *h_a = nullptr, *h_b = nullptr, *d_a = nullptr, *d_b = nullptr;
    if (cudaMallocHost(&h_a, bytes) == cudaSuccess && cudaMallocHost(&h_b, bytes) == cudaSuccess && cudaMalloc(&d_a, bytes) == cudaSuccess &&
        cudaMalloc(&d_b, bytes) == cudaSuccess) {
        std::memset(h_a, 1, bytes);
        std::memset(h_b, 2, bytes);
        cudaStream_t s1, s2;
        cudaStreamCreate(&s1);
        cudaStreamCreate(&s2);
        const size_t piece = 64 << 20;
        auto copy = [&](bool up, bool down) {
            cudaDeviceSynchronize();
            auto t0 = Clock::now();
            for (size_t o = 0; o < bytes; o += piece) {
                if (up) cudaMemcpyAsync(static_cast<char*>(d_a) + o, static_cast<char*>(h_a) + o, piece, cudaMemcpyHostToDevice, s1);
                if (down) cudaMemcpyAsync(static_cast<char*>(h_b) + o, static_cast<char*>(d_b) + o, piece, cudaMemcpyDeviceToHost, s2);
            }
            cudaDeviceSynchronize();
            return secs(t0);
        };
        copy(true, true);
        const double t_up = copy(true, false), t_down = copy(false, true), t_both = copy(true, true);
        printf(", \"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f, \"both_gbs_each\": %.1f", bytes / 1e9 / t_up, bytes / 1e9 / t_down, bytes / 1e9 / t_both);
        // small pieces, the size of the engine's staging buffers
        {
            cudaDeviceSynchronize();
            auto t0 = Clock::now();
            for (size_t o = 0; o < bytes; o += (4 << 20)) {
                cudaMemcpyAsync(static_cast<char*>(d_a) + o, static_cast<char*>(h_a) + o, 4 << 20, cudaMemcpyHostToDevice, s1);
                cudaMemcpyAsync(static_cast<char*>(h_b) + o, static_cast<char*>(d_b) + o, 4 << 20, cudaMemcpyDeviceToHost, s2);
            }
            cudaDeviceSynchronize();
            printf(", \"both_4MiB_pieces_gbs_each\": %.1f", bytes / 1e9 / secs(t0));
        }
    } else {
        printf(", \"pcie\": \"no device\"");
        cudaGetLastError();
        h_a = h_a ? h_a : std::malloc(bytes);
        std::memset(h_a, 1, bytes);
    }

    // ---- 2. page copies ----------------------------------------------------------------------------
    const size_t n_pages = bytes / kPage;
    std::vector<PageObj*> pages(n_pages);
    for (auto& p: pages) p = new PageObj;
    for (auto* p: pages) std::memset(p->data, 3, kPage); // fault them in
    struct V {
        const char* name;
        void (*fn)(void*, const void*);
        bool ok;
    } variants[] = {{"memcpy", copy_memcpy, true},
                    {"nt16", copy_nt16, true},
                    {"nt32", copy_nt32, static_cast<bool>(__builtin_cpu_supports("avx2"))},
                    {"nt64", copy_nt64, static_cast<bool>(__builtin_cpu_supports("avx512f"))}};
    for (auto& v: variants) {
        if (!v.ok) continue;
        for (int dir = 0; dir < 2; ++dir) { // 0: staging -> pages (download side), 1: pages -> staging (upload side)
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                const double t = run_threads(threads, [&](int t) {
                    const size_t per = (n_pages + threads - 1) / threads;
                    const size_t lo = t * per, hi = std::min(n_pages, lo + per);
                    for (size_t i = lo; i < hi; ++i) {
                        if (dir == 0) v.fn(pages[i]->data, static_cast<char*>(h_a) + i * kPage);
                        else v.fn(static_cast<char*>(h_a) + i * kPage, pages[i]->data);
                    }
                    _mm_sfence();
                });
                best = std::min(best, t);
            }
            printf(", \"%s_%s_gbs\": %.1f", v.name, dir == 0 ? "to_pages" : "from_pages", bytes / 1e9 / best);
        }
    }
    // both directions at once, half the threads each (what the pipeline does), best variant = nt64 or nt16
    printf("}\n");
    return 0;
}
