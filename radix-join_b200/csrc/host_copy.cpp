// Host-side copy of one 8 KB page (the gather into / scatter out of the pinned staging buffers of
// rj_execute_pages, engine.cu).  Compiled by the host compiler alone so that the AVX2 / AVX-512
// variants can use target attributes; the variant is picked once per process from the CPU's features.
//
// Why not memcpy: neither the staging buffers (read next by the DMA engine) nor fresh result pages (read
// by the caller much later) are wanted in the cache, and a plain store makes the core read the
// destination line first.  Non-temporal stores cut the memory traffic of every page copy by a third;
// measured on the B200 box's 16 host cores (tools/host_probe.cpp, 2 GiB of pages, all cores):
// memcpy 52 GB/s, 16-byte NT 78, 32-byte NT 83, 64-byte NT 87 GB/s.  RJ_NT_COPY=0 forces memcpy.
#include <immintrin.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace rj {
namespace {

constexpr size_t kPage = 8192;

void copy_memcpy(void* d, const void* s) { std::memcpy(d, s, kPage); }

// destination 16-byte aligned (what malloc / cudaMallocHost return)
void copy_nt16(void* d, const void* s) {
    const char* a = static_cast<const char*>(s);
    char*       b = static_cast<char*>(d);
    for (size_t i = 0; i < kPage; i += 64) {
        const __m128i x0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(a + i));
        const __m128i x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(a + i + 16));
        const __m128i x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(a + i + 32));
        const __m128i x3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(a + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(b + i), x0);
        _mm_stream_si128(reinterpret_cast<__m128i*>(b + i + 16), x1);
        _mm_stream_si128(reinterpret_cast<__m128i*>(b + i + 32), x2);
        _mm_stream_si128(reinterpret_cast<__m128i*>(b + i + 48), x3);
    }
}

__attribute__((target("avx512f"))) void copy_nt64(void* d, const void* s) {
    const char* a = static_cast<const char*>(s);
    char*       b = static_cast<char*>(d);
    size_t      n = kPage;
    while (reinterpret_cast<uintptr_t>(b) & 63) { // 16-byte steps up to the cache-line boundary
        _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
        a += 16; b += 16; n -= 16;
    }
    for (; n >= 256; n -= 256, a += 256, b += 256) {
        const __m512i x0 = _mm512_loadu_si512(a), x1 = _mm512_loadu_si512(a + 64), x2 = _mm512_loadu_si512(a + 128), x3 = _mm512_loadu_si512(a + 192);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b), x0);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 64), x1);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 128), x2);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(b + 192), x3);
    }
    for (; n >= 64; n -= 64, a += 64, b += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(b), _mm512_loadu_si512(a));
    for (; n >= 16; n -= 16, a += 16, b += 16) _mm_stream_si128(reinterpret_cast<__m128i*>(b), _mm_loadu_si128(reinterpret_cast<const __m128i*>(a)));
}

using CopyFn = void (*)(void*, const void*);
CopyFn pick() {
    const char* env = std::getenv("RJ_NT_COPY");
    if (env && std::atoi(env) == 0) return copy_memcpy;
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && !(env && std::atoi(env) == 16)) return copy_nt64;
    return copy_nt16;
}
const CopyFn g_copy = pick();

} // namespace

void copy_page(void* dst, const void* src) {
    if (reinterpret_cast<uintptr_t>(dst) & 15u) { // not what an allocator returns: take the safe road
        std::memcpy(dst, src, kPage);
        return;
    }
    g_copy(dst, src);
}

// non-temporal stores are weakly ordered: fence before another agent (DMA engine, another thread) reads
void copy_fence() { _mm_sfence(); }

} // namespace rj
