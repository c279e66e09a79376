// Host-side logic of the fused root join's launch sizing (radix-join_b200/csrc/rj_internal.h), compiled and run by
// tests/test_host_logic.py without a GPU: grid, emitting warps and the page budget must be consistent for every
// probe size, because the kernel writes a chunk's pages without bounds checks.
#include "../../radix-join_b200/csrc/rj_internal.h"

#include <cstdio>

int main() {
    using namespace rj;
    int bad = 0;
    const int sms[] = {1, 16, 132, 148};
    const uint64_t probes[] = {0, 1, 31, 1983, 1984, 1985, 100000, 10000000, 19000000, 536870912, 4294967294ull};
    const uint64_t parts[] = {1, 2, 4, 128, 512, 4096, 32768};
    for (int sm: sms)
        for (uint64_t np: probes)
            for (uint64_t p: parts)
            for (uint32_t mc: {kEmitMinChunksPerWarp, kEmitMinChunksResident}) {
                const unsigned grid = join_emit_grid(np, p, sm);
                const uint32_t act  = join_emit_active_warps(np, grid, mc);
                const uint64_t cap  = join_emit_max_chunks(np, p, sm);
                // every emitting warp may leave one partly filled chunk and one reserved (empty) chunk behind
                const uint64_t need = np / kEmitChunkRows + 2ull * act * grid;
                if (grid < 1 || grid > 2u * sm || grid > (p > 1 ? p : 1) || act < 1 || act > kEmitWarps - 1 || cap < need) {
                    std::printf("bad: sm %d np %llu parts %llu -> grid %u active %u cap %llu need %llu\n", sm, (unsigned long long)np,
                                (unsigned long long)p, grid, act, (unsigned long long)cap, (unsigned long long)need);
                    ++bad;
                }
                // the partly filled chunks stay a small fraction of the result once the probe side is large
                if (np >= 100000000ull && 2ull * act * grid * mc > np / kEmitChunkRows * 2) {
                    std::printf("waste: sm %d np %llu parts %llu -> grid %u active %u\n", sm, (unsigned long long)np, (unsigned long long)p, grid, act);
                    ++bad;
                }
            }
    // radix bits: at most kMaxTotalBits, two passes of at most kMaxPassBits
    static_assert(kMaxTotalBits <= 2 * kMaxPassBits, "two scatter passes must cover every partition count");
    std::printf("%s\n", bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
