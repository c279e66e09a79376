// Radix partitioning -- the GPU replacement of the reference's serial histogram / prefix / scatter
// (src/execute.cpp:85-92 bucket count, :124-132 histogram, :169-184 prefix sum + scatter of row ids).
//
// Differences by design (none observable in results):
//   * the fan-out is sized for SHARED MEMORY, not a 1 MiB CPU L2: partitions target <= 2048 build
//     tuples so that the build side of one partition fits an 8192-slot table in one CTA's smem;
//   * (key, row id) pairs are scattered, not bare row ids, so build/probe never gather keys again;
//   * up to 16 radix bits in at most two passes of <= 8 bits; ONE histogram kernel over all bits
//     yields every offset of both passes;
//   * NULL keys are dropped here (execute.cpp:61-83: a NULL key never matches).
//
// Scatter kernel: a CTA takes a tile of 4096 tuples, ranks every tuple inside its partition with a
// shared-memory atomic (warp-aggregated through __match_any_sync when a warp is dominated by one
// partition -- skewed keys), stages the tile in shared memory IN PARTITION ORDER (software
// write-combining), reserves each partition's run with one global atomic, and streams the runs out
// with coalesced stores: consecutive threads write consecutive addresses inside a run.
#include "rj_common.cuh"
#include "rj_internal.h"

#include <type_traits>

namespace rj {
namespace {

constexpr int kHistThreads    = 512;
constexpr int kScatterThreads = 256;
template <typename K>
struct ScatterCfg {
    // tuples per thread: 16 x 4-byte keys or 8 x 8-byte keys keep the kernel under 64 registers
    static constexpr int      kItems = sizeof(K) == 4 ? 16 : 8;
    static constexpr uint32_t kTile  = kItems * kScatterThreads;
};
constexpr int kPlanThreads    = 1024;

template <typename K>
__global__ void __launch_bounds__(1024)
    radix_hist_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ valid, uint64_t n, int shift,
                      int bits, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_hist[];
    const uint32_t nb = 1u << bits, mask = nb - 1;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (sizeof(K) == 4 && (reinterpret_cast<uintptr_t>(keys) & 15u) == 0) {
        // 4-byte keys: 16-byte loads, four of them in flight per thread (64 KB per SM in flight: one CTA of 1024
        // threads per SM has to cover the whole bandwidth-delay product); four consecutive keys share one
        // validity word
        const uint4*   keys4 = reinterpret_cast<const uint4*>(keys);
        const uint64_t n4 = n / 4;
        auto count4 = [&](const uint4& k, uint64_t j) {
            uint32_t vb = 0xfu;
            if (valid) vb = (valid[(j * 4) >> 5] >> ((j * 4) & 31u)) & 0xfu;
            if (vb & 1u) atomicAdd(&s_hist[(hash_key(k.x) >> shift) & mask], 1u);
            if (vb & 2u) atomicAdd(&s_hist[(hash_key(k.y) >> shift) & mask], 1u);
            if (vb & 4u) atomicAdd(&s_hist[(hash_key(k.z) >> shift) & mask], 1u);
            if (vb & 8u) atomicAdd(&s_hist[(hash_key(k.w) >> shift) & mask], 1u);
        };
        uint64_t j = i;
        for (; j + 3 * stride < n4; j += 4 * stride) {
            const uint4 a = ld_stream_u128(keys4 + j), b = ld_stream_u128(keys4 + j + stride), c = ld_stream_u128(keys4 + j + 2 * stride),
                        d = ld_stream_u128(keys4 + j + 3 * stride);
            count4(a, j);
            count4(b, j + stride);
            count4(c, j + 2 * stride);
            count4(d, j + 3 * stride);
        }
        for (; j < n4; j += stride) count4(ld_stream_u128(keys4 + j), j);
        // the last n % 4 keys
        for (uint64_t t = n4 * 4 + i; t < n; t += stride)
            if (!valid || test_bit(valid, t)) atomicAdd(&s_hist[(hash_key(keys[t]) >> shift) & mask], 1u);
        i = n; // nothing left for the scalar loops below
    }
    // 4 independent loads in flight per thread
    for (; i + 3 * stride < n; i += 4 * stride) {
        K k0 = keys[i], k1 = keys[i + stride], k2 = keys[i + 2 * stride], k3 = keys[i + 3 * stride];
        bool v0 = true, v1 = true, v2 = true, v3 = true;
        if (valid) {
            v0 = test_bit(valid, i);
            v1 = test_bit(valid, i + stride);
            v2 = test_bit(valid, i + 2 * stride);
            v3 = test_bit(valid, i + 3 * stride);
        }
        if (v0) atomicAdd(&s_hist[(hash_key(k0) >> shift) & mask], 1u);
        if (v1) atomicAdd(&s_hist[(hash_key(k1) >> shift) & mask], 1u);
        if (v2) atomicAdd(&s_hist[(hash_key(k2) >> shift) & mask], 1u);
        if (v3) atomicAdd(&s_hist[(hash_key(k3) >> shift) & mask], 1u);
    }
    for (; i < n; i += stride) {
        if (!valid || test_bit(valid, i)) atomicAdd(&s_hist[(hash_key(keys[i]) >> shift) & mask], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        uint32_t c = s_hist[b];
        if (c) atomicAdd(&hist[b], c);
    }
}

// ---- single-block planner ---------------------------------------------------------------------------
// exclusive scan of f(i) for i in [0, n) into out[0..n] (out[n] = total); whole block cooperates
template <class F>
__device__ void block_exclusive_scan(uint32_t n, uint32_t* out, F f) {
    __shared__ uint32_t warp_sums[kPlanThreads / 32];
    __shared__ uint32_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += kPlanThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? f(i) : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        uint32_t prefix = 0, total = 0;
        for (uint32_t w = 0; w < kPlanThreads / 32; ++w) {
            uint32_t s = warp_sums[w];
            if (w < warp) prefix += s;
            total += s;
        }
        const uint32_t carry = s_carry;
        if (i < n) out[i] = carry + prefix + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
    __syncthreads();
}

__global__ void __launch_bounds__(kPlanThreads)
    partition_plan_kernel(const uint32_t* __restrict__ hist_b, const uint32_t* __restrict__ hist_p, uint32_t flat_b,
                          uint32_t flat_p, int total_bits, int pass1_bits, uint32_t tile, PartitionPlanDev plan, uint32_t build_cap,
                          uint32_t probe_chunk) {
    const uint32_t nparts = 1u << total_bits;
    // final offsets (partition-major layout of the fully partitioned relations)
    if (total_bits == 0) {
        if (threadIdx.x == 0) {
            plan.off_b[0] = 0;
            plan.off_b[1] = flat_b;
            plan.off_p[0] = 0;
            plan.off_p[1] = flat_p;
        }
        __syncthreads();
    } else {
        block_exclusive_scan(nparts, plan.off_b, [&](uint32_t i) { return hist_b[i]; });
        block_exclusive_scan(nparts, plan.off_p, [&](uint32_t i) { return hist_p[i]; });
        for (uint32_t i = threadIdx.x; i < nparts; i += kPlanThreads) {
            plan.cur_b[i] = plan.off_b[i];
            plan.cur_p[i] = plan.off_p[i];
        }
    }
    if (pass1_bits > 0) {
        // pass-1 regions = groups of 2^(total-pass1) consecutive final partitions
        const uint32_t nreg = 1u << pass1_bits;
        const int      b2   = total_bits - pass1_bits;
        for (uint32_t r = threadIdx.x; r <= nreg; r += kPlanThreads) {
            plan.reg_b[r] = plan.off_b[r << b2];
            plan.reg_p[r] = plan.off_p[r << b2];
            if (r < nreg) {
                plan.cur1_b[r] = plan.off_b[r << b2];
                plan.cur1_p[r] = plan.off_p[r << b2];
            }
        }
        __syncthreads();
        block_exclusive_scan(nreg, plan.tile_b, [&](uint32_t r) {
            return (plan.reg_b[r + 1] - plan.reg_b[r] + tile - 1) / tile;
        });
        block_exclusive_scan(nreg, plan.tile_p, [&](uint32_t r) {
            return (plan.reg_p[r + 1] - plan.reg_p[r] + tile - 1) / tile;
        });
    }
    // join work units: (build chunk) x (probe chunk) per partition; none if either side is empty
    block_exclusive_scan(nparts, plan.unit_start, [&](uint32_t i) {
        const uint32_t nb = plan.off_b[i + 1] - plan.off_b[i];
        const uint32_t np = plan.off_p[i + 1] - plan.off_p[i];
        if (nb == 0 || np == 0) return 0u;
        return ((nb + build_cap - 1) / build_cap) * ((np + probe_chunk - 1) / probe_chunk);
    });
    // unit -> partition for the first 2 * nparts units: the join kernels look a unit up with one load instead of a
    // search over unit_start
    for (uint32_t i = threadIdx.x; i < nparts; i += kPlanThreads) {
        const uint32_t hi = plan.unit_start[i + 1] < 2 * nparts ? plan.unit_start[i + 1] : 2 * nparts;
        for (uint32_t u = plan.unit_start[i]; u < hi; ++u) plan.unit_part[u] = i;
    }
}

// ---- scatter: flat pass, segmented pass 2, multi-GPU exchange ---------------------------------------
// Written around the instruction budget: the first version of this kernel executed ~170 thread
// instructions per tuple (ncu source view: 64-bit index arithmetic, per-tuple bounds predicates, two
// bitmap probes through 64-bit addresses, the hash recomputed at copy-out).  Here
//   * a tile that is full (all but the last of a relation / region) runs without bounds predicates, and
//     full copy-out batches run without per-position predicates;
//   * everything inside a tile is addressed by 32-bit offsets from the tile's base pointers;
//   * the staged word is  tile offset | partition << kOffBits | flags << 30, so the copy-out neither
//     re-hashes the key nor loads a separate partition id, and the row id is  tile start + offset;
//   * validity of the key and of up to two carried columns is read as ONE warp-uniform bitmap word per
//     (warp, item) in the load phase -- bit = lane -- and carried through the staged word;
//   * the skew detector runs on every fourth item.
struct TileFlags {
    int         n = 0;
    const void* src[2] = {nullptr, nullptr}; // flat pass: validity bitmaps by row; pass 2: bytes by position
    uint8_t*    dst[2] = {nullptr, nullptr}; // flat pass: one byte per scattered tuple
};

// kMulti (the multi-GPU exchange, rj_radix_scatter_multi): every partition has its own output bases,
// possibly in another GPU's memory; `dsts` holds them per output array and partition
struct MultiDsts {
    static constexpr int kKeys = 0, kRows = 1, kFlag0 = 2, kPay0 = 4, kArrays = 4 + ScatterPayload::kMax;
    void* p[kArrays][8];
};

template <typename K, bool kRegions, bool kMulti = false>
__global__ void __launch_bounds__(kScatterThreads, 4)
    scatter_tile_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ valid,
                        const uint32_t* __restrict__ idx_in, uint64_t n, const uint32_t* __restrict__ region_start,
                        const uint32_t* __restrict__ tile_start, uint32_t n_regions, int shift, int bits,
                        uint32_t* __restrict__ cursor, K* __restrict__ keys_out, uint32_t* __restrict__ idx_out,
                        ScatterPayload pay, TileFlags flags, int n_tma, const MultiDsts* __restrict__ dsts = nullptr) {
    static_assert(!(kRegions && kMulti), "the exchange is a flat pass");
    constexpr int      kItems     = ScatterCfg<K>::kItems;
    constexpr uint32_t kTile      = ScatterCfg<K>::kTile;
    constexpr int      kPartShift = 12; // staged word: tile offset | bucket << 12 | flags << 30
    constexpr int      kWarps     = kScatterThreads / 32;
    constexpr int      kDepth     = 8;  // staged positions per thread in flight at copy-out
    constexpr uint32_t kWinBytes  = kTile * 8 + 16; // one TMA window in s_pay (16 bytes of slack: unaligned region tiles)
    static_assert(kTile <= (1u << kPartShift), "tile offset and rank must fit 12 bits");
    // static shared memory: every address below is base + compile-time constant
    __shared__ __align__(16) K s_keys[kTile];
    __shared__ uint32_t s_idx[kTile];
    __shared__ uint32_t s_count[257]; // [nb] tuples of this tile per partition, [nb] = dropped tuples
    __shared__ uint32_t s_start[257]; // exclusive prefix inside the tile; the dropped ones go last
    __shared__ uint32_t s_gbase[256]; // global run start minus s_start
    __shared__ uint32_t s_warp_sums[kWarps];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_region_start[kRegions ? 258 : 1];
    __shared__ uint32_t s_tile_start[kRegions ? 258 : 1];
    // flat pass: the row windows of the first n_tma carried columns, brought in by one TMA bulk copy per
    // tile and column (kTile * 8 bytes each).  Gathering them from global memory instead costs one L1
    // wavefront per tuple (32 distinct lines per warp load) -- that, not DRAM, bounded the pass.
    extern __shared__ __align__(128) uint8_t s_pay[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ void* s_dst[kMulti ? MultiDsts::kArrays : 1][8];
    if (kMulti) {
        for (uint32_t i = threadIdx.x; i < MultiDsts::kArrays * 8; i += kScatterThreads) s_dst[i >> 3][i & 7] = dsts->p[i >> 3][i & 7];
    }

    const uint32_t nb = 1u << bits, mask = nb - 1;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    uint32_t tma_phase = 0;
    if (n_tma > 0 && tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }

    uint64_t n_tiles;
    if (kRegions) {
        for (uint32_t r = tid; r <= n_regions; r += kScatterThreads) {
            s_region_start[r] = region_start[r];
            s_tile_start[r]   = tile_start[r];
        }
        __syncthreads();
        n_tiles = s_tile_start[n_regions];
    } else {
        n_tiles = (n + kTile - 1) / kTile;
    }
    for (uint32_t b = tid; b <= nb; b += kScatterThreads) s_count[b] = 0;
    __syncthreads();

    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint64_t lo;
        uint32_t cnt, cursor_base = 0;
        bool     staged = false;
        if (kRegions) {
            uint32_t a = 0, b = n_regions;
            while (b - a > 1) {
                uint32_t m = (a + b) >> 1;
                if (s_tile_start[m] <= t) a = m; else b = m;
            }
            lo = static_cast<uint64_t>(s_region_start[a]) + (t - s_tile_start[a]) * kTile;
            const uint64_t left = s_region_start[a + 1] - lo;
            cnt = left < kTile ? static_cast<uint32_t>(left) : kTile;
            cursor_base = a << bits;
        } else {
            lo = t * kTile;
            cnt = n - lo < kTile ? static_cast<uint32_t>(n - lo) : kTile;
        }
        // carried value columns: the tile's row window of the first n_tma of them arrives by one TMA bulk
        // copy each.  A region tile starts at an arbitrary tuple, so its window starts at the 16-byte
        // boundary below it (win_off[c] = elements skipped) and is 16 bytes longer.
        uint32_t win_off[2] = {0, 0};
        staged = n_tma > 0 && cnt == kTile;
        if (staged) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (c < n_tma) win_off[c] = kRegions ? static_cast<uint32_t>(lo & (16u / static_cast<uint32_t>(pay.width[c]) - 1u)) : 0u;
            if (tid == 0) {
                // every read of the previous tile's windows is behind the __syncthreads that ended it
                uint32_t bytes = 0;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (c < n_tma) bytes += kTile * static_cast<uint32_t>(pay.width[c]) + (kRegions ? 16u : 0u);
                mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c < n_tma)
                        tma_load_1d(s_pay + c * kWinBytes, static_cast<const char*>(pay.src[c]) + (lo - win_off[c]) * pay.width[c],
                                    kTile * static_cast<uint32_t>(pay.width[c]) + (kRegions ? 16u : 0u), &s_bar);
                }
            }
        }
        // the other carried columns are gathered from this tile's row window at copy-out: start
        // pulling the window into L2 now
        if (idx_in == nullptr) {
#pragma unroll
            for (int c = 0; c < ScatterPayload::kMax; ++c) {
                if (c < pay.n && pay.width[c] > 1 && !(staged && c < n_tma)) {
                    const uint32_t bytes = cnt * static_cast<uint32_t>(pay.width[c]);
                    const char*    w0    = static_cast<const char*>(pay.src[c]) + lo * pay.width[c];
                    for (uint32_t o = tid * 128; o < bytes; o += kScatterThreads * 128) prefetch_l2(w0 + o);
                }
            }
        }
        const K* __restrict__ tkeys = keys + lo;

        // 1) load the tile; bucket per tuple (nb = dropped: NULL key or past the end of the tile)
        K        key[kItems];
        uint32_t pr[kItems]; // bucket, later bucket << 12 | rank inside the bucket
        uint32_t fl = 0;     // 2 carried-validity flags per item
        auto load_tile = [&](auto full_c) {
            constexpr bool kFull = decltype(full_c)::value;
            bool ok[kItems];
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t off = k * kScatterThreads + tid;
                ok[k]  = kFull || off < cnt;
                key[k] = ok[k] ? tkeys[off] : K(0);
            }
            if (kRegions) {
                if (flags.n > 0) {
                    const uint8_t* __restrict__ f0 = static_cast<const uint8_t*>(flags.src[0]) + lo;
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (ok[k] && f0[k * kScatterThreads + tid]) fl |= 1u << (2 * k);
                }
                if (flags.n > 1) {
                    const uint8_t* __restrict__ f1 = static_cast<const uint8_t*>(flags.src[1]) + lo;
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (ok[k] && f1[k * kScatterThreads + tid]) fl |= 2u << (2 * k);
                }
            } else {
                // lo is a multiple of the tile: item k of this warp is bit `lane` of word k * kWarps + warp
                const uint64_t w0 = (lo >> 5) + warp;
                if (flags.n > 0) {
                    const uint32_t* __restrict__ f0 = static_cast<const uint32_t*>(flags.src[0]) + w0;
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (ok[k]) fl |= ((f0[k * kWarps] >> lane) & 1u) << (2 * k);
                }
                if (flags.n > 1) {
                    const uint32_t* __restrict__ f1 = static_cast<const uint32_t*>(flags.src[1]) + w0;
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (ok[k]) fl |= ((f1[k * kWarps] >> lane) & 1u) << (2 * k + 1);
                }
                if (valid != nullptr) {
                    const uint32_t* __restrict__ v = valid + w0;
#pragma unroll
                    for (int k = 0; k < kItems; ++k)
                        if (ok[k]) ok[k] = (v[k * kWarps] >> lane) & 1u;
                }
            }
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t digit = (hash_key(key[k]) >> shift) & mask;
                pr[k] = ok[k] ? digit : nb;
            }
        };
        if (cnt == kTile) load_tile(std::true_type{}); else load_tile(std::false_type{});

        // rank inside the bucket: shared-memory atomic, warp-aggregated when the warp is skewed.  Every
        // tuple takes a rank (the dropped ones in the bucket that is staged last and never copied out),
        // so neither this loop nor the staging below carries a per-tuple predicate.
        bool skewed = false;
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t part = pr[k];
            if ((k & 3) == 0) {
                const uint32_t nbr = __shfl_xor_sync(RJ_FULL_MASK, part, 1);
                skewed = __popc(__ballot_sync(RJ_FULL_MASK, nbr == part)) >= 4;
            }
            uint32_t rank;
            if (skewed) {
                const uint32_t peers  = __match_any_sync(RJ_FULL_MASK, part);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&s_count[part], static_cast<uint32_t>(__popc(peers)));
                base = __shfl_sync(RJ_FULL_MASK, base, leader);
                rank = base + __popc(peers & lt);
            } else {
                rank = atomicAdd(&s_count[part], 1u);
            }
            pr[k] = (part << kPartShift) | rank;
        }
        __syncthreads();

        // 2) exclusive scan of the per-partition counts (nb <= kScatterThreads), reserve global runs
        {
            const uint32_t c = tid < nb ? s_count[tid] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) s_warp_sums[warp] = inc;
            __syncthreads();
            uint32_t prefix = 0;
#pragma unroll
            for (uint32_t w = 0; w < kWarps; ++w) prefix += w < warp ? s_warp_sums[w] : 0u;
            const uint32_t start = prefix + inc - c;
            if (tid < nb) {
                s_start[tid] = start;
                uint32_t g = 0;
                if (c) g = atomicAdd(&cursor[cursor_base + tid], c);
                s_gbase[tid] = g - start;
                s_count[tid] = 0; // ready for the next tile
            }
            if (tid == kScatterThreads - 1) {
                s_total     = prefix + inc;
                s_start[nb] = prefix + inc; // dropped tuples: behind everything that is copied out
                s_count[nb] = 0;
            }
        }
        __syncthreads();

        // 3) stage the tile in partition order
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t pos = s_start[pr[k] >> kPartShift] + (pr[k] & ((1u << kPartShift) - 1));
            s_keys[pos] = key[k];
            s_idx[pos]  = (pr[k] & (0x1ffu << kPartShift)) | (k * kScatterThreads + tid) |
                          ((fl << (30 - 2 * k)) & 0xc0000000u);
        }
        __syncthreads();

        // 4) stream the runs out: thread -> staged position, kDepth positions per thread in flight
        const uint32_t total = s_total;
        const uint32_t lo32  = static_cast<uint32_t>(lo);
        if (staged) {
            mbar_wait(&s_bar, tma_phase);
            tma_phase ^= 1;
        }
        auto copy_out = [&](auto pred_c, uint32_t base) {
            constexpr bool kPred = decltype(pred_c)::value;
            K        kk[kDepth];
            uint32_t rr[kDepth], dd[kDepth], ww[kDepth];
            bool     in[kDepth];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) {
                const uint32_t pos = base + j * kScatterThreads + tid;
                in[j] = !kPred || pos < total;
                kk[j] = in[j] ? s_keys[pos] : K(0);
                ww[j] = in[j] ? s_idx[pos] : 0u;
                dd[j] = s_gbase[(ww[j] >> kPartShift) & 0xffu] + pos;
                rr[j] = lo32 + (ww[j] & (kTile - 1));
            }
            if (!kRegions && idx_in != nullptr) {
#pragma unroll
                for (int j = 0; j < kDepth; ++j) rr[j] = in[j] ? idx_in[rr[j]] : 0u;
            }
            // destination of array `a` for staged position j: one base, or one base per partition
            auto out = [&](int a, void* single, int j) -> void* {
                if (kMulti) return s_dst[a][(ww[j] >> kPartShift) & 7u];
                return single;
            };
#pragma unroll
            for (int j = 0; j < kDepth; ++j)
                if (in[j]) static_cast<K*>(out(MultiDsts::kKeys, keys_out, j))[dd[j]] = kk[j];
            if (kMulti ? s_dst[MultiDsts::kRows][0] != nullptr : idx_out != nullptr) { // row ids / positions nobody reads are dropped
#pragma unroll
                for (int j = 0; j < kDepth; ++j)
                    if (in[j]) static_cast<uint32_t*>(out(MultiDsts::kRows, idx_out, j))[dd[j]] = kRegions ? (rr[j] | (ww[j] & 0xc0000000u)) : rr[j];
            }
            if (flags.n > 0 && (kMulti || flags.dst[0] != nullptr)) {
#pragma unroll
                for (int j = 0; j < kDepth; ++j) if (in[j]) static_cast<uint8_t*>(out(MultiDsts::kFlag0, flags.dst[0], j))[dd[j]] = (ww[j] >> 30) & 1u;
            }
            if (flags.n > 1 && (kMulti || flags.dst[1] != nullptr)) {
#pragma unroll
                for (int j = 0; j < kDepth; ++j) if (in[j]) static_cast<uint8_t*>(out(MultiDsts::kFlag0 + 1, flags.dst[1], j))[dd[j]] = ww[j] >> 31;
            }
            // carried payload columns: the reads stay inside this tile's row window (prefetched above)
#pragma unroll
            for (int c = 0; c < ScatterPayload::kMax; ++c) {
                if (c < pay.n) {
                    const int w = pay.width[c];
                    if (c < 2 && staged && c < n_tma) {
                        // the window is in shared memory: row offset inside the tile = low bits of ww
                        if (w == 8) {
                            const uint64_t* win = reinterpret_cast<const uint64_t*>(s_pay + c * kWinBytes) + win_off[c];
#pragma unroll
                            for (int j = 0; j < kDepth; ++j)
                                if (in[j]) static_cast<uint64_t*>(out(MultiDsts::kPay0 + c, pay.dst[c], j))[dd[j]] = win[ww[j] & (kTile - 1)];
                        } else {
                            const uint32_t* win = reinterpret_cast<const uint32_t*>(s_pay + c * kWinBytes) + win_off[c];
#pragma unroll
                            for (int j = 0; j < kDepth; ++j)
                                if (in[j]) static_cast<uint32_t*>(out(MultiDsts::kPay0 + c, pay.dst[c], j))[dd[j]] = win[ww[j] & (kTile - 1)];
                        }
                    } else if (w == 8) {
                        uint64_t v[kDepth];
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) v[j] = in[j] ? static_cast<const uint64_t*>(pay.src[c])[rr[j]] : 0ull;
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) if (in[j]) static_cast<uint64_t*>(out(MultiDsts::kPay0 + c, pay.dst[c], j))[dd[j]] = v[j];
                    } else if (w == 1) { // a validity bitmap travels as one byte per tuple
                        uint32_t v[kDepth];
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) v[j] = in[j] ? static_cast<const uint32_t*>(pay.src[c])[rr[j] >> 5] : 0u;
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) if (in[j]) static_cast<uint8_t*>(out(MultiDsts::kPay0 + c, pay.dst[c], j))[dd[j]] = (v[j] >> (rr[j] & 31)) & 1u;
                    } else {
                        uint32_t v[kDepth];
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) v[j] = in[j] ? static_cast<const uint32_t*>(pay.src[c])[rr[j]] : 0u;
#pragma unroll
                        for (int j = 0; j < kDepth; ++j) if (in[j]) static_cast<uint32_t*>(out(MultiDsts::kPay0 + c, pay.dst[c], j))[dd[j]] = v[j];
                    }
                }
            }
        };
        uint32_t base = 0;
        for (; base + kDepth * kScatterThreads <= total; base += kDepth * kScatterThreads) copy_out(std::false_type{}, base);
        if (base < total) copy_out(std::true_type{}, base);
        __syncthreads();
    }
}

// persistent grid = the CTAs that are resident at once (the windows in dynamic shared memory change that)
template <typename Kern>
unsigned resident_grid(Kern kern, size_t smem, uint64_t n_tiles, int sm_count) {
    int per_sm = 0;
    RJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kScatterThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * per_sm;
    return static_cast<unsigned>(n_tiles < cap ? n_tiles : cap);
}

} // namespace

size_t partition_plan_words(int total_bits, int pass1_bits) {
    const size_t nparts = size_t(1) << total_bits;
    const size_t nreg   = pass1_bits > 0 ? (size_t(1) << pass1_bits) : 0;
    return 2 * (nparts + 1) + 2 * nparts + (nparts + 1) + 2 * nparts + (nreg ? 2 * (nreg + 1) + 2 * nreg + 2 * (nreg + 1) : 0) + 16;
}

void partition_plan_carve(uint32_t* base, int total_bits, int pass1_bits, PartitionPlanDev* plan) {
    const size_t nparts = size_t(1) << total_bits;
    const size_t nreg   = pass1_bits > 0 ? (size_t(1) << pass1_bits) : 0;
    uint32_t* p = base;
    plan->off_b = p; p += nparts + 1;
    plan->off_p = p; p += nparts + 1;
    plan->cur_b = p; p += nparts;
    plan->cur_p = p; p += nparts;
    plan->unit_start = p; p += nparts + 1;
    plan->unit_cursor = p; p += 1;
    plan->unit_part = p; p += 2 * nparts;
    if (nreg) {
        plan->reg_b = p; p += nreg + 1;
        plan->reg_p = p; p += nreg + 1;
        plan->cur1_b = p; p += nreg;
        plan->cur1_p = p; p += nreg;
        plan->tile_b = p; p += nreg + 1;
        plan->tile_p = p; p += nreg + 1;
    } else {
        plan->reg_b = plan->reg_p = plan->cur1_b = plan->cur1_p = plan->tile_b = plan->tile_p = nullptr;
    }
}

void launch_radix_histogram(const void* keys, const uint32_t* valid, uint64_t n, int key_bytes, int shift, int bits,
                            uint32_t* hist, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    const size_t smem = sizeof(uint32_t) << bits;
    // a 2^15-bin histogram is 128 KB of shared memory: one CTA per SM, so give it 1024 threads
    const unsigned threads = smem > 48 * 1024 ? 1024 : kHistThreads;
    const unsigned per_sm  = smem > 96 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
    uint64_t want = (n + threads * 8 - 1) / (threads * 8);
    unsigned blocks = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * per_sm ? want : static_cast<uint64_t>(sm_count) * per_sm);
    if (blocks == 0) blocks = 1;
    if (key_bytes == 4) {
        static SmemConfigured cfg;
        cfg.ensure(radix_hist_kernel<uint32_t>, smem);
        radix_hist_kernel<uint32_t><<<blocks, threads, smem, s>>>(static_cast<const uint32_t*>(keys), valid, n, shift, bits, hist);
    } else {
        static SmemConfigured cfg;
        cfg.ensure(radix_hist_kernel<uint64_t>, smem);
        radix_hist_kernel<uint64_t><<<blocks, threads, smem, s>>>(static_cast<const uint64_t*>(keys), valid, n, shift, bits, hist);
    }
    RJ_LAUNCH_CHECK();
}

void launch_partition_plan(const uint32_t* hist_b, const uint32_t* hist_p, uint32_t flat_b, uint32_t flat_p,
                           int total_bits, int pass1_bits, int key_bytes, const PartitionPlanDev& plan,
                           cudaStream_t s, uint32_t build_cap, uint32_t probe_chunk) {
    partition_plan_kernel<<<1, kPlanThreads, 0, s>>>(hist_b, hist_p, flat_b, flat_p, total_bits, pass1_bits,
                                                     scatter_tile(key_bytes), plan, build_cap, probe_chunk);
    RJ_LAUNCH_CHECK();
}

void launch_radix_scatter(const void* keys, const uint32_t* valid, const uint32_t* idx_in, uint64_t n, int key_bytes,
                          int shift, int bits, uint32_t* cursor, void* keys_out, uint32_t* idx_out,
                          const ScatterPayload& payload, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    const uint32_t tile = scatter_tile(key_bytes);
    uint64_t n_tiles = (n + tile - 1) / tile;
    unsigned blocks = static_cast<unsigned>(n_tiles < static_cast<uint64_t>(sm_count) * 4 ? n_tiles : static_cast<uint64_t>(sm_count) * 4);
    // row id == input position: up to two carried validity bitmaps take the cheap path (one bitmap word
    // per warp and item in the load phase) instead of a gather per tuple at copy-out
    ScatterPayload pay;
    TileFlags      flags;
    int            n_tma = 0;
    auto push = [&](int c) {
        pay.src[pay.n] = payload.src[c];
        pay.dst[pay.n] = payload.dst[c];
        pay.width[pay.n] = payload.width[c];
        ++pay.n;
    };
    // up to two value columns travel through shared-memory windows (TMA); they come first
    auto tma_ok = [&](int c) {
        return idx_in == nullptr && payload.width[c] >= 4 && reinterpret_cast<uintptr_t>(payload.src[c]) % 16 == 0;
    };
    for (int c = 0; c < payload.n; ++c) {
        if (tma_ok(c) && n_tma < 2) {
            push(c);
            ++n_tma;
        }
    }
    {
        int taken = 0;
        for (int c = 0; c < payload.n; ++c) {
            if (tma_ok(c) && taken < 2) {
                ++taken;
            } else if (payload.width[c] == 1 && idx_in == nullptr && flags.n < 2) {
                flags.src[flags.n] = payload.src[c];
                flags.dst[flags.n] = static_cast<uint8_t*>(payload.dst[c]);
                ++flags.n;
            } else {
                push(c);
            }
        }
    }
    const size_t smem = static_cast<size_t>(n_tma) * (tile * 8 + 16);
    if (key_bytes == 4) {
        static SmemConfigured cfg;
        cfg.ensure(scatter_tile_kernel<uint32_t, false>, smem);
        blocks = resident_grid(scatter_tile_kernel<uint32_t, false>, smem, n_tiles, sm_count);
        scatter_tile_kernel<uint32_t, false><<<blocks, kScatterThreads, smem, s>>>(
            static_cast<const uint32_t*>(keys), valid, idx_in, n, nullptr, nullptr, 0, shift, bits, cursor,
            static_cast<uint32_t*>(keys_out), idx_out, pay, flags, n_tma);
    } else {
        static SmemConfigured cfg;
        cfg.ensure(scatter_tile_kernel<uint64_t, false>, smem);
        blocks = resident_grid(scatter_tile_kernel<uint64_t, false>, smem, n_tiles, sm_count);
        scatter_tile_kernel<uint64_t, false><<<blocks, kScatterThreads, smem, s>>>(
            static_cast<const uint64_t*>(keys), valid, idx_in, n, nullptr, nullptr, 0, shift, bits, cursor,
            static_cast<uint64_t*>(keys_out), idx_out, pay, flags, n_tma);
    }
    RJ_LAUNCH_CHECK();
}

void launch_radix_scatter_multi(const void* keys, const uint32_t* valid, uint64_t n, int key_bytes, int shift, int bits,
                                uint32_t* cursor, const rj_scatter_multi_t& out, int sm_count, cudaStream_t s) {
    if (n == 0) return;
    // Same slot assignment as the flat pass: up to two value columns through TMA windows (first slots),
    // up to two validity bitmaps through the per-warp bitmap words, the rest gathered.
    ScatterPayload pay;
    TileFlags      flags;
    MultiDsts      h{};
    int            n_tma = 0;
    for (int p = 0; p < 8; ++p) {
        h.p[MultiDsts::kKeys][p] = out.keys_out[p];
        h.p[MultiDsts::kRows][p] = out.rows_out[p];
    }
    auto push = [&](uint32_t c) {
        pay.src[pay.n] = out.pay_src[c];
        pay.width[pay.n] = out.pay_width[c];
        for (int p = 0; p < 8; ++p) h.p[MultiDsts::kPay0 + pay.n][p] = out.pay_dst[c][p];
        ++pay.n;
    };
    auto tma_ok = [&](uint32_t c) { return out.pay_width[c] >= 4 && reinterpret_cast<uintptr_t>(out.pay_src[c]) % 16 == 0; };
    for (uint32_t c = 0; c < out.n_payload; ++c) {
        if (tma_ok(c) && n_tma < 2) {
            push(c);
            ++n_tma;
        }
    }
    int taken = 0;
    for (uint32_t c = 0; c < out.n_payload; ++c) {
        if (tma_ok(c) && taken < 2) {
            ++taken;
        } else if (out.pay_width[c] == 1 && flags.n < 2) {
            flags.src[flags.n] = out.pay_src[c];
            for (int p = 0; p < 8; ++p) h.p[MultiDsts::kFlag0 + flags.n][p] = out.pay_dst[c][p];
            ++flags.n;
        } else {
            push(c);
        }
    }
    // the destination table (~900 bytes of pointers) lives in device memory for the duration of the launch
    static thread_local MultiDsts* d_dsts_of[64] = {};
    int dev = 0;
    RJ_CUDA(cudaGetDevice(&dev));
    MultiDsts*& d_dsts = d_dsts_of[dev & 63];
    if (!d_dsts) RJ_CUDA(cudaMalloc(&d_dsts, sizeof(MultiDsts)));
    RJ_CUDA(cudaMemcpyAsync(d_dsts, &h, sizeof(MultiDsts), cudaMemcpyHostToDevice, s));
    const uint32_t tile = scatter_tile(key_bytes);
    const uint64_t n_tiles = (n + tile - 1) / tile;
    const size_t   smem = static_cast<size_t>(n_tma) * (tile * 8 + 16);
    if (key_bytes == 4) {
        auto kern = scatter_tile_kernel<uint32_t, false, true>;
        static SmemConfigured cfg;
        cfg.ensure(kern, smem);
        const unsigned blocks = resident_grid(kern, smem, n_tiles, sm_count);
        kern<<<blocks, kScatterThreads, smem, s>>>(static_cast<const uint32_t*>(keys), valid, nullptr, n, nullptr, nullptr, 0, shift,
                                                   bits, cursor, nullptr, nullptr, pay, flags, n_tma, d_dsts);
    } else {
        auto kern = scatter_tile_kernel<uint64_t, false, true>;
        static SmemConfigured cfg;
        cfg.ensure(kern, smem);
        const unsigned blocks = resident_grid(kern, smem, n_tiles, sm_count);
        kern<<<blocks, kScatterThreads, smem, s>>>(static_cast<const uint64_t*>(keys), valid, nullptr, n, nullptr, nullptr, 0, shift,
                                                   bits, cursor, nullptr, nullptr, pay, flags, n_tma, d_dsts);
    }
    RJ_LAUNCH_CHECK();
}

void launch_radix_scatter_regions(const void* keys, const uint32_t* idx_in, const uint32_t* region_start,
                                  const uint32_t* tile_start, uint32_t n_regions, uint64_t n_upper, int key_bytes,
                                  int shift, int bits, uint32_t* cursor, void* keys_out, uint32_t* idx_out,
                                  const RegionFlags& flags, const ScatterPayload& payload, int sm_count, cudaStream_t s) {
    if (n_upper == 0) return;
    // the exact tile count lives on the device (tile_start[n_regions]); size the persistent grid from
    // its upper bound so no host synchronisation is needed
    const uint32_t tile = scatter_tile(key_bytes);
    uint64_t tiles_upper = (n_upper + tile - 1) / tile + n_regions;
    TileFlags tf;
    tf.n = flags.n;
    for (int c = 0; c < flags.n; ++c) {
        tf.src[c] = flags.src[c];
        tf.dst[c] = flags.dst[c];
    }
    // value columns that move with the tuples (positions are the input order of this pass): the first
    // two whose source is 8-byte aligned go through TMA windows, the rest is gathered from the tile's window
    ScatterPayload pay;
    int            n_tma = 0;
    auto push = [&](int c) {
        pay.src[pay.n] = payload.src[c];
        pay.dst[pay.n] = payload.dst[c];
        pay.width[pay.n] = payload.width[c];
        ++pay.n;
    };
    auto tma_ok = [&](int c) { return payload.width[c] >= 4 && reinterpret_cast<uintptr_t>(payload.src[c]) % 16 == 0; };
    for (int c = 0; c < payload.n; ++c)
        if (tma_ok(c) && n_tma < 2) {
            push(c);
            ++n_tma;
        }
    {
        int taken = 0;
        for (int c = 0; c < payload.n; ++c) {
            if (tma_ok(c) && taken < 2) ++taken; else push(c);
        }
    }
    const size_t smem = static_cast<size_t>(n_tma) * (tile * 8 + 16);
    if (key_bytes == 4) {
        auto kern = scatter_tile_kernel<uint32_t, true>;
        static SmemConfigured cfg;
        cfg.ensure(kern, smem);
        const unsigned blocks = resident_grid(kern, smem, tiles_upper, sm_count);
        kern<<<blocks, kScatterThreads, smem, s>>>(static_cast<const uint32_t*>(keys), nullptr, nullptr, n_upper, region_start, tile_start,
                                                   n_regions, shift, bits, cursor, static_cast<uint32_t*>(keys_out), idx_out, pay, tf, n_tma, nullptr);
    } else {
        auto kern = scatter_tile_kernel<uint64_t, true>;
        static SmemConfigured cfg;
        cfg.ensure(kern, smem);
        const unsigned blocks = resident_grid(kern, smem, tiles_upper, sm_count);
        kern<<<blocks, kScatterThreads, smem, s>>>(static_cast<const uint64_t*>(keys), nullptr, nullptr, n_upper, region_start, tile_start,
                                                   n_regions, shift, bits, cursor, static_cast<uint64_t*>(keys_out), idx_out, pay, tf, n_tma, nullptr);
    }
    RJ_LAUNCH_CHECK();
    (void)idx_in;
}

} // namespace rj
