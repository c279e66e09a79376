// Radix scatter that moves whole tuples -- the key AND the columns the root will output -- for the fused
// root join (k_join_emit.cu).  Replaces the reference's prefix sum + scatter of row ids
// (src/execute.cpp:169-184) on the path where late materialisation loses: with 512 Mi probe tuples every
// value gathered through a row id costs its own 128-byte line (profiles/r1_notes.md, "Random gathers"),
// so the columns travel with the keys through both passes instead and the join reads them sequentially.
//
// A tile-at-a-time software-write-combining scatter like scatter_tile_kernel (k_partition.cu), rebuilt
// around what bounded that kernel in ncu (long-scoreboard stalls in the load phase, three CTAs per SM
// stepping through load -> rank -> scan -> stage -> copy-out with nothing in flight between phases):
//   * every input of a tile arrives by 1-D TMA bulk copies into shared-memory windows: the key window is
//     double-buffered and the NEXT tile's keys are requested before this tile is ranked; the value /
//     validity windows are requested as soon as the previous copy-out has finished reading them, and land
//     while the tile is being ranked and staged.  No thread ever waits on a global load.
//   * the tile is not staged itself: a permutation word per output position (tile offset | partition)
//     is, and the copy-out reads keys, values and validity out of the windows through it.
//   * the global reservation of the partitions' runs (one atomic per non-empty partition) is issued
//     before the staging loop and consumed after it.
//   * 512 threads per CTA, two CTAs per SM: 32 warps hide the shared-memory latency of the ranking atomics and
//     of the permuted reads (the 256-thread version sat on short-scoreboard stalls at 31 % issue).
// Output stores are runs of consecutive tuples per partition (software write combining); L2 merges the
// partial lines of neighbouring runs.
//
// Two shapes: the FLAT pass (first pass, or the only one) reads decoded columns in row order, drops tuples
// with a NULL key (src/execute.cpp:61-83: NULL never matches) and turns the validity BITMAPS of the carried
// columns into one byte per tuple; the REGION pass (second pass) re-partitions every pass-1 region in
// place order, reading and writing value arrays and validity bytes by position.
#include "rj_common.cuh"
#include "rj_internal.h"

#include <type_traits>

namespace rj {
namespace {

constexpr int      kT     = 512;
constexpr int      kItems = 8;
constexpr uint32_t kTile  = kItems * kT; // 4096 tuples
constexpr int      kWarps = kT / 32;
constexpr int      kDepth = 4;           // output positions per thread in flight at copy-out
constexpr int      kOffBits = 12;
static_assert(kTile == (1u << kOffBits), "tile offset must fill kOffBits");

struct CarryArgs {
    const uint32_t* keys;
    const uint32_t* valid; // flat: validity bitmap of the keys (NULL: none is NULL)
    uint64_t        n;
    const uint32_t* region_start;
    const uint32_t* tile_start;
    uint32_t        n_regions;
    int             shift, bits;
    uint32_t*       cursor;
    uint32_t*       keys_out;
    const void*     val_src[2];
    void*           val_dst[2];
    int             n_flag;
    const void*     flag_src[2]; // flat: bitmaps by row; regions: bytes by position
    uint8_t*        flag_dst[2];
    // kMulti (the multi-GPU exchange): partition p belongs to owner p >> owner_shift, whose arrays -- possibly
    // in another GPU's memory, mapped through NVLink -- are listed in `dsts`; cursor[p] indexes the owner's arrays
    const struct MultiCarryDsts* dsts;
    int             owner_shift;
    // region pass reading SCATTERED inputs (the multi-GPU pull: a region's tuples lie in several peers' pass-1
    // arrays): src_tab[x][0..4] = byte addresses of (keys, value 0, value 1, flag 0, flag 1) of region x, biased so
    // that element index = the region's virtual position; region_group[x] = which cursor group x feeds
    const uint64_t* src_tab;
    const uint32_t* region_group;
    int             val_tab[2]; // table column (0 / 1) of the kernel's value 0 / 1 (the launcher puts the wider one first)
};

struct MultiCarryDsts {
    static constexpr int kKeys = 0, kVal0 = 1, kFlag0 = 3, kArrays = 5;
    void* p[kArrays][8];
};

template <int W>
struct ValT {
    using type = uint32_t;
};
template <>
struct ValT<8> {
    using type = uint64_t;
};

template <bool kRegions, int W0, int W1>
struct Layout {
    static constexpr uint32_t kKeyWin  = kTile * 4 + 16;
    static constexpr uint32_t kV0      = W0 ? kTile * W0 + 16 : 0;
    static constexpr uint32_t kV1      = W1 ? kTile * W1 + 16 : 0;
    static constexpr uint32_t kFlagWin = kRegions ? kTile + 16 : 0; // flat pass: validity bits ride in registers
    static constexpr uint32_t oKey     = 0;
    static constexpr uint32_t oV0      = 2 * kKeyWin;
    static constexpr uint32_t oV1      = oV0 + kV0;
    static constexpr uint32_t oFlag    = oV1 + kV1;
    static constexpr uint32_t oPerm    = oFlag + 2 * kFlagWin;
    static constexpr uint32_t kBytes   = oPerm + kTile * 4;
};

struct Tile {
    uint32_t lo, cnt, cbase, reg;
};

// T = threads per CTA: 512 for the flat pass (32 warps per SM hide the latency of the ranking atomics), 256 for the
// region pass (measured: 3.02 vs 3.25 ms on config 2's probe side -- its windows arrive later, fewer threads per
// barrier help more than more warps)
template <bool kRegions, int W0, int W1, bool kMulti = false, int T = 512>
__global__ void __launch_bounds__(T, (W0 + W1 > 12) ? 1 : 2) scatter_carry_kernel(const CarryArgs a) {
    static_assert(!(kRegions && kMulti), "the exchange is a flat pass");
    constexpr int kT = T, kItems = static_cast<int>(kTile) / T, kWarps = T / 32, kDepth = T == 256 ? 8 : 4;
    using L  = Layout<kRegions, W0, W1>;
    using T0 = typename ValT<W0>::type;
    using T1 = typename ValT<W1>::type;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* const perm = reinterpret_cast<uint32_t*>(smem + L::oPerm);
    __shared__ uint32_t s_count[257]; // [nb] = dropped tuples
    __shared__ uint32_t s_start[257];
    __shared__ uint32_t s_gbase[256];
    __shared__ uint32_t s_warp_sums[kWarps];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_region_start[kRegions ? 258 : 1];
    __shared__ uint32_t s_tile_start[kRegions ? 258 : 1];
    __shared__ Tile s_next_tile[2]; // regions: the next tile's description, by parity of the iteration
    __shared__ __align__(8) uint64_t s_kbar[2];
    __shared__ __align__(8) uint64_t s_vbar;
    __shared__ void* s_dst[kMulti ? MultiCarryDsts::kArrays : 1][8];
    if (kMulti) {
        for (uint32_t i = threadIdx.x; i < MultiCarryDsts::kArrays * 8; i += kT) s_dst[i >> 3][i & 7] = a.dsts->p[i >> 3][i & 7];
    }

    const uint32_t nb = 1u << a.bits, mask = nb - 1;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    const int      n_flag = a.n_flag;
    const bool     has_win = W0 != 0 || W1 != 0 || (kRegions && n_flag > 0); // windows that arrive by TMA

    if (tid == 0) {
        mbar_init(&s_kbar[0], 1);
        mbar_init(&s_kbar[1], 1);
        mbar_init(&s_vbar, 1);
        fence_mbar_init();
    }
    uint64_t n_tiles;
    if (kRegions) {
        for (uint32_t r = tid; r <= a.n_regions; r += kT) {
            s_region_start[r] = a.region_start[r];
            s_tile_start[r]   = a.tile_start[r];
        }
        __syncthreads();
        n_tiles = s_tile_start[a.n_regions];
    } else {
        n_tiles = (a.n + kTile - 1) / kTile;
    }
    for (uint32_t b = tid; b <= nb; b += kT) s_count[b] = 0;
    __syncthreads();

    auto describe = [&](uint64_t t) -> Tile {
        Tile d;
        if (kRegions) {
            uint32_t x = 0, y = a.n_regions;
            while (y - x > 1) {
                const uint32_t m = (x + y) >> 1;
                if (s_tile_start[m] <= t) x = m; else y = m;
            }
            d.lo = s_region_start[x] + static_cast<uint32_t>(t - s_tile_start[x]) * kTile;
            const uint32_t left = s_region_start[x + 1] - d.lo;
            d.cnt   = left < kTile ? left : kTile;
            d.reg   = x;
            d.cbase = (a.region_group != nullptr ? a.region_group[x] : x) << a.bits;
        } else {
            d.lo = static_cast<uint32_t>(t * kTile);
            const uint64_t left = a.n - t * kTile;
            d.cnt   = left < kTile ? static_cast<uint32_t>(left) : kTile;
            d.cbase = 0;
            d.reg   = 0;
        }
        return d;
    };
    // Source address of a tile's first element in array `which` (0 keys, 1 / 2 values, 3 / 4 flags) of width w.
    // A region tile starts at an arbitrary element: its window is loaded from the 16-byte boundary below
    // (skew = elements skipped) and is up to 16 bytes longer.
    auto src_addr = [&](const Tile& d, int which, uint32_t w) -> uint64_t {
        uint64_t base;
        if (kRegions && a.src_tab != nullptr) base = a.src_tab[static_cast<uint64_t>(d.reg) * 5 + ((which == 1 || which == 2) ? 1 + a.val_tab[which - 1] : which)];
        else base = reinterpret_cast<uint64_t>(which == 0 ? static_cast<const void*>(a.keys) : (which < 3 ? a.val_src[which - 1] : a.flag_src[which - 3]));
        return base + static_cast<uint64_t>(d.lo) * w;
    };
    auto skew_of = [&](uint64_t addr, uint32_t w) -> uint32_t { return kRegions ? static_cast<uint32_t>(addr & 15u) / w : 0u; };
    auto round16 = [](uint32_t bytes) -> uint32_t { return (bytes + 15u) & ~15u; };

    auto issue_keys = [&](const Tile& d, int b) { // one thread
        const uint64_t addr  = src_addr(d, 0, 4);
        const uint32_t off   = skew_of(addr, 4);
        const uint32_t bytes = round16((d.cnt + off) * 4);
        mbar_arrive_expect_tx(&s_kbar[b], bytes);
        tma_load_1d(smem + L::oKey + b * L::kKeyWin, reinterpret_cast<const void*>(addr - off * 4u), bytes, &s_kbar[b]);
    };
    auto issue_wins = [&](const Tile& d) { // one thread
        uint32_t bytes[4] = {0, 0, 0, 0};
        uint64_t addr[4] = {0, 0, 0, 0};
        if (W0) { addr[0] = src_addr(d, 1, W0 ? W0 : 4); bytes[0] = round16((d.cnt + skew_of(addr[0], W0 ? W0 : 4)) * W0); }
        if (W1) { addr[1] = src_addr(d, 2, W1 ? W1 : 4); bytes[1] = round16((d.cnt + skew_of(addr[1], W1 ? W1 : 4)) * W1); }
#pragma unroll
        for (int f = 0; f < 2; ++f)
            if (kRegions && f < n_flag) { addr[2 + f] = src_addr(d, 3 + f, 1); bytes[2 + f] = round16(d.cnt + skew_of(addr[2 + f], 1)); }
        mbar_arrive_expect_tx(&s_vbar, bytes[0] + bytes[1] + bytes[2] + bytes[3]);
        if (W0) tma_load_1d(smem + L::oV0, reinterpret_cast<const void*>(addr[0] & ~15ull), bytes[0], &s_vbar);
        if (W1) tma_load_1d(smem + L::oV1, reinterpret_cast<const void*>(addr[1] & ~15ull), bytes[1], &s_vbar);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            if (kRegions && f < n_flag)
                tma_load_1d(smem + L::oFlag + f * L::kFlagWin, reinterpret_cast<const void*>(addr[2 + f] & ~15ull), bytes[2 + f], &s_vbar);
        }
    };

    uint64_t t = blockIdx.x;
    if (t >= n_tiles) return;
    Tile cur = describe(t);
    if (tid == 0) {
        issue_keys(cur, 0);
        if (has_win) issue_wins(cur);
    }
    for (uint32_t it = 0; t < n_tiles; t += gridDim.x, ++it) {
        const int      b = it & 1;
        const uint64_t tn = t + gridDim.x;
        const bool     has_next = tn < n_tiles;
        Tile           nxt = cur;
        if (has_next) {
            // regions: finding a tile's region is a search over up to 256 region starts -- done by the one thread
            // that needs the answer now (to request the keys); the others pick it up behind the tile's last barrier
            if (!kRegions || tid == 0) nxt = describe(tn);
            if (kRegions && tid == 0) s_next_tile[b] = nxt;
            // buffer b^1 held the previous tile's keys: their last read is behind the barrier that ended its copy-out
            if (tid == 0) issue_keys(nxt, b ^ 1);
        }
        const uint32_t* __restrict__ kw = reinterpret_cast<const uint32_t*>(smem + L::oKey + b * L::kKeyWin) + skew_of(src_addr(cur, 0, 4), 4);
        const uint32_t cnt = cur.cnt;
        mbar_wait(&s_kbar[b], (it >> 1) & 1);

        // flat pass: validity of the carried columns.  Item k of this warp covers the 32 rows of bitmap word
        // k * kWarps + warp of the tile (lo is a multiple of the tile), bit = lane: lane k fetches that word now
        // (a window of a streamed table starts on a word, not a 16-byte, boundary: no TMA here) and hands it
        // round at staging time.
        uint32_t fword[2] = {0u, 0u};
        if (!kRegions) {
#pragma unroll
            for (int f = 0; f < 2; ++f)
                if (f < n_flag && lane < kItems && (lane * kWarps + warp) * 32 < cnt)
                    fword[f] = static_cast<const uint32_t*>(a.flag_src[f])[(cur.lo >> 5) + lane * kWarps + warp];
        }

        // 1) partition of every tuple (nb = dropped: NULL key, or past the end of a partial tile) and its
        //    rank inside the partition: shared-memory atomic, warp-aggregated when the warp is skewed
        uint32_t pr[kItems];
        auto classify = [&](auto full_c) {
            constexpr bool kFull = decltype(full_c)::value;
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t off = k * kT + tid;
                bool ok = kFull || off < cnt;
                const uint32_t key = kw[off]; // past cnt: stale bytes of an earlier tile, never used
                if (!kRegions && a.valid != nullptr) {
                    // lo is a multiple of the tile: the 32 tuples of this warp and item share one bitmap word
                    const uint32_t w = a.valid[(cur.lo >> 5) + k * kWarps + warp];
                    ok = ok && ((w >> lane) & 1u);
                }
                const uint32_t digit = (hash_key(key) >> a.shift) & mask;
                pr[k] = ok ? digit : nb;
            }
        };
        if (cnt == kTile) classify(std::true_type{}); else classify(std::false_type{});
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
            pr[k] = (part << kOffBits) | rank;
        }
        __syncthreads();

        // 2) exclusive scan of the counts; the global runs are reserved now and their bases used after staging.  Only
        //    the warps that hold a bin take part (4 of 16 at 7 bits); the others go straight to the barriers
        uint32_t g = 0, start = 0;
        {
            const uint32_t scan_warps = (nb + 31u) >> 5;
            uint32_t c = 0, inc = 0;
            if (warp < scan_warps) {
                c = tid < nb ? s_count[tid] : 0u;
                inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(RJ_FULL_MASK, inc, d);
                    if (lane >= d) inc += o;
                }
                if (lane == 31) s_warp_sums[warp] = inc;
                if (tid < nb && c) g = atomicAdd(&a.cursor[cur.cbase + tid], c);
            }
            __syncthreads();
            if (warp < scan_warps) {
                uint32_t prefix = 0;
                for (uint32_t w = 0; w < warp; ++w) prefix += s_warp_sums[w];
                start = prefix + inc - c;
                if (tid < nb) {
                    s_start[tid] = start;
                    s_count[tid] = 0; // ready for the next tile
                }
                if (tid == nb - 1) {
                    s_total     = prefix + inc;
                    s_start[nb] = prefix + inc; // dropped tuples: behind everything that is copied out
                    s_count[nb] = 0;
                }
            }
        }
        __syncthreads();

        // 3) the permutation: output position inside the tile -> (tile offset | partition << 12 | validity << 30)
        if (kRegions && has_win && n_flag > 0) mbar_wait(&s_vbar, it & 1); // the validity bytes are read here
        uint32_t fsk[2] = {0u, 0u};
#pragma unroll
        for (int f = 0; f < 2; ++f)
            if (kRegions && f < n_flag) fsk[f] = skew_of(src_addr(cur, 3 + f, 1), 1);
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t part = pr[k] >> kOffBits;
            const uint32_t pos  = s_start[part] + (pr[k] & (kTile - 1));
            uint32_t fl = 0;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                if (f < n_flag) {
                    if (kRegions) fl |= (smem[L::oFlag + f * L::kFlagWin + fsk[f] + k * kT + tid] ? 1u : 0u) << f;
                    else fl |= ((__shfl_sync(RJ_FULL_MASK, fword[f], k) >> lane) & 1u) << f;
                }
            }
            perm[pos] = (fl << 30) | (part << kOffBits) | (k * kT + tid);
        }
        if (tid < nb) s_gbase[tid] = g - start;
        __syncthreads();

        // 4) copy out: consecutive threads write consecutive positions of a partition's run
        const uint32_t total = s_total;
        if (has_win && !(kRegions && n_flag > 0)) mbar_wait(&s_vbar, it & 1);
        const T0* __restrict__ v0 = reinterpret_cast<const T0*>(smem + L::oV0) + (W0 ? skew_of(src_addr(cur, 1, W0 ? W0 : 4), W0 ? W0 : 4) : 0u);
        const T1* __restrict__ v1 = reinterpret_cast<const T1*>(smem + L::oV1) + (W1 ? skew_of(src_addr(cur, 2, W1 ? W1 : 4), W1 ? W1 : 4) : 0u);
        auto copy_out = [&](auto pred_c, uint32_t base) {
            constexpr bool kPred = decltype(pred_c)::value;
            uint32_t off[kDepth], dd[kDepth], fl[kDepth], own[kDepth];
            bool     in[kDepth];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) {
                const uint32_t pos = base + j * kT + tid;
                in[j] = !kPred || pos < total;
                const uint32_t w = in[j] ? perm[pos] : 0u;
                off[j] = w & (kTile - 1);
                fl[j]  = w >> 30;
                dd[j]  = s_gbase[(w >> kOffBits) & 0xffu] + pos;
                own[j] = kMulti ? (((w >> kOffBits) & 0xffu) >> a.owner_shift) & 7u : 0u;
            }
            {
                uint32_t kk[kDepth];
#pragma unroll
                for (int j = 0; j < kDepth; ++j) kk[j] = kw[off[j]];
#pragma unroll
                for (int j = 0; j < kDepth; ++j)
                    if (in[j]) (kMulti ? static_cast<uint32_t*>(s_dst[MultiCarryDsts::kKeys][own[j]]) : a.keys_out)[dd[j]] = kk[j];
            }
            if (W0) {
                T0 vv[kDepth];
#pragma unroll
                for (int j = 0; j < kDepth; ++j) vv[j] = v0[off[j]];
#pragma unroll
                for (int j = 0; j < kDepth; ++j)
                    if (in[j]) static_cast<T0*>(kMulti ? s_dst[MultiCarryDsts::kVal0][own[j]] : a.val_dst[0])[dd[j]] = vv[j];
            }
            if (W1) {
                T1 vv[kDepth];
#pragma unroll
                for (int j = 0; j < kDepth; ++j) vv[j] = v1[off[j]];
#pragma unroll
                for (int j = 0; j < kDepth; ++j)
                    if (in[j]) static_cast<T1*>(kMulti ? s_dst[MultiCarryDsts::kVal0 + 1][own[j]] : a.val_dst[1])[dd[j]] = vv[j];
            }
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                if (f < n_flag) {
#pragma unroll
                    for (int j = 0; j < kDepth; ++j)
                        // single GPU: the launcher has filled the validity bytes with 1, only the NULLs are stored (a
                        // byte store costs the LSU as much as a key store: 12 % of this kernel's stall samples when every
                        // tuple stored its byte)
                        if (in[j] && (kMulti || !((fl[j] >> f) & 1u)))
                            (kMulti ? static_cast<uint8_t*>(s_dst[MultiCarryDsts::kFlag0 + f][own[j]]) : a.flag_dst[f])[dd[j]] = static_cast<uint8_t>((fl[j] >> f) & 1u);
                }
            }
        };
        uint32_t base = 0;
        for (; base + kDepth * kT <= total; base += kDepth * kT) copy_out(std::false_type{}, base);
        if (base < total) copy_out(std::true_type{}, base);
        __syncthreads(); // every read of the windows and of perm is done
        if (kRegions && has_next) nxt = s_next_tile[b];
        if (has_next && has_win && tid == 0) issue_wins(nxt);
        cur = nxt;
    }
}

template <bool kRegions, int W0, int W1, bool kMulti = false>
void launch_one(const CarryArgs& a, uint64_t tiles_upper, int sm_count, cudaStream_t s) {
    constexpr int kT = kRegions ? 256 : 512;
    auto         kern = scatter_carry_kernel<kRegions, W0, W1, kMulti, kT>;
    const size_t smem = Layout<kRegions, W0, W1>::kBytes;
    static SmemConfigured cfg;
    cfg.ensure(kern, smem);
    int per_sm = 0;
    RJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kT, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * per_sm;
    const unsigned blocks = static_cast<unsigned>(tiles_upper < cap ? tiles_upper : cap);
    kern<<<blocks, kT, smem, s>>>(a);
    RJ_LAUNCH_CHECK();
}

template <bool kRegions, bool kMulti = false>
void dispatch(const CarryArgs& a, int w0, int w1, uint64_t tiles_upper, int sm_count, cudaStream_t s) {
    const int code = w0 * 10 + w1;
    switch (code) {
    case 0:  launch_one<kRegions, 0, 0, kMulti>(a, tiles_upper, sm_count, s); break;
    case 40: launch_one<kRegions, 4, 0, kMulti>(a, tiles_upper, sm_count, s); break;
    case 80: launch_one<kRegions, 8, 0, kMulti>(a, tiles_upper, sm_count, s); break;
    case 44: launch_one<kRegions, 4, 4, kMulti>(a, tiles_upper, sm_count, s); break;
    case 84: launch_one<kRegions, 8, 4, kMulti>(a, tiles_upper, sm_count, s); break;
    case 88: launch_one<kRegions, 8, 8, kMulti>(a, tiles_upper, sm_count, s); break;
    default: throw CudaError("scatter_carry: unsupported column widths");
    }
}

} // namespace

void launch_scatter_carry(const CarryScatter& c, int sm_count, cudaStream_t s) {
    if (c.n == 0) return;
    if (c.bits < 1 || c.bits > kMaxPassBits) throw CudaError("scatter_carry: 1..8 radix bits per pass");
    if (c.n_val < 0 || c.n_val > 2 || c.n_flag < 0 || c.n_flag > 2) throw CudaError("scatter_carry: at most two value and two validity columns");
    CarryArgs a{};
    a.keys = c.keys; a.valid = c.valid; a.n = c.n;
    a.region_start = c.region_start; a.tile_start = c.tile_start; a.n_regions = c.n_regions;
    a.shift = c.shift; a.bits = c.bits; a.cursor = c.cursor; a.keys_out = c.keys_out;
    a.n_flag = c.n_flag;
    a.src_tab = c.src_tab;
    a.region_group = c.region_group;
    for (int f = 0; f < c.n_flag; ++f) {
        a.flag_src[f] = c.flag_src[f];
        a.flag_dst[f] = c.flag_dst[f];
    }
    // the wider column first (the kernel is instantiated for w0 >= w1)
    int order[2] = {0, 1};
    if (c.n_val == 2 && c.val_width[1] > c.val_width[0]) { order[0] = 1; order[1] = 0; }
    int w[2] = {0, 0};
    for (int i = 0; i < c.n_val; ++i) {
        const int k = order[i];
        if (c.val_width[k] != 4 && c.val_width[k] != 8) throw CudaError("scatter_carry: value columns are 4 or 8 bytes wide");
        if (c.src_tab == nullptr && reinterpret_cast<uintptr_t>(c.val_src[k]) % 16 != 0) throw CudaError("scatter_carry: value columns must be 16-byte aligned");
        a.val_src[i] = c.val_src[k];
        a.val_dst[i] = c.val_dst[k];
        a.val_tab[i] = k;
        w[i] = c.val_width[k];
    }
    if (c.src_tab == nullptr && reinterpret_cast<uintptr_t>(c.keys) % 16 != 0) throw CudaError("scatter_carry: keys must be 16-byte aligned");
    if (c.src_tab != nullptr && c.region_start == nullptr) throw CudaError("scatter_carry: a source table needs regions");
    const bool regions = c.region_start != nullptr;
    // regions: the exact tile count lives on the device (tile_start[n_regions]); size the grid from its upper bound
    const uint64_t tiles_upper = (c.n + kTile - 1) / kTile + (regions ? c.n_regions : 0);
    if (c.n_owners > 0) {
        // the multi-GPU exchange: the destination table (~320 bytes of pointers) lives in device memory for the launch
        if (regions) throw CudaError("scatter_carry: the exchange is a flat pass");
        if (c.n_owners > 8) throw CudaError("scatter_carry: at most 8 owners");
        MultiCarryDsts h{};
        for (int o = 0; o < c.n_owners; ++o) {
            h.p[MultiCarryDsts::kKeys][o] = c.keys_dst_multi[o];
            for (int i = 0; i < c.n_val; ++i) h.p[MultiCarryDsts::kVal0 + i][o] = c.val_dst_multi[order[i]][o];
            for (int f = 0; f < c.n_flag; ++f) h.p[MultiCarryDsts::kFlag0 + f][o] = c.flag_dst_multi[f][o];
        }
        static thread_local MultiCarryDsts* d_dsts_of[64] = {};
        int dev = 0;
        RJ_CUDA(cudaGetDevice(&dev));
        MultiCarryDsts*& d_dsts = d_dsts_of[dev & 63];
        if (!d_dsts) RJ_CUDA(cudaMalloc(&d_dsts, 2 * sizeof(MultiCarryDsts)));
        // two slots, used alternately: the previous launch may still be reading its table
        static thread_local unsigned flip[64] = {};
        MultiCarryDsts* slot = d_dsts + (flip[dev & 63]++ & 1u);
        RJ_CUDA(cudaMemcpyAsync(slot, &h, sizeof(MultiCarryDsts), cudaMemcpyHostToDevice, s));
        a.dsts = slot;
        a.owner_shift = c.owner_shift;
        dispatch<false, true>(a, w[0], w[1], tiles_upper, sm_count, s);
        return;
    }
    // validity bytes: all ones, the kernel stores the NULLs (see copy_out)
    for (int f = 0; f < c.n_flag; ++f) RJ_CUDA(cudaMemsetAsync(c.flag_dst[f], 1, c.n, s));
    if (regions) dispatch<true>(a, w[0], w[1], tiles_upper, sm_count, s);
    else dispatch<false>(a, w[0], w[1], tiles_upper, sm_count, s);
}

} // namespace rj
