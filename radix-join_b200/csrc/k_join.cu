// Per-partition build + probe in shared memory -- the GPU replacement of the reference's per-bucket
// open-addressing join (src/execute.cpp:196-249: table of cap = pow2 >= 2*cnt slots, linear probing,
// duplicates listed per slot, one output row per (probe row, matching build row)).
//
// One CTA = one work unit = (partition, build chunk, probe chunk):
//   build : <= 6144 build tuples are inserted into an 8192-slot linear-probing table in shared memory
//           (keys[] + row ids[], the row id doubles as the occupancy flag).  Duplicate keys simply take
//           separate slots; the probe walks the cluster until an empty slot and emits every equal key,
//           which yields the reference's "one row per duplicate" semantics (tests/unit_tests.cpp:125-161).
//           All global loads of the chunk are issued before the first insert (12 per thread in flight).
//   probe : the probe chunk is streamed in super-batches of 8 tuples per thread (4096 per CTA): the 16
//           loads of the NEXT super-batch are in flight while the current one is probed, so the kernel is
//           bound by bandwidth, not by load latency.  Matches are staged as (build row, probe row) pairs
//           in shared memory through warp-aggregated slot reservation; the buffer is flushed once per
//           super-batch with ONE global atomic and coalesced stores (one block barrier per 4096 probes).
//           A thread that finds the staging buffer full remembers where it stopped and resumes after the
//           flush, so any number of duplicates per probe tuple is handled.
// Partitions whose build side exceeds one table are processed as several build chunks against the
// same probe tuples (the union of the chunk joins is the join) -- the overflow path.
// The slot hash uses the hash bits ABOVE the ones consumed by partitioning, so tuples of one
// partition (which share the low bits) still spread over the table.
#include "rj_common.cuh"
#include "rj_internal.h"

namespace rj {
namespace {

constexpr int      kJoinThreads = 512;
constexpr uint32_t kOutCap      = 4096;        // staged pairs per CTA
constexpr uint32_t kEmpty       = 0xffffffffu; // row ids are < 2^32 - 1
constexpr int      kBuildItems  = kJoinBuildCap / kJoinThreads; // 12

template <typename K>
struct JoinCfg {
    static constexpr int kProbeItems = sizeof(K) == 4 ? 8 : 4; // tuples per thread per super-batch
};

struct JoinArgs {
    const void*     bkeys;
    const uint32_t* bidx;
    const uint32_t* bvalid;
    const void*     pkeys;
    const uint32_t* pidx;
    const uint32_t* pvalid;
    const uint32_t* off_b;
    const uint32_t* off_p;
    const uint32_t* unit_start;
    uint32_t        nparts;
    int             part_bits;
    uint32_t*       out_b;
    uint32_t*       out_p;
    unsigned long long capacity;
    unsigned long long* out_count;
};

template <typename K>
__global__ void __launch_bounds__(kJoinThreads, 2) join_kernel(JoinArgs a) {
    constexpr int      kItems = JoinCfg<K>::kProbeItems;
    constexpr uint32_t kBatch = kItems * kJoinThreads;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    K*        s_keys  = reinterpret_cast<K*>(smem_raw);
    uint32_t* s_rows  = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * kJoinSlots);
    uint32_t* s_out_b = s_rows + kJoinSlots;
    uint32_t* s_out_p = s_out_b + kOutCap;
    __shared__ uint32_t           s_out_n;
    __shared__ unsigned long long s_flush_base;

    const K* __restrict__ bkeys = static_cast<const K*>(a.bkeys);
    const K* __restrict__ pkeys = static_cast<const K*>(a.pkeys);
    constexpr uint32_t kSlotMask = kJoinSlots - 1;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt   = lanemask_lt();
    const uint32_t n_units = a.unit_start[a.nparts];
    const bool     do_write = a.out_b != nullptr;
    const int      part_bits = a.part_bits;

    if (threadIdx.x == 0) s_out_n = 0;

    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        // ---- which (partition, build chunk, probe chunk)? --------------------------------------------
        uint32_t lo = 0, hi = a.nparts;
        while (hi - lo > 1) {
            uint32_t m = (lo + hi) >> 1;
            if (a.unit_start[m] <= u) lo = m; else hi = m;
        }
        const uint32_t part  = lo;
        const uint32_t local = u - a.unit_start[part];
        const uint32_t b_lo = a.off_b[part], b_hi = a.off_b[part + 1];
        const uint32_t p_lo = a.off_p[part], p_hi = a.off_p[part + 1];
        const uint32_t n_pchunks = (p_hi - p_lo + kJoinProbeChunk - 1) / kJoinProbeChunk;
        const uint32_t bc = local / n_pchunks, pc = local - bc * n_pchunks;
        const uint32_t bs = b_lo + bc * kJoinBuildCap;
        const uint32_t be = (b_hi - bs > kJoinBuildCap) ? bs + kJoinBuildCap : b_hi;
        const uint32_t ps = p_lo + pc * kJoinProbeChunk;
        const uint32_t pe = (p_hi - ps > kJoinProbeChunk) ? ps + kJoinProbeChunk : p_hi;

        // first probe super-batch: issue its loads before anything else so they overlap the build
        K        nkey[kItems];
        uint32_t nrow[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t i = ps + k * kJoinThreads + threadIdx.x;
            nkey[k] = K(0);
            nrow[k] = kEmpty; // kEmpty = no tuple / NULL key
            if (i < pe && (a.pvalid == nullptr || test_bit(a.pvalid, i))) {
                nkey[k] = pkeys[i];
                nrow[k] = a.pidx != nullptr ? a.pidx[i] : i;
            }
        }

        // ---- build -----------------------------------------------------------------------------------
        {
            K        bkey[kBuildItems];
            uint32_t brow[kBuildItems];
#pragma unroll
            for (int k = 0; k < kBuildItems; ++k) {
                const uint32_t i = bs + k * kJoinThreads + threadIdx.x;
                bkey[k] = K(0);
                brow[k] = kEmpty;
                if (i < be && (a.bvalid == nullptr || test_bit(a.bvalid, i))) {
                    bkey[k] = bkeys[i];
                    brow[k] = a.bidx != nullptr ? a.bidx[i] : i;
                }
            }
            __syncthreads(); // previous unit is done with the table
            for (uint32_t s = threadIdx.x; s < kJoinSlots; s += kJoinThreads) s_rows[s] = kEmpty;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kBuildItems; ++k) {
                if (brow[k] != kEmpty) {
                    uint32_t slot = (hash_key(bkey[k]) >> part_bits) & kSlotMask;
                    while (atomicCAS(&s_rows[slot], kEmpty, brow[k]) != kEmpty) slot = (slot + 1) & kSlotMask;
                    s_keys[slot] = bkey[k];
                }
                __syncwarp(); // keep the warp converged from one insert to the next
            }
            __syncthreads();
        }

        // ---- probe -----------------------------------------------------------------------------------
        for (uint32_t base = ps; base < pe; base += kBatch) {
            K        key[kItems];
            uint32_t row[kItems];
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                key[k] = nkey[k];
                row[k] = nrow[k];
            }
            // loads of the next super-batch
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const uint32_t i = base + kBatch + k * kJoinThreads + threadIdx.x;
                nkey[k] = K(0);
                nrow[k] = kEmpty;
                if (i < pe && (a.pvalid == nullptr || test_bit(a.pvalid, i))) {
                    nkey[k] = pkeys[i];
                    nrow[k] = a.pidx != nullptr ? a.pidx[i] : i;
                }
            }
            int      item = 0;           // first item of this thread that is not finished
            uint32_t slot = 0xffffffffu; // its current slot (0xffffffff = start at the home slot)
            for (;;) {
                bool stalled = false;
                // Items are walked in LOCKSTEP by the warp: the loop is fully unrolled (static register
                // indices) and every item ends in __syncwarp(), so lanes whose cluster walk is short wait
                // for the others instead of running ahead into the next item -- without it the warp
                // splits into 32 independent instruction streams (measured: 3 active threads / instr).
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    if (k >= item && !stalled && row[k] != kEmpty) {
                        const K mykey = key[k];
                        if (slot == 0xffffffffu) slot = (hash_key(mykey) >> part_bits) & kSlotMask;
                        for (;;) {
                            const uint32_t brow = s_rows[slot];
                            if (brow == kEmpty) break;
                            if (s_keys[slot] == mykey) {
                                // warp-aggregated reservation among the lanes that found a match right now
                                const uint32_t active = __activemask();
                                const uint32_t leader = __ffs(active) - 1;
                                uint32_t       pos = 0;
                                if (lane == leader) pos = atomicAdd(&s_out_n, static_cast<uint32_t>(__popc(active)));
                                pos = __shfl_sync(active, pos, leader) + __popc(active & lt);
                                if (pos >= kOutCap) {
                                    stalled = true; // staging buffer full: resume at this slot after the flush
                                    break;
                                }
                                s_out_b[pos] = brow;
                                s_out_p[pos] = row[k];
                            }
                            slot = (slot + 1) & kSlotMask;
                        }
                        if (stalled) {
                            item = k;
                        } else {
                            item = k + 1;
                            slot = 0xffffffffu;
                        }
                    } else if (k >= item && !stalled) {
                        item = k + 1; // no tuple / NULL key
                    }
                    __syncwarp();
                }
                const int any_stalled = __syncthreads_or(stalled ? 1 : 0);
                const uint32_t staged = s_out_n < kOutCap ? s_out_n : kOutCap;
                const bool last_batch = base + kBatch >= pe;
                // the next super-batch can add up to kBatch pairs without stalling only if there is room
                if (any_stalled || last_batch || staged + kBatch > kOutCap) {
                    // ---- flush: one global atomic, coalesced stores ---------------------------------
                    if (threadIdx.x == 0) s_flush_base = atomicAdd(a.out_count, static_cast<unsigned long long>(staged));
                    __syncthreads();
                    const unsigned long long gbase = s_flush_base;
                    if (do_write && gbase + staged <= a.capacity) {
                        for (uint32_t k = threadIdx.x; k < staged; k += kJoinThreads) {
                            a.out_b[gbase + k] = s_out_b[k];
                            a.out_p[gbase + k] = s_out_p[k];
                        }
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) s_out_n = 0;
                    __syncthreads();
                }
                if (!any_stalled) break;
            }
        }
    }
}

template <typename K>
void run_join(const JoinLaunch& L, int sm_count, cudaStream_t s) {
    const size_t smem = (sizeof(K) + 4) * kJoinSlots + 8 * kOutCap;
    static bool  configured = false;
    if (!configured) {
        RJ_CUDA(cudaFuncSetAttribute(join_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    JoinArgs a;
    a.bkeys = L.bkeys; a.bidx = L.bidx; a.bvalid = L.bvalid;
    a.pkeys = L.pkeys; a.pidx = L.pidx; a.pvalid = L.pvalid;
    a.off_b = L.off_b; a.off_p = L.off_p; a.unit_start = L.unit_start;
    a.nparts = L.nparts; a.part_bits = L.part_bits;
    a.out_b = L.out_b; a.out_p = L.out_p; a.capacity = L.capacity; a.out_count = L.out_count;
    // persistent grid: 2 CTAs per SM pull work units in a strided order
    join_kernel<K><<<static_cast<unsigned>(sm_count) * 2, kJoinThreads, smem, s>>>(a);
    RJ_LAUNCH_CHECK();
}

} // namespace

void launch_join(const JoinLaunch& a, int sm_count, cudaStream_t s) {
    if (a.key_bytes == 4) {
        run_join<uint32_t>(a, sm_count, s);
    } else {
        run_join<uint64_t>(a, sm_count, s);
    }
}

} // namespace rj
