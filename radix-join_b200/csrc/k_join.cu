// Per-partition build + probe in shared memory -- the GPU replacement of the reference's per-bucket
// open-addressing join (src/execute.cpp:196-249: table of cap = pow2 >= 2*cnt slots, linear probing,
// duplicates listed per slot, one output row per (probe row, matching build row)).
//
// One CTA = one work unit = (partition, build chunk, probe chunk):
//   build : <= 6144 build tuples are inserted into an 8192-slot linear-probing table in shared memory
//           (keys[] + row ids[], the row id doubles as the occupancy flag).  Duplicate keys simply take
//           separate slots; the probe walks the cluster until an empty slot and emits every equal key,
//           which yields the reference's "one row per duplicate" semantics (tests/unit_tests.cpp:125-161).
//   probe : the probe chunk is streamed from global memory (coalesced, software-prefetched), every
//           thread looks its key up in the table and stages (build row, probe row) pairs in a
//           shared-memory buffer through warp-aggregated slot reservation; full buffers are flushed
//           with one global atomic per flush and coalesced stores.
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

        // ---- build -----------------------------------------------------------------------------------
        __syncthreads(); // previous unit is done with the table
        for (uint32_t s = threadIdx.x; s < kJoinSlots; s += kJoinThreads) s_rows[s] = kEmpty;
        __syncthreads();
        for (uint32_t i = bs + threadIdx.x; i < be; i += kJoinThreads) {
            if (a.bvalid != nullptr && !test_bit(a.bvalid, i)) continue;
            const K        key = bkeys[i];
            const uint32_t row = a.bidx != nullptr ? a.bidx[i] : i;
            uint32_t slot = (hash_key(key) >> a.part_bits) & kSlotMask;
            while (atomicCAS(&s_rows[slot], kEmpty, row) != kEmpty) slot = (slot + 1) & kSlotMask;
            s_keys[slot] = key;
        }
        __syncthreads();

        // ---- probe -----------------------------------------------------------------------------------
        // software prefetch: the tuple of the next batch is loaded before the current one is probed
        uint32_t i_next = ps + threadIdx.x;
        K        key_next = K(0);
        uint32_t row_next = 0;
        bool     ok_next  = false;
        if (i_next < pe) {
            ok_next = a.pvalid == nullptr || test_bit(a.pvalid, i_next);
            key_next = pkeys[i_next];
            row_next = a.pidx != nullptr ? a.pidx[i_next] : i_next;
        }
        for (uint32_t base = ps; base < pe; base += kJoinThreads) {
            const K        key = key_next;
            const uint32_t row = row_next;
            bool           pending = ok_next;
            i_next = base + kJoinThreads + threadIdx.x;
            ok_next = false;
            if (i_next < pe) {
                ok_next = a.pvalid == nullptr || test_bit(a.pvalid, i_next);
                key_next = pkeys[i_next];
                row_next = a.pidx != nullptr ? a.pidx[i_next] : i_next;
            }
            uint32_t slot = (hash_key(key) >> a.part_bits) & kSlotMask;
            for (;;) {
                // walk the cluster; suspend when the staging buffer is full
                while (pending) {
                    const uint32_t brow = s_rows[slot];
                    if (brow == kEmpty) {
                        pending = false;
                        break;
                    }
                    if (s_keys[slot] == key) {
                        // warp-aggregated reservation among the lanes that found a match right now
                        const uint32_t active = __activemask();
                        const uint32_t leader = __ffs(active) - 1;
                        uint32_t       pos = 0;
                        if (lane == leader) pos = atomicAdd(&s_out_n, static_cast<uint32_t>(__popc(active)));
                        pos = __shfl_sync(active, pos, leader) + __popc(active & lt);
                        if (pos >= kOutCap) break; // retry this slot after the flush
                        s_out_b[pos] = brow;
                        s_out_p[pos] = row;
                    }
                    slot = (slot + 1) & kSlotMask;
                }
                const int any_pending = __syncthreads_or(pending ? 1 : 0);
                const uint32_t staged = s_out_n < kOutCap ? s_out_n : kOutCap;
                const bool last_batch = base + kJoinThreads >= pe;
                if (any_pending || last_batch || staged + kJoinThreads > kOutCap) {
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
                if (!any_pending) break;
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
