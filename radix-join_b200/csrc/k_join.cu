// Per-partition build + probe in shared memory -- the GPU replacement of the reference's per-bucket
// open-addressing join (src/execute.cpp:196-249: table of cap = pow2 >= 2*cnt slots, linear probing,
// duplicates listed per slot, one output row per (probe row, matching build row)).  This is the GENERAL join:
// any key type (4-byte keys, or 8-byte keys / VARCHAR hashes), duplicates on both sides, output = (build
// position, probe position) pairs that the caller gathers through.  Root joins of two scans on an INT32 key
// take the fused kernel instead (k_join_emit.cuh).
//
// One CTA = one work unit = (partition, build chunk, probe chunk), handed out IN ORDER by a global atomic
// cursor (see below why):
//   build : <= 6144 build tuples (2048 on average: 25 % fill) go into an 8192-slot open-addressing table in
//           shared memory with DOUBLE hashing (the stride of a key's probe sequence is a second hash of the
//           key).  For 4-byte keys a slot is ONE 64-bit word (key | row id << 32) claimed with a 64-bit CAS,
//           so a probe step is a single LDS.64 and a failed CAS tells the inserter whether it collided with
//           an EQUAL key.  The table holds every distinct key once; further tuples of that key hang off the
//           slot as a chain (dup_head / dup_next, private to the CTA) -- the counterpart of the reference's
//           slot_idxs vectors (src/execute.cpp:203-223): a key with d duplicates costs O(d) to build and only
//           probes of that very key walk them.  8-byte keys keep duplicates as separate entries along the
//           key's own sequence.  (Details at `struct Table`.)
//   probe : the probe chunk (a.probe_chunk tuples: 16384, fewer when the probe side is too small to give
//           every SM a unit) is streamed in super-batches of 4-8 tuples per thread whose global loads are in
//           flight one super-batch ahead.  The warp walks in lockstep ROUNDS: every lane walks to its next
//           match (or the end of its sequence / chain), then the warp emits all matches of the round with one
//           ballot + one shared-memory atomic (the first version emitted inside the divergent walk: ncu
//           showed 3-8 active threads per instruction).  Pairs are staged in shared memory and flushed with
//           ONE global atomic and coalesced stores.  A lane that finds the staging buffer full remembers item
//           + position and resumes after the flush, so any number of duplicates per probe tuple is handled.
// Partitions whose build side exceeds one table are processed as several build chunks against the same
// probe tuples (the union of the chunk joins is the join) -- the overflow path.
// The slot hash uses the hash bits ABOVE the ones consumed by partitioning, so tuples of one partition
// (which share the low bits) still spread over the table.
#include "rj_common.cuh"
#include "rj_internal.h"

#include <stdexcept>

namespace rj {
namespace {

constexpr int      kJoinThreads = 512;
constexpr uint32_t kOutCap      = 4096;        // staged pairs per CTA
constexpr uint32_t kEmpty       = 0xffffffffu; // row ids are < 2^32 - 1
constexpr int      kBuildItems  = kJoinBuildCap / kJoinThreads; // 12
constexpr uint32_t kSlotMask    = kJoinSlots - 1;
// Units are handed out IN ORDER through a global cursor (one atomic per unit): the units in flight at
// any instant are the ~2 x #SM most recently started ones, i.e. about one pass-1 region, so everything
// the matches of that moment refer to (row ids, keys, carried payloads in position order) is an
// L2-sized window.  A static round-robin lets CTAs drift apart under skew: measured, 4096 consecutive
// output rows then span 20 regions and every gathered value costs a DRAM line.

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
    uint32_t        probe_chunk; // probe tuples per work unit
    uint32_t*       unit_cursor;
    uint32_t        nparts;
    int             part_bits;
    uint32_t*       out_b;
    uint32_t*       out_p;
    unsigned long long capacity;
    unsigned long long* out_count;
    uint32_t*       dup_next; // [gridDim.x * kJoinBuildCap] duplicate chains of 4-byte keys: next node + 1, per CTA
    uint32_t*       dup_head; // [gridDim.x * kJoinSlots]    first node + 1 per table slot, per CTA
};

// ---- the shared-memory table, specialised on the key width -----------------------------------------
// Open addressing with DOUBLE hashing: the stride of a key's probe sequence is a second hash of the key
// (odd, so the sequence visits every slot).
//
// Duplicate build keys.  4-byte keys: the table holds every DISTINCT key once (key | row of the tuple
// that won the slot); the other tuples of that key hang off the slot as a chain through global memory
// (dup_head: first node per slot, dup_next: next node per build tuple; a node is a build tuple's index
// inside the unit's build chunk, its row id is re-read from the build arrays; both arrays are private to
// the CTA, because several CTAs build the same chunk at once -- one per probe chunk -- each in its own
// order; the heads are read with ld.cg since they are written by atomics, past the L1).  Pushing a node is one atomicExch, so a key with d
// duplicates costs O(d) to build and only probes of that very key walk them.  The first version stored
// duplicates as separate entries of a linear-probing table: d duplicates cost O(d^2) CAS attempts on one
// contended frontier slot and formed a run that every probe hashing into it walked to its end -- JOB
// plans with 4 000 copies of one movie_id on the table side spent 6-13 ms per join (results unchanged).
// 8-byte keys (INT64, VARCHAR hashes) still keep duplicates as separate entries along the key's own
// sequence.
__device__ __forceinline__ uint32_t probe_step(uint32_t k) { return ((k * 0x9E3779B1u) >> 19) | 1u; }
__device__ __forceinline__ uint32_t probe_step(uint64_t k) { return probe_step(static_cast<uint32_t>(k ^ (k >> 32))); }

template <typename K>
struct Table;

template <>
struct Table<uint32_t> {
    static constexpr int    kProbeItems = 8;
    static constexpr size_t kBytes = sizeof(uint64_t) * kJoinSlots;
    unsigned long long*     slots; // key | row << 32; all ones = empty
    __device__ explicit Table(uint8_t* smem): slots(reinterpret_cast<unsigned long long*>(smem)) {}
    __device__ void clear() {
        for (uint32_t s = threadIdx.x; s < kJoinSlots; s += kJoinThreads) slots[s] = ~0ull;
    }
    // returns true when the key is in the table already: the tuple is NOT inserted (it joins the key's chain)
    __device__ bool insert(uint32_t key, uint32_t row, uint32_t slot) {
        const unsigned long long mine = static_cast<unsigned long long>(key) | (static_cast<unsigned long long>(row) << 32);
        const uint32_t step = probe_step(key);
        for (;;) {
            unsigned long long cur = slots[slot];
            if (cur == ~0ull) cur = atomicCAS(&slots[slot], ~0ull, mine);
            if (cur == ~0ull) return false;
            if (static_cast<uint32_t>(cur) == key) return true;
            slot = (slot + step) & kSlotMask;
        }
    }
    // slot of a key that is known to be in the table
    __device__ uint32_t find(uint32_t key, uint32_t slot) const {
        const uint32_t step = probe_step(key);
        while (static_cast<uint32_t>(slots[slot]) != key) slot = (slot + step) & kSlotMask;
        return slot;
    }
    __device__ void post_build(uint32_t, uint32_t, uint32_t, int*) {}
    // row id stored at `slot` (kEmpty if free) and whether its key equals `key`
    __device__ uint32_t load(uint32_t slot, uint32_t key, bool* equal) const {
        const unsigned long long e = slots[slot];
        *equal = static_cast<uint32_t>(e) == key;
        return static_cast<uint32_t>(e >> 32);
    }
};

template <>
struct Table<uint64_t> {
    static constexpr int    kProbeItems = 4;
    static constexpr size_t kBytes = (sizeof(uint64_t) + sizeof(uint32_t)) * kJoinSlots;
    uint64_t* keys;
    uint32_t* rows;
    __device__ explicit Table(uint8_t* smem)
        : keys(reinterpret_cast<uint64_t*>(smem)), rows(reinterpret_cast<uint32_t*>(smem + sizeof(uint64_t) * kJoinSlots)) {}
    __device__ void clear() {
        for (uint32_t s = threadIdx.x; s < kJoinSlots; s += kJoinThreads) rows[s] = kEmpty;
    }
    __device__ bool insert(uint64_t key, uint32_t row, uint32_t slot) {
        const uint32_t step = probe_step(key);
        while (atomicCAS(&rows[slot], kEmpty, row) != kEmpty) slot = (slot + step) & kSlotMask;
        keys[slot] = key;
        return false; // the key of a colliding slot may not be written yet: checked in post_build
    }
    // after the build barrier: did an equal key land between my home slot and my own slot?
    __device__ void post_build(uint64_t key, uint32_t row, uint32_t slot, int* dups) {
        const uint32_t step = probe_step(key);
        while (rows[slot] != row) {
            if (keys[slot] == key) {
                *dups = 1;
                return;
            }
            slot = (slot + step) & kSlotMask;
        }
    }
    __device__ uint32_t load(uint32_t slot, uint64_t key, bool* equal) const {
        const uint32_t r = rows[slot];
        *equal = r != kEmpty && keys[slot] == key;
        return r;
    }
    __device__ uint32_t find(uint64_t, uint32_t slot) const { return slot; } // chains are for 4-byte keys
};

template <typename K>
// 4-byte keys: two CTAs of 96 KB per SM; 8-byte keys need 128 KB of table + staging: one CTA per SM, and no reason to
// cap it at 64 registers
__global__ void __launch_bounds__(kJoinThreads, sizeof(K) == 4 ? 2 : 1) join_kernel(JoinArgs a) {
    constexpr int      kItems = Table<K>::kProbeItems;
    constexpr uint32_t kBatch = kItems * kJoinThreads;
    constexpr bool     kChains = sizeof(K) == 4; // duplicates of a key hang off its slot (see Table)
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Table<K>  table(smem_raw);
    uint32_t* s_out_b = reinterpret_cast<uint32_t*>(smem_raw + Table<K>::kBytes);
    uint32_t* s_out_p = s_out_b + kOutCap;
    __shared__ uint32_t           s_out_n;
    __shared__ int                s_dups;
    __shared__ uint32_t           s_unit, s_part;
    __shared__ unsigned long long s_flush_base;

    const K* __restrict__ bkeys = static_cast<const K*>(a.bkeys);
    const K* __restrict__ pkeys = static_cast<const K*>(a.pkeys);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt   = lanemask_lt();
    const uint32_t n_units = a.unit_start[a.nparts];
    const bool     do_write = a.out_b != nullptr;
    const int      part_bits = a.part_bits;

    if (threadIdx.x == 0) s_out_n = 0;

    for (;;) {
        // ---- next work unit, in global order ---------------------------------------------------------
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 takes the next unit and finds its partition with a 32-ary search: each lane probes
            // one splitter per step, so 2^15 partitions take 3 dependent loads instead of 15
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(a.unit_cursor, 1u);
            u = __shfl_sync(RJ_FULL_MASK, u, 0);
            uint32_t lo = 0, hi = a.nparts; // unit_start[lo] <= u < unit_start[hi]
            if (u < n_units) {
                while (hi - lo > 1) {
                    const uint32_t span = hi - lo;
                    const uint32_t step = (span + 31) / 32;
                    const uint32_t probe = lo + (lane + 1) * step; // splitters lo+step, lo+2*step, ...
                    const bool     le = probe < hi && a.unit_start[probe] <= u;
                    const uint32_t m = __ballot_sync(RJ_FULL_MASK, le);
                    const uint32_t k = __popc(m); // unit_start is non-decreasing: the lanes that hold are a prefix
                    const uint32_t nlo = lo + k * step;
                    const uint32_t nhi = (k < 32 && lo + (k + 1) * step < hi) ? lo + (k + 1) * step : hi;
                    lo = nlo;
                    hi = nhi;
                }
            }
            if (lane == 0) {
                s_unit = u;
                s_part = lo;
            }
        }
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units) break;
        {
            const uint32_t part = s_part;
            const uint32_t local = u - a.unit_start[part];
            const uint32_t b_lo = a.off_b[part], b_hi = a.off_b[part + 1];
            const uint32_t p_lo = a.off_p[part], p_hi = a.off_p[part + 1];
            const uint32_t n_pchunks = (p_hi - p_lo + a.probe_chunk - 1) / a.probe_chunk;
            const uint32_t bc = local / n_pchunks, pc = local - bc * n_pchunks;
            const uint32_t bs = b_lo + bc * kJoinBuildCap;
            const uint32_t be = (b_hi - bs > kJoinBuildCap) ? bs + kJoinBuildCap : b_hi;
            const uint32_t ps = p_lo + pc * a.probe_chunk;
            const uint32_t pe = (p_hi - ps > a.probe_chunk) ? ps + a.probe_chunk : p_hi;

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

            // ---- build -------------------------------------------------------------------------------
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
                table.clear();
                if (threadIdx.x == 0) s_dups = 0;
                __syncthreads();
                uint32_t dup = 0; // bit k: build item k met its key in the table
#pragma unroll
                for (int k = 0; k < kBuildItems; ++k) {
                    if (brow[k] != kEmpty && table.insert(bkey[k], brow[k], (hash_key(bkey[k]) >> part_bits) & kSlotMask)) dup |= 1u << k;
                    __syncwarp(); // keep the warp converged from one insert to the next
                }
                if (dup) s_dups = 1;
                __syncthreads();
                if (kChains && s_dups) {
                    // duplicates exist: reset this CTA's chain heads, then every duplicate pushes itself
                    // onto the chain of its key's slot (one atomicExch each)
                    uint32_t* head = a.dup_head + static_cast<size_t>(blockIdx.x) * kJoinSlots;
                    uint32_t* chain = a.dup_next + static_cast<size_t>(blockIdx.x) * kJoinBuildCap;
                    for (uint32_t sl = threadIdx.x; sl < kJoinSlots; sl += kJoinThreads) head[sl] = 0;
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < kBuildItems; ++k) {
                        if (dup & (1u << k)) {
                            const uint32_t li = k * kJoinThreads + threadIdx.x; // index inside the build chunk
                            const uint32_t sl = table.find(bkey[k], (hash_key(bkey[k]) >> part_bits) & kSlotMask);
                            chain[li] = atomicExch(&head[sl], li + 1);
                        }
                    }
                    __syncthreads();
                }
                if (sizeof(K) == 8) {
                    int found = 0;
#pragma unroll
                    for (int k = 0; k < kBuildItems; ++k) {
                        if (brow[k] != kEmpty && !found)
                            table.post_build(bkey[k], brow[k], (hash_key(bkey[k]) >> part_bits) & kSlotMask, &found);
                        __syncwarp();
                    }
                    if (found) s_dups = 1;
                    __syncthreads();
                }
            }
            const bool unique = s_dups == 0; // every build key of this table is distinct

            // ---- probe -------------------------------------------------------------------------------
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
                int      item = 0;           // first item of this lane that is not finished
                uint32_t slot = 0xffffffffu; // where its cluster walk continues (0xffffffff = home slot)
                for (;;) {
                    bool stalled = false;
                    if (unique) {
                        // ---- fast path: no duplicate build keys in this table --------------------------
                        // at most one match per probe tuple, so a super-batch stages <= kBatch pairs into a
                        // buffer that is flushed after every super-batch: no overflow, no resume logic.
                        // All items are walked first (lockstep per item), then the warp reserves its
                        // staging slots ONCE: per-item ballots give conflict-free, item-major positions.
                        uint32_t mrow[kItems];
                        uint32_t bal[kItems];
                        uint32_t total = 0;
#pragma unroll
                        for (int k = 0; k < kItems; ++k) {
                            mrow[k] = kEmpty;
                            if (row[k] != kEmpty) {
                                uint32_t sl = (hash_key(key[k]) >> part_bits) & kSlotMask;
                                const uint32_t step = probe_step(key[k]);
                                for (;;) {
                                    bool           eq;
                                    const uint32_t r = table.load(sl, key[k], &eq);
                                    if (r == kEmpty) break;
                                    if (eq) {
                                        mrow[k] = r;
                                        break;
                                    }
                                    sl = (sl + step) & kSlotMask;
                                }
                            }
                            __syncwarp();
                            bal[k] = __ballot_sync(RJ_FULL_MASK, mrow[k] != kEmpty);
                            total += __popc(bal[k]);
                        }
                        uint32_t off = 0;
                        if (lane == 0 && total) off = atomicAdd(&s_out_n, total);
                        off = __shfl_sync(RJ_FULL_MASK, off, 0);
#pragma unroll
                        for (int k = 0; k < kItems; ++k) {
                            if (mrow[k] != kEmpty) {
                                const uint32_t pos = off + __popc(bal[k] & lt);
                                s_out_b[pos] = mrow[k];
                                s_out_p[pos] = row[k];
                            }
                            off += __popc(bal[k]);
                        }
                    } else if (kChains) {
                        // ---- duplicates as chains: find the key once, then follow its nodes ---------------
                        // `slot` is the lane's position in the chain of item `item`: 0xffffffff = the key has
                        // not been looked up yet, otherwise the next node + 1 (0 = end of the chain)
                        const uint32_t* head = a.dup_head + static_cast<size_t>(blockIdx.x) * kJoinSlots;
                        const uint32_t* chain = a.dup_next + static_cast<size_t>(blockIdx.x) * kJoinBuildCap;
#pragma unroll
                        for (int k = 0; k < kItems; ++k) {
                            bool walking = k >= item && !stalled && row[k] != kEmpty;
                            for (;;) { // rounds: at most one match per lane per round
                                uint32_t match = kEmpty, next = 0;
                                if (walking) {
                                    if (slot == 0xffffffffu) {
                                        uint32_t       sl = (hash_key(key[k]) >> part_bits) & kSlotMask;
                                        const uint32_t step = probe_step(key[k]);
                                        for (;;) {
                                            bool           eq;
                                            const uint32_t r = table.load(sl, key[k], &eq);
                                            if (r == kEmpty) break; // the key is not in the table
                                            if (eq) {
                                                match = r;
                                                next = __ldcg(&head[sl]);
                                                break;
                                            }
                                            sl = (sl + step) & kSlotMask;
                                        }
                                        if (match == kEmpty) walking = false;
                                    } else {
                                        const uint32_t gi = bs + slot - 1;
                                        match = a.bidx != nullptr ? a.bidx[gi] : gi;
                                        next = chain[slot - 1]; // written by this CTA's own stores: L1 is coherent with them
                                    }
                                }
                                __syncwarp();
                                const uint32_t m = __ballot_sync(RJ_FULL_MASK, match != kEmpty);
                                if (m == 0) break; // every lane reached the end of its chain
                                uint32_t pos = 0;
                                if (lane == 0) pos = atomicAdd(&s_out_n, static_cast<uint32_t>(__popc(m)));
                                pos = __shfl_sync(RJ_FULL_MASK, pos, 0) + __popc(m & lt);
                                if (match != kEmpty) {
                                    if (pos < kOutCap) {
                                        s_out_b[pos] = match;
                                        s_out_p[pos] = row[k];
                                        slot = next;
                                        if (next == 0) walking = false;
                                    } else {
                                        // staging buffer full: `slot` still names this match, redo it after the flush
                                        walking = false;
                                        stalled = true;
                                        item = k;
                                    }
                                }
                            }
                            if (!stalled && k >= item) {
                                item = k + 1;
                                slot = 0xffffffffu;
                            }
                        }
                    } else {
#pragma unroll
                    for (int k = 0; k < kItems; ++k) {
                        // lanes that already finished item k (before a flush), have no tuple, or stalled
                        // on an earlier item sit this item out but keep in step with the warp
                        bool walking = k >= item && !stalled && row[k] != kEmpty;
                        if (walking && slot == 0xffffffffu) slot = (hash_key(key[k]) >> part_bits) & kSlotMask;
                        const uint32_t step = probe_step(key[k]);
                        for (;;) { // rounds: at most one match per lane per round
                            uint32_t match = kEmpty;
                            if (walking) {
                                for (;;) {
                                    bool           eq;
                                    const uint32_t r = table.load(slot, key[k], &eq);
                                    if (r == kEmpty) { // end of the cluster
                                        walking = false;
                                        break;
                                    }
                                    slot = (slot + step) & kSlotMask;
                                    if (eq) {
                                        match = r;
                                        break;
                                    }
                                }
                            }
                            __syncwarp();
                            const uint32_t m = __ballot_sync(RJ_FULL_MASK, match != kEmpty);
                            if (m == 0) break; // every lane reached the end of its cluster
                            uint32_t pos = 0;
                            if (lane == 0) pos = atomicAdd(&s_out_n, static_cast<uint32_t>(__popc(m)));
                            pos = __shfl_sync(RJ_FULL_MASK, pos, 0) + __popc(m & lt);
                            if (match != kEmpty) {
                                if (pos < kOutCap) {
                                    s_out_b[pos] = match;
                                    s_out_p[pos] = row[k];
                                } else {
                                    // staging buffer full: step back onto the match, resume after the flush
                                    slot = (slot - step) & kSlotMask;
                                    walking = false;
                                    stalled = true;
                                    item = k;
                                }
                            }
                        }
                        if (!stalled && k >= item) {
                            item = k + 1;
                            slot = 0xffffffffu;
                        }
                    }
                    }
                    const int any_stalled = __syncthreads_or(stalled ? 1 : 0);
                    const uint32_t staged = s_out_n < kOutCap ? s_out_n : kOutCap;
                    const bool last_batch = base + kBatch >= pe;
                    // the next super-batch can add up to kBatch pairs without stalling only if there is room
                    if (any_stalled || last_batch || staged + kBatch > kOutCap) {
                        // ---- flush: one global atomic, coalesced stores -----------------------------
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
}

template <typename K>
void run_join(const JoinLaunch& L, int sm_count, cudaStream_t s) {
    const size_t smem = Table<K>::kBytes + 8 * kOutCap;
    static SmemConfigured cfg;
    cfg.ensure(join_kernel<K>, smem);
    JoinArgs a;
    a.bkeys = L.bkeys; a.bidx = L.bidx; a.bvalid = L.bvalid;
    a.pkeys = L.pkeys; a.pidx = L.pidx; a.pvalid = L.pvalid;
    a.off_b = L.off_b; a.off_p = L.off_p; a.unit_start = L.unit_start; a.unit_cursor = L.unit_cursor;
    a.probe_chunk = L.probe_chunk ? L.probe_chunk : kJoinProbeChunk;
    a.nparts = L.nparts; a.part_bits = L.part_bits;
    a.out_b = L.out_b; a.out_p = L.out_p; a.capacity = L.capacity; a.out_count = L.out_count;
    a.dup_next = L.dup_next; a.dup_head = L.dup_head;
    if (sizeof(K) == 4 && (!a.dup_next || !a.dup_head)) throw std::runtime_error("join: duplicate-chain scratch missing");
    // persistent grid: 2 CTAs per SM pull batches of work units in a strided order
    const unsigned grid = sizeof(K) == 4 ? join_grid(sm_count) : static_cast<unsigned>(sm_count); // resident CTAs only
    join_kernel<K><<<grid, kJoinThreads, smem, s>>>(a);
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
