// Launcher of the fused root join (kernel: k_join_emit.cuh) + its instantiations without carried build
// columns.  Shared-memory layout, grid sizing and the dispatch over (build columns, probe columns, nullable
// columns) live here.
#include "k_join_emit.cuh"

namespace rj {

using emit::EmitArgs;
using emit::kBatch;

bool launch_join_emit_b0(const EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s) {
    return emit::launch_emit_nb<0>(a, n_ppay, null_mask, width_mask, smem, grid, s);
}

static size_t join_emit_layout(const JoinEmitLaunch& L, EmitArgs* a) {
    size_t off = sizeof(uint64_t) * kEmitSlots; // offsets are relative to the table (emit::kBarBytes of mbarriers lie in front)
    auto   take = [&](size_t bytes) {
        off = (off + 15) / 16 * 16;
        const size_t at = off;
        off += bytes;
        return static_cast<uint32_t>(at);
    };
    EmitArgs tmp{};
    EmitArgs& r = a ? *a : tmp;
    for (int c = 0; c < L.n_bpay; ++c) r.sm_bpay[c] = take(size_t(kEmitBuildCap) * L.bwidth[c] + 16);
    for (int c = 0; c < L.n_bpay; ++c) r.sm_bvalid[c] = L.bvalid[c] ? take(kEmitBuildCap + 16) : 0;
    // kStages probe buffers of identical layout
    r.sm_pkeys = take(size_t(kBatch) * 4 + 16);
    for (int c = 0; c < L.n_ppay; ++c) r.sm_ppay[c] = take(size_t(kBatch) * L.pwidth[c] + 16);
    for (int c = 0; c < L.n_ppay; ++c) r.sm_pvalid[c] = L.pvalid[c] ? take(kBatch + 16) : 0;
    off = (off + 15) / 16 * 16;
    r.sm_pstride = static_cast<uint32_t>(off - r.sm_pkeys);
    return off + size_t(emit::kStages - 1) * r.sm_pstride + emit::kBarBytes;
}

bool join_emit_fits(const JoinEmitLaunch& L) {
    if (L.n_out < 1 || L.n_out > kEmitMaxOut || L.n_bpay > kEmitMaxPay || L.n_ppay > kEmitMaxPay) return false;
    return join_emit_layout(L, nullptr) <= 226 * 1024;
}

void launch_join_emit(const JoinEmitLaunch& L, uint64_t n_probe, int sm_count, cudaStream_t s) {
    EmitArgs a{};
    a.bkeys = L.bkeys; a.pkeys = L.pkeys; a.off_b = L.off_b; a.off_p = L.off_p;
    a.unit_start = L.unit_start; a.unit_cursor = L.unit_cursor; a.nparts = L.nparts; a.part_bits = L.part_bits;
    a.unit_part = L.unit_part; a.unit_part_cap = L.unit_part ? 2 * L.nparts : 0;
    a.probe_chunk = kEmitProbeChunk;
    int null_mask = 0, width_mask = 0;
    for (int c = 0; c < kEmitMaxPay; ++c) {
        a.bpay[c] = L.bpay[c]; a.bvalid[c] = L.bvalid[c]; a.bwidth[c] = L.bwidth[c] ? L.bwidth[c] : 4;
        if (c < L.n_bpay && L.bvalid[c]) null_mask |= 1 << c;
        if (c < L.n_bpay && a.bwidth[c] == 8) width_mask |= 1 << c;
        a.ppay[c] = L.ppay[c]; a.pvalid[c] = L.pvalid[c]; a.pwidth[c] = L.pwidth[c] ? L.pwidth[c] : 4;
        if (c < L.n_ppay && L.pvalid[c]) null_mask |= 1 << (kEmitMaxPay + c);
        if (c < L.n_ppay && a.pwidth[c] == 8) width_mask |= 1 << (kEmitMaxPay + c);
    }
    for (int j = 0; j < L.n_out; ++j) {
        a.out_pages[j] = L.out_pages[j];
        int src = 0;
        if (L.out_src[j] == 1) src = 1 + L.out_idx[j];
        else if (L.out_src[j] == 2) src = 1 + kEmitMaxPay + L.out_idx[j];
        const int width = L.out_src[j] == 0 ? 4 : (L.out_src[j] == 1 ? a.bwidth[L.out_idx[j]] : a.pwidth[L.out_idx[j]]);
        if (L.out_width[j] != width) throw CudaError("join_emit: an output column's width differs from its source's");
        if (a.src_pages[src] == nullptr) a.src_pages[src] = L.out_pages[j];
        else a.src_rest[src] |= 1u << j;
    }
    // the rank-structure table needs the hash to be a bijection with at most 17 bits left (see the kernel)
    a.direct = L.part_bits >= 15 && L.part_bits <= 27 ? 1 : 0;
    a.all_once = 1;
    for (int src = 0; src < emit::kSources; ++src) {
        const bool used = src == 0 || (src <= kEmitMaxPay ? src - 1 < L.n_bpay : src - 1 - kEmitMaxPay < L.n_ppay);
        if (used && (a.src_pages[src] == nullptr || a.src_rest[src] != 0)) a.all_once = 0;
    }
    a.chunk_counter = L.chunk_counter; a.row_counter = L.row_counter; a.abort_flag = L.abort_flag;
    const size_t smem = join_emit_layout(L, &a);
    if (smem > 226 * 1024) throw CudaError("join_emit: the columns do not fit shared memory");
    const unsigned grid = join_emit_grid(n_probe, L.nparts, sm_count);
    a.n_active = join_emit_active_warps(n_probe, grid, L.min_chunks_per_warp);
    bool launched = false;
    switch (L.n_bpay) {
    case 0: launched = launch_join_emit_b0(a, L.n_ppay, null_mask, width_mask, smem, grid, s); break;
    case 1: launched = launch_join_emit_b1(a, L.n_ppay, null_mask, width_mask, smem, grid, s); break;
    case 2: launched = launch_join_emit_b2(a, L.n_ppay, null_mask, width_mask, smem, grid, s); break;
    default: break;
    }
    if (!launched) throw CudaError("join_emit: unsupported column configuration");
    RJ_LAUNCH_CHECK();
}

} // namespace rj
