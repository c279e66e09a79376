// join_emit_kernel instantiations with TWO carried build columns (see k_join_emit.cuh)
#include "k_join_emit.cuh"

namespace rj {
bool launch_join_emit_b2(const emit::EmitArgs& a, int n_ppay, int null_mask, int width_mask, size_t smem, unsigned grid, cudaStream_t s) {
    return emit::launch_emit_nb<2>(a, n_ppay, null_mask, width_mask, smem, grid, s);
}
} // namespace rj
