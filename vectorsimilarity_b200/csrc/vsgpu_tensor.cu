// Tensor-core coarse pass + exact re-rank. Placeholder until the tcgen05 kernel lands:
// reports "not supported" so every query takes the exact path.
#include "vsgpu_internal.cuh"

namespace vsgpu {
bool tensor_path_supported(const vsgpu_store *, size_t, size_t) { return false; }
int tensor_topk(vsgpu_store *, const void *, size_t, size_t, const float *, size_t, uint32_t *, void *, uint64_t *) {
    set_error("tensor path not built");
    return VSGPU_ERR_ARG;
}
int tensor_sync_mirrors(vsgpu_store *) { return VSGPU_OK; }
void tensor_release(vsgpu_store *) {}
} // namespace vsgpu
