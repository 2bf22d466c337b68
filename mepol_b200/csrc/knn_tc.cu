// K1 (tensor-core form) -- placeholder until the tcgen05 pruner lands; AUTO resolves to SIMT.
#include "knn_internal.cuh"

namespace mepol {
bool tc_supported(int, int) { return false; }
int launch_prune_tc(const KnnPlan &, cudaStream_t) {
    set_error("knn: tensor-core pruner not built");
    return MEPOL_E_UNSUPPORTED;
}
}  // namespace mepol
