// tcgen05 / TMA contraction kernel -- placeholder until the tensor-core path lands.
#include "gemm.cuh"

namespace zutis {

size_t gemm_tcgen05_workspace_bytes(int, long, int, int, int) { return 0; }
bool gemm_tcgen05_supports(const GemmParams&, int, int) { return false; }
int launch_gemm_tcgen05(const GemmParams&, int, int, void*, size_t, cudaStream_t) {
    return fail(ZUTIS_ERR_UNSUPPORTED, "tcgen05 contraction kernel not built");
}

}  // namespace zutis
