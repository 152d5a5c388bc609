// gemm.cu -- K2 placeholder (replaced by the tcgen05 kernel).
#include "common.cuh"
namespace cmw {
int encode_bf16_tmap(Store*) { return -1; }
bool gemm_supported(const Store*) { return false; }
int launch_gemm(const GemmArgs&, cudaStream_t) {
    set_error("K2 (tcgen05 GEMM) is not built in");
    return -1;
}
}  // namespace cmw
