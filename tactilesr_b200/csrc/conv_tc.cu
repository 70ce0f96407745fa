// placeholder until the tcgen05 kernels land (next commit)
#include "common.cuh"
extern "C" {
int tsr_pack_conv_weight_bf16(const float*, void*, void*, int, int, int, cudaStream_t) { tsr_set_error("tc path not built"); return TSR_ERR_UNSUPPORTED; }
int tsr_conv2d_tc(const void*, int, const void*, const float*, const void*, int, void*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t) { tsr_set_error("tc path not built"); return TSR_ERR_UNSUPPORTED; }
size_t tsr_conv2d_tc_workspace(int, int, int, int, int, int) { return 0; }
int tsr_conv2d_wgrad_tc(const void*, int, const void*, int, float*, void*, size_t, int, int, int, int, int, int, int, cudaStream_t) { tsr_set_error("tc path not built"); return TSR_ERR_UNSUPPORTED; }
size_t tsr_conv2d_wgrad_tc_workspace(int, int, int, int, int, int) { return 0; }
int tsr_tc_selftest(int, void*, void*, cudaStream_t) { tsr_set_error("tc path not built"); return TSR_ERR_UNSUPPORTED; }
}
