// tcgen05 implicit-GEMM conv3d (placeholder until the kernel lands: reports "unsupported" so AUTO picks SIMT).
#include "common.cuh"
namespace cfun {
bool tc_supported(const cfun_conv3d_desc*, int) { return false; }
size_t tc_workspace(const cfun_conv3d_desc*, int) { return 0; }
int tc_conv_fwd(const cfun_conv3d_desc*, const float*, const float*, const float*, float*, int, int, void*, size_t, cudaStream_t) { return CFUN_ERR_INVALID; }
int tc_conv_bwd_data(const cfun_conv3d_desc*, const float*, const float*, float*, int, void*, size_t, cudaStream_t) { return CFUN_ERR_INVALID; }
int tc_conv_bwd_weight(const cfun_conv3d_desc*, const float*, const float*, float*, float*, int, void*, size_t, cudaStream_t) { return CFUN_ERR_INVALID; }
int cfun_pack_split_bf16_impl() { return 0; }
}  // namespace cfun
extern "C" int cfun_pack_split_bf16(const float*, void*, void*, long long, int, int, void*) { return CFUN_ERR_INVALID; }
