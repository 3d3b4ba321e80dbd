// conv3d entry points: argument validation and algorithm selection (CUDA-core implicit GEMM vs tcgen05 implicit GEMM).
#include "common.cuh"
#include <cstdlib>

namespace cfun {
size_t simt_workspace(const cfun_conv3d_desc* d, int pass);
int simt_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, void* ws,
                  size_t ws_bytes, cudaStream_t st);
int simt_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, void* ws, size_t ws_bytes,
                       cudaStream_t st);
int simt_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, void* ws,
                         size_t ws_bytes, cudaStream_t st);
int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

bool tc_supported(const cfun_conv3d_desc* d, int pass);
bool tc_preferred(const cfun_conv3d_desc* d, int pass);
size_t tc_workspace(const cfun_conv3d_desc* d, int pass);
int tc_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, int nsplit,
                void* ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, int nsplit, void* ws,
                     size_t ws_bytes, cudaStream_t st);
int tc_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st);

// conv_c1.cu: direct CUDA-core kernels for single-channel inputs (stem, first U-Net conv); part of the SIMT algorithm
bool c1_supported(const cfun_conv3d_desc* d, int pass);
int c1_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, cudaStream_t st);
int c1_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st);

// conv_pw.cu: streaming kernels for 1x1x1 convs with <= 8 output channels (segmentation / RPN heads); part of SIMT
bool pw_supported(const cfun_conv3d_desc* d, int pass);
int pw_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, cudaStream_t st);
int pw_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);
int pw_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st);

static int resolve(const cfun_conv3d_desc* d, int pass, int algo) {
  if (algo == CFUN_CONV_ALGO_AUTO) {
    const char* e = getenv("CFUN_CONV_ALGO");  // "simt" pins the CUDA-core path (debug / A-B measurements)
    if (e && e[0] == 's') return CFUN_CONV_ALGO_SIMT;
    const char* m = getenv("CFUN_TC_PASSES");   // bit mask of passes allowed on tensor cores (1 fwd, 2 dgrad, 4 wgrad)
    if (m && !((atoi(m) >> pass) & 1)) return CFUN_CONV_ALGO_SIMT;
    return tc_preferred(d, pass) ? CFUN_CONV_ALGO_TC : CFUN_CONV_ALGO_SIMT;
  }
  return algo;
}
}  // namespace cfun

using namespace cfun;

extern "C" int cfun_conv3d_supported(const cfun_conv3d_desc* d, int pass, int algo) {
  if (!d) return 0;
  if (algo == CFUN_CONV_ALGO_SIMT || algo == CFUN_CONV_ALGO_AUTO) return 1;
  return tc_supported(d, pass) ? 1 : 0;
}

extern "C" int cfun_conv3d_pick_algo(const cfun_conv3d_desc* d, int pass) { return resolve(d, pass, CFUN_CONV_ALGO_AUTO); }

extern "C" size_t cfun_conv3d_workspace_size(const cfun_conv3d_desc* d, int pass, int algo) {
  if (!d) return 0;
  int a = resolve(d, pass, algo);
  if (a == CFUN_CONV_ALGO_SIMT) return simt_workspace(d, pass);
  return tc_workspace(d, pass);
}

extern "C" int cfun_conv3d_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                               int epi_flags, int algo, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(d != nullptr);
  int a = resolve(d, CFUN_PASS_FWD, algo);
  if (a == CFUN_CONV_ALGO_SIMT && c1_supported(d, CFUN_PASS_FWD)) return c1_conv_fwd(d, x, w, bias, y, epi_flags, as_stream(stream));
  if (a == CFUN_CONV_ALGO_SIMT && pw_supported(d, CFUN_PASS_FWD)) return pw_conv_fwd(d, x, w, bias, y, epi_flags, as_stream(stream));
  if (a == CFUN_CONV_ALGO_SIMT) return simt_conv_fwd(d, x, w, bias, y, epi_flags, ws, ws_bytes, as_stream(stream));
  CFUN_CHECK_ARG(tc_supported(d, CFUN_PASS_FWD));
  return tc_conv_fwd(d, x, w, bias, y, epi_flags, a == CFUN_CONV_ALGO_TC1 ? 1 : 3, ws, ws_bytes, as_stream(stream));
}

extern "C" int cfun_conv3d_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, int algo, void* ws,
                                    size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(d != nullptr);
  int a = resolve(d, CFUN_PASS_BWD_DATA, algo);
  if (a == CFUN_CONV_ALGO_SIMT && pw_supported(d, CFUN_PASS_BWD_DATA)) return pw_conv_bwd_data(d, dy, w, dx, as_stream(stream));
  if (a == CFUN_CONV_ALGO_SIMT) return simt_conv_bwd_data(d, dy, w, dx, ws, ws_bytes, as_stream(stream));
  CFUN_CHECK_ARG(tc_supported(d, CFUN_PASS_BWD_DATA));
  return tc_conv_bwd_data(d, dy, w, dx, a == CFUN_CONV_ALGO_TC1 ? 1 : 3, ws, ws_bytes, as_stream(stream));
}

extern "C" int cfun_conv3d_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                                      int algo, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(d != nullptr);
  int a = resolve(d, CFUN_PASS_BWD_WEIGHT, algo);
  if (a == CFUN_CONV_ALGO_SIMT && c1_supported(d, CFUN_PASS_BWD_WEIGHT)) return c1_conv_bwd_weight(d, x, dy, dw, dbias, as_stream(stream));
  if (a == CFUN_CONV_ALGO_SIMT && pw_supported(d, CFUN_PASS_BWD_WEIGHT)) return pw_conv_bwd_weight(d, x, dy, dw, dbias, as_stream(stream));
  if (a == CFUN_CONV_ALGO_SIMT) return simt_conv_bwd_weight(d, x, dy, dw, dbias, ws, ws_bytes, as_stream(stream));
  CFUN_CHECK_ARG(tc_supported(d, CFUN_PASS_BWD_WEIGHT));
  return tc_conv_bwd_weight(d, x, dy, dw, dbias, a == CFUN_CONV_ALGO_TC1 ? 1 : 3, ws, ws_bytes, as_stream(stream));
}
