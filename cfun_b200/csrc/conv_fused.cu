// Fused backward of the 3x3x3 / stride-1 tensor-core convolutions.
//
// The halo-family kernels (conv_tc_halo.cu, conv_tc_hx.cu, conv_tc_wgrad_ds.cu) all consume the same operand format: a
// group-planar split-bf16 pack [C/8][N*(D+2)][H][W][8ch] (hi and lo).  A train step used to write it four times per conv
// (forward: X; data gradient: dY; weight gradient: X and dY again), ~0.3 ms per 40-channel 96^3 tensor each time.  Here
//   * cfun_conv3d_fwd_keep_pack   runs the forward and leaves the X pack in a caller-owned buffer,
//   * cfun_conv3d_bwd_fused       packs dY once and feeds it to the data-gradient kernel AND (with the kept X pack) to the
//                                 d-stacked weight-gradient kernel.
// Replaces the same nn.Conv3d forward / autograd backward as cfun_conv3d_{fwd,bwd_data,bwd_weight} (mask_branch.py:23-89,
// model.py:131-148,713 and loss.backward() at model.py:1640) for the shapes where all three passes run on these kernels.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace cfun {
bool hl_supported(const cfun_conv3d_desc* d, int pass);
bool hx_supported(const cfun_conv3d_desc* d, int pass);
bool ds_supported(const cfun_conv3d_desc* d);
size_t hl_pack_bytes(const cfun_conv3d_desc* d, int pass);
size_t hx_pack_bytes(const cfun_conv3d_desc* d, int pass);
size_t tc_workspace(const cfun_conv3d_desc* d, int pass);
int hl_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               double* stat_acc);
int hx_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               int tapmask, double* stat_acc);
int ds_bwd_weight_packed(const cfun_conv3d_desc* d, __nv_bfloat16* yh, __nv_bfloat16* yl, int gy_pack, __nv_bfloat16* xh,
                         __nv_bfloat16* xl, float* dw, cudaStream_t st);
int launch_pack_act_gp_pad(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, int P,
                           cudaStream_t st);
int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);
bool pack_preact_supported(int C, int G);
int launch_pack_preact_gp_pad(const float* x, const float* scale, float slope, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H,
                              int W, int C, int G, int P, cudaStream_t st);
bool pack_cat_supported(int C1, int C2, int G);
int launch_pack_cat_gp_pad(const float* a, int C1, const float* b, int C2, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H,
                           int W, int G, int P, cudaStream_t st);

static bool fused_ok(const cfun_conv3d_desc* d) {
  const char* e = getenv("CFUN_CONV_FUSED");      // "0": separate fwd / dgrad / wgrad calls (A/B measurements)
  if (e && e[0] == '0') return false;
  if (!d) return false;
  for (int pass = 0; pass < 3; ++pass)
    if (cfun_conv3d_pick_algo(d, pass) != CFUN_CONV_ALGO_TC) return false;
  if (!hl_supported(d, CFUN_PASS_FWD) && !hx_supported(d, CFUN_PASS_FWD)) return false;
  if (!hl_supported(d, CFUN_PASS_BWD_DATA) && !hx_supported(d, CFUN_PASS_BWD_DATA)) return false;
  return ds_supported(d);
}
static size_t act_bytes(const cfun_conv3d_desc* d, int pass) {
  return hl_supported(d, pass) ? hl_pack_bytes(d, pass) : hx_pack_bytes(d, pass);
}
static int run_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
                    void* ws, size_t ws_bytes, __nv_bfloat16* hi, __nv_bfloat16* lo, bool ready, cudaStream_t st,
                    double* stat_acc = nullptr) {
  if (hl_supported(d, pass)) return hl_conv_ex(d, pass, src, w, bias, dst, epi, 3, ws, ws_bytes, hi, lo, ready, st, stat_acc);
  return hx_conv_ex(d, pass, src, w, bias, dst, epi, 3, ws, ws_bytes, hi, lo, ready, st, 0x7FFFFFF, stat_acc);
}
}  // namespace cfun

using namespace cfun;

extern "C" size_t cfun_conv3d_pack_bytes(const cfun_conv3d_desc* d) {
  if (!fused_ok(d)) return 0;
  return 2 * act_bytes(d, CFUN_PASS_FWD);
}

extern "C" size_t cfun_conv3d_bwd_fused_workspace_size(const cfun_conv3d_desc* d) {
  if (!fused_ok(d)) return 0;
  return 2 * act_bytes(d, CFUN_PASS_BWD_DATA) + tc_workspace(d, CFUN_PASS_BWD_DATA) + 4096;
}

extern "C" int cfun_conv3d_fwd_keep_pack(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                                         int epi_flags, void* xpack, size_t xpack_bytes, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(fused_ok(d));
  CFUN_CHECK_ARG(x && w && y && xpack && ws);
  CFUN_CHECK_ARG(!(epi_flags & CFUN_EPI_BIAS) || bias);
  const size_t act = act_bytes(d, CFUN_PASS_FWD);
  CFUN_CHECK_ARG(xpack_bytes >= 2 * act && ((size_t)xpack & 127) == 0);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xpack);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(xpack) + act);
  return run_conv(d, CFUN_PASS_FWD, x, w, bias, y, epi_flags, ws, ws_bytes, hi, lo, false, as_stream(stream));
}

// cfun_conv3d_fwd_keep_pack whose epilogue also accumulates the InstanceNorm statistics of y: stat_acc [N][Cout][2] doubles
// (sum, sum of squares over the sample's voxels), zeroed here -- what cfun_instnorm_stats would compute in a separate pass
// over y; cfun_instnorm_finalize turns them into mean / rstd.
extern "C" int cfun_conv3d_fwd_stats(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                                     int epi_flags, void* xpack, size_t xpack_bytes, double* stat_acc, void* ws, size_t ws_bytes,
                                     void* stream) {
  CFUN_CHECK_ARG(fused_ok(d));
  CFUN_CHECK_ARG(x && w && y && xpack && ws && stat_acc);
  CFUN_CHECK_ARG(!(epi_flags & CFUN_EPI_BIAS) || bias);
  const size_t act = act_bytes(d, CFUN_PASS_FWD);
  CFUN_CHECK_ARG(xpack_bytes >= 2 * act && ((size_t)xpack & 127) == 0);
  cudaStream_t st = as_stream(stream);
  CFUN_CUDA(cudaMemsetAsync(stat_acc, 0, sizeof(double) * 2 * (size_t)d->N * d->Cout, st));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xpack);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(xpack) + act);
  return run_conv(d, CFUN_PASS_FWD, x, w, bias, y, epi_flags, ws, ws_bytes, hi, lo, false, st, stat_acc);
}

// cfun_conv3d_fwd_stats whose input is the channel concatenation [a (C1 channels) | b (C2 channels)], C1 + C2 == d->Cin (the
// U-Net decoder's torch.cat((up, skip), dim=1) -> conv, mask_branch.py:185-205): the operand pack is built straight from the two
// tensors, the concatenated fp32 tensor never exists.  cfun_conv3d_cat_supported: 1 if this call handles the shape.
extern "C" int cfun_conv3d_cat_supported(const cfun_conv3d_desc* d, int C1, int C2) {
  if (!fused_ok(d) || C1 + C2 != d->Cin) return 0;
  return pack_cat_supported(C1, C2, (int)align_up((size_t)d->Cin, 16) / 8) ? 1 : 0;
}
extern "C" int cfun_conv3d_fwd_stats_cat(const cfun_conv3d_desc* d, const float* a, int C1, const float* b, int C2, const float* w,
                                         float* y, void* xpack, size_t xpack_bytes, double* stat_acc, void* ws, size_t ws_bytes,
                                         void* stream) {
  CFUN_CHECK_ARG(cfun_conv3d_cat_supported(d, C1, C2));
  CFUN_CHECK_ARG(a && b && w && y && xpack && ws);
  const size_t act = act_bytes(d, CFUN_PASS_FWD);
  CFUN_CHECK_ARG(xpack_bytes >= 2 * act && ((size_t)xpack & 127) == 0);
  cudaStream_t st = as_stream(stream);
  if (stat_acc) CFUN_CUDA(cudaMemsetAsync(stat_acc, 0, sizeof(double) * 2 * (size_t)d->N * d->Cout, st));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xpack);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(xpack) + act);
  int rc = launch_pack_cat_gp_pad(a, C1, b, C2, hi, lo, d->N, d->Din, d->Hin, d->Win, (int)align_up((size_t)d->Cin, 16) / 8, d->kD / 2, st);
  if (rc != CFUN_OK) return rc;
  return run_conv(d, CFUN_PASS_FWD, nullptr, w, nullptr, y, 0, ws, ws_bytes, hi, lo, true, st, stat_acc);
}

// cfun_conv3d_fwd_keep_pack on leaky_relu(x * scale[n][c], slope) (scale NULL = 1): the activation / Dropout3d channel scale that
// precede the conv (mask_branch.py:127-131: conv3d_c1_2(lrelu(out)), lrelu_conv_c1(lrelu(dropout(out)))) are applied on the
// way into the operand pack; the activated tensor is never written.  The kept pack holds the activated values, so
// cfun_conv3d_bwd_fused returns the gradient w.r.t. the ACTIVATED input; cfun_affine_act_bwd(x, scale, 0, ...) finishes it.
extern "C" int cfun_conv3d_preact_supported(const cfun_conv3d_desc* d) {
  if (!fused_ok(d)) return 0;
  return pack_preact_supported(d->Cin, (int)align_up((size_t)d->Cin, 16) / 8) ? 1 : 0;
}
extern "C" int cfun_conv3d_fwd_keep_pack_preact(const cfun_conv3d_desc* d, const float* x, const float* scale, float slope,
                                                const float* w, const float* bias, float* y, int epi_flags, void* xpack,
                                                size_t xpack_bytes, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(cfun_conv3d_preact_supported(d));
  CFUN_CHECK_ARG(x && w && y && xpack && ws);
  CFUN_CHECK_ARG(!(epi_flags & CFUN_EPI_BIAS) || bias);
  const size_t act = act_bytes(d, CFUN_PASS_FWD);
  CFUN_CHECK_ARG(xpack_bytes >= 2 * act && ((size_t)xpack & 127) == 0);
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xpack);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(xpack) + act);
  int rc = launch_pack_preact_gp_pad(x, scale, slope, hi, lo, d->N, d->Din, d->Hin, d->Win, d->Cin, (int)align_up((size_t)d->Cin, 16) / 8,
                                     d->kD / 2, st);
  if (rc != CFUN_OK) return rc;
  return run_conv(d, CFUN_PASS_FWD, nullptr, w, bias, y, epi_flags, ws, ws_bytes, hi, lo, true, st);
}

// cfun_conv3d_fwd_stats on an X pack the caller has already filled (cfun_instnorm_up2_pack: the decoder's norm -> lrelu ->
// upsample -> conv, where the upsampled fp32 tensor never exists); stat_acc may be NULL
extern "C" int cfun_conv3d_fwd_stats_packed(const cfun_conv3d_desc* d, void* xpack, size_t xpack_bytes, const float* w, float* y,
                                            double* stat_acc, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(fused_ok(d));
  CFUN_CHECK_ARG(w && y && xpack && ws);
  const size_t act = act_bytes(d, CFUN_PASS_FWD);
  CFUN_CHECK_ARG(xpack_bytes >= 2 * act && ((size_t)xpack & 127) == 0);
  cudaStream_t st = as_stream(stream);
  if (stat_acc) CFUN_CUDA(cudaMemsetAsync(stat_acc, 0, sizeof(double) * 2 * (size_t)d->N * d->Cout, st));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(xpack);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(xpack) + act);
  return run_conv(d, CFUN_PASS_FWD, nullptr, w, nullptr, y, 0, ws, ws_bytes, hi, lo, true, st, stat_acc);
}

// geometry of the dY pack cfun_conv3d_bwd_fused builds internally, for callers that produce it themselves
// (cfun_instnorm_bwd_apply_pack): groups of 8 channels, zero planes per side, bytes of hi + lo (0 = shape not eligible)
extern "C" size_t cfun_conv3d_dy_pack_geometry(const cfun_conv3d_desc* d, int* groups, int* pad_planes) {
  if (!fused_ok(d)) return 0;
  if (groups) *groups = (int)align_up((size_t)d->Cout, 16) / 8;
  if (pad_planes) *pad_planes = d->kD / 2;
  return 2 * act_bytes(d, CFUN_PASS_BWD_DATA);
}

// cfun_conv3d_bwd_fused with the dY pack already made (ypack: hi then lo, each half of cfun_conv3d_dy_pack_geometry's bytes)
extern "C" int cfun_conv3d_bwd_fused_packed(const cfun_conv3d_desc* d, const void* xpack, size_t xpack_bytes, void* ypack,
                                            size_t ypack_bytes, const float* w, float* dx, float* dw, void* ws, size_t ws_bytes,
                                            void* stream) {
  CFUN_CHECK_ARG(fused_ok(d));
  CFUN_CHECK_ARG(ypack && w && ws && (dx || dw));
  cudaStream_t st = as_stream(stream);
  const size_t act_y = act_bytes(d, CFUN_PASS_BWD_DATA);
  CFUN_CHECK_ARG(ypack_bytes >= 2 * act_y && ((size_t)ypack & 127) == 0);
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(ypack);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(ypack) + act_y);
  const int gy = (int)align_up((size_t)d->Cout, 16) / 8;
  int rc;
  if (dx) {
    if ((rc = run_conv(d, CFUN_PASS_BWD_DATA, nullptr, w, nullptr, dx, 0, ws, ws_bytes, yh, yl, true, st)) != CFUN_OK) return rc;
  }
  if (dw) {
    const size_t act_x = act_bytes(d, CFUN_PASS_FWD);
    CFUN_CHECK_ARG(xpack && xpack_bytes >= 2 * act_x);
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(xpack));
    __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(const_cast<void*>(xpack)) + act_x);
    if ((rc = ds_bwd_weight_packed(d, yh, yl, gy, xh, xl, dw, st)) != CFUN_OK) return rc;
  }
  return CFUN_OK;
}

extern "C" int cfun_conv3d_bwd_fused(const cfun_conv3d_desc* d, const void* xpack, size_t xpack_bytes, const float* dy,
                                     const float* w, float* dx, float* dw, float* dbias, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(fused_ok(d));
  CFUN_CHECK_ARG(dy && w && ws && (dx || dw || dbias));
  cudaStream_t st = as_stream(stream);
  const size_t act_y = act_bytes(d, CFUN_PASS_BWD_DATA);
  const size_t inner = tc_workspace(d, CFUN_PASS_BWD_DATA);
  const size_t base = align_up((size_t)ws, 1024);
  if (base + 2 * act_y + inner > (size_t)ws + ws_bytes) { set_error("conv3d fused backward: workspace too small"); return CFUN_ERR_WORKSPACE; }
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(base + act_y);
  void* iws = reinterpret_cast<void*>(base + 2 * act_y);
  const int gy = (int)align_up((size_t)d->Cout, 16) / 8;
  int rc;
  if (dx || dw) {
    // zero planes before / after every sample = the kernel's padding (1 for 3^3, 2 for 5^3)
    if ((rc = launch_pack_act_gp_pad(dy, yh, yl, d->N, d->Dout, d->Hout, d->Wout, d->Cout, gy, d->kD / 2, st)) != CFUN_OK) return rc;
  }
  if (dx) {
    if ((rc = run_conv(d, CFUN_PASS_BWD_DATA, nullptr, w, nullptr, dx, 0, iws, ws_bytes - (2 * act_y + (base - (size_t)ws)), yh, yl, true, st)) != CFUN_OK) return rc;
  }
  if (dw) {
    const size_t act_x = act_bytes(d, CFUN_PASS_FWD);
    CFUN_CHECK_ARG(xpack && xpack_bytes >= 2 * act_x);
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(xpack));
    __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(const_cast<void*>(xpack)) + act_x);
    if ((rc = ds_bwd_weight_packed(d, yh, yl, gy, xh, xl, dw, st)) != CFUN_OK) return rc;
  }
  if (dbias) return simt_bias_grad(dy, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}
