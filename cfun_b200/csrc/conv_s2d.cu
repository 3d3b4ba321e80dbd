// Stride-2 3x3x3 / pad 1 convolutions (the four down-sampling convs of the U-Net context pathway, mask_branch.py:33-52)
// on the tensor-core path, by space-to-depth:
//
//   y[o] = sum_{k in 0..2} x[2 o + k - 1] w[k]   per axis;   write 2 o + k - 1 = 2 (o + s) + q  with parity q in {0,1}:
//       k = 0 -> (s,q) = (-1,1),   k = 1 -> (0,0),   k = 2 -> (0,1),   (s,q) = (-1,0) never occurs.
//
// So with X'[n, (qd,qh,qw,ci), d', h', w'] = X[n, ci, 2d'+qd, 2h'+qh, 2w'+qw]  (8 Cin channels, half the extent) the
// strided conv IS a stride-1, 2x2x2-tap, pad-1 convolution of X' whose weights W'[co, (q,ci), k'] hold the 27 original
// taps and 37 structural zeros (k' = s + 1).  Forward, data gradient and weight gradient of that dense stride-1 problem
// run on the existing tcgen05 implicit-GEMM kernels (conv_tc.cu / conv_tc_wgrad.cu); the 64/27 extra MMA work is cheap
// next to the CUDA-core kernels it replaces (20->40 @ 96^3 data gradient: 4.3 ms -> see DESIGN.md).
// The permutations X <-> X' are pure fp32 copies (NDHWC rows of Cin floats move as a block); when 8 Cin exceeds the
// weight-gradient kernel's 256-channel limit X' is written as P parity-group tensors of 8 Cin / P channels.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace cfun {

bool tc_capable(const cfun_conv3d_desc* d, int pass);          // conv_tc.cu: geometry check without the size policy
size_t tc_workspace(const cfun_conv3d_desc* d, int pass);
int tc_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, int nsplit,
                void* ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, int nsplit, void* ws,
                     size_t ws_bytes, cudaStream_t st);
int tc_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st);
int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);
bool ds_masked_supported(const cfun_conv3d_desc* d, int tap_mask);          // conv_tc_wgrad_ds.cu
size_t ds_masked_workspace(const cfun_conv3d_desc* d, int tap_mask);
int ds_conv_bwd_weight_masked(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, int tap_mask, int kd_mask,
                              void* ws, size_t ws_bytes, cudaStream_t st);
bool hx_supported(const cfun_conv3d_desc* d, int pass);                      // conv_tc_hx.cu
size_t hx_workspace(const cfun_conv3d_desc* d, int pass);
int hx_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               int tapmask, double* stat_acc = nullptr);
constexpr int S2D_MASK27 = 0x361B;    // taps (kd,kh,kw) in {0,1}^3 of the embedding 3x3x3 kernel: bits {0,1,3,4,9,10,12,13}
constexpr int S2D_TAPMASK = 0x1B;     // (kh,kw) in {0,1}^2 -> kh*3+kw in {0,1,3,4}
constexpr int S2D_KDMASK = 0x3;       // kd in {0,1}

// X (N,D,H,W,C) -> P tensors (N,D/2,H/2,W/2, 8C/P); parity class q = (qd*2+qh)*2+qw lives in tensor q / (8/P), channel
// block q % (8/P).  inverse = true copies the other way (the data gradient's dX' -> dX).
template <bool INV>
__global__ void __launch_bounds__(256) s2d_kernel(float* __restrict__ x, float* __restrict__ xp, int N, int D, int H, int W, int C,
                                                  int P) {
  const int c4 = C >> 2;
  const long long total = (long long)N * D * H * W * c4;
  const int qpt = 8 / P;                                  // parity classes per output tensor
  const long long rows_half = (long long)N * (D / 2) * (H / 2) * (W / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4);
    long long r = i / c4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const int n = (int)(r / D);
    const int q = ((d & 1) * 2 + (h & 1)) * 2 + (w & 1);
    const long long row2 = (((long long)n * (D / 2) + (d >> 1)) * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
    float4* src = reinterpret_cast<float4*>(x) + i;
    float4* dst = reinterpret_cast<float4*>(xp + ((long long)(q / qpt) * rows_half + row2) * (long long)(qpt * C) + (q % qpt) * C) + cc;
    if (INV) *src = *dst;
    else *dst = *src;
  }
}

// W (Cout, Cin, 27) -> P tensors (Cout, 8Cin/P, 8) with structural zeros; inverse gathers dW' back into dW
template <bool INV>
__global__ void __launch_bounds__(256) w_s2d_kernel(float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int P) {
  const int qpt = 8 / P;
  const long long total = (long long)Cout * 8 * Cin * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(i % 8);                          // k' = (k'd, k'h, k'w)
    long long r = i / 8;
    const int ci = (int)(r % Cin); r /= Cin;
    const int q = (int)(r % 8);
    const int co = (int)(r / 8);
    int k[3];
    bool live = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ka = (kp >> (2 - a)) & 1, qa = (q >> (2 - a)) & 1;
      if (ka == 0) { live = live && qa == 1; k[a] = 0; }
      else k[a] = qa ? 2 : 1;
    }
    const long long o = ((long long)(q / qpt) * Cout + co) * (long long)(qpt * Cin) * 8 + ((long long)(q % qpt) * Cin + ci) * 8 + kp;
    const long long s = ((long long)co * Cin + ci) * 27 + (k[0] * 3 + k[1]) * 3 + k[2];
    if (INV) { if (live) w[s] = wp[o]; }
    else wp[o] = live ? w[s] : 0.f;
  }
}

// W (Cout, Cin, 27) -> W3 (Cout, 8 Cin, 27): the 2x2x2 space-to-depth kernel as the {0,1}^3 corner of a 3x3x3 / pad-1 kernel
// (W3[co][(q,ci)][k'] = W[co][ci][k] with per axis (k',q) -> k: (0,1)->0, (1,0)->1, (1,1)->2; zero elsewhere)
__global__ void __launch_bounds__(256) w_s2d3_embed_kernel(const float* __restrict__ w, float* __restrict__ w3, int Cout, int Cin) {
  const long long total = (long long)Cout * 8 * Cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t3 = (int)(i % 27);
    long long r = i / 27;
    const int ci = (int)(r % Cin); r /= Cin;
    const int q = (int)(r % 8);
    const int co = (int)(r / 8);
    const int kp[3] = {t3 / 9, (t3 / 3) % 3, t3 % 3};
    bool live = true;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int qa = (q >> (2 - a)) & 1;
      int ka = 0;
      if (kp[a] == 0) { live = live && qa == 1; ka = 0; }
      else if (kp[a] == 1) ka = qa ? 2 : 1;
      else live = false;
      k = k * 3 + ka;
    }
    w3[i] = live ? w[((long long)co * Cin + ci) * 27 + k] : 0.f;
  }
}

// dW3 (Cout, 8 Cin, 27) of the embedding 3^3 kernel (only the taps with kd,kh,kw in {0,1} are filled) -> dW (Cout, Cin, 27)
__global__ void __launch_bounds__(256) w_s2d3_gather_kernel(const float* __restrict__ w3, float* __restrict__ w, int Cout, int Cin) {
  const long long total = (long long)Cout * Cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 27);
    long long r = i / 27;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    const int k[3] = {tap / 9, (tap / 3) % 3, tap % 3};
    int q = 0, t3 = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {          // k = 0 -> (k',q) = (0,1); k = 1 -> (1,0); k = 2 -> (1,1)
      const int kp = k[a] == 0 ? 0 : 1, qa = k[a] == 1 ? 0 : 1;
      q = q * 2 + qa;
      t3 = t3 * 3 + kp;
    }
    w[i] = w3[((long long)co * 8 * Cin + (long long)q * Cin + ci) * 27 + t3];
  }
}

struct S2dPlan {
  cfun_conv3d_desc d2;       // the stride-1 problem on X' (8 Cin channels): 2x2x2, or its 3x3x3 embedding when `embed`
  int P;
  bool embed;                // forward / data gradient through the tap-masked halo kernel (conv_tc_hx.cu) on the 3^3 embedding
  size_t act, wgt, inner, total;
};

static bool make_s2d_plan(const cfun_conv3d_desc* d, int pass, S2dPlan& pl) {
  if (!d) return false;
  if (d->sD != 2 || d->sH != 2 || d->sW != 2 || d->kD != 3 || d->kH != 3 || d->kW != 3 || d->pD != 1 || d->pH != 1 || d->pW != 1) return false;
  if ((d->Din | d->Hin | d->Win) & 1) return false;
  if (d->Dout != d->Din / 2 || d->Hout != d->Hin / 2 || d->Wout != d->Win / 2) return false;
  if (d->Cin & 3) return false;
  pl.P = 1;
  cfun_conv3d_desc& e = pl.d2;
  e = *d;
  e.Cin = 8 * d->Cin;
  e.Din = d->Din / 2; e.Hin = d->Hin / 2; e.Win = d->Win / 2;
  e.sD = e.sH = e.sW = 1;
  e.pD = e.pH = e.pW = 1;
  pl.act = align_up((size_t)d->N * d->Din * d->Hin * d->Win * d->Cin * 4, 1024);
  if (pass == CFUN_PASS_BWD_WEIGHT) {
    // weight gradient: the 2x2x2 kernel is the {0,1}^3 corner of a 3x3x3 / pad-1 kernel -> masked d-stacked kernel
    e.kD = e.kH = e.kW = 3;
    if (8 * d->Cin > 320) return false;       // >= 8 channel slices: the CUDA-core kernel wins (80->160: 0.27 vs 0.34 ms)
    if (!ds_masked_supported(&e, S2D_TAPMASK)) return false;
    pl.wgt = align_up((size_t)d->Cout * 8 * d->Cin * 27 * 4, 1024);
    pl.inner = ds_masked_workspace(&e, S2D_TAPMASK);
    pl.total = pl.act + pl.wgt + pl.inner + 2048;
    return pl.inner > 0;
  }
  // forward / data gradient: the halo kernel on the 3^3 embedding with a tap mask (the activation halo is loaded once
  // per K chunk instead of once per tap: 20->40 @ 96^3 forward 0.98 -> see DESIGN.md); else the per-tap kernel on 2^3
  pl.embed = false;
  {
    const char* x = getenv("CFUN_TC_S2D");
    e.kD = e.kH = e.kW = 3;
    // (only up to 8 Cin = 320 channels: beyond, 80->160 @ 24^3 measures 0.22 vs 0.16 ms on the per-tap kernel)
    if (!(x && x[0] == 'p') && 8 * d->Cin <= 320 && hx_supported(&e, pass)) {
      pl.embed = true;
      pl.wgt = align_up((size_t)d->Cout * 8 * d->Cin * 27 * 4, 1024);
      pl.inner = hx_workspace(&e, pass);
      pl.total = pl.act + pl.wgt + pl.inner + 2048;
      return pl.inner > 0;
    }
  }
  e.kD = e.kH = e.kW = 2;
  if (!tc_capable(&e, pass)) return false;
  pl.wgt = align_up((size_t)d->Cout * 8 * d->Cin * 8 * 4, 1024);
  pl.inner = tc_workspace(&e, pass);
  pl.total = pl.act + pl.wgt + pl.inner + 2048;
  return pl.inner > 0;
}

bool s2d_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_TC_S2D");           // "0" keeps strided convs on CUDA cores (A/B measurements)
  if (e && e[0] == '0') return false;
  if (pass == CFUN_PASS_BWD_WEIGHT && e && e[0] == 'n') return false;      // "nw": forward / data gradient only
  S2dPlan pl;
  return make_s2d_plan(d, pass, pl);
}
size_t s2d_workspace(const cfun_conv3d_desc* d, int pass) {
  S2dPlan pl;
  return make_s2d_plan(d, pass, pl) ? pl.total : 0;
}

static inline unsigned grid_for(long long total) { return (unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()); }

int s2d_conv(const cfun_conv3d_desc* d, int pass, const float* a, const float* b, const float* bias, float* out, float* dbias,
             int epi, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  S2dPlan pl;
  CFUN_CHECK_ARG(make_s2d_plan(d, pass, pl));
  CFUN_CHECK_ARG(a && b && out && ws);
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d s2d: workspace too small"); return CFUN_ERR_WORKSPACE; }
  float* xp = reinterpret_cast<float*>(base);
  float* wp = reinterpret_cast<float*>(base + pl.act);
  void* inner = reinterpret_cast<void*>(base + pl.act + pl.wgt);
  const long long xtotal = (long long)d->N * d->Din * d->Hin * d->Win * (d->Cin >> 2);
  const long long wtotal = (long long)d->Cout * 8 * d->Cin * 8;
  int rc;
  if (pl.embed && pass != CFUN_PASS_BWD_WEIGHT) {
    const long long w3total = (long long)d->Cout * 8 * d->Cin * 27;
    w_s2d3_embed_kernel<<<grid_for(w3total), 256, 0, st>>>(b, wp, d->Cout, d->Cin);
    CFUN_LAUNCH_CHECK();
    if (pass == CFUN_PASS_FWD) {
      s2d_kernel<false><<<grid_for(xtotal), 256, 0, st>>>(const_cast<float*>(a), xp, d->N, d->Din, d->Hin, d->Win, d->Cin, 1);
      CFUN_LAUNCH_CHECK();
      return hx_conv_ex(&pl.d2, CFUN_PASS_FWD, xp, wp, bias, out, epi, nsplit, inner, pl.inner, nullptr, nullptr, false, st, S2D_MASK27);
    }
    if ((rc = hx_conv_ex(&pl.d2, CFUN_PASS_BWD_DATA, a, wp, nullptr, xp, 0, nsplit, inner, pl.inner, nullptr, nullptr, false, st, S2D_MASK27)) != CFUN_OK) return rc;
    s2d_kernel<true><<<grid_for(xtotal), 256, 0, st>>>(out, xp, d->N, d->Din, d->Hin, d->Win, d->Cin, 1);
    CFUN_LAUNCH_CHECK();
    return CFUN_OK;
  }
  if (pass == CFUN_PASS_FWD) {                     // a = x, b = w
    s2d_kernel<false><<<grid_for(xtotal), 256, 0, st>>>(const_cast<float*>(a), xp, d->N, d->Din, d->Hin, d->Win, d->Cin, 1);
    CFUN_LAUNCH_CHECK();
    w_s2d_kernel<false><<<grid_for(wtotal), 256, 0, st>>>(const_cast<float*>(b), wp, d->Cout, d->Cin, 1);
    CFUN_LAUNCH_CHECK();
    return tc_conv_fwd(&pl.d2, xp, wp, bias, out, epi, nsplit, inner, pl.inner, st);
  }
  if (pass == CFUN_PASS_BWD_DATA) {                // a = dy, b = w, out = dx
    w_s2d_kernel<false><<<grid_for(wtotal), 256, 0, st>>>(const_cast<float*>(b), wp, d->Cout, d->Cin, 1);
    CFUN_LAUNCH_CHECK();
    if ((rc = tc_conv_bwd_data(&pl.d2, a, wp, xp, nsplit, inner, pl.inner, st)) != CFUN_OK) return rc;
    s2d_kernel<true><<<grid_for(xtotal), 256, 0, st>>>(out, xp, d->N, d->Din, d->Hin, d->Win, d->Cin, 1);
    CFUN_LAUNCH_CHECK();
    return CFUN_OK;
  }
  // weight gradient: a = x, b = dy, out = dw
  s2d_kernel<false><<<grid_for(xtotal), 256, 0, st>>>(const_cast<float*>(a), xp, d->N, d->Din, d->Hin, d->Win, d->Cin, 1);
  CFUN_LAUNCH_CHECK();
  if ((rc = ds_conv_bwd_weight_masked(&pl.d2, xp, b, wp, S2D_TAPMASK, S2D_KDMASK, inner, pl.inner, st)) != CFUN_OK) return rc;
  {
    const long long tot = (long long)d->Cout * d->Cin * 27;
    w_s2d3_gather_kernel<<<grid_for(tot), 256, 0, st>>>(wp, out, d->Cout, d->Cin);
    CFUN_LAUNCH_CHECK();
  }
  if (dbias) return simt_bias_grad(b, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}

}  // namespace cfun
