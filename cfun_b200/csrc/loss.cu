// Mask cross-entropy of the U-Net head, forward and backward, one pass each.
// Replaces nn.CrossEntropyLoss over the mask logits (reference model.py:909-935: argmax of the one-hot target vs the logits
// [P, ncls, d, h, w]; LiTS_2017/model.py:926 adds the class weights [1, 1, 100]): torch runs log_softmax (read + write of the
// 113 MB logits tensor), nll_loss (read) and two more passes backward.  Here: forward reads logits + target once and reduces;
// backward reads them once more and writes dlogits = g * w_y * (softmax - onehot) / sum(w_y).
// Layout: logits channels-last [V][C] fp32 (V = P*d*h*w voxels), target int64 [V]; C <= 16.
#include "common.cuh"

namespace cfun {

constexpr int CE_MAX_C = 16;

template <bool BWD>
__global__ void __launch_bounds__(256) mask_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target, long long V,
                                                      int C, const float* __restrict__ weight, double* __restrict__ acc,
                                                      const float* __restrict__ gscale, float* __restrict__ dlogits) {
  double num = 0.0, den = 0.0;
  float scale = 0.f;
  if (BWD) scale = (*gscale) / (float)acc[1];
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const float* x = logits + v * C;
    float xv[CE_MAX_C];
    if ((C & 3) == 0) {
#pragma unroll
      for (int c4 = 0; c4 < CE_MAX_C / 4; ++c4)
        if (c4 * 4 < C) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(x) + c4);
          xv[c4 * 4] = t.x; xv[c4 * 4 + 1] = t.y; xv[c4 * 4 + 2] = t.z; xv[c4 * 4 + 3] = t.w;
        }
    } else {
#pragma unroll
      for (int c = 0; c < CE_MAX_C; ++c)
        if (c < C) xv[c] = __ldg(x + c);
    }
    const int y = (int)target[v];
    float m = xv[0];
#pragma unroll
    for (int c = 1; c < CE_MAX_C; ++c)
      if (c < C) m = fmaxf(m, xv[c]);
    float s = 0.f, xy = 0.f;
#pragma unroll
    for (int c = 0; c < CE_MAX_C; ++c)
      if (c < C) {
        s += expf(xv[c] - m);
        if (c == y) xy = xv[c];
      }
    const float wy = weight ? __ldg(weight + y) : 1.f;
    if (!BWD) {
      const float lse = m + logf(s);
      num += (double)((lse - xy) * wy);
      den += (double)wy;
    } else {
      const float k = scale * wy, inv = 1.f / s;
      float* o = dlogits + v * C;
#pragma unroll
      for (int c = 0; c < CE_MAX_C; ++c)
        if (c < C) o[c] = k * (expf(xv[c] - m) * inv - (c == y ? 1.f : 0.f));
    }
  }
  if (!BWD) {
    num = warp_sum(num);
    den = warp_sum(den);
    if ((threadIdx.x & 31) == 0) { atomicAdd(acc, num); atomicAdd(acc + 1, den); }
  }
}

__global__ void mask_ce_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)(acc[0] / acc[1]); }

}  // namespace cfun

using namespace cfun;

// acc: 2 doubles (sum of weighted losses, sum of weights), zeroed here; kept by the caller for the backward
extern "C" int cfun_mask_ce_fwd(const float* logits, const long long* target, long long V, int C, const float* weight, double* acc,
                                float* loss, void* stream) {
  CFUN_CHECK_ARG(logits && target && acc && loss && V > 0 && C >= 2 && C <= CE_MAX_C);
  cudaStream_t st = as_stream(stream);
  CFUN_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  mask_ce_kernel<false><<<(unsigned)std::min<long long>(cdiv(V, 256), 16LL * num_sms()), 256, 0, st>>>(logits, target, V, C, weight, acc, nullptr, nullptr);
  CFUN_LAUNCH_CHECK();
  mask_ce_finalize_kernel<<<1, 1, 0, st>>>(acc, loss);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_mask_ce_bwd(const float* logits, const long long* target, long long V, int C, const float* weight, const double* acc,
                                const float* grad_scale, float* dlogits, void* stream) {
  CFUN_CHECK_ARG(logits && target && acc && grad_scale && dlogits && V > 0 && C >= 2 && C <= CE_MAX_C);
  mask_ce_kernel<true><<<(unsigned)std::min<long long>(cdiv(V, 256), 16LL * num_sms()), 256, 0, as_stream(stream)>>>(
      logits, target, V, C, weight, const_cast<double*>(acc), grad_scale, dlogits);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
