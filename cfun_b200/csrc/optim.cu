// Optimizer tail over one flat fp32 buffer: global-norm clip (clip_grad_norm_(params, 5.0), model.py:1641) fused with
// SGD + momentum + selective weight decay (model.py:1538-1545, 1643).
#include "common.cuh"

namespace cfun {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = g[i];
    s += (double)v * (double)v;
  }
  s = warp_sum(s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom,
                                                  const unsigned char* __restrict__ wd_mask, long long n,
                                                  const double* __restrict__ sumsq, float max_norm, float lr, float momentum,
                                                  float weight_decay) {
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
  float norm = (float)sqrt(*sumsq);
  float coef = max_norm / (norm + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float d = g[i] * coef;
    if (wd_mask == nullptr || wd_mask[i]) d = fmaf(weight_decay, w, d);
    float b = fmaf(momentum, mom[i], d);
    mom[i] = b;
    p[i] = w - lr * b;
  }
}

}  // namespace cfun

using namespace cfun;

extern "C" int cfun_sumsq(const float* g, long long n, double* out_acc, void* stream) {
  CFUN_CHECK_ARG(n >= 0 && out_acc);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(g);
  sumsq_kernel<<<(unsigned)std::min<long long>(cdiv(n, 1024), 8LL * num_sms()), 256, 0, as_stream(stream)>>>(g, n, out_acc);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_sgd_clip_step(float* p, const float* g, float* mom, const unsigned char* wd_mask, long long n,
                                  const double* sumsq, float max_norm, float lr, float momentum, float weight_decay,
                                  int first_step, void* stream) {
  (void)first_step;  // momentum buffers start at zero, which makes the first step buf = d_p like torch.optim.SGD
  CFUN_CHECK_ARG(n >= 0 && sumsq);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(p && g && mom);
  sgd_kernel<<<(unsigned)std::min<long long>(cdiv(n, 1024), 8LL * num_sms()), 256, 0, as_stream(stream)>>>(
      p, g, mom, wd_mask, n, sumsq, max_norm, lr, momentum, weight_decay);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
