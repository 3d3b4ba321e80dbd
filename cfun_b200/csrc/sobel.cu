// 3-D Sobel edge loss, forward and backward, fused over all (RoI, class) planes.
// Replaces model.compute_mrcnn_mask_edge_loss (model.py:938-981): 56 tiny conv3d launches + ~10 elementwise passes
// per plane in the reference.  Quirks preserved: magnitude = sqrt(g0^2 + g1^2 + g0^2) (response 0 twice, the third
// Sobel response unused, model.py:969-972); classes are the literal range(7) -> channels 1..7; divide by #positives.
#include "common.cuh"

namespace cfun {

// Sobel bank as built in model.py:947-952 (cross-correlation weights [kD][kH][kW]):
//   k0[a][b][c] = s[a] * d[b] * s[c]   (derivative along H),  k1[a][b][c] = d[a] * s[b] * s[c]  (derivative along D)
__device__ __forceinline__ float sm3(int i) { return i == 1 ? 2.f : 1.f; }
__device__ __forceinline__ float df3(int i) { return i == 0 ? 1.f : (i == 1 ? 0.f : -1.f); }

constexpr int NCLS_MAX = 8;

template <bool BWD>
__global__ void __launch_bounds__(128) sobel_pass1_kernel(const float* __restrict__ pred, const long long* __restrict__ tgt,
                                                          int P, int M, int ncls, double* __restrict__ loss_acc,
                                                          const float* __restrict__ grad_scale, float* __restrict__ A) {
  const int Mo = M - 2;
  const long long per = (long long)Mo * Mo * Mo;
  const long long total = (long long)P * per;
  double local = 0.0;
  const float gs = BWD ? (*grad_scale) * (2.0f / ((float)per * (float)P)) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i / per);
    long long r = i % per;
    int x = (int)(r % Mo);
    int y = (int)((r / Mo) % Mo);
    int z = (int)(r / ((long long)Mo * Mo));
    float gp0[7], gp1[7], gt0[7], gt1[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) gp0[j] = gp1[j] = gt0[j] = gt1[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float w0 = sm3(a) * df3(b) * sm3(c);
          const float w1 = df3(a) * sm3(b) * sm3(c);
          if (w0 == 0.f && w1 == 0.f) continue;
          long long vox = (((long long)p * M + (z + a)) * M + (y + b)) * M + (x + c);
          const float* pv = pred + vox * ncls;
          int t = (int)tgt[vox];
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            float pj = __ldg(pv + j + 1);
            float tj = (t == j + 1) ? 1.f : 0.f;
            gp0[j] = fmaf(w0, pj, gp0[j]); gp1[j] = fmaf(w1, pj, gp1[j]);
            gt0[j] = fmaf(w0, tj, gt0[j]); gt1[j] = fmaf(w1, tj, gt1[j]);
          }
        }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      float mp = sqrtf(gp0[j] * gp0[j] + gp1[j] * gp1[j] + gp0[j] * gp0[j]);
      float mt = sqrtf(gt0[j] * gt0[j] + gt1[j] * gt1[j] + gt0[j] * gt0[j]);
      float diff = mp - mt;
      if (!BWD) {
        local += (double)diff * (double)diff;
      } else {
        // autograd of mse(sqrt(g0^2+g1^2+g0^2)): dmag/(2*mag) * (4*g0 , 2*g1); 0/0 -> NaN exactly like torch
        float dm = gs * diff;
        float h = dm / (2.f * mp);
        A[(i * 7 + j) * 2 + 0] = h * (2.f * gp0[j]) + h * (2.f * gp0[j]);
        A[(i * 7 + j) * 2 + 1] = h * (2.f * gp1[j]);
      }
    }
  }
  if (!BWD) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_acc, local);
  }
}

__global__ void sobel_finalize_kernel(const double* __restrict__ acc, double denom, float* __restrict__ loss) {
  *loss = (float)(*acc / denom);
}

// transposed stencil: dpred[p, v, j+1] = sum_taps k0[tap]*A0[v - tap] + k1[tap]*A1[v - tap]
__global__ void __launch_bounds__(128) sobel_pass2_kernel(const float* __restrict__ A, int P, int M, int ncls,
                                                          float* __restrict__ dpred) {
  const int Mo = M - 2;
  const long long per_in = (long long)M * M * M, per_out = (long long)Mo * Mo * Mo;
  const long long total = (long long)P * per_in;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i / per_in);
    long long r = i % per_in;
    int x = (int)(r % M);
    int y = (int)((r / M) % M);
    int z = (int)(r / ((long long)M * M));
    float g[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) g[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float w0 = sm3(a) * df3(b) * sm3(c);
          const float w1 = df3(a) * sm3(b) * sm3(c);
          if (w0 == 0.f && w1 == 0.f) continue;
          int oz = z - a, oy = y - b, ox = x - c;
          if ((unsigned)oz >= (unsigned)Mo || (unsigned)oy >= (unsigned)Mo || (unsigned)ox >= (unsigned)Mo) continue;
          const float* ap = A + (((long long)p * per_out + ((long long)oz * Mo + oy) * Mo + ox) * 7) * 2;
#pragma unroll
          for (int j = 0; j < 7; ++j) g[j] += w0 * __ldg(ap + 2 * j) + w1 * __ldg(ap + 2 * j + 1);
        }
    float* o = dpred + i * ncls;
    o[0] = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) o[j + 1] = g[j];
    for (int j = 8; j < ncls; ++j) o[j] = 0.f;
  }
}

}  // namespace cfun

using namespace cfun;

extern "C" size_t cfun_sobel_edge_workspace_size(int P, int M, int ncls) {
  (void)ncls;
  if (P <= 0 || M <= 2) return 256;
  size_t per = (size_t)(M - 2) * (M - 2) * (M - 2);
  return 256 + (size_t)P * per * 7 * 2 * sizeof(float);
}

extern "C" int cfun_sobel_edge_loss_fwd(const float* pred, const long long* tgt_index, int P, int M, int ncls, float* loss,
                                        void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(pred && tgt_index && loss && ws && P > 0 && M > 2 && ncls >= 8 && ws_bytes >= 256);
  cudaStream_t st = as_stream(stream);
  double* acc = reinterpret_cast<double*>(align_up((size_t)ws, 16));
  CFUN_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  long long per = (long long)(M - 2) * (M - 2) * (M - 2);
  long long total = (long long)P * per;
  sobel_pass1_kernel<false><<<(unsigned)std::min<long long>(cdiv(total, 128), 64LL * num_sms()), 128, 0, st>>>(
      pred, tgt_index, P, M, ncls, acc, nullptr, nullptr);
  CFUN_LAUNCH_CHECK();
  sobel_finalize_kernel<<<1, 1, 0, st>>>(acc, (double)per * (double)P, loss);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_sobel_edge_loss_bwd(const float* pred, const long long* tgt_index, int P, int M, int ncls,
                                        const float* grad_scale, float* dpred, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(pred && tgt_index && grad_scale && dpred && ws && P > 0 && M > 2 && ncls >= 8);
  if (ws_bytes < cfun_sobel_edge_workspace_size(P, M, ncls)) { set_error("sobel workspace too small"); return CFUN_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  float* A = reinterpret_cast<float*>(align_up((size_t)ws, 16) + 64);
  long long per = (long long)(M - 2) * (M - 2) * (M - 2);
  long long total = (long long)P * per;
  sobel_pass1_kernel<true><<<(unsigned)std::min<long long>(cdiv(total, 128), 64LL * num_sms()), 128, 0, st>>>(
      pred, tgt_index, P, M, ncls, nullptr, grad_scale, A);
  CFUN_LAUNCH_CHECK();
  long long tin = (long long)P * M * M * M;
  sobel_pass2_kernel<<<(unsigned)std::min<long long>(cdiv(tin, 128), 64LL * num_sms()), 128, 0, st>>>(A, P, M, ncls, dpred);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
