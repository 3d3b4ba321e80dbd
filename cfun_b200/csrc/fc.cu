// Classifier.conv1 (model.py:758): a conv whose kernel equals its input extent, i.e. y[M,Nout] = x[M,K] . w[Nout,K]^T with
// K = Cin*12^3 = 221184, M = #RoIs (12 in training, 64 at inference).  28.3 M of the model's 41.35 M parameters live
// here: all three passes are weight-bandwidth bound (113 MB), so each reads / writes w exactly once.
#include "common.cuh"

namespace cfun {

__global__ void fc_init_kernel(float* __restrict__ y, const float* __restrict__ bias, int M, int Nout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M * Nout) y[i] = bias ? bias[i % Nout] : 0.f;
}

// block: one k-slab of KSLAB, one 128-wide n tile; thread (n = t % 128, mg = t / 128) owns rows m = mg + 2*i
template <int MT>  // number of 16-row tiles of M
__global__ void __launch_bounds__(256) fc_fwd_kernel(int M, int Nout, long long K, const float* __restrict__ x,
                                                     const float* __restrict__ w, float* __restrict__ y, int kslab) {
  constexpr int KT = 32;
  __shared__ float Ws[KT][129];
  __shared__ float Xs[16 * MT][KT + 1];
  const int t = threadIdx.x;
  const int n_local = t & 127, mg = t >> 7;
  const int n0 = blockIdx.y * 128;
  const long long kbeg = (long long)blockIdx.x * kslab;
  const long long kend = kbeg + kslab < K ? kbeg + kslab : K;
  float acc[8 * MT];
#pragma unroll
  for (int i = 0; i < 8 * MT; ++i) acc[i] = 0.f;
  for (long long k0 = kbeg; k0 < kend; k0 += KT) {
    // W tile: 128 rows x 32 k, coalesced along k
    for (int e = t; e < 128 * KT; e += 256) {
      int r = e / KT, kk = e % KT;
      float v = 0.f;
      if (n0 + r < Nout && k0 + kk < kend) v = __ldg(w + (long long)(n0 + r) * K + k0 + kk);
      Ws[kk][r] = v;
    }
    for (int e = t; e < 16 * MT * KT; e += 256) {
      int r = e / KT, kk = e % KT;
      float v = 0.f;
      if (r < M && k0 + kk < kend) v = __ldg(x + (long long)r * K + k0 + kk);
      Xs[r][kk] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KT; ++kk) {
      float wv = Ws[kk][n_local];
#pragma unroll
      for (int i = 0; i < 8 * MT; ++i) acc[i] = fmaf(Xs[mg + 2 * i][kk], wv, acc[i]);
    }
    __syncthreads();
  }
  if (n0 + n_local < Nout) {
#pragma unroll
    for (int i = 0; i < 8 * MT; ++i) {
      int m = mg + 2 * i;
      if (m < M) atomicAdd(y + (long long)m * Nout + n0 + n_local, acc[i]);
    }
  }
}

// dx[m,k] = sum_n dy[m,n] w[n,k]   (M <= 16)
__global__ void __launch_bounds__(256) fc_bwd_data_kernel(int M, int Nout, long long K, const float* __restrict__ dy,
                                                          const float* __restrict__ w, float* __restrict__ dx) {
  extern __shared__ float dys[];  // [Nout][16]
  for (int e = threadIdx.x; e < Nout * 16; e += blockDim.x) {
    int n = e / 16, m = e % 16;
    dys[e] = m < M ? dy[(long long)m * Nout + n] : 0.f;
  }
  __syncthreads();
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    for (int n = 0; n < Nout; ++n) {
      float wv = __ldg(w + (long long)n * K + k);
      const float4* d4 = reinterpret_cast<const float4*>(dys + n * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 d = d4[q];
        acc[4 * q] = fmaf(d.x, wv, acc[4 * q]); acc[4 * q + 1] = fmaf(d.y, wv, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(d.z, wv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(d.w, wv, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (m < M) dx[(long long)m * K + k] = acc[m];
  }
}

// dw[n,k] (+)= sum_m dy[m,n] x[m,k]   (M <= 16 rows per launch; accumulate = 1 adds to what an earlier row chunk wrote)
__global__ void __launch_bounds__(256) fc_bwd_weight_kernel(int M, int Nout, long long K, const float* __restrict__ dy,
                                                            const float* __restrict__ x, float* __restrict__ dw, int accumulate) {
  extern __shared__ float dys[];  // [Nout][16]
  for (int e = threadIdx.x; e < Nout * 16; e += blockDim.x) {
    int n = e / 16, m = e % 16;
    dys[e] = m < M ? dy[(long long)m * Nout + n] : 0.f;
  }
  __syncthreads();
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
    float xv[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) xv[m] = m < M ? __ldg(x + (long long)m * K + k) : 0.f;
    for (int n = 0; n < Nout; ++n) {
      const float4* d4 = reinterpret_cast<const float4*>(dys + n * 16);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 d = d4[q];
        s = fmaf(d.x, xv[4 * q], s); s = fmaf(d.y, xv[4 * q + 1], s);
        s = fmaf(d.z, xv[4 * q + 2], s); s = fmaf(d.w, xv[4 * q + 3], s);
      }
      float* o = dw + (long long)n * K + k;
      *o = accumulate ? *o + s : s;
    }
  }
}

__global__ void fc_dbias_kernel(int M, int Nout, const float* __restrict__ dy, float* __restrict__ db) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nout) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += dy[(long long)m * Nout + n];
  db[n] = s;
}

}  // namespace cfun

using namespace cfun;

extern "C" int cfun_fc_fwd(int M, int Nout, long long K, const float* x, const float* w, const float* bias, float* y,
                           void* stream) {
  CFUN_CHECK_ARG(M >= 0 && Nout > 0 && K > 0);
  if (M == 0) return CFUN_OK;
  CFUN_CHECK_ARG(x && w && y);
  cudaStream_t st = as_stream(stream);
  fc_init_kernel<<<(unsigned)cdiv((long long)M * Nout, 256), 256, 0, st>>>(y, bias, M, Nout);
  CFUN_LAUNCH_CHECK();
  long long slabs = std::min<long long>(cdiv(K, 32), 6LL * num_sms());
  int kslab = (int)align_up((size_t)cdiv(K, slabs), 32);
  dim3 grid((unsigned)cdiv(K, kslab), (unsigned)cdiv(Nout, 128));
  // any number of rows (the reference has no limit: TRAIN_ROIS_PER_IMAGE is 200 in the base Config, POST_NMS_ROIS_INFERENCE
  // 1000): chunks of <= 128 rows, each one pass over the weights
  for (int m0 = 0; m0 < M; m0 += 128) {
    const int Mc = std::min(128, M - m0);
    const float* xc = x + (long long)m0 * K;
    float* yc = y + (long long)m0 * Nout;
    if (Mc <= 16) fc_fwd_kernel<1><<<grid, 256, 0, st>>>(Mc, Nout, K, xc, w, yc, kslab);
    else if (Mc <= 32) fc_fwd_kernel<2><<<grid, 256, 0, st>>>(Mc, Nout, K, xc, w, yc, kslab);
    else if (Mc <= 64) fc_fwd_kernel<4><<<grid, 256, 0, st>>>(Mc, Nout, K, xc, w, yc, kslab);
    else fc_fwd_kernel<8><<<grid, 256, 0, st>>>(Mc, Nout, K, xc, w, yc, kslab);
    CFUN_LAUNCH_CHECK();
  }
  return CFUN_OK;
}

extern "C" int cfun_fc_bwd_data(int M, int Nout, long long K, const float* dy, const float* w, float* dx, void* stream) {
  CFUN_CHECK_ARG(M >= 0 && Nout > 0 && K > 0);
  if (M == 0) return CFUN_OK;
  CFUN_CHECK_ARG(dy && w && dx && Nout <= 512);
  for (int m0 = 0; m0 < M; m0 += 16) {      // row chunks of 16 (the kernel's register tile); rows are independent
    fc_bwd_data_kernel<<<(unsigned)std::min<long long>(cdiv(K, 256), 8LL * num_sms()), 256, Nout * 16 * sizeof(float), as_stream(stream)>>>(
        std::min(16, M - m0), Nout, K, dy + (long long)m0 * Nout, w, dx + (long long)m0 * K);
    CFUN_LAUNCH_CHECK();
  }
  return CFUN_OK;
}

extern "C" int cfun_fc_bwd_weight(int M, int Nout, long long K, const float* dy, const float* x, float* dw, float* dbias,
                                  void* stream) {
  CFUN_CHECK_ARG(M >= 0 && Nout > 0 && K > 0 && dw);
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Nout * K, st));
    if (dbias) CFUN_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Nout, st));
    return CFUN_OK;
  }
  CFUN_CHECK_ARG(dy && x && Nout <= 512);
  for (int m0 = 0; m0 < M; m0 += 16) {      // row chunks of 16, accumulated into dw
    fc_bwd_weight_kernel<<<(unsigned)std::min<long long>(cdiv(K, 256), 8LL * num_sms()), 256, Nout * 16 * sizeof(float), st>>>(
        std::min(16, M - m0), Nout, K, dy + (long long)m0 * Nout, x + (long long)m0 * K, dw, m0 > 0 ? 1 : 0);
    CFUN_LAUNCH_CHECK();
  }
  if (dbias) {
    fc_dbias_kernel<<<(unsigned)cdiv(Nout, 128), 128, 0, st>>>(M, Nout, dy, dbias);
    CFUN_LAUNCH_CHECK();
  }
  return CFUN_OK;
}
