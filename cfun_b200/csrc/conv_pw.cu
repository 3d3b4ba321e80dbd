// Pointwise (1x1x1, stride 1) convolutions with at most 8 output channels: the segmentation / deep-supervision heads of
// the mask U-Net (mask_branch.py:72,86-87: conv3d_l4 40->8 @ 96^3, ds2_1x1_conv3d, ds3_1x1_conv3d) and the RPN class /
// bbox heads (model.py:714-715).  They move 0.7 GB per pass at 96^3 and do 0.3 GFLOP: HBM-bound streaming kernels, which
// the generic implicit GEMM ran 3-6x off the bandwidth roofline (its tiles assume Cout >= 16).
//   forward        y[v][co]  = sum_ci x[v][ci] w[co][ci] (+ bias, ReLU)   one thread per voxel pair, weights in shared memory
//   data gradient  dx[v][ci] = sum_co dy[v][co] w[co][ci]                 same
//   weight gradient dw[co][ci] = sum_v dy[v][co] x[v][ci]                 thread = (ci, co), persistent blocks, atomic flush
//                                                                         (opt-in only: slower than the generic kernel)
// Arithmetic is plain fp32 FMA (the exact-mode CUDA-core path).
#include "common.cuh"

namespace cfun {

constexpr int PW_MAX_CIN = 320;
constexpr int PW_THREADS = 256;

// weights staged as wsm[ci][8] (output channels beyond Cout are zero)
__global__ void __launch_bounds__(PW_THREADS) pw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, long long M, int Cin,
                                                            int Cout, int epi) {
  __shared__ __align__(16) float wsm[PW_MAX_CIN * 8];
  for (int i = threadIdx.x; i < Cin * 8; i += PW_THREADS) {
    const int ci = i >> 3, co = i & 7;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int c4n = Cin >> 2;
  for (long long v = (long long)blockIdx.x * PW_THREADS + threadIdx.x; v < M; v += (long long)gridDim.x * PW_THREADS) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(x + v * Cin);
#pragma unroll 2
    for (int c4 = 0; c4 < c4n; ++c4) {
      const float4 xv = __ldg(xr + c4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8 + 4);
        acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
        acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
        acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
        acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((epi & CFUN_EPI_BIAS) && j < Cout) acc[j] += __ldg(bias + j);
      if (epi & CFUN_EPI_RELU) acc[j] = fmaxf(acc[j], 0.f);
    }
    float* yo = y + v * Cout;
    if (Cout == 8) {
      reinterpret_cast<float4*>(yo)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(yo)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < Cout) yo[j] = acc[j];
    }
  }
}

// weights staged as wsm[co][Cin] (rows beyond Cout are zero)
__global__ void __launch_bounds__(PW_THREADS) pw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, long long M, int Cin, int Cout) {
  __shared__ __align__(16) float wsm[8 * PW_MAX_CIN];
  for (int i = threadIdx.x; i < 8 * Cin; i += PW_THREADS) {
    const int co = i / Cin, ci = i - co * Cin;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int c4n = Cin >> 2;
  for (long long v = (long long)blockIdx.x * PW_THREADS + threadIdx.x; v < M; v += (long long)gridDim.x * PW_THREADS) {
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = j < Cout ? __ldg(dy + v * Cout + j) : 0.f;
    float4* xo = reinterpret_cast<float4*>(dx + v * Cin);
#pragma unroll 2
    for (int c4 = 0; c4 < c4n; ++c4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(wsm + j * Cin + c4 * 4);
        o.x = fmaf(g[j], wv.x, o.x); o.y = fmaf(g[j], wv.y, o.y); o.z = fmaf(g[j], wv.z, o.z); o.w = fmaf(g[j], wv.w, o.w);
      }
      xo[c4] = o;
    }
  }
}

// thread = (ci, co) with ci fastest: x[v][ci] loads are coalesced over ci, dy[v][co] is a broadcast within each co group
__global__ void __launch_bounds__(1024) pw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                        long long M, int Cin, int Cout, long long rows_per_block) {
  const int ci = threadIdx.x % Cin, co = threadIdx.x / Cin;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long long v = r0;
  for (; v + 3 < r1; v += 4) {
    const float x0 = __ldg(x + v * Cin + ci), x1 = __ldg(x + (v + 1) * Cin + ci), x2 = __ldg(x + (v + 2) * Cin + ci),
                x3 = __ldg(x + (v + 3) * Cin + ci);
    const float g0 = __ldg(dy + v * Cout + co), g1 = __ldg(dy + (v + 1) * Cout + co), g2 = __ldg(dy + (v + 2) * Cout + co),
                g3 = __ldg(dy + (v + 3) * Cout + co);
    a0 = fmaf(x0, g0, a0); a1 = fmaf(x1, g1, a1); a2 = fmaf(x2, g2, a2); a3 = fmaf(x3, g3, a3);
  }
  for (; v < r1; ++v) a0 = fmaf(__ldg(x + v * Cin + ci), __ldg(dy + v * Cout + co), a0);
  atomicAdd(dw + (long long)co * Cin + ci, (a0 + a1) + (a2 + a3));
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

bool pw_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_CONV_PW");          // "0": generic implicit GEMM (A/B measurements)
  if (e && e[0] == '0') return false;
  if (!d || d->kD != 1 || d->kH != 1 || d->kW != 1 || d->sD != 1 || d->sH != 1 || d->sW != 1 || d->pD || d->pH || d->pW) return false;
  if (d->Cout > 8 || d->Cout < 1 || (d->Cin & 3) || d->Cin > PW_MAX_CIN || d->Cin < 4) return false;
  // measured on B200 (tools/conv_cases.py pw): the forward wins everywhere (40->8 @ 4x96^3: 0.34 -> 0.18 ms), the data
  // gradient only for narrow inputs (0.46 -> 0.27 ms at 40 channels, slower from 80 up), the weight gradient nowhere
  // (0.69 -> 1.0 ms) -- it stays on the generic kernel unless CFUN_CONV_PW=w asks for it
  if (pass == CFUN_PASS_BWD_DATA && d->Cin > 48) return false;
  if (pass == CFUN_PASS_BWD_WEIGHT && (!(e && e[0] == 'w') || d->Cin * d->Cout > 1024)) return false;
  return true;
}

static inline long long pw_rows(const cfun_conv3d_desc* d) { return (long long)d->N * d->Dout * d->Hout * d->Wout; }

int pw_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_FWD) && x && w && y);
  CFUN_CHECK_ARG(!(epi & CFUN_EPI_BIAS) || bias);
  const long long M = pw_rows(d);
  const unsigned grid = (unsigned)std::min<long long>(cdiv(M, PW_THREADS), 16LL * num_sms());
  pw_fwd_kernel<<<grid, PW_THREADS, 0, st>>>(x, w, bias, y, M, d->Cin, d->Cout, epi);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int pw_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_BWD_DATA) && dy && w && dx);
  const long long M = pw_rows(d);
  const unsigned grid = (unsigned)std::min<long long>(cdiv(M, PW_THREADS), 16LL * num_sms());
  pw_dgrad_kernel<<<grid, PW_THREADS, 0, st>>>(dy, w, dx, M, d->Cin, d->Cout);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int pw_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_BWD_WEIGHT) && x && dy && dw);
  const long long M = pw_rows(d);
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin, st));
  const long long blocks = std::max<long long>(1, std::min<long long>(cdiv(M, 256), 4LL * num_sms()));
  const long long rpb = cdiv(M, blocks);
  pw_wgrad_kernel<<<(unsigned)cdiv(M, rpb), d->Cin * d->Cout, 0, st>>>(x, dy, dw, M, d->Cin, d->Cout, rpb);
  CFUN_LAUNCH_CHECK();
  if (dbias) return simt_bias_grad(dy, M, d->Cout, dbias, st);
  return CFUN_OK;
}

}  // namespace cfun
