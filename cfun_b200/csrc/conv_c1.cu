// Direct CUDA-core convolution for single-channel inputs: the P3D stem (backbone.py:124: Conv3d 1->16, k (3,7,7), stride 2,
// pad (1,3,3); LiTS: 1->24, k (5,7,7)) and the first conv of the mask U-Net (mask_branch.py:23: Conv3d 1->20, 3^3, pad 1) --
// forward and weight gradient (a one-channel image gets no data gradient).
//
// With Cin = 1 the contraction is K = taps (147 / 27): no tensor-core shape, and the generic implicit GEMM gathers one
// strided scalar per (voxel, tap) from L2 (stem forward 1.75 ms for 9.9 GFLOP).  Here a block stages the input tile
// ((TD-1)s+kD) x ((TH-1)s+kH) x ((TW-1)s+kW) floats once in shared memory; in the forward each thread keeps 4 adjacent
// output voxels x all Cout accumulators in registers and walks the taps with one row of ((4-1)s+kW) inputs per (kd,kh);
// in the weight gradient a thread owns (kd, kh, 4 output channels) x all kw (<= 28 accumulators), blocks are persistent,
// and the partial sums are flushed once per block with fp32 atomics.  Arithmetic is plain fp32 FMA (exact-mode path).
#include "common.cuh"

namespace cfun {

constexpr int C1_TD = 4, C1_TH = 4, C1_TW = 32, C1_VW = 4;
constexpr int C1_FWD_THREADS = C1_TD * C1_TH * (C1_TW / C1_VW);      // 128

struct C1Params {
  int N, D, H, W;            // input extents
  int Do, Ho, Wo;            // output extents
  int pD, pH, pW;
  int tilesD, tilesH, tilesW;
  long long ntiles;
  int epi;
};

template <int KD, int KH, int KW, int S>
struct C1Geo {
  static constexpr int ID = (C1_TD - 1) * S + KD, IH = (C1_TH - 1) * S + KH, IW = (C1_TW - 1) * S + KW;
  static constexpr int IWP = IW + 1;                     // +1: rows of consecutive h do not start on the same bank
  static constexpr int XS = ID * IH * IWP;
  static constexpr int ROW = (C1_VW - 1) * S + KW;       // inputs one thread needs from a tile row
};

template <int KD, int KH, int KW, int S>
__device__ __forceinline__ void c1_load_tile(const float* __restrict__ x, float* xs, const C1Params& p, int n, int d0, int h0, int w0,
                                             int nthreads) {
  using G = C1Geo<KD, KH, KW, S>;
  const int id0 = d0 * S - p.pD, ih0 = h0 * S - p.pH, iw0 = w0 * S - p.pW;
  for (int i = threadIdx.x; i < G::ID * G::IH * G::IW; i += nthreads) {
    const int iw = i % G::IW, r = i / G::IW, ih = r % G::IH, id = r / G::IH;
    const int gd = id0 + id, gh = ih0 + ih, gw = iw0 + iw;
    float v = 0.f;
    if ((unsigned)gd < (unsigned)p.D && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W)
      v = __ldg(x + (((long long)n * p.D + gd) * p.H + gh) * p.W + gw);
    xs[(id * G::IH + ih) * G::IWP + iw] = v;
  }
}

template <int CO, int KD, int KH, int KW, int S>
__global__ void __launch_bounds__(C1_FWD_THREADS) c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y, const C1Params p) {
  using G = C1Geo<KD, KH, KW, S>;
  constexpr int TAPS = KD * KH * KW;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                   // input tile
  float* wsm = smem + ((G::XS + 3) & ~3);             // [tap][CO]
  long long t = blockIdx.x;
  const int tw = (int)(t % p.tilesW); t /= p.tilesW;
  const int th = (int)(t % p.tilesH); t /= p.tilesH;
  const int td = (int)(t % p.tilesD);
  const int n = (int)(t / p.tilesD);
  const int d0 = td * C1_TD, h0 = th * C1_TH, w0 = tw * C1_TW;
  for (int i = threadIdx.x; i < TAPS * CO; i += C1_FWD_THREADS) {
    const int tap = i / CO, c = i - tap * CO;
    wsm[i] = __ldg(w + (long long)c * TAPS + tap);     // (Cout, 1, taps) -> [tap][co]
  }
  c1_load_tile<KD, KH, KW, S>(x, xs, p, n, d0, h0, w0, C1_FWD_THREADS);
  __syncthreads();
  const int lw = (threadIdx.x % (C1_TW / C1_VW)) * C1_VW;
  const int lh = (threadIdx.x / (C1_TW / C1_VW)) % C1_TH;
  const int ld = threadIdx.x / ((C1_TW / C1_VW) * C1_TH);
  float acc[C1_VW][CO];
#pragma unroll
  for (int v = 0; v < C1_VW; ++v)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[v][c] = 0.f;
#pragma unroll 1
  for (int kd = 0; kd < KD; ++kd) {
#pragma unroll 1
    for (int kh = 0; kh < KH; ++kh) {
      const float* row = xs + ((ld * S + kd) * G::IH + (lh * S + kh)) * G::IWP + lw * S;
      float r[G::ROW];
#pragma unroll
      for (int i = 0; i < G::ROW; ++i) r[i] = row[i];
      const float4* wr = reinterpret_cast<const float4*>(wsm + (kd * KH + kh) * KW * CO);
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) {
          const float4 wv = wr[kw * (CO / 4) + c4];
#pragma unroll
          for (int v = 0; v < C1_VW; ++v) {
            const float xv = r[v * S + kw];
            acc[v][c4 * 4 + 0] = fmaf(xv, wv.x, acc[v][c4 * 4 + 0]);
            acc[v][c4 * 4 + 1] = fmaf(xv, wv.y, acc[v][c4 * 4 + 1]);
            acc[v][c4 * 4 + 2] = fmaf(xv, wv.z, acc[v][c4 * 4 + 2]);
            acc[v][c4 * 4 + 3] = fmaf(xv, wv.w, acc[v][c4 * 4 + 3]);
          }
        }
      }
    }
  }
  const int od = d0 + ld, oh = h0 + lh;
  if (od >= p.Do || oh >= p.Ho) return;
#pragma unroll
  for (int v = 0; v < C1_VW; ++v) {
    const int ow = w0 + lw + v;
    if (ow >= p.Wo) break;
    float4* yo = reinterpret_cast<float4*>(y + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * CO);
#pragma unroll
    for (int c4 = 0; c4 < CO / 4; ++c4) {
      float4 o = make_float4(acc[v][c4 * 4], acc[v][c4 * 4 + 1], acc[v][c4 * 4 + 2], acc[v][c4 * 4 + 3]);
      if (p.epi & CFUN_EPI_BIAS) {
        o.x += __ldg(bias + c4 * 4); o.y += __ldg(bias + c4 * 4 + 1); o.z += __ldg(bias + c4 * 4 + 2); o.w += __ldg(bias + c4 * 4 + 3);
      }
      if (p.epi & CFUN_EPI_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      yo[c4] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: dW[co][kd][kh][kw] = sum_v dY[v][co] * X[v*s + k - p]
// ---------------------------------------------------------------------------------------------------------
constexpr int C1_WG_GROUPS = 4;      // voxel groups per block (each group runs the whole (kd,kh,co/4) work-item set)

template <int CO, int KD, int KH, int KW, int S>
struct C1Wg {
  static constexpr int WI = KD * KH * (CO / 4);               // work items: (kd, kh, 4 output channels)
  static constexpr int TPG = (WI + 31) / 32 * 32;             // threads per voxel group
  static constexpr int THREADS = TPG * C1_WG_GROUPS;
};

template <int CO, int KD, int KH, int KW, int S>
__global__ void __launch_bounds__(C1Wg<CO, KD, KH, KW, S>::THREADS)
c1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, const C1Params p) {
  using G = C1Geo<KD, KH, KW, S>;
  using WGp = C1Wg<CO, KD, KH, KW, S>;
  constexpr int TAPS = KD * KH * KW;
  constexpr int TILE = C1_TD * C1_TH * C1_TW;           // 512 output voxels
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* dys = smem + ((G::XS + 3) & ~3);               // [voxel][CO]
  const int grp = threadIdx.x / WGp::TPG, wi = threadIdx.x % WGp::TPG;
  const bool active = wi < WGp::WI;
  const int c4 = wi % (CO / 4), khd = wi / (CO / 4), kh = khd % KH, kd = khd / KH;
  float acc[KW][4];
#pragma unroll
  for (int k = 0; k < KW; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }

  for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    long long t = tile;
    const int tw = (int)(t % p.tilesW); t /= p.tilesW;
    const int th = (int)(t % p.tilesH); t /= p.tilesH;
    const int td = (int)(t % p.tilesD);
    const int n = (int)(t / p.tilesD);
    const int d0 = td * C1_TD, h0 = th * C1_TH, w0 = tw * C1_TW;
    __syncthreads();                                     // previous tile fully consumed
    c1_load_tile<KD, KH, KW, S>(x, xs, p, n, d0, h0, w0, WGp::THREADS);
    for (int i = threadIdx.x; i < TILE * (CO / 4); i += WGp::THREADS) {
      const int q = i % (CO / 4), v = i / (CO / 4);
      const int lw = v % C1_TW, lh = (v / C1_TW) % C1_TH, ld = v / (C1_TW * C1_TH);
      const int od = d0 + ld, oh = h0 + lh, ow = w0 + lw;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (od < p.Do && oh < p.Ho && ow < p.Wo)
        g = __ldg(reinterpret_cast<const float4*>(dy + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * CO) + q);
      reinterpret_cast<float4*>(dys)[i] = g;
    }
    __syncthreads();
    if (active) {
      // this group's share of the tile: strips of C1_VW voxels along w, strip index s = grp, grp + GROUPS, ...
      constexpr int STRIPS = TILE / C1_VW;
#pragma unroll 1
      for (int s = grp; s < STRIPS; s += C1_WG_GROUPS) {
        const int lw = (s % (C1_TW / C1_VW)) * C1_VW, lh = (s / (C1_TW / C1_VW)) % C1_TH, ld = s / ((C1_TW / C1_VW) * C1_TH);
        const float* row = xs + ((ld * S + kd) * G::IH + (lh * S + kh)) * G::IWP + lw * S;
        float r[G::ROW];
#pragma unroll
        for (int i = 0; i < G::ROW; ++i) r[i] = row[i];
        const int v0 = (ld * C1_TH + lh) * C1_TW + lw;
#pragma unroll
        for (int v = 0; v < C1_VW; ++v) {
          const float4 g = reinterpret_cast<const float4*>(dys)[(v0 + v) * (CO / 4) + c4];
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            const float xv = r[v * S + k];
            acc[k][0] = fmaf(xv, g.x, acc[k][0]);
            acc[k][1] = fmaf(xv, g.y, acc[k][1]);
            acc[k][2] = fmaf(xv, g.z, acc[k][2]);
            acc[k][3] = fmaf(xv, g.w, acc[k][3]);
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < KW; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dw + (long long)(c4 * 4 + j) * TAPS + (kd * KH + kh) * KW + k, acc[k][j]);
  }
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

static bool c1_fill(const cfun_conv3d_desc* d, C1Params& p) {
  if (!d || d->Cin != 1) return false;
  p.N = d->N; p.D = d->Din; p.H = d->Hin; p.W = d->Win;
  p.Do = d->Dout; p.Ho = d->Hout; p.Wo = d->Wout;
  p.pD = d->pD; p.pH = d->pH; p.pW = d->pW;
  p.tilesD = (int)cdiv(p.Do, C1_TD); p.tilesH = (int)cdiv(p.Ho, C1_TH); p.tilesW = (int)cdiv(p.Wo, C1_TW);
  p.ntiles = (long long)p.N * p.tilesD * p.tilesH * p.tilesW;
  p.epi = 0;
  return p.ntiles > 0 && p.ntiles < 0x7fffffffLL;
}

// the instantiated (Cout, kernel, stride) combinations: heart stem, LiTS stem, U-Net first conv (20 / 32 base filters)
#define C1_FOR_EACH(X) X(16, 3, 7, 7, 2) X(24, 5, 7, 7, 2) X(20, 3, 3, 3, 1) X(32, 3, 3, 3, 1)

static bool c1_match(const cfun_conv3d_desc* d, int CO, int KD, int KH, int KW, int S) {
  return d->Cout == CO && d->kD == KD && d->kH == KH && d->kW == KW && d->sD == S && d->sH == S && d->sW == S;
}

bool c1_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_CONV_C1");         // "0": generic implicit GEMM (A/B measurements)
  if (e && e[0] == '0') return false;
  if (!d || d->Cin != 1 || pass == CFUN_PASS_BWD_DATA) return false;
#define X(CO, KD, KH, KW, S) if (c1_match(d, CO, KD, KH, KW, S)) return true;
  C1_FOR_EACH(X)
#undef X
  return false;
}

template <int CO, int KD, int KH, int KW, int S>
static int c1_run_fwd(const C1Params& p, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  using G = C1Geo<KD, KH, KW, S>;
  const size_t smem = (size_t)(((G::XS + 3) & ~3) + KD * KH * KW * CO) * sizeof(float);
  static bool attr = false;
  if (!attr) { CFUN_CUDA(cudaFuncSetAttribute(c1_fwd_kernel<CO, KD, KH, KW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  c1_fwd_kernel<CO, KD, KH, KW, S><<<(unsigned)p.ntiles, C1_FWD_THREADS, smem, st>>>(x, w, bias, y, p);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

template <int CO, int KD, int KH, int KW, int S>
static int c1_run_wgrad(const C1Params& p, const float* x, const float* dy, float* dw, cudaStream_t st) {
  using G = C1Geo<KD, KH, KW, S>;
  using WGp = C1Wg<CO, KD, KH, KW, S>;
  const size_t smem = (size_t)(((G::XS + 3) & ~3) + C1_TD * C1_TH * C1_TW * CO) * sizeof(float);
  static bool attr = false;
  if (!attr) { CFUN_CUDA(cudaFuncSetAttribute(c1_wgrad_kernel<CO, KD, KH, KW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)CO * KD * KH * KW, st));
  const unsigned grid = (unsigned)std::min<long long>(p.ntiles, 2LL * num_sms());
  c1_wgrad_kernel<CO, KD, KH, KW, S><<<grid, WGp::THREADS, smem, st>>>(x, dy, dw, p);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int c1_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, cudaStream_t st) {
  C1Params p;
  CFUN_CHECK_ARG(c1_fill(d, p) && x && w && y);
  CFUN_CHECK_ARG(!(epi & CFUN_EPI_BIAS) || bias);
  p.epi = epi;
#define X(CO, KD, KH, KW, S) if (c1_match(d, CO, KD, KH, KW, S)) return c1_run_fwd<CO, KD, KH, KW, S>(p, x, w, bias, y, st);
  C1_FOR_EACH(X)
#undef X
  set_error("c1_conv_fwd: no instantiation for this shape");
  return CFUN_ERR_INVALID;
}

int c1_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st) {
  C1Params p;
  CFUN_CHECK_ARG(c1_fill(d, p) && x && dy && dw);
  int rc = CFUN_ERR_INVALID;
#define X(CO, KD, KH, KW, S) if (c1_match(d, CO, KD, KH, KW, S)) rc = c1_run_wgrad<CO, KD, KH, KW, S>(p, x, dy, dw, st);
  C1_FOR_EACH(X)
#undef X
  if (rc != CFUN_OK) return rc;
  if (dbias) return simt_bias_grad(dy, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}

}  // namespace cfun
