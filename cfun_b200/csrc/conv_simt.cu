// fp32 CUDA-core implicit-GEMM 3-D convolution (forward, data gradient, weight gradient) for NDHWC tensors.
//
// This is the any-shape path (Cin = 1 stem, asymmetric P3D kernels, strided convs, 5^3 / 1^3 kernels, tiny maps);
// the dense stride-1 convs that carry the FLOPs are routed to the tcgen05 kernels in conv_tc.cu.
//
//   forward   y[m, co]  = sum_kk  A[m, kk] * Bf[kk, co]      kk = tap * Cin  + ci   (A = implicit im2col of x)
//   dgrad     dx[m, ci] = sum_kk  A'[m, kk] * Bd[kk, ci]     kk = tap * Cout + co   (A' gathers dy)
//   wgrad     dW[kk, co] = sum_m  A[m, kk] * dy[m, co]       (split over m, fp32 atomics into a packed scratch)
//
// Replaces nn.Conv3d forward/backward everywhere on the CFUN path (backbone.py:14-55,124; model.py:131-134,713-717,
// 760; mask_branch.py:23-89).
#include "common.cuh"

namespace cfun {

struct ConvGeo {
  int N, Cs, Ds, Hs, Ws;  // source tensor (gathered operand)
  int Ct, Dt, Ht, Wt;     // target tensor (one GEMM row per target voxel)
  int kD, kH, kW, sD, sH, sW, pD, pH, pW;
  int Ktot;  // taps * Cs
  int cls;   // 1: rows are ordered by stride-parity class (strided data gradient), see row_to_voxel()
  long long Mc;  // rows per parity class
  int ldb;   // leading dimension of the packed weight matrix
  long long M;
};

template <int MODE>
__device__ __forceinline__ bool src_coord(int t, int k, int s, int p, int size_src, int& out) {
  if (MODE == 0) {
    out = t * s - p + k;
    return (unsigned)out < (unsigned)size_src;
  } else {
    int q = t + p - k;
    if (q < 0) return false;
    if (s != 1) {
      if (q % s) return false;
      q /= s;
    }
    out = q;
    return q < size_src;
  }
}

// GEMM row -> target voxel.  For a strided data gradient the rows are grouped by the parity class of the input voxel
// ((td+pD)%s, (th+pH)%s, (tw+pW)%s with s = 2): all rows of a class share the same set of contributing taps (1/8 of the
// 27 on average), so a block whose rows sit in one class can skip the other taps' K-chunks altogether.
__device__ __forceinline__ void row_to_voxel(const ConvGeo& g, long long m, int& n, int& td, int& th, int& tw) {
  if (!g.cls) {
    tw = (int)(m % g.Wt); m /= g.Wt;
    th = (int)(m % g.Ht); m /= g.Ht;
    td = (int)(m % g.Dt); m /= g.Dt;
    n = (int)m;
  } else {
    const int c = (int)(m / g.Mc);
    long long r = m % g.Mc;
    const int Wq = g.Wt >> 1, Hq = g.Ht >> 1, Dq = g.Dt >> 1;
    tw = 2 * (int)(r % Wq) + (c & 1); r /= Wq;
    th = 2 * (int)(r % Hq) + ((c >> 1) & 1); r /= Hq;
    td = 2 * (int)(r % Dq) + ((c >> 2) & 1); r /= Dq;
    n = (int)r;
  }
}
// does K-chunk kt (16 consecutive kk) touch a tap that can contribute to parity class c?
__device__ __forceinline__ bool chunk_live(const ConvGeo& g, int kt, int c) {
  const int t0 = (kt * 16) / g.Cs, t1 = min((kt * 16 + 15) / g.Cs, g.kD * g.kH * g.kW - 1);
  const int khw = g.kH * g.kW;
  for (int t = t0; t <= t1; ++t) {
    const int kd = t / khw, r = t - kd * khw, kh = r / g.kW, kw = r - kh * g.kW;
    if ((((c >> 2) & 1) + g.pD - kd) % 2 == 0 && (((c >> 1) & 1) + g.pH - kh) % 2 == 0 && ((c & 1) + g.pW - kw) % 2 == 0) return true;
  }
  return false;
}

// address (in floats) of source element for target voxel (n,td,th,tw) and reduction index kk; -1 when it is padding
template <int MODE>
__device__ __forceinline__ long long gather_addr(const ConvGeo& g, int n, int td, int th, int tw, int kk) {
  int tap = kk / g.Cs;
  int c = kk - tap * g.Cs;
  int khw = g.kH * g.kW;
  int kd = tap / khw;
  int r = tap - kd * khw;
  int kh = r / g.kW;
  int kw = r - kh * g.kW;
  int zd, zh, zw;
  if (!src_coord<MODE>(td, kd, g.sD, g.pD, g.Ds, zd)) return -1;
  if (!src_coord<MODE>(th, kh, g.sH, g.pH, g.Hs, zh)) return -1;
  if (!src_coord<MODE>(tw, kw, g.sW, g.pW, g.Ws, zw)) return -1;
  return ((((long long)n * g.Ds + zd) * g.Hs + zh) * g.Ws + zw) * g.Cs + c;
}

// ---------------------------------------------------------------------------------------------------------
// forward / dgrad implicit GEMM
// ---------------------------------------------------------------------------------------------------------
template <int MODE, int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) igemm_kernel(ConvGeo g, const float* __restrict__ src,
                                                    const float* __restrict__ Bm, const float* __restrict__ bias,
                                                    float* __restrict__ dst, int epi) {
  constexpr int BK = 16;
  constexpr int TX = BN / TN;
  constexpr int TY = BM / TM;
  static_assert(TX * TY == 256, "tile config");
  constexpr int APT = BM * BK / 256;  // A elements per thread (8 or 16)
  constexpr int TPR = BK / APT;       // threads per A row
  constexpr int BPT = BK * BN / 256;  // B elements per thread (4, 2 or 1)
  constexpr int LDA = BM + 4;

  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const long long m0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  // A loader: one row, APT consecutive kk
  const int arow = tid / TPR;
  const int akofs = (tid % TPR) * APT;
  const long long am = m0 + arow;
  const bool arow_ok = am < g.M;
  int an = 0, atd = 0, ath = 0, atw = 0;
  if (arow_ok) row_to_voxel(g, am, an, atd, ath, atw);
  // block-uniform parity class (strided dgrad only): -1 = rows of several classes, no K-chunk skipping
  int bcls = -1;
  if (MODE == 1 && g.cls) {
    const long long mlast = min(g.M - 1, m0 + BM - 1);
    if (m0 / g.Mc == mlast / g.Mc) bcls = (int)(m0 / g.Mc);
  }
  // B loader
  const int bk = (tid * BPT) / BN;
  const int bn = (tid * BPT) % BN;

  float areg[APT];
  float breg[BPT];

  auto load_tile = [&](int kt) {
    const int kk0 = kt * BK + akofs;
    if (VEC) {
#pragma unroll
      for (int j = 0; j < APT; j += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int kk = kk0 + j;
        if (arow_ok && kk < g.Ktot) {
          long long a = gather_addr<MODE>(g, an, atd, ath, atw, kk);
          if (a >= 0) v = __ldg(reinterpret_cast<const float4*>(src + a));
        }
        areg[j] = v.x; areg[j + 1] = v.y; areg[j + 2] = v.z; areg[j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < APT; ++j) {
        float v = 0.f;
        int kk = kk0 + j;
        if (arow_ok && kk < g.Ktot) {
          long long a = gather_addr<MODE>(g, an, atd, ath, atw, kk);
          if (a >= 0) v = __ldg(src + a);
        }
        areg[j] = v;
      }
    }
    const float* bp = Bm + (long long)(kt * BK + bk) * g.ldb + col0 + bn;  // rows/cols are zero padded
    if (BPT == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(bp));
      breg[0] = v.x; breg[1] = v.y; breg[2] = v.z; breg[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < BPT; ++j) breg[j] = __ldg(bp + j);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < APT; ++j) As[buf][akofs + j][arow] = areg[j];
#pragma unroll
    for (int j = 0; j < BPT; ++j) Bs[buf][bk][bn + j] = breg[j];
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (g.Ktot + BK - 1) / BK;
  auto next_live = [&](int k) {
    if (MODE == 1 && bcls >= 0)
      while (k < nk && !chunk_live(g, k, bcls)) ++k;
    return k;
  };
  int kt = next_live(0);
  if (kt < nk) { load_tile(kt); store_tile(0); }
  __syncthreads();
  int buf = 0;
  while (kt < nk) {
    const int kn = next_live(kt + 1);
    if (kn < nk) load_tile(kn);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kn < nk) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    kt = kn;
  }

  const bool vec_out = (g.Ct % 4 == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= g.M) continue;
    if (g.cls) {   // parity-class row order -> real voxel index
      int n, td, th, tw;
      row_to_voxel(g, m, n, td, th, tw);
      m = (((long long)n * g.Dt + td) * g.Ht + th) * g.Wt + tw;
    }
    float* orow = dst + m * g.Ct;
    int c = col0 + tx * TN;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float t = acc[i][j];
      if ((epi & CFUN_EPI_BIAS) && c + j < g.Ct) t += __ldg(bias + c + j);
      if (epi & CFUN_EPI_RELU) t = fmaxf(t, 0.f);
      v[j] = t;
    }
    if (vec_out && c + TN <= g.Ct) {
#pragma unroll
      for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(orow + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (c + j < g.Ct) orow[c + j] = v[j];
    }
  }
}

// packed weight matrices.  w is (Cout, Cin, taps).  The kernels write the whole zero-padded matrix [rows_pad][ldb] (the
// GEMM kernels read the padding), so no memset precedes them.
__global__ void pack_w_fwd_kernel(const float* __restrict__ w, float* __restrict__ Bm, int Cout, int Cin, int taps, int ldb,
                                  int rows_pad) {
  const int total = rows_pad * ldb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / ldb, co = i - r * ldb;
    const int tap = r / Cin, ci = r - tap * Cin;
    Bm[i] = (tap < taps && co < Cout) ? __ldg(w + ((long long)co * Cin + ci) * taps + tap) : 0.f;
  }
}
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, float* __restrict__ Bm, int Cout, int Cin, int taps, int ldb,
                                    int rows_pad) {
  const int total = rows_pad * ldb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / ldb, ci = i - r * ldb;
    const int tap = r / Cout, co = r - tap * Cout;
    Bm[i] = (tap < taps && ci < Cin) ? __ldg(w + ((long long)co * Cin + ci) * taps + tap) : 0.f;
  }
}
__global__ void unpack_dw_kernel(const float* __restrict__ dWm, float* __restrict__ dw, int Cout, int Cin, int taps, int ldb) {
  long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    long long r = i / taps;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    dw[i] = dWm[((long long)tap * Cin + ci) * ldb + co];
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: 64 (kk) x BNC (co) tile per block, reduction over a slice of the M rows
// ---------------------------------------------------------------------------------------------------------
template <int BNC, bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(ConvGeo g, const float* __restrict__ x, const float* __restrict__ dy,
                                                    float* __restrict__ dWm, long long rows_per_split) {
  constexpr int BKK = 64, BR = 16;
  constexpr int TN = BNC / 16;
  __shared__ __align__(16) float As[2][BR][BKK];
  __shared__ __align__(16) float Ds[2][BR][BNC];

  const int tid = threadIdx.x;
  const int tkk = tid % 16, tcol = tid / 16;
  const int kk0 = blockIdx.x * BKK;
  const int col0 = blockIdx.y * BNC;
  const long long mbeg = (long long)blockIdx.z * rows_per_split;
  long long mend = mbeg + rows_per_split;
  if (mend > g.M) mend = g.M;

  // A loader: row = tid / 16, 4 consecutive kk starting at kk0 + (tid % 16) * 4 (fixed for the whole loop)
  const int arow = tid / 16;
  const int akk = kk0 + (tid % 16) * 4;
  int a_c[4], a_kd[4], a_kh[4], a_kw[4];
  bool a_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int kk = akk + j;
    a_ok[j] = kk < g.Ktot;
    int tap = a_ok[j] ? kk / g.Cs : 0;
    a_c[j] = a_ok[j] ? kk - tap * g.Cs : 0;
    int khw = g.kH * g.kW;
    a_kd[j] = tap / khw;
    int r = tap - a_kd[j] * khw;
    a_kh[j] = r / g.kW;
    a_kw[j] = r - a_kh[j] * g.kW;
  }
  constexpr int DPT = (BR * BNC + 255) / 256;  // dy elements per thread

  float areg[4];
  float dreg[DPT];
  const int Cout = g.Ct;

  auto load_tile = [&](long long mb) {
    long long m = mb + arow;
    areg[0] = areg[1] = areg[2] = areg[3] = 0.f;
    if (m < mend) {
      long long t = m;
      int tw = (int)(t % g.Wt); t /= g.Wt;
      int th = (int)(t % g.Ht); t /= g.Ht;
      int td = (int)(t % g.Dt); t /= g.Dt;
      int n = (int)t;
      if (VEC) {
        if (a_ok[0]) {
          int zd = td * g.sD - g.pD + a_kd[0], zh = th * g.sH - g.pH + a_kh[0], zw = tw * g.sW - g.pW + a_kw[0];
          if ((unsigned)zd < (unsigned)g.Ds && (unsigned)zh < (unsigned)g.Hs && (unsigned)zw < (unsigned)g.Ws) {
            float4 v = __ldg(reinterpret_cast<const float4*>(x + ((((long long)n * g.Ds + zd) * g.Hs + zh) * g.Ws + zw) * g.Cs + a_c[0]));
            areg[0] = v.x; areg[1] = v.y; areg[2] = v.z; areg[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!a_ok[j]) continue;
          int zd = td * g.sD - g.pD + a_kd[j], zh = th * g.sH - g.pH + a_kh[j], zw = tw * g.sW - g.pW + a_kw[j];
          if ((unsigned)zd < (unsigned)g.Ds && (unsigned)zh < (unsigned)g.Hs && (unsigned)zw < (unsigned)g.Ws)
            areg[j] = __ldg(x + ((((long long)n * g.Ds + zd) * g.Hs + zh) * g.Ws + zw) * g.Cs + a_c[j]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < DPT; ++e) {
      int idx = tid + e * 256;
      float v = 0.f;
      if (idx < BR * BNC) {
        int r = idx / BNC, c = idx % BNC;
        long long mm = mb + r;
        if (mm < mend && col0 + c < Cout) v = __ldg(dy + mm * Cout + col0 + c);
      }
      dreg[e] = v;
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][arow][(tid % 16) * 4]) = make_float4(areg[0], areg[1], areg[2], areg[3]);
#pragma unroll
    for (int e = 0; e < DPT; ++e) {
      int idx = tid + e * 256;
      if (idx < BR * BNC) Ds[buf][idx / BNC][idx % BNC] = dreg[e];
    }
  };

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const long long nsteps = (mend - mbeg + BR - 1) / BR;
  if (nsteps <= 0) return;
  load_tile(mbeg);
  store_tile(0);
  __syncthreads();
  for (long long s = 0; s < nsteps; ++s) {
    const int buf = (int)(s & 1);
    if (s + 1 < nsteps) load_tile(mbeg + (s + 1) * BR);
#pragma unroll
    for (int k = 0; k < BR; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][tkk * 4]);
      float a[4] = {av.x, av.y, av.z, av.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ds[buf][k][tcol * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (s + 1 < nsteps) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = kk0 + tkk * 4 + i;
    if (kk >= g.Ktot) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = col0 + tcol * TN + j;
      if (c < Cout) atomicAdd(dWm + (long long)kk * g.ldb + c, acc[i][j]);
    }
  }
}

// column sums of a [M, C] matrix (bias gradient)
__global__ void colsum_kernel(const float* __restrict__ dy, long long M, int C, float* __restrict__ out, long long rows_per_block) {
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(dy + r * C + c);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

static bool make_geo(const cfun_conv3d_desc* d, int pass, ConvGeo& g) {
  if (!d) return false;
  if (d->N <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->kD <= 0 || d->kH <= 0 || d->kW <= 0) return false;
  if (d->sD <= 0 || d->sH <= 0 || d->sW <= 0 || d->pD < 0 || d->pH < 0 || d->pW < 0) return false;
  if (d->Dout != (d->Din + 2 * d->pD - d->kD) / d->sD + 1) return false;
  if (d->Hout != (d->Hin + 2 * d->pH - d->kH) / d->sH + 1) return false;
  if (d->Wout != (d->Win + 2 * d->pW - d->kW) / d->sW + 1) return false;
  if (d->Dout <= 0 || d->Hout <= 0 || d->Wout <= 0) return false;
  g.N = d->N;
  g.kD = d->kD; g.kH = d->kH; g.kW = d->kW;
  g.sD = d->sD; g.sH = d->sH; g.sW = d->sW;
  g.pD = d->pD; g.pH = d->pH; g.pW = d->pW;
  int taps = d->kD * d->kH * d->kW;
  if (pass == CFUN_PASS_BWD_DATA) {
    g.Cs = d->Cout; g.Ds = d->Dout; g.Hs = d->Hout; g.Ws = d->Wout;
    g.Ct = d->Cin; g.Dt = d->Din; g.Ht = d->Hin; g.Wt = d->Win;
  } else {
    g.Cs = d->Cin; g.Ds = d->Din; g.Hs = d->Hin; g.Ws = d->Win;
    g.Ct = d->Cout; g.Dt = d->Dout; g.Ht = d->Hout; g.Wt = d->Wout;
  }
  long long kt = (long long)taps * g.Cs;
  if (kt > 0x7fffffffLL) return false;
  g.Ktot = (int)kt;
  g.ldb = (int)align_up((size_t)g.Ct, 64);
  g.M = (long long)g.N * g.Dt * g.Ht * g.Wt;
  g.cls = 0;
  g.Mc = g.M;
  if (pass == CFUN_PASS_BWD_DATA && d->sD == 2 && d->sH == 2 && d->sW == 2 && (g.Dt % 2 == 0) && (g.Ht % 2 == 0) && (g.Wt % 2 == 0)) {
    g.cls = 1;
    g.Mc = g.M / 8;
  }
  return true;
}

size_t simt_workspace(const cfun_conv3d_desc* d, int pass) {
  ConvGeo g;
  if (!make_geo(d, pass, g)) return 0;
  size_t rows = align_up((size_t)g.Ktot, 64);
  return rows * (size_t)g.ldb * sizeof(float) + 256;
}

template <int MODE>
static int launch_igemm(const ConvGeo& g, const float* src, const float* Bm, const float* bias, float* dst, int epi,
                        cudaStream_t st) {
  const bool vec = (g.Cs % 4 == 0);
  if (g.Ct <= 16) {
    dim3 grid((unsigned)cdiv(g.M, 256), (unsigned)cdiv(g.Ct, 16));
    if (vec) igemm_kernel<MODE, 256, 16, 4, 4, true><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
    else igemm_kernel<MODE, 256, 16, 4, 4, false><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
  } else if (g.Ct <= 32) {
    if (cdiv(g.M, 256) < num_sms()) {      // less than one wave of 256-row tiles (16^3 .. 32^3 backbone stages): half-height tiles
      dim3 grid((unsigned)cdiv(g.M, 128), (unsigned)cdiv(g.Ct, 32));
      if (vec) igemm_kernel<MODE, 128, 32, 4, 4, true><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
      else igemm_kernel<MODE, 128, 32, 4, 4, false><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
    } else {
      dim3 grid((unsigned)cdiv(g.M, 256), (unsigned)cdiv(g.Ct, 32));
      if (vec) igemm_kernel<MODE, 256, 32, 8, 4, true><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
      else igemm_kernel<MODE, 256, 32, 8, 4, false><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
    }
  } else if (cdiv(g.M, 128) * cdiv(g.Ct, 64) < num_sms()) {
    dim3 grid((unsigned)cdiv(g.M, 64), (unsigned)cdiv(g.Ct, 64));
    if (vec) igemm_kernel<MODE, 64, 64, 4, 4, true><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
    else igemm_kernel<MODE, 64, 64, 4, 4, false><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
  } else {
    dim3 grid((unsigned)cdiv(g.M, 128), (unsigned)cdiv(g.Ct, 64));
    if (vec) igemm_kernel<MODE, 128, 64, 8, 4, true><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
    else igemm_kernel<MODE, 128, 64, 8, 4, false><<<grid, 256, 0, st>>>(g, src, Bm, bias, dst, epi);
  }
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int simt_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi,
                  void* ws, size_t ws_bytes, cudaStream_t st) {
  ConvGeo g;
  CFUN_CHECK_ARG(make_geo(d, CFUN_PASS_FWD, g));
  CFUN_CHECK_ARG(x && w && y && ws);
  CFUN_CHECK_ARG(!(epi & CFUN_EPI_BIAS) || bias);
  size_t need = simt_workspace(d, CFUN_PASS_FWD);
  if (ws_bytes < need) { set_error("conv3d_fwd: workspace %zu < %zu", ws_bytes, need); return CFUN_ERR_WORKSPACE; }
  float* Bm = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  int taps = g.kD * g.kH * g.kW;
  const int rows_pad = (int)align_up((size_t)g.Ktot, 64);
  CFUN_CHECK_ARG((long long)rows_pad * g.ldb < (1LL << 31));
  pack_w_fwd_kernel<<<(unsigned)std::min<long long>(cdiv((long long)rows_pad * g.ldb, 256), 4096), 256, 0, st>>>(w, Bm, d->Cout, d->Cin, taps, g.ldb, rows_pad);
  CFUN_LAUNCH_CHECK();
  return launch_igemm<0>(g, x, Bm, bias, y, epi, st);
}

int simt_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  ConvGeo g;
  CFUN_CHECK_ARG(make_geo(d, CFUN_PASS_BWD_DATA, g));
  CFUN_CHECK_ARG(dy && w && dx && ws);
  size_t need = simt_workspace(d, CFUN_PASS_BWD_DATA);
  if (ws_bytes < need) { set_error("conv3d_bwd_data: workspace %zu < %zu", ws_bytes, need); return CFUN_ERR_WORKSPACE; }
  float* Bm = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  int taps = g.kD * g.kH * g.kW;
  const int rows_pad = (int)align_up((size_t)g.Ktot, 64);
  CFUN_CHECK_ARG((long long)rows_pad * g.ldb < (1LL << 31));
  pack_w_dgrad_kernel<<<(unsigned)std::min<long long>(cdiv((long long)rows_pad * g.ldb, 256), 4096), 256, 0, st>>>(w, Bm, d->Cout, d->Cin, taps, g.ldb, rows_pad);
  CFUN_LAUNCH_CHECK();
  return launch_igemm<1>(g, dy, Bm, nullptr, dx, 0, st);
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st) {
  CFUN_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, st));
  long long rpb = std::max<long long>(64, cdiv(M, 4 * num_sms()));
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(M, rpb));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(dy, M, C, dbias, rpb);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int simt_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  ConvGeo g;
  CFUN_CHECK_ARG(make_geo(d, CFUN_PASS_BWD_WEIGHT, g));
  CFUN_CHECK_ARG(x && dy && dw && ws);
  size_t need = simt_workspace(d, CFUN_PASS_BWD_WEIGHT);
  if (ws_bytes < need) { set_error("conv3d_bwd_weight: workspace %zu < %zu", ws_bytes, need); return CFUN_ERR_WORKSPACE; }
  float* dWm = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  CFUN_CUDA(cudaMemsetAsync(dWm, 0, align_up((size_t)g.Ktot, 64) * (size_t)g.ldb * sizeof(float), st));
  const int Cout = g.Ct;
  const int bnc = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
  const unsigned gx = (unsigned)cdiv(g.Ktot, 64), gy = (unsigned)cdiv(Cout, bnc);
  // enough splits for ~6 waves, each at least 128 rows (8 tiles of 16: the 16^3 backbone stages would otherwise run on 40 blocks)
  long long want = cdiv(6LL * num_sms(), (long long)gx * gy);
  long long splits = std::max<long long>(1, std::min<long long>(want, cdiv(g.M, 128)));
  splits = std::min<long long>(splits, 65535);
  long long rps = align_up((size_t)cdiv(g.M, splits), 16);
  splits = cdiv(g.M, rps);
  dim3 grid(gx, gy, (unsigned)splits);
  const bool vec = (g.Cs % 4 == 0);
#define LAUNCH_WG(BNC)                                                                         \
  do {                                                                                         \
    if (vec) wgrad_kernel<BNC, true><<<grid, 256, 0, st>>>(g, x, dy, dWm, rps);                \
    else wgrad_kernel<BNC, false><<<grid, 256, 0, st>>>(g, x, dy, dWm, rps);                   \
  } while (0)
  if (bnc == 16) LAUNCH_WG(16);
  else if (bnc == 32) LAUNCH_WG(32);
  else LAUNCH_WG(64);
#undef LAUNCH_WG
  CFUN_LAUNCH_CHECK();
  int taps = g.kD * g.kH * g.kW;
  long long tot = (long long)d->Cout * d->Cin * taps;
  unpack_dw_kernel<<<(unsigned)std::min<long long>(cdiv(tot, 256), 4096), 256, 0, st>>>(dWm, dw, d->Cout, d->Cin, taps, g.ldb);
  CFUN_LAUNCH_CHECK();
  if (dbias) return simt_bias_grad(dy, g.M, Cout, dbias, st);
  return CFUN_OK;
}

}  // namespace cfun
