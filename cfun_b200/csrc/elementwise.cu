// Bandwidth-bound passes around the convolutions, all on fp32 NDHWC tensors:
//   instance-norm statistics, fused (affine -> +residual -> leaky-relu -> nearest x2 -> concat-slice) forward and
//   backward, instance-norm backward, 2x2x2 max-pool, CT-volume molding.
// Replaces InstanceNorm3d / LeakyReLU / Dropout3d / Upsample / cat (mask_branch.py:18-20,91-122,124-220), frozen
// BatchNorm3d + ReLU + residual add (backbone.py:26-114), MaxPool3d (backbone.py:127), mold_image (model.py:1902).
#include "common.cuh"

#include <cuda_bf16.h>

namespace cfun {

int launch_pack_zero_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int G, int P, cudaStream_t st);   // conv_tc_halo.cu

// hi = bf16(v), lo = bf16(v - hi): the operand split of the tcgen05 convs (tc_ptx.cuh split_bf16)
__device__ __forceinline__ void split_bf16_ew(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Thread layout shared by the per-(n,c) reduction kernels: a block owns `rows_per_block` consecutive voxels of one
// sample; thread t handles vector-channel q = t % CV (V channels each) of rows r = t / CV, r + R, ...
struct RowMap {
  int CV;  // number of channel vectors per row
  int R;   // rows per iteration
};
static RowMap make_rowmap(int C, int V) {
  RowMap m;
  m.CV = C / V;
  m.R = 256 / m.CV;
  if (m.R < 1) m.R = 1;
  return m;
}

template <int V>
struct Vec;
template <>
struct Vec<4> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<1> {
  static __device__ __forceinline__ void load(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float* v) { p[0] = v[0]; }
};

// ---------------------------------------------------------------------------------------------------------
// instance-norm statistics
// ---------------------------------------------------------------------------------------------------------
// Block-level tail of the per-(n,c) reductions: thread (q, rr) parks its V partial sums of each of the NS statistics in
// shared memory [NS][R][C]; the K = NS*C columns are then summed over the R rows in double by blockDim/K threads each
// (two stages), and one atomic per column leaves the block.  Shared memory: NS*R*C floats + 256 doubles.
template <int V, int NS>
__device__ __forceinline__ void block_reduce_stats(float* sm, const float (*part)[V], int q, int rr, int R, int C,
                                                   bool active, double* __restrict__ acc_n) {
  if (active) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int j = 0; j < V; ++j) sm[(s * R + rr) * C + q * V + j] = part[s][j];
  }
  __syncthreads();
  const int K = NS * C;
  int groups = (int)blockDim.x / K;
  if (groups > R) groups = R;
  if (groups > 1) {
    double* sm2 = reinterpret_cast<double*>(sm + NS * R * C);   // NS*R*C is even for NS = 2
    const int t = threadIdx.x;
    if (t < groups * K) {
      const int col = t % K, grp = t / K;
      const int s = col / C, c = col - s * C;
      double a = 0.0;
      for (int k = grp; k < R; k += groups) a += (double)sm[(s * R + k) * C + c];
      sm2[grp * K + col] = a;
    }
    __syncthreads();
    if (t < K) {
      const int s = t / C, c = t - s * C;
      double a = 0.0;
      for (int g = 0; g < groups; ++g) a += sm2[g * K + t];
      atomicAdd(&acc_n[c * 2 + s], a);
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      const int s = i / C, c = i - s * C;
      double a = 0.0;
      for (int k = 0; k < R; ++k) a += (double)sm[(s * R + k) * C + c];
      atomicAdd(&acc_n[c * 2 + s], a);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ x, long long S, int C, int CV, int R,
                                                       long long rows_per_block, double* __restrict__ acc) {
  extern __shared__ float smf[];  // [2][R][C]
  constexpr int U = 8;
  const int n = blockIdx.y;
  const int q = threadIdx.x % CV, rr = threadIdx.x / CV;
  const bool active = rr < R;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  float part[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) part[0][j] = part[1][j] = 0.f;
  if (active) {
    long long row = r0 + rr;
    const float* xp = x + ((long long)n * S + row) * C + q * V;
    const long long sx = (long long)R * C;
    for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) {
      float v[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<V>::load(xp + u * sx, v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) { part[0][j] += v[u][j]; part[1][j] = fmaf(v[u][j], v[u][j], part[1][j]); }
      xp += U * sx;
    }
    for (; row < r1; row += R) {
      float v[V];
      Vec<V>::load(xp, v);
#pragma unroll
      for (int j = 0; j < V; ++j) { part[0][j] += v[j]; part[1][j] = fmaf(v[j], v[j], part[1][j]); }
      xp += sx;
    }
  }
  block_reduce_stats<V, 2>(smf, part, q, rr, R, C, active, acc + (long long)n * C * 2);
}

__global__ void in_finalize_kernel(const double* __restrict__ acc, long long NC, double invS, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= NC) return;
  double m = acc[2 * i] * invS;
  double var = acc[2 * i + 1] * invS - m * m;
  if (var < 0) var = 0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---------------------------------------------------------------------------------------------------------
// fused affine + residual + leaky relu (+ nearest x2 upsample, + concat slice) : forward
// ---------------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) affine_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                             const float* __restrict__ b, int a_nstride,
                                                             const float* __restrict__ r, float* __restrict__ y, int N,
                                                             int D, int H, int W, int C, int Ctot, int c_off, int up,
                                                             float slope) {
  const int CV = C / V;
  const long long total = (long long)N * D * H * W * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % CV);
    long long vox = i / CV;
    int c = q * V;
    float v[V];
    Vec<V>::load(x + vox * C + c, v);
    long long t = vox;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H); t /= H;
    int d = (int)(t % D); t /= D;
    int n = (int)t;
    if (a) {
      float av[V], bv[V];
      Vec<V>::load(a + (long long)n * a_nstride + c, av);
      Vec<V>::load(b + (long long)n * a_nstride + c, bv);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaf(v[j], av[j], bv[j]);
    }
    if (r) {
      float rv[V];
      Vec<V>::load(r + vox * C + c, rv);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] += rv[j];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    if (up == 1) {
      Vec<V>::store(y + vox * Ctot + c_off + c, v);
    } else {
      const int H2 = H * 2, W2 = W * 2;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy_ = 0; dy_ < 2; ++dy_)
#pragma unroll
          for (int dx_ = 0; dx_ < 2; ++dx_) {
            long long o = (((long long)n * (D * 2) + (2 * d + dz)) * H2 + (2 * h + dy_)) * W2 + (2 * w + dx_);
            Vec<V>::store(y + o * Ctot + c_off + c, v);
          }
    }
  }
}

// up == 1 fast path: the RowMap layout of the reduction kernels (no per-element division, the per-(n,c) constants in
// registers, U independent 16-byte loads per operand in flight per thread)
template <int V, int U>
__global__ void __launch_bounds__(256) affine_act_rows_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                              const float* __restrict__ b, int a_nstride,
                                                              const float* __restrict__ r, float* __restrict__ y,
                                                              long long S, int C, int Ctot, int c_off, float slope,
                                                              int CV, int R, long long rows_per_block) {
  const int n = blockIdx.y;
  const int q = threadIdx.x % CV, rr = threadIdx.x / CV;
  if (rr >= R) return;
  const int c = q * V;
  float av[V], bv[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { av[j] = 1.f; bv[j] = 0.f; }
  if (a) {
    Vec<V>::load(a + (long long)n * a_nstride + c, av);
    Vec<V>::load(b + (long long)n * a_nstride + c, bv);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  long long row = r0 + rr;
  const long long vox0 = (long long)n * S + row;
  const float* xp = x + vox0 * C + c;
  const float* rp = r ? r + vox0 * C + c : nullptr;
  float* yp = y + vox0 * Ctot + c_off + c;
  const long long sx = (long long)R * C, sy = (long long)R * Ctot;
  for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) {
    float v[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) Vec<V>::load(xp + u * sx, v[u]);
    if (rp) {
      float rv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<V>::load(rp + u * sx, rv[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) v[u][j] = fmaf(v[u][j], av[j], bv[j]) + rv[u][j];
      rp += U * sx;
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) v[u][j] = fmaf(v[u][j], av[j], bv[j]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[u][j] = v[u][j] > 0.f ? v[u][j] : v[u][j] * slope;
      Vec<V>::store(yp + u * sy, v[u]);
    }
    xp += U * sx;
    yp += U * sy;
  }
  for (; row < r1; row += R) {
    float v[V];
    Vec<V>::load(xp, v);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaf(v[j], av[j], bv[j]);
    if (rp) {
      float rv[V];
      Vec<V>::load(rp, rv);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] += rv[j];
      rp += sx;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    Vec<V>::store(yp, v);
    xp += sx;
    yp += sy;
  }
}

// backward: per-sample blocks so that instance-norm reductions can be fused
struct ActBwdArgs {
  const float* x; const float* r; const float* dy;
  float* dx; float* dr;
  int D, H, W, C, Ctot, c_off;
  float slope;
  bool stats;
};

// UU consecutive rows (stride R) of one thread: all loads first, then the arithmetic and the stores
template <int V, int UU, bool UP2>
__device__ __forceinline__ void act_bwd_rows(const ActBwdArgs& p, int n, long long S, long long row, int R, int c,
                                             const float* av, const float* bv, float (*part)[V]) {
  float xv[UU][V], rv[UU][V], g[UU][V];
#pragma unroll
  for (int u = 0; u < UU; ++u) Vec<V>::load(p.x + ((long long)n * S + row + (long long)u * R) * p.C + c, xv[u]);
  if (p.r) {
#pragma unroll
    for (int u = 0; u < UU; ++u) Vec<V>::load(p.r + ((long long)n * S + row + (long long)u * R) * p.C + c, rv[u]);
  }
#pragma unroll
  for (int u = 0; u < UU; ++u) {
    const long long rw = row + (long long)u * R;
    if (!UP2) {
      Vec<V>::load(p.dy + ((long long)n * S + rw) * p.Ctot + p.c_off + c, g[u]);
    } else {
      long long t = rw;
      const int w = (int)(t % p.W); t /= p.W;
      const int h = (int)(t % p.H); t /= p.H;
      const int d = (int)t;
      const int H2 = p.H * 2, W2 = p.W * 2;
      float t8[8][V];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long o = (((long long)n * (p.D * 2) + (2 * d + (k >> 2))) * H2 + (2 * h + ((k >> 1) & 1))) * W2 +
                            (2 * w + (k & 1));
        Vec<V>::load(p.dy + o * p.Ctot + p.c_off + c, t8[k]);
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += t8[k][j];  // same summation order as the z, y, x loop nest
        g[u][j] = acc;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UU; ++u) {
    const long long vox = (long long)n * S + row + (long long)u * R;
    float xh[V], o[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      xh[j] = fmaf(xv[u][j], av[j], bv[j]);
      const float pre = p.r ? xh[j] + rv[u][j] : xh[j];
      g[u][j] = pre > 0.f ? g[u][j] : g[u][j] * p.slope;
    }
    if (p.dr) Vec<V>::store(p.dr + vox * p.C + c, g[u]);
    if (p.stats) {
#pragma unroll
      for (int j = 0; j < V; ++j) { part[0][j] += g[u][j]; part[1][j] = fmaf(g[u][j], xh[j], part[1][j]); }
      Vec<V>::store(p.dx + vox * p.C + c, g[u]);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = g[u][j] * av[j];
      Vec<V>::store(p.dx + vox * p.C + c, o);
    }
  }
}

template <int V, bool UP2>
__global__ void __launch_bounds__(256) affine_act_bwd_kernel(const ActBwdArgs p, const float* __restrict__ a,
                                                             const float* __restrict__ b, int a_nstride,
                                                             double* __restrict__ stat_acc, int CV, int R,
                                                             long long rows_per_block) {
  extern __shared__ float smf[];  // [2][R][C] when stat_acc
  constexpr int U = UP2 ? 1 : 4;
  const int n = blockIdx.y;
  const long long S = (long long)p.D * p.H * p.W;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  const int q = threadIdx.x % CV, rr = threadIdx.x / CV;
  const bool active = rr < R;
  float part[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) part[0][j] = part[1][j] = 0.f;
  if (active) {
    const int c = q * V;
    float av[V], bv[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { av[j] = 1.f; bv[j] = 0.f; }
    if (a) {
      Vec<V>::load(a + (long long)n * a_nstride + c, av);
      Vec<V>::load(b + (long long)n * a_nstride + c, bv);
    }
    long long row = r0 + rr;
    for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) act_bwd_rows<V, U, UP2>(p, n, S, row, R, c, av, bv, part);
    if (U > 1)
      for (; row < r1; row += R) act_bwd_rows<V, 1, UP2>(p, n, S, row, R, c, av, bv, part);
  }
  if (stat_acc) block_reduce_stats<V, 2>(smf, part, q, rr, R, p.C, active, stat_acc + (long long)n * p.C * 2);
}

// dx = a * (g - s1/S - xhat * s2/S),  xhat = x*a + b, g stored in dx
// EXTRA: + leaky_relu'(x) * de -- the gradient of a second consumer lrelu(x) of the norm's input (the U-Net's level-1
// context branch, cfun_add_act_stats), so that the two gradients are never summed in a pass of their own
template <int V, bool EXTRA>
__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                           const float* __restrict__ b,
                                                           const double* __restrict__ stat_acc, float* __restrict__ dx,
                                                           long long S, int C, int CV, int R, long long rows_per_block,
                                                           const float* __restrict__ de, float slope) {
  constexpr int U = 4;
  const int n = blockIdx.y;
  const int q = threadIdx.x % CV, rr = threadIdx.x / CV;
  if (rr >= R) return;
  const int c = q * V;
  const double invS = 1.0 / (double)S;
  float av[V], bv[V], m1[V], m2[V];
  Vec<V>::load(a + (long long)n * C + c, av);
  Vec<V>::load(b + (long long)n * C + c, bv);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    m1[j] = (float)(stat_acc[((long long)n * C + c + j) * 2 + 0] * invS);
    m2[j] = (float)(stat_acc[((long long)n * C + c + j) * 2 + 1] * invS);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  long long row = r0 + rr;
  const long long off0 = ((long long)n * S + row) * C + c;
  const float* xp = x + off0;
  float* gp = dx + off0;
  const float* ep = EXTRA ? de + off0 : nullptr;
  const long long sx = (long long)R * C;
  for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) {
    float xv[U][V], g[U][V], ev[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) Vec<V>::load(xp + u * sx, xv[u]);
    if (EXTRA) {
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<V>::load(ep + u * sx, ev[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 t;
      if (V == 4) {
        t = *reinterpret_cast<const float4*>(gp + u * sx);
        g[u][0] = t.x; g[u][1 % V] = t.y; g[u][2 % V] = t.z; g[u][3 % V] = t.w;
      } else {
        g[u][0] = gp[u * sx];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float xh = fmaf(xv[u][j], av[j], bv[j]);
        o[j] = av[j] * (g[u][j] - m1[j] - xh * m2[j]);
        if (EXTRA) o[j] += xv[u][j] > 0.f ? ev[u][j] : ev[u][j] * slope;
      }
      Vec<V>::store(gp + u * sx, o);
    }
    xp += U * sx;
    gp += U * sx;
    if (EXTRA) ep += U * sx;
  }
  for (; row < r1; row += R) {
    float xv[V], g[V], o[V], ev[V];
    Vec<V>::load(xp, xv);
    if (EXTRA) Vec<V>::load(ep, ev);
#pragma unroll
    for (int j = 0; j < V; ++j) g[j] = gp[j];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xh = fmaf(xv[j], av[j], bv[j]);
      o[j] = av[j] * (g[j] - m1[j] - xh * m2[j]);
      if (EXTRA) o[j] += xv[j] > 0.f ? ev[j] : ev[j] * slope;
    }
    Vec<V>::store(gp, o);
    xp += sx;
    gp += sx;
    if (EXTRA) ep += sx;
  }
}

// s = a + b, ctx = leaky_relu(s), and the InstanceNorm statistics of s, in one pass (the U-Net's level-1 residual sum feeds
// both the skip connection and the norm: mask_branch.py:132-136)
template <int V>
__global__ void __launch_bounds__(256) add_act_stats_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ s, float* __restrict__ ctx, float slope, long long S,
                                                            int C, int CV, int R, long long rows_per_block, double* __restrict__ acc) {
  extern __shared__ float smf[];  // [2][R][C] + 256 doubles
  constexpr int U = 4;
  const int n = blockIdx.y;
  const int q = threadIdx.x % CV, rr = threadIdx.x / CV;
  const bool active = rr < R;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  float part[2][V];
#pragma unroll
  for (int j = 0; j < V; ++j) part[0][j] = part[1][j] = 0.f;
  if (active) {
    long long row = r0 + rr;
    const long long off0 = ((long long)n * S + row) * C + q * V;
    const float* ap = a + off0;
    const float* bp = b + off0;
    float* sp = s + off0;
    float* cp = ctx + off0;
    const long long sx = (long long)R * C;
    auto one = [&](const float* av, const float* bv, long long o) {
      float sv[V], cv[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        sv[j] = av[j] + bv[j];
        cv[j] = sv[j] > 0.f ? sv[j] : sv[j] * slope;
        part[0][j] += sv[j];
        part[1][j] = fmaf(sv[j], sv[j], part[1][j]);
      }
      Vec<V>::store(sp + o, sv);
      Vec<V>::store(cp + o, cv);
    };
    for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) {
      float av[U][V], bv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<V>::load(ap + u * sx, av[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<V>::load(bp + u * sx, bv[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) one(av[u], bv[u], u * sx);
      ap += U * sx; bp += U * sx; sp += U * sx; cp += U * sx;
    }
    for (; row < r1; row += R) {
      float av[V], bv[V];
      Vec<V>::load(ap, av);
      Vec<V>::load(bp, bv);
      one(av, bv, 0);
      ap += sx; bp += sx; sp += sx; cp += sx;
    }
  }
  block_reduce_stats<V, 2>(smf, part, q, rr, R, C, active, acc + (long long)n * C * 2);
}

// in_bwd_apply whose result goes straight into the split-bf16 group-planar operand pack of the producing conv's backward
// (conv_fused.cu) instead of an fp32 tensor: hi/lo [G][N*(D+2P)][H][W][8]; thread = one 16-byte pack row (8 channels of one
// voxel, the second quad or the whole row may lie in the pack's zero padding beyond C).  g: the un-normalised gradient that
// affine_act_bwd left in its dx buffer (read only).
__global__ void __launch_bounds__(256) in_bwd_apply_pack_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                                const float* __restrict__ b, const double* __restrict__ stat_acc,
                                                                const float* __restrict__ g, long long S, int C, int G, int R,
                                                                long long rows_per_block, __nv_bfloat16* __restrict__ hi,
                                                                __nv_bfloat16* __restrict__ lo, long long plane_rows, long long vox_p,
                                                                long long pad_rows) {
  constexpr int U = 2;
  const int n = blockIdx.y;
  const int q = threadIdx.x % G, rr = threadIdx.x / G;
  if (rr >= R) return;
  const int c = q * 8;
  const int nq = c >= C ? 0 : (c + 4 >= C ? 1 : 2);          // real channel quads of this row (C is a multiple of 4)
  const double invS = 1.0 / (double)S;
  float av[8], bv[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) av[j] = bv[j] = m1[j] = m2[j] = 0.f;
  for (int h = 0; h < nq; ++h) {
    Vec<4>::load(a + (long long)n * C + c + 4 * h, av + 4 * h);
    Vec<4>::load(b + (long long)n * C + c + 4 * h, bv + 4 * h);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m1[4 * h + j] = (float)(stat_acc[((long long)n * C + c + 4 * h + j) * 2 + 0] * invS);
      m2[4 * h + j] = (float)(stat_acc[((long long)n * C + c + 4 * h + j) * 2 + 1] * invS);
    }
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > S) r1 = S;
  // pack row of (sample n, voxel row): group q, position n * plane_rows + pad_rows + row
  uint4* ohi = reinterpret_cast<uint4*>(hi) + (long long)q * vox_p + (long long)n * plane_rows + pad_rows;
  uint4* olo = reinterpret_cast<uint4*>(lo) + (long long)q * vox_p + (long long)n * plane_rows + pad_rows;
  auto emit = [&](long long row, const float* xv, const float* gv) {
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xv[j], av[j], bv[j]);
      split_bf16_ew(av[j] * (gv[j] - m1[j] - xh * m2[j]), h[j], l[j]);      // padding channels: av = 0 -> +0
    }
    ohi[row] = *reinterpret_cast<const uint4*>(h);
    olo[row] = *reinterpret_cast<const uint4*>(l);
  };
  long long row = r0 + rr;
  const long long off0 = ((long long)n * S + row) * C + c;
  const float* xp = x + off0;
  const float* gp = g + off0;
  const long long sx = (long long)R * C;
  for (; row + (long long)(U - 1) * R < r1; row += (long long)U * R) {
    float xv[U][8], gv[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[u][j] = gv[u][j] = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (nq > 0) { Vec<4>::load(xp + u * sx, xv[u]); Vec<4>::load(gp + u * sx, gv[u]); }
      if (nq > 1) { Vec<4>::load(xp + u * sx + 4, xv[u] + 4); Vec<4>::load(gp + u * sx + 4, gv[u] + 4); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) emit(row + (long long)u * R, xv[u], gv[u]);
    xp += U * sx;
    gp += U * sx;
  }
  for (; row < r1; row += R) {
    float xv[8], gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = gv[j] = 0.f;
    if (nq > 0) { Vec<4>::load(xp, xv); Vec<4>::load(gp, gv); }
    if (nq > 1) { Vec<4>::load(xp + 4, xv + 4); Vec<4>::load(gp + 4, gv + 4); }
    emit(row, xv, gv);
    xp += sx;
    gp += sx;
  }
}

// InstanceNorm apply + LeakyReLU + nearest x2 upsampling written straight into the split-bf16 group-planar operand pack of the
// conv that follows (the U-Net decoder's norm -> lrelu -> upsample -> conv, mask_branch.py:91-103): thread = (low-res voxel,
// 8-channel group); the activated values are split once and the same 16-byte hi / lo rows go to the 2x2x2 output voxels.
// The 8x larger upsampled fp32 tensor is never written.  hi/lo: [G][N*(2D+2P)][2H][2W][8].
__global__ void __launch_bounds__(256) affine_act_up2_pack_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                                  const float* __restrict__ b, int N, int D, int H, int W, int C,
                                                                  float slope, __nv_bfloat16* __restrict__ hi,
                                                                  __nv_bfloat16* __restrict__ lo, int G, int P) {
  const long long vox = (long long)N * D * H * W;
  const long long total = vox * G;
  const int H2 = 2 * H, W2 = 2 * W;
  const long long plane2 = (long long)H2 * W2;
  const long long vox_p = (long long)N * (2 * D + 2 * P) * plane2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i / vox);
    long long v = i - (long long)g * vox;
    const int w = (int)(v % W); long long t = v / W;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    const int n = (int)(t / D);
    const int c = 8 * g;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (c + 4 * q < C) {
        float xv[4], av[4], bv[4];
        Vec<4>::load(x + v * C + c + 4 * q, xv);
        Vec<4>::load(a + (long long)n * C + c + 4 * q, av);
        Vec<4>::load(b + (long long)n * C + c + 4 * q, bv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float y = fmaf(xv[j], av[j], bv[j]);
          f[4 * q + j] = y > 0.f ? y : y * slope;
        }
      }
    }
    __align__(16) __nv_bfloat16 hh[8];
    __align__(16) __nv_bfloat16 ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16_ew(f[j], hh[j], ll[j]);
    const uint4 hv = *reinterpret_cast<const uint4*>(hh), lv = *reinterpret_cast<const uint4*>(ll);
    uint4* oh = reinterpret_cast<uint4*>(hi) + (long long)g * vox_p;
    uint4* ol = reinterpret_cast<uint4*>(lo) + (long long)g * vox_p;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long pos = ((long long)n * (2 * D + 2 * P) + P + 2 * d + (k >> 2)) * plane2 + (long long)(2 * h + ((k >> 1) & 1)) * W2 +
                            (2 * w + (k & 1));
      oh[pos] = hv;
      ol[pos] = lv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2x2x2 max pool, stride 2 (even extents)
// ---------------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D,
                                                           int H, int W, int C) {
  const int CV = C / V, Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % CV);
    long long t = i / CV;
    int w = (int)(t % Wo); t /= Wo;
    int h = (int)(t % Ho); t /= Ho;
    int d = (int)(t % Do); t /= Do;
    int n = (int)t;
    float m[V];
#pragma unroll
    for (int j = 0; j < V; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy_ = 0; dy_ < 2; ++dy_)
#pragma unroll
        for (int dx_ = 0; dx_ < 2; ++dx_) {
          float v[V];
          Vec<V>::load(x + ((((long long)n * D + 2 * d + dz) * H + 2 * h + dy_) * W + 2 * w + dx_) * C + q * V, v);
#pragma unroll
          for (int j = 0; j < V; ++j) m[j] = (v[j] > m[j] || v[j] != v[j]) ? v[j] : m[j];
        }
    Vec<V>::store(y + (i / CV) * C + q * V, m);
  }
}

template <int V>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int N, int D, int H, int W, int C) {
  const int CV = C / V, Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % CV);
    long long t = i / CV;
    int w = (int)(t % Wo); t /= Wo;
    int h = (int)(t % Ho); t /= Ho;
    int d = (int)(t % Do); t /= Do;
    int n = (int)t;
    float m[V], g[V];
    int arg[V];
    Vec<V>::load(dy + (i / CV) * C + q * V, g);
#pragma unroll
    for (int j = 0; j < V; ++j) { m[j] = -INFINITY; arg[j] = 0; }
    float v[8][V];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int dz = k >> 2, dy_ = (k >> 1) & 1, dx_ = k & 1;
      Vec<V>::load(x + ((((long long)n * D + 2 * d + dz) * H + 2 * h + dy_) * W + 2 * w + dx_) * C + q * V, v[k]);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (v[k][j] > m[j] || v[k][j] != v[k][j]) { m[j] = v[k][j]; arg[j] = k; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int dz = k >> 2, dy_ = (k >> 1) & 1, dx_ = k & 1;
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = arg[j] == k ? g[j] : 0.f;
      Vec<V>::store(dx + ((((long long)n * D + 2 * d + dz) * H + 2 * h + dy_) * W + 2 * w + dx_) * C + q * V, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// CT volume molding: int16 [H,W,D] -> fp32 [D,H,W], (x - mean) / std
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) i16_stats_kernel(const short* __restrict__ v, long long n, double* __restrict__ acc) {
  long long s = 0, ss = 0;  // exact integer accumulation
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long t = v[i];
    s += t;
    ss += t * t;
  }
  double ds = warp_sum((double)s), dss = warp_sum((double)ss);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, ds);
    atomicAdd(acc + 1, dss);
  }
}
__global__ void mold_transpose_kernel(const short* __restrict__ v, int H, int W, int D, const double* __restrict__ acc,
                                      float* __restrict__ out) {
  __shared__ float tile[32][33];
  const double cnt = (double)H * W * D;
  const double mean = acc[0] / cnt;
  double var = acc[1] / cnt - mean * mean;
  if (var < 0) var = 0;
  const float fm = (float)mean, fr = (float)(1.0 / sqrt(var));
  const int h = blockIdx.z;
  const int w0 = blockIdx.y * 32, d0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int w = w0 + j, d = d0 + threadIdx.x;
    if (w < W && d < D) tile[j][threadIdx.x] = ((float)v[((long long)h * W + w) * D + d] - fm) * fr;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int d = d0 + j, w = w0 + threadIdx.x;
    if (w < W && d < D) out[((long long)d * H + h) * W + w] = tile[threadIdx.x][j];
  }
}

// rows of one sample per block for the RowMap kernels: about `per_sm` blocks per SM over the whole launch (16 for the
// streaming kernels; 6 for the reductions, whose blocks each end in 2C double atomics on the same N*C addresses), at
// least 8 iterations of R rows each
static long long rows_per_block(long long S, int N, int R, int per_sm = 16) {
  const long long blocks_per_sample = std::max<long long>(1, (long long)per_sm * num_sms() / N);
  return std::max<long long>((long long)R * 8, cdiv(S, blocks_per_sample));
}

static int pick_blocks(long long total) {
  long long b = cdiv(total, 256);
  long long cap = 32LL * num_sms();
  return (int)std::max<long long>(1, std::min(b, cap));
}

}  // namespace cfun

using namespace cfun;

extern "C" int cfun_instnorm_stats(const float* x, int N, long long S, int C, float eps, double* acc, float* mean,
                                   float* rstd, void* stream) {
  CFUN_CHECK_ARG(x && acc && mean && rstd && N > 0 && S > 0 && C > 0 && C <= 1024);
  cudaStream_t st = as_stream(stream);
  CFUN_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * (size_t)N * C, st));
  const int V = (C % 4 == 0) ? 4 : 1;
  CFUN_CHECK_ARG(C / V <= 256);
  RowMap rm = make_rowmap(C, V);
  long long rpb = rows_per_block(S, N, rm.R, 6);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  size_t smem = sizeof(float) * 2 * rm.R * C + sizeof(double) * 256;
  if (V == 4) in_stats_kernel<4><<<grid, 256, smem, st>>>(x, S, C, rm.CV, rm.R, rpb, acc);
  else in_stats_kernel<1><<<grid, 256, smem, st>>>(x, S, C, rm.CV, rm.R, rpb, acc);
  CFUN_LAUNCH_CHECK();
  long long NC = (long long)N * C;
  in_finalize_kernel<<<(unsigned)cdiv(NC, 256), 256, 0, st>>>(acc, NC, 1.0 / (double)S, eps, mean, rstd);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// mean / rstd from statistics accumulated elsewhere (the conv epilogue of cfun_conv3d_fwd_stats)
extern "C" int cfun_instnorm_finalize(const double* acc, int N, long long S, int C, float eps, float* mean, float* rstd,
                                      void* stream) {
  CFUN_CHECK_ARG(acc && mean && rstd && N > 0 && S > 0 && C > 0);
  const long long NC = (long long)N * C;
  in_finalize_kernel<<<(unsigned)cdiv(NC, 256), 256, 0, as_stream(stream)>>>(acc, NC, 1.0 / (double)S, eps, mean, rstd);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_affine_act_fwd(const float* x, const float* a, const float* b, int a_nstride, const float* r,
                                   float* y, int N, int D, int H, int W, int C, int Ctot, int c_off, int up, float slope,
                                   void* stream) {
  CFUN_CHECK_ARG(x && y && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && Ctot >= c_off + C && c_off >= 0);
  CFUN_CHECK_ARG(up == 1 || up == 2);
  CFUN_CHECK_ARG((a == nullptr) == (b == nullptr));
  cudaStream_t st = as_stream(stream);
  const bool v4 = (C % 4 == 0) && (Ctot % 4 == 0) && (c_off % 4 == 0) && (a_nstride % 4 == 0);
  long long total = (long long)N * D * H * W * (v4 ? C / 4 : C);
  if (up == 1 && C / (v4 ? 4 : 1) <= 256) {
    RowMap rm = make_rowmap(C, v4 ? 4 : 1);
    const long long S = (long long)D * H * W;
    long long rpb = rows_per_block(S, N, rm.R);
    dim3 grid((unsigned)cdiv(S, rpb), N);
    if (v4) affine_act_rows_kernel<4, 4><<<grid, 256, 0, st>>>(x, a, b, a_nstride, r, y, S, C, Ctot, c_off, slope, rm.CV, rm.R, rpb);
    else affine_act_rows_kernel<1, 4><<<grid, 256, 0, st>>>(x, a, b, a_nstride, r, y, S, C, Ctot, c_off, slope, rm.CV, rm.R, rpb);
  } else if (v4) {
    affine_act_fwd_kernel<4><<<pick_blocks(total), 256, 0, st>>>(x, a, b, a_nstride, r, y, N, D, H, W, C, Ctot, c_off, up, slope);
  } else {
    affine_act_fwd_kernel<1><<<pick_blocks(total), 256, 0, st>>>(x, a, b, a_nstride, r, y, N, D, H, W, C, Ctot, c_off, up, slope);
  }
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_affine_act_bwd(const float* x, const float* a, const float* b, int a_nstride, const float* r,
                                   const float* dy, float* dx, float* dr, double* stat_acc, int N, int D, int H, int W,
                                   int C, int Ctot, int c_off, int up, float slope, void* stream) {
  CFUN_CHECK_ARG(x && dy && dx && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && Ctot >= c_off + C && c_off >= 0);
  CFUN_CHECK_ARG(up == 1 || up == 2);
  CFUN_CHECK_ARG((a == nullptr) == (b == nullptr));
  CFUN_CHECK_ARG(!stat_acc || (a && a_nstride == C && !r));
  cudaStream_t st = as_stream(stream);
  if (stat_acc) CFUN_CUDA(cudaMemsetAsync(stat_acc, 0, sizeof(double) * 2 * (size_t)N * C, st));
  const bool v4 = (C % 4 == 0) && (Ctot % 4 == 0) && (c_off % 4 == 0) && (a_nstride % 4 == 0);
  const int V = v4 ? 4 : 1;
  CFUN_CHECK_ARG(C / V <= 256);
  RowMap rm = make_rowmap(C, V);
  const long long S = (long long)D * H * W;
  long long rpb = rows_per_block(S, N, rm.R, stat_acc ? 6 : 16);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  size_t smem = stat_acc ? sizeof(float) * 2 * rm.R * C + sizeof(double) * 256 : 0;
  ActBwdArgs p{x, r, dy, dx, dr, D, H, W, C, Ctot, c_off, slope, stat_acc != nullptr};
  if (v4 && up == 1) affine_act_bwd_kernel<4, false><<<grid, 256, smem, st>>>(p, a, b, a_nstride, stat_acc, rm.CV, rm.R, rpb);
  else if (v4) affine_act_bwd_kernel<4, true><<<grid, 256, smem, st>>>(p, a, b, a_nstride, stat_acc, rm.CV, rm.R, rpb);
  else if (up == 1) affine_act_bwd_kernel<1, false><<<grid, 256, smem, st>>>(p, a, b, a_nstride, stat_acc, rm.CV, rm.R, rpb);
  else affine_act_bwd_kernel<1, true><<<grid, 256, smem, st>>>(p, a, b, a_nstride, stat_acc, rm.CV, rm.R, rpb);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_instnorm_bwd_apply(const float* x, const float* a, const float* b, const double* stat_acc,
                                       float* dx, int N, long long S, int C, void* stream) {
  CFUN_CHECK_ARG(x && a && b && stat_acc && dx && N > 0 && S > 0 && C > 0);
  cudaStream_t st = as_stream(stream);
  const bool v4 = (C % 4 == 0);
  CFUN_CHECK_ARG(C / (v4 ? 4 : 1) <= 256);
  RowMap rm = make_rowmap(C, v4 ? 4 : 1);
  long long rpb = rows_per_block(S, N, rm.R);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  if (v4) in_bwd_apply_kernel<4, false><<<grid, 256, 0, st>>>(x, a, b, stat_acc, dx, S, C, rm.CV, rm.R, rpb, nullptr, 0.f);
  else in_bwd_apply_kernel<1, false><<<grid, 256, 0, st>>>(x, a, b, stat_acc, dx, S, C, rm.CV, rm.R, rpb, nullptr, 0.f);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// cfun_instnorm_bwd_apply + leaky_relu'(x, slope) * dextra: the norm's input x also fed a LeakyReLU whose output gradient is
// dextra (cfun_add_act_stats) -- both gradients of x in one pass
extern "C" int cfun_instnorm_bwd_apply_extra(const float* x, const float* a, const float* b, const double* stat_acc, float* dx,
                                             int N, long long S, int C, const float* dextra, float slope, void* stream) {
  CFUN_CHECK_ARG(x && a && b && stat_acc && dx && dextra && N > 0 && S > 0 && C > 0);
  cudaStream_t st = as_stream(stream);
  const bool v4 = (C % 4 == 0);
  CFUN_CHECK_ARG(C / (v4 ? 4 : 1) <= 256);
  RowMap rm = make_rowmap(C, v4 ? 4 : 1);
  long long rpb = rows_per_block(S, N, rm.R);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  if (v4) in_bwd_apply_kernel<4, true><<<grid, 256, 0, st>>>(x, a, b, stat_acc, dx, S, C, rm.CV, rm.R, rpb, dextra, slope);
  else in_bwd_apply_kernel<1, true><<<grid, 256, 0, st>>>(x, a, b, stat_acc, dx, S, C, rm.CV, rm.R, rpb, dextra, slope);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// s = a + b, ctx = leaky_relu(s, slope), InstanceNorm statistics of s (acc, mean, rstd as cfun_instnorm_stats) in one pass
extern "C" int cfun_add_act_stats(const float* a, const float* b, float* s, float* ctx, int N, long long S, int C, float slope,
                                  float eps, double* acc, float* mean, float* rstd, void* stream) {
  CFUN_CHECK_ARG(a && b && s && ctx && acc && mean && rstd && N > 0 && S > 0 && C > 0 && C <= 1024);
  cudaStream_t st = as_stream(stream);
  CFUN_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * (size_t)N * C, st));
  const int V = (C % 4 == 0) ? 4 : 1;
  CFUN_CHECK_ARG(C / V <= 256);
  RowMap rm = make_rowmap(C, V);
  long long rpb = rows_per_block(S, N, rm.R, 6);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  size_t smem = sizeof(float) * 2 * rm.R * C + sizeof(double) * 256;
  if (V == 4) add_act_stats_kernel<4><<<grid, 256, smem, st>>>(a, b, s, ctx, slope, S, C, rm.CV, rm.R, rpb, acc);
  else add_act_stats_kernel<1><<<grid, 256, smem, st>>>(a, b, s, ctx, slope, S, C, rm.CV, rm.R, rpb, acc);
  CFUN_LAUNCH_CHECK();
  long long NC = (long long)N * C;
  in_finalize_kernel<<<(unsigned)cdiv(NC, 256), 256, 0, st>>>(acc, NC, 1.0 / (double)S, eps, mean, rstd);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// channel concatenation of two NDHWC tensors (torch.cat(dim=1) of the U-Net skip connections, mask_branch.py:189,197,
// 204,211) and its backward (split): float4 rows, one pass, instead of torch's strided copy kernels
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) cat2_kernel(float* __restrict__ a, float* __restrict__ b, float* __restrict__ o, long long M,
                                                   int C1q, int C2q) {
  const int Cq = C1q + C2q;
  const long long total = M * Cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / Cq;
    const int c = (int)(i - v * Cq);
    float4* src = c < C1q ? reinterpret_cast<float4*>(a) + v * C1q + c : reinterpret_cast<float4*>(b) + v * C2q + (c - C1q);
    float4* cat = reinterpret_cast<float4*>(o) + i;
    if (SPLIT) *src = *cat;
    else *cat = *src;
  }
}

extern "C" int cfun_instnorm_bwd_apply_pack(const float* x, const float* a, const float* b, const double* stat_acc, const float* g,
                                            int N, int D, int H, int W, int C, void* hi, void* lo, int G, int P, void* stream) {
  CFUN_CHECK_ARG(x && a && b && stat_acc && g && hi && lo && N > 0 && D > 0 && H > 0 && W > 0 && C > 0);
  CFUN_CHECK_ARG((C & 3) == 0 && G * 8 >= C && G <= 256 && P >= 1 && P <= 2);
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(lo);
  int rc = launch_pack_zero_planes(h, l, N, D, H, W, G, P, st);
  if (rc != CFUN_OK) return rc;
  const long long S = (long long)D * H * W, HW = (long long)H * W;
  const int R = 256 / G;                        // thread = (row, channel group): G threads per voxel row
  long long rpb = rows_per_block(S, N, R);
  dim3 grid((unsigned)cdiv(S, rpb), N);
  in_bwd_apply_pack_kernel<<<grid, 256, 0, st>>>(x, a, b, stat_acc, g, S, C, G, R, rpb, h, l, (long long)(D + 2 * P) * HW,
                                                 (long long)N * (D + 2 * P) * HW, (long long)P * HW);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_instnorm_up2_pack(const float* x, const float* a, const float* b, int N, int D, int H, int W, int C, float slope,
                                      void* hi, void* lo, int G, int P, void* stream) {
  CFUN_CHECK_ARG(x && a && b && hi && lo && N > 0 && D > 0 && H > 0 && W > 0 && C > 0);
  CFUN_CHECK_ARG((C & 3) == 0 && G * 8 >= C && P >= 1 && P <= 2);
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(lo);
  int rc = launch_pack_zero_planes(h, l, N, 2 * D, 2 * H, 2 * W, G, P, st);
  if (rc != CFUN_OK) return rc;
  const long long total = (long long)N * D * H * W * G;
  affine_act_up2_pack_kernel<<<pick_blocks(total), 256, 0, st>>>(x, a, b, N, D, H, W, C, slope, h, l, G, P);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_cat2_channels(const float* a, int C1, const float* b, int C2, float* out, long long M, void* stream) {
  CFUN_CHECK_ARG(a && b && out && M > 0 && C1 > 0 && C2 > 0 && C1 % 4 == 0 && C2 % 4 == 0);
  cat2_kernel<false><<<pick_blocks(M * ((C1 + C2) / 4)), 256, 0, as_stream(stream)>>>(const_cast<float*>(a), const_cast<float*>(b), out, M, C1 / 4, C2 / 4);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_split2_channels(const float* cat, int C1, int C2, float* a, float* b, long long M, void* stream) {
  CFUN_CHECK_ARG(a && b && cat && M > 0 && C1 > 0 && C2 > 0 && C1 % 4 == 0 && C2 % 4 == 0);
  cat2_kernel<true><<<pick_blocks(M * ((C1 + C2) / 4)), 256, 0, as_stream(stream)>>>(a, b, const_cast<float*>(cat), M, C1 / 4, C2 / 4);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_maxpool2_fwd(const float* x, float* y, int N, int D, int H, int W, int C, void* stream) {
  CFUN_CHECK_ARG(x && y && N > 0 && C > 0 && D > 0 && H > 0 && W > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0);
  cudaStream_t st = as_stream(stream);
  const bool v4 = C % 4 == 0;
  long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (v4 ? C / 4 : C);
  if (v4) maxpool2_fwd_kernel<4><<<pick_blocks(total), 256, 0, st>>>(x, y, N, D, H, W, C);
  else maxpool2_fwd_kernel<1><<<pick_blocks(total), 256, 0, st>>>(x, y, N, D, H, W, C);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_maxpool2_bwd(const float* x, const float* y, const float* dy, float* dx, int N, int D, int H, int W,
                                 int C, void* stream) {
  (void)y;
  CFUN_CHECK_ARG(x && dy && dx && N > 0 && C > 0 && D > 0 && H > 0 && W > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0);
  cudaStream_t st = as_stream(stream);
  const bool v4 = C % 4 == 0;
  long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (v4 ? C / 4 : C);
  if (v4) maxpool2_bwd_kernel<4><<<pick_blocks(total), 256, 0, st>>>(x, dy, dx, N, D, H, W, C);
  else maxpool2_bwd_kernel<1><<<pick_blocks(total), 256, 0, st>>>(x, dy, dx, N, D, H, W, C);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_mold_volume_i16(const short* vol_hwd, int H, int W, int D, double* acc, float* out_dhw, void* stream) {
  CFUN_CHECK_ARG(vol_hwd && acc && out_dhw && H > 0 && W > 0 && D > 0);
  cudaStream_t st = as_stream(stream);
  CFUN_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  long long n = (long long)H * W * D;
  i16_stats_kernel<<<pick_blocks(n / 4 + 1), 256, 0, st>>>(vol_hwd, n, acc);
  CFUN_LAUNCH_CHECK();
  dim3 grid((unsigned)cdiv(D, 32), (unsigned)cdiv(W, 32), (unsigned)H);
  mold_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(vol_hwd, H, W, D, acc, out_dhw);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
