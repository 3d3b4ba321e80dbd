// 3-D Sobel edge loss, forward and backward, fused over all (RoI, class) planes, shared-memory tiled.
// Replaces model.compute_mrcnn_mask_edge_loss (model.py:938-981): 56 tiny conv3d launches + ~10 elementwise passes per plane
// in the reference.  Two variants (mode):
//   0  heart (model.py:969-975): magnitude = sqrt(g0^2 + g1^2 + g0^2) -- response 0 twice, the third Sobel response unused --
//      of prediction and target, MSE between the magnitudes;
//   1  LiTS (LiTS_2017/model.py:967-975): MSE between the RAW three responses (the magnitude lines are commented out there).
// In both the classes are 1..ncls-1 (the reference hard-codes range(7) / target_masks[:, 1:]), the per-plane MSE means are
// summed and divided by the number of positives.  The crop may be non-cubic (LiTS MASK_SHAPE = (32, 80, 80)).
//
// Round 2: the round-1 kernels read every prediction voxel 24 times from L1 with 4-byte loads at a 32-byte stride (8 wavefronts
// per warp load) and ran at 5 % of the HBM roofline at 4 x 192^3 (4.6 ms forward).  Now a block stages a (4+2) x (8+2) x (32+2)
// voxel tile of all classes in shared memory once (coalesced 32-byte-per-voxel rows, transposed to class planes) and the 27
// taps are conflict-free shared-memory reads.
#include "common.cuh"

namespace cfun {

// Sobel bank as built in model.py:947-952 (cross-correlation weights [kD=a][kH=b][kW=c]), s = (1,2,1), d = (1,0,-1):
//   k0 = s[a] d[b] s[c] (derivative along H),  k1 = d[a] s[b] s[c] (along D),  k2 = s[a] s[b] d[c] (along W)
__device__ __forceinline__ float sm3(int i) { return i == 1 ? 2.f : 1.f; }
__device__ __forceinline__ float df3(int i) { return i == 0 ? 1.f : (i == 1 ? 0.f : -1.f); }

constexpr int SB_TZ = 4, SB_TY = 8, SB_TX = 32;              // output tile of one block (256 threads: x, y; z looped)
constexpr int SB_IZ = SB_TZ + 2, SB_IY = SB_TY + 2, SB_IX = SB_TX + 2;
constexpr int SB_PLANE = SB_IZ * SB_IY * SB_IX;              // 2040 voxels
constexpr int SB_MAXC = 7;                                   // foreground classes handled per pass

struct SobelGeo {
  int P, Md, Mh, Mw, ncls, mode;
  int tz, ty, tx;           // tiles per dim over the OUTPUT (valid) volume (M - 2)
};

// responses of one class plane at output (z, y, x) of the tile from shared memory s[z][y][x]
__device__ __forceinline__ void sobel3(const float* __restrict__ s, int z, int y, int x, float& g0, float& g1, float& g2) {
  g0 = g1 = g2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = s[((z + a) * SB_IY + (y + b)) * SB_IX + (x + c)];
        g0 = fmaf(sm3(a) * df3(b) * sm3(c), v, g0);
        g1 = fmaf(df3(a) * sm3(b) * sm3(c), v, g1);
        g2 = fmaf(sm3(a) * sm3(b) * df3(c), v, g2);
      }
}

// pass 1: forward loss (BWD = false) or the per-output derivative coefficients A (BWD = true), planar
// A[p][class][r][oz][oy][ox], r < (mode ? 3 : 2)
template <bool BWD>
__global__ void __launch_bounds__(256) sobel_pass1_kernel(const float* __restrict__ pred, const long long* __restrict__ tgt, SobelGeo g,
                                                          double* __restrict__ loss_acc, const float* __restrict__ grad_scale,
                                                          float denom, float* __restrict__ A) {
  extern __shared__ float sm[];                       // [nfg][SB_PLANE] predictions, then SB_PLANE target ids (as float)
  const int nfg = g.ncls - 1;
  float* st = sm + (size_t)nfg * SB_PLANE;
  const int Do = g.Md - 2, Ho = g.Mh - 2, Wo = g.Mw - 2;
  long long b = blockIdx.x;
  const int bx = (int)(b % g.tx); b /= g.tx;
  const int by = (int)(b % g.ty); b /= g.ty;
  const int bz = (int)(b % g.tz);
  const int p = (int)(b / g.tz);
  const int z0 = bz * SB_TZ, y0 = by * SB_TY, x0 = bx * SB_TX;
  // stage the input tile: voxel rows are ncls contiguous floats (channels-last); out-of-range voxels are zero
  for (int i = threadIdx.x; i < SB_PLANE; i += blockDim.x) {
    const int x = i % SB_IX, y = (i / SB_IX) % SB_IY, z = i / (SB_IX * SB_IY);
    const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
    const bool ok = gz < g.Md && gy < g.Mh && gx < g.Mw;
    const long long vox = (((long long)p * g.Md + gz) * g.Mh + gy) * g.Mw + gx;
    const float* pv = pred + vox * g.ncls;
    for (int c = 0; c < nfg; ++c) sm[(size_t)c * SB_PLANE + i] = ok ? __ldg(pv + c + 1) : 0.f;
    st[i] = ok ? (float)tgt[vox] : -1.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ox = x0 + tx, oy = y0 + ty;
  const float gs = BWD ? (*grad_scale) * (2.0f / denom) : 0.f;
  const long long per = (long long)Do * Ho * Wo;
  const int R = g.mode ? 3 : 2;
  double local = 0.0;
  for (int z = 0; z < SB_TZ; ++z) {
    const int oz = z0 + z;
    const bool live = oz < Do && oy < Ho && ox < Wo;
    for (int c = 0; c < nfg; ++c) {
      float gp0, gp1, gp2, gt0 = 0.f, gt1 = 0.f, gt2 = 0.f;
      sobel3(sm + (size_t)c * SB_PLANE, z, ty, tx, gp0, gp1, gp2);
      const float cls = (float)(c + 1);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            const float v = st[((z + a) * SB_IY + (ty + bb)) * SB_IX + (tx + cc)] == cls ? 1.f : 0.f;
            gt0 = fmaf(sm3(a) * df3(bb) * sm3(cc), v, gt0);
            gt1 = fmaf(df3(a) * sm3(bb) * sm3(cc), v, gt1);
            gt2 = fmaf(sm3(a) * sm3(bb) * df3(cc), v, gt2);
          }
      if (!live) continue;
      if (g.mode == 0) {
        const float mp = sqrtf(gp0 * gp0 + gp1 * gp1 + gp0 * gp0);
        const float mt = sqrtf(gt0 * gt0 + gt1 * gt1 + gt0 * gt0);
        const float diff = mp - mt;
        if (!BWD) local += (double)diff * (double)diff;
        else {
          // autograd of mse(sqrt(g0^2+g1^2+g0^2)): dmag/(2*mag) * (4*g0 , 2*g1); 0/0 -> NaN exactly like torch
          const float h = (gs * diff) / (2.f * mp);
          float* a0 = A + ((((long long)p * nfg + c) * R + 0) * per) + ((long long)oz * Ho + oy) * Wo + ox;
          a0[0] = h * (2.f * gp0) + h * (2.f * gp0);
          a0[per] = h * (2.f * gp1);
        }
      } else {
        const float d0 = gp0 - gt0, d1 = gp1 - gt1, d2 = gp2 - gt2;
        if (!BWD) local += (double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2;
        else {
          float* a0 = A + ((((long long)p * nfg + c) * R + 0) * per) + ((long long)oz * Ho + oy) * Wo + ox;
          a0[0] = gs * d0; a0[per] = gs * d1; a0[2 * per] = gs * d2;
        }
      }
    }
  }
  if (!BWD) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(loss_acc, local);
  }
}

__global__ void sobel_finalize_kernel(const double* __restrict__ acc, double denom, float* __restrict__ loss) {
  *loss = (float)(*acc / denom);
}

// pass 2, transposed stencil: dpred[p, v, c+1] = sum_taps sum_r k_r[tap] * A_r[v - tap]; one class at a time through shared memory
__global__ void __launch_bounds__(256) sobel_pass2_kernel(const float* __restrict__ A, SobelGeo g, float* __restrict__ dpred) {
  extern __shared__ float sm[];                       // [R][SB_PLANE]: A tile with origin (v0 - 2)
  const int nfg = g.ncls - 1;
  const int Do = g.Md - 2, Ho = g.Mh - 2, Wo = g.Mw - 2;
  const long long per = (long long)Do * Ho * Wo;
  const int R = g.mode ? 3 : 2;
  long long b = blockIdx.x;
  const int bx = (int)(b % g.tx); b /= g.tx;
  const int by = (int)(b % g.ty); b /= g.ty;
  const int bz = (int)(b % g.tz);
  const int p = (int)(b / g.tz);
  const int z0 = bz * SB_TZ, y0 = by * SB_TY, x0 = bx * SB_TX;       // tile origin over the INPUT volume
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[SB_TZ][SB_MAXC];
#pragma unroll
  for (int z = 0; z < SB_TZ; ++z)
#pragma unroll
    for (int c = 0; c < SB_MAXC; ++c) acc[z][c] = 0.f;
#pragma unroll 1
  for (int c = 0; c < nfg; ++c) {
    __syncthreads();
    for (int i = threadIdx.x; i < SB_PLANE * R; i += blockDim.x) {
      const int r = i / SB_PLANE, j = i - r * SB_PLANE;
      const int x = j % SB_IX, y = (j / SB_IX) % SB_IY, z = j / (SB_IX * SB_IY);
      const int oz = z0 - 2 + z, oy = y0 - 2 + y, ox = x0 - 2 + x;
      float v = 0.f;
      if ((unsigned)oz < (unsigned)Do && (unsigned)oy < (unsigned)Ho && (unsigned)ox < (unsigned)Wo)
        v = __ldg(A + ((((long long)p * nfg + c) * R + r) * per) + ((long long)oz * Ho + oy) * Wo + ox);
      sm[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int z = 0; z < SB_TZ; ++z) {
      float s = 0.f;
      // input voxel (z,ty,tx) of the tile gets A[v - tap]: tile coordinate (v + 2 - tap)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            const int j = ((z + 2 - a) * SB_IY + (ty + 2 - bb)) * SB_IX + (tx + 2 - cc);
            s = fmaf(sm3(a) * df3(bb) * sm3(cc), sm[j], s);
            s = fmaf(df3(a) * sm3(bb) * sm3(cc), sm[SB_PLANE + j], s);
            if (R == 3) s = fmaf(sm3(a) * sm3(bb) * df3(cc), sm[2 * SB_PLANE + j], s);
          }
#pragma unroll
      for (int k = 0; k < SB_MAXC; ++k)
        if (k == c) acc[z][k] = s;
    }
  }
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx < g.Mw && gy < g.Mh) {
#pragma unroll
    for (int z = 0; z < SB_TZ; ++z) {
      const int gz = z0 + z;
      if (gz >= g.Md) continue;
      float* o = dpred + ((((long long)p * g.Md + gz) * g.Mh + gy) * g.Mw + gx) * g.ncls;
      o[0] = 0.f;
#pragma unroll
      for (int k = 0; k < SB_MAXC; ++k)
        if (k < nfg) o[k + 1] = acc[z][k];
    }
  }
}

static bool make_geo(int P, int Md, int Mh, int Mw, int ncls, int mode, bool over_input, SobelGeo& g) {
  if (P <= 0 || Md <= 2 || Mh <= 2 || Mw <= 2 || ncls < 2 || ncls - 1 > SB_MAXC || (mode != 0 && mode != 1)) return false;
  g.P = P; g.Md = Md; g.Mh = Mh; g.Mw = Mw; g.ncls = ncls; g.mode = mode;
  const int d = over_input ? Md : Md - 2, h = over_input ? Mh : Mh - 2, w = over_input ? Mw : Mw - 2;
  g.tz = (int)cdiv(d, SB_TZ); g.ty = (int)cdiv(h, SB_TY); g.tx = (int)cdiv(w, SB_TX);
  return true;
}

}  // namespace cfun

using namespace cfun;

extern "C" size_t cfun_sobel_edge_workspace_size(int P, int Md, int Mh, int Mw, int ncls, int mode) {
  if (P <= 0 || Md <= 2 || Mh <= 2 || Mw <= 2 || ncls < 2) return 256;
  const size_t per = (size_t)(Md - 2) * (Mh - 2) * (Mw - 2);
  return 256 + (size_t)P * per * (ncls - 1) * (mode ? 3 : 2) * sizeof(float);
}

extern "C" int cfun_sobel_edge_loss_fwd(const float* pred, const long long* tgt_index, int P, int Md, int Mh, int Mw, int ncls,
                                        int mode, float* loss, void* ws, size_t ws_bytes, void* stream) {
  SobelGeo g;
  CFUN_CHECK_ARG(pred && tgt_index && loss && ws && ws_bytes >= 256 && make_geo(P, Md, Mh, Mw, ncls, mode, false, g));
  cudaStream_t st = as_stream(stream);
  double* acc = reinterpret_cast<double*>(align_up((size_t)ws, 16));
  CFUN_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const size_t smem = (size_t)ncls * SB_PLANE * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CFUN_CUDA(cudaFuncSetAttribute(sobel_pass1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB_PLANE * 4));
    CFUN_CUDA(cudaFuncSetAttribute(sobel_pass1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB_PLANE * 4));
    attr = true;
  }
  const double per = (double)(Md - 2) * (Mh - 2) * (Mw - 2);
  const double denom = per * (double)P * (mode ? 3.0 : 1.0);
  sobel_pass1_kernel<false><<<(unsigned)((long long)P * g.tz * g.ty * g.tx), 256, smem, st>>>(pred, tgt_index, g, acc, nullptr, (float)denom, nullptr);
  CFUN_LAUNCH_CHECK();
  sobel_finalize_kernel<<<1, 1, 0, st>>>(acc, denom, loss);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_sobel_edge_loss_bwd(const float* pred, const long long* tgt_index, int P, int Md, int Mh, int Mw, int ncls,
                                        int mode, const float* grad_scale, float* dpred, void* ws, size_t ws_bytes, void* stream) {
  SobelGeo g, gi;
  CFUN_CHECK_ARG(pred && tgt_index && grad_scale && dpred && ws && make_geo(P, Md, Mh, Mw, ncls, mode, false, g) &&
                 make_geo(P, Md, Mh, Mw, ncls, mode, true, gi));
  if (ws_bytes < cfun_sobel_edge_workspace_size(P, Md, Mh, Mw, ncls, mode)) { set_error("sobel workspace too small"); return CFUN_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  float* A = reinterpret_cast<float*>(align_up((size_t)ws, 16) + 64);
  const size_t smem = (size_t)ncls * SB_PLANE * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CFUN_CUDA(cudaFuncSetAttribute(sobel_pass1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB_PLANE * 4));
    CFUN_CUDA(cudaFuncSetAttribute(sobel_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SB_PLANE * 4));
    attr = true;
  }
  const double per = (double)(Md - 2) * (Mh - 2) * (Mw - 2);
  const double denom = per * (double)P * (mode ? 3.0 : 1.0);
  sobel_pass1_kernel<true><<<(unsigned)((long long)P * g.tz * g.ty * g.tx), 256, smem, st>>>(pred, tgt_index, g, nullptr, grad_scale, (float)denom, A);
  CFUN_LAUNCH_CHECK();
  sobel_pass2_kernel<<<(unsigned)((long long)P * gi.tz * gi.ty * gi.tx), 256, (size_t)(mode ? 3 : 2) * SB_PLANE * sizeof(float), st>>>(A, gi, dpred);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
