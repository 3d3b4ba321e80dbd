// 3-D RoI crop + trilinear (align_corners=True) resize, forward and backward.
// Replaces model.RoI_Align (model.py:265-289): denormalise by the feature-map size, floor the lower / ceil the upper
// corner, python-slice crop, F.interpolate(mode='trilinear', align_corners=True); empty crops leave zero rows (the
// reference's bare `except`).  One launch covers every box of both pyramid levels (model.pyramid_roi_align,
// model.py:334-368, runs one Python pass per level and one interpolate call per box).
#include "common.cuh"

namespace cfun {

struct RoiMaps {
  const float* f[2];
  float* df[2];
  int D[2], H[2], W[2];
};

struct Crop {
  int z1, y1, x1, cd, ch, cw;
};

__device__ __forceinline__ int slice_bound(float v, int size) {
  long long t = (long long)v;  // .long() truncation of an already floored / ceiled float
  if (t < 0) t += size;
  if (t < 0) t = 0;
  if (t > size) t = size;
  return (int)t;
}

__device__ __forceinline__ Crop crop_of(const float* box, int D, int H, int W) {
  // utils.denorm_boxes_graph (utils.py:160-174): boxes * [D,H,W,D,H,W] in fp32; model.py:272-278 floor / ceil / long
  Crop c;
  int z1 = slice_bound(floorf(__fmul_rn(box[0], (float)D)), D), z2 = slice_bound(ceilf(__fmul_rn(box[3], (float)D)), D);
  int y1 = slice_bound(floorf(__fmul_rn(box[1], (float)H)), H), y2 = slice_bound(ceilf(__fmul_rn(box[4], (float)H)), H);
  int x1 = slice_bound(floorf(__fmul_rn(box[2], (float)W)), W), x2 = slice_bound(ceilf(__fmul_rn(box[5], (float)W)), W);
  c.z1 = z1; c.y1 = y1; c.x1 = x1;
  c.cd = z2 - z1; c.ch = y2 - y1; c.cw = x2 - x1;
  return c;
}

// align_corners=True source index / weights (ATen UpSample.h area_pixel_compute_scale / source_index)
__device__ __forceinline__ void lin_coord(int o, int in, int out, int& i0, int& i1, float& l0, float& l1) {
  float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  float src = scale * (float)o;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  l0 = 1.f - l1;
}

template <bool BWD>
__global__ void __launch_bounds__(128) roi_kernel(RoiMaps maps, int C, const float* __restrict__ boxes,
                                                  const int* __restrict__ level, int n, int pd, int ph, int pw,
                                                  float* __restrict__ out, const float* __restrict__ dout, int out_ncdhw) {
  const long long per = (long long)pd * ph * pw;
  const int b = blockIdx.y;
  const int lv = level ? level[b] : 0;
  const int D = maps.D[lv], H = maps.H[lv], W = maps.W[lv];
  const Crop cr = crop_of(boxes + (long long)b * 6, D, H, W);
  const bool empty = cr.cd <= 0 || cr.ch <= 0 || cr.cw <= 0;
  const float* __restrict__ f = maps.f[lv];
  float* __restrict__ df = maps.df[lv];
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < per; v += (long long)gridDim.x * blockDim.x) {
    int ox = (int)(v % pw);
    int oy = (int)((v / pw) % ph);
    int oz = (int)(v / ((long long)pw * ph));
    long long obase = out_ncdhw ? ((long long)b * C) * per + v : ((long long)b * per + v) * C;
    long long ostride = out_ncdhw ? per : 1;
    if (empty) {
      if (!BWD)
        for (int c = 0; c < C; ++c) out[obase + c * ostride] = 0.f;
      continue;
    }
    int z0, z1, y0, y1, x0, x1;
    float wz0, wz1, wy0, wy1, wx0, wx1;
    lin_coord(oz, cr.cd, pd, z0, z1, wz0, wz1);
    lin_coord(oy, cr.ch, ph, y0, y1, wy0, wy1);
    lin_coord(ox, cr.cw, pw, x0, x1, wx0, wx1);
    z0 += cr.z1; z1 += cr.z1; y0 += cr.y1; y1 += cr.y1; x0 += cr.x1; x1 += cr.x1;
    const long long a000 = (((long long)z0 * H + y0) * W + x0) * C, a001 = (((long long)z0 * H + y0) * W + x1) * C;
    const long long a010 = (((long long)z0 * H + y1) * W + x0) * C, a011 = (((long long)z0 * H + y1) * W + x1) * C;
    const long long a100 = (((long long)z1 * H + y0) * W + x0) * C, a101 = (((long long)z1 * H + y0) * W + x1) * C;
    const long long a110 = (((long long)z1 * H + y1) * W + x0) * C, a111 = (((long long)z1 * H + y1) * W + x1) * C;
    if (!BWD) {
      for (int c = 0; c < C; ++c) {
        float v00 = wx0 * __ldg(f + a000 + c) + wx1 * __ldg(f + a001 + c);
        float v01 = wx0 * __ldg(f + a010 + c) + wx1 * __ldg(f + a011 + c);
        float v10 = wx0 * __ldg(f + a100 + c) + wx1 * __ldg(f + a101 + c);
        float v11 = wx0 * __ldg(f + a110 + c) + wx1 * __ldg(f + a111 + c);
        out[obase + c * ostride] = wz0 * (wy0 * v00 + wy1 * v01) + wz1 * (wy0 * v10 + wy1 * v11);
      }
    } else {
      for (int c = 0; c < C; ++c) {
        float g = __ldg(dout + obase + c * ostride);
        atomicAdd(df + a000 + c, g * wz0 * wy0 * wx0);
        atomicAdd(df + a001 + c, g * wz0 * wy0 * wx1);
        atomicAdd(df + a010 + c, g * wz0 * wy1 * wx0);
        atomicAdd(df + a011 + c, g * wz0 * wy1 * wx1);
        atomicAdd(df + a100 + c, g * wz1 * wy0 * wx0);
        atomicAdd(df + a101 + c, g * wz1 * wy0 * wx1);
        atomicAdd(df + a110 + c, g * wz1 * wy1 * wx0);
        atomicAdd(df + a111 + c, g * wz1 * wy1 * wx1);
      }
    }
  }
}

}  // namespace cfun

using namespace cfun;

static int roi_launch(bool bwd, const float* f0, float* df0, int D0, int H0, int W0, const float* f1, float* df1, int D1,
                      int H1, int W1, int C, const float* boxes, const int* level, int n, int pd, int ph, int pw, float* out,
                      const float* dout, int out_ncdhw, void* stream) {
  CFUN_CHECK_ARG(n >= 0 && C > 0 && pd > 0 && ph > 0 && pw > 0);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(boxes && D0 > 0 && H0 > 0 && W0 > 0);
  CFUN_CHECK_ARG(bwd ? (df0 && dout) : (f0 && out));
  CFUN_CHECK_ARG(!level || (bwd ? df1 != nullptr : f1 != nullptr));
  RoiMaps m;
  m.f[0] = f0; m.f[1] = f1 ? f1 : f0;
  m.df[0] = df0; m.df[1] = df1 ? df1 : df0;
  m.D[0] = D0; m.H[0] = H0; m.W[0] = W0;
  m.D[1] = f1 || df1 ? D1 : D0; m.H[1] = f1 || df1 ? H1 : H0; m.W[1] = f1 || df1 ? W1 : W0;
  long long per = (long long)pd * ph * pw;
  dim3 grid((unsigned)std::min<long long>(cdiv(per, 128), 4096), (unsigned)n);
  if (bwd) roi_kernel<true><<<grid, 128, 0, as_stream(stream)>>>(m, C, boxes, level, n, pd, ph, pw, nullptr, dout, out_ncdhw);
  else roi_kernel<false><<<grid, 128, 0, as_stream(stream)>>>(m, C, boxes, level, n, pd, ph, pw, out, nullptr, out_ncdhw);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_roi_crop_resize_fwd(const float* fmap0, int D0, int H0, int W0, const float* fmap1, int D1, int H1, int W1,
                                        int C, const float* boxes, const int* level, int n, int pd, int ph, int pw,
                                        float* out, int out_ncdhw, void* stream) {
  return roi_launch(false, fmap0, nullptr, D0, H0, W0, fmap1, nullptr, D1, H1, W1, C, boxes, level, n, pd, ph, pw, out,
                    nullptr, out_ncdhw, stream);
}

extern "C" int cfun_roi_crop_resize_bwd(float* dfmap0, int D0, int H0, int W0, float* dfmap1, int D1, int H1, int W1, int C,
                                        const float* boxes, const int* level, int n, int pd, int ph, int pw,
                                        const float* dout, int out_ncdhw, void* stream) {
  return roi_launch(true, nullptr, dfmap0, D0, H0, W0, nullptr, dfmap1, D1, H1, W1, C, boxes, level, n, pd, ph, pw, nullptr,
                    dout, out_ncdhw, stream);
}
