// Inference pre / post-processing on the device: the two steps either side of predict('inference')
// that the reference runs on the host around a GPU round trip.
//   * cfun_resize_linear3d   -- utils.resize_image mode 'self' (reference utils.py:389-393): order-1 resize of the raw scan to
//     [IMAGE_MAX_DIM, IMAGE_MAX_DIM, IMAGE_MIN_DIM] as skimage >= 0.19 evaluates it (scipy.ndimage.zoom, order 1, mode
//     'grid-constant', grid_mode=True: source coordinate (o + 0.5) in/out - 0.5, neighbours outside the volume are 0),
//     float64 arithmetic in the same (H, W, D) separable order and with unfused multiplies / adds so that the cast back to
//     the scan's integer dtype (C truncation) lands on the same integer.
//   * cfun_unmold_mask_argmax -- utils.unmold_mask + np.argmax (reference utils.py:443-460, model.py:1851-1853): trilinear
//     (align_corners=False, PyTorch's source-index rule: negative coordinates clamp to 0) resize of the class-probability
//     crop to the detected box, pasted into a zero volume, argmax over classes -- fused, so the 8-channel full-size float
//     volume the reference materialises (8 x 320 x 320 x 192 x 4 B = 629 MB) never exists; output is the class-id volume in
//     the [H, W, D] order unmold_detections returns.
#include "common.cuh"

namespace cfun {

template <typename T> __device__ __forceinline__ T cast_trunc(double v);
template <> __device__ __forceinline__ short cast_trunc<short>(double v) { return (short)__double2int_rz(v); }
template <> __device__ __forceinline__ float cast_trunc<float>(double v) { return (float)v; }

// one axis of the separable order-1 zoom: index of the lower neighbour (may be -1 or n_in - 1) and the upper weight
__device__ __forceinline__ void zoom_coord(int o, int n_in, int n_out, int& i0, double& t) {
  const double cc = __dadd_rn(__dmul_rn((double)o + 0.5, (double)n_in / (double)n_out), -0.5);
  const double f = floor(cc);
  i0 = (int)f;
  t = __dadd_rn(cc, -f);
}

template <typename T>
__global__ void __launch_bounds__(256) resize_linear3d_kernel(const T* __restrict__ src, int H, int W, int D, T* __restrict__ dst,
                                                              int H2, int W2, int D2) {
  const long long total = (long long)H2 * W2 * D2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D2);
    const int w = (int)((i / D2) % W2);
    const int h = (int)(i / ((long long)D2 * W2));
    int h0, w0, d0;
    double th, tw, td;
    zoom_coord(h, H, H2, h0, th);
    zoom_coord(w, W, W2, w0, tw);
    zoom_coord(d, D, D2, d0, td);
    auto at = [&](int hh, int ww, int dd) -> double {
      if ((unsigned)hh >= (unsigned)H || (unsigned)ww >= (unsigned)W || (unsigned)dd >= (unsigned)D) return 0.0;
      return (double)src[((long long)hh * W + ww) * D + dd];
    };
    auto lerp = [](double a, double b, double t) { return __dadd_rn(__dmul_rn(a, __dadd_rn(1.0, -t)), __dmul_rn(b, t)); };
    // separable order of the restatement (H, then W, then D): H blend of the four (w, d) corners, W blend, D blend
    double c[2][2];
#pragma unroll
    for (int jw = 0; jw < 2; ++jw)
#pragma unroll
      for (int jd = 0; jd < 2; ++jd) c[jw][jd] = lerp(at(h0, w0 + jw, d0 + jd), at(h0 + 1, w0 + jw, d0 + jd), th);
    double vw[2];
    vw[0] = lerp(c[0][0], c[1][0], tw);
    vw[1] = lerp(c[0][1], c[1][1], tw);
    dst[i] = cast_trunc<T>(lerp(vw[0], vw[1], td));
  }
}

// PyTorch's upsample_trilinear3d source index (align_corners = False): scale * (o + 0.5) - 0.5, negative -> 0
__device__ __forceinline__ void tri_coord(int o, float scale, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float r = __fadd_rn(__fmul_rn(scale, __fadd_rn((float)o, 0.5f)), -0.5f);
  if (r < 0.f) r = 0.f;
  i0 = (int)r;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = fminf(fmaxf(__fadd_rn(r, -(float)i0), 0.f), 1.f);
  l0 = __fadd_rn(1.f, -l1);
}

// mask: class probabilities of ONE detection, element (c, z, y, x) at mask[c * cs + ((z * mh + y) * mw + x) * vs]
__global__ void __launch_bounds__(256) unmold_argmax_kernel(const float* __restrict__ mask, int ncls, int md, int mh, int mw,
                                                            long long cs, long long vs, int z1, int y1, int x1, int z2, int y2,
                                                            int x2, int D, int H, int W, unsigned char* __restrict__ out) {
  const long long total = (long long)H * W * D;
  const float sd = (float)md / (float)(z2 - z1), sh = (float)mh / (float)(y2 - y1), sw = (float)mw / (float)(x2 - x1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % D);                       // output order [H, W, D]
    const int x = (int)((i / D) % W);
    const int y = (int)(i / ((long long)D * W));
    unsigned char best = 0;
    if (z >= z1 && z < z2 && y >= y1 && y < y2 && x >= x1 && x < x2) {
      int a0, a1, b0, b1, c0, c1;
      float la0, la1, lb0, lb1, lc0, lc1;
      tri_coord(z - z1, sd, md, a0, a1, la0, la1);
      tri_coord(y - y1, sh, mh, b0, b1, lb0, lb1);
      tri_coord(x - x1, sw, mw, c0, c1, lc0, lc1);
      float bv = 0.f;
      for (int c = 0; c < ncls; ++c) {
        const float* m = mask + (long long)c * cs;
        auto v = [&](int zz, int yy, int xx) { return __ldg(m + (((long long)zz * mh + yy) * mw + xx) * vs); };
        auto row = [&](int zz, int yy) { return __fadd_rn(__fmul_rn(lc0, v(zz, yy, c0)), __fmul_rn(lc1, v(zz, yy, c1))); };
        auto pl = [&](int zz) { return __fadd_rn(__fmul_rn(lb0, row(zz, b0)), __fmul_rn(lb1, row(zz, b1))); };
        const float val = __fadd_rn(__fmul_rn(la0, pl(a0)), __fmul_rn(la1, pl(a1)));
        if (c == 0 || val > bv) { bv = val; best = (unsigned char)c; }      // np.argmax: first maximum wins
      }
      // outside the box the reference's volume is all zeros (argmax 0); inside, values are the interpolated probabilities
    }
    out[i] = best;
  }
}

}  // namespace cfun

using namespace cfun;

extern "C" int cfun_resize_linear3d(const void* src, int H, int W, int D, void* dst, int H2, int W2, int D2, int dtype, void* stream) {
  CFUN_CHECK_ARG(src && dst && H > 0 && W > 0 && D > 0 && H2 > 0 && W2 > 0 && D2 > 0 && (dtype == 0 || dtype == 1));
  const long long total = (long long)H2 * W2 * D2;
  const unsigned grid = (unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms());
  if (dtype == 0) resize_linear3d_kernel<short><<<grid, 256, 0, as_stream(stream)>>>((const short*)src, H, W, D, (short*)dst, H2, W2, D2);
  else resize_linear3d_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, H, W, D, (float*)dst, H2, W2, D2);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_unmold_mask_argmax(const float* mask, int ncls, int md, int mh, int mw, long long class_stride,
                                       long long voxel_stride, const int* box6_host, int D, int H, int W, unsigned char* out,
                                       void* stream) {
  CFUN_CHECK_ARG(mask && box6_host && out && ncls > 0 && ncls <= 255 && md > 0 && mh > 0 && mw > 0 && D > 0 && H > 0 && W > 0);
  const int z1 = box6_host[0], y1 = box6_host[1], x1 = box6_host[2], z2 = box6_host[3], y2 = box6_host[4], x2 = box6_host[5];
  CFUN_CHECK_ARG(z2 > z1 && y2 > y1 && x2 > x1 && z1 >= 0 && y1 >= 0 && x1 >= 0 && z2 <= D && y2 <= H && x2 <= W);
  const long long total = (long long)H * W * D;
  unmold_argmax_kernel<<<(unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()), 256, 0, as_stream(stream)>>>(
      mask, ncls, md, mh, mw, class_stride, voxel_stride, z1, y1, x1, z2, y2, x2, D, H, W, out);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
