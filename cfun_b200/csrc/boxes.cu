// Box pipeline on device: bitonic score sort, anchor decode + clip, bit-exact greedy 3-D NMS, IoU overlaps,
// box refinement targets, GT-mask crop + nearest resize, RoI pyramid level.
// Replaces model.proposal_layer (model.py:199-258), apply_box_deltas (:155-182), clip_boxes (:185-196),
// utils.non_max_suppression / compute_iou (utils.py:122-157, 50-70), bbox_overlaps (model.py:377-411),
// utils.box_refinement (utils.py:92-119), the mask-target loop of detection_target_layer (model.py:481-493)
// and the level rule of pyramid_roi_align (model.py:322-332).
//
// All box arithmetic that decides an *index* (IoU vs threshold) is written with explicit round-to-nearest
// intrinsics so nvcc cannot contract it into FMAs: the reference computes it with unfused fp32 numpy ops.
#include "common.cuh"
#include <climits>

namespace cfun {

// ---------------------------------------------------------------------------------------------------------
// bitonic sort, order = (key descending, index ascending)
// ---------------------------------------------------------------------------------------------------------
constexpr int SORT_CHUNK = 2048;

__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) { return ka > kb || (ka == kb && ia < ib); }

__global__ void sort_init_kernel(const float* __restrict__ scores, int n, int P, float* __restrict__ keys, int* __restrict__ idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    keys[i] = i < n ? scores[i] : -INFINITY;
    idx[i] = i < n ? i : INT_MAX;
  }
}

// runs the bitonic sub-stages j = jstart .. 1 of stage k for one SORT_CHUNK-sized chunk held in shared memory;
// when full != 0 runs all stages k = 2 .. kmax (kmax <= SORT_CHUNK) instead
__global__ void __launch_bounds__(1024) sort_local_kernel(float* __restrict__ keys, int* __restrict__ idx, int P, int k_in,
                                                          int jstart, int full) {
  __shared__ float sk[SORT_CHUNK];
  __shared__ int si[SORT_CHUNK];
  const int base = blockIdx.x * SORT_CHUNK;
  const int len = min(SORT_CHUNK, P - base);
  for (int t = threadIdx.x; t < len; t += blockDim.x) { sk[t] = keys[base + t]; si[t] = idx[base + t]; }
  __syncthreads();
  const int kbeg = full ? 2 : k_in, kend = full ? min(P, SORT_CHUNK) : k_in;
  for (int k = kbeg; k <= kend; k <<= 1) {
    for (int j = full ? (k >> 1) : jstart; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < len / 2; t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j cleared
        int p = i | j;
        bool asc = (((base + i) & k) == 0);
        float ka = sk[i], kb = sk[p];
        int ia = si[i], ib = si[p];
        bool swap = asc ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) { sk[i] = kb; sk[p] = ka; si[i] = ib; si[p] = ia; }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < len; t += blockDim.x) { keys[base + t] = sk[t]; idx[base + t] = si[t]; }
}

__global__ void sort_global_kernel(float* __restrict__ keys, int* __restrict__ idx, int P, int k, int j) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P / 2; t += gridDim.x * blockDim.x) {
    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    int p = i | j;
    bool asc = ((i & k) == 0);
    float ka = keys[i], kb = keys[p];
    int ia = idx[i], ib = idx[p];
    bool swap = asc ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
    if (swap) { keys[i] = kb; keys[p] = ka; idx[i] = ib; idx[p] = ia; }
  }
}

__global__ void copy_int_kernel(const int* __restrict__ src, int* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

static int next_pow2(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

// ---------------------------------------------------------------------------------------------------------
// decode + clip   (model.py:155-196, 216-241)
// ---------------------------------------------------------------------------------------------------------
struct F6 { float v[6]; };

__global__ void decode_clip_kernel(const float* __restrict__ anchors, const float* __restrict__ deltas,
                                   const float* __restrict__ scores, int sstride, int soff, const int* __restrict__ order,
                                   int k, F6 sd, F6 win, float* __restrict__ boxes_out, float* __restrict__ scores_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  int o = order ? order[i] : i;
  float a[6], d[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    a[j] = anchors[(long long)o * 6 + j];
    d[j] = __fmul_rn(deltas[(long long)o * 6 + j], sd.v[j]);
  }
  float sz[3], ct[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    sz[j] = __fsub_rn(a[3 + j], a[j]);
    ct[j] = __fadd_rn(a[j], __fmul_rn(0.5f, sz[j]));
    ct[j] = __fadd_rn(ct[j], __fmul_rn(d[j], sz[j]));
    sz[j] = __fmul_rn(sz[j], expf(d[3 + j]));
  }
  float b[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    b[j] = __fsub_rn(ct[j], __fmul_rn(0.5f, sz[j]));
    b[3 + j] = __fadd_rn(b[j], sz[j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float lo = win.v[j % 3], hi = win.v[3 + j % 3];
    boxes_out[(long long)i * 6 + j] = fminf(fmaxf(b[j], lo), hi);
  }
  if (scores_out) scores_out[i] = scores[(long long)o * sstride + soff];
}

// ---------------------------------------------------------------------------------------------------------
// NMS   (utils.py:50-70, 122-157)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float vol_rn(const float* b) {
  return __fmul_rn(__fmul_rn(__fsub_rn(b[3], b[0]), __fsub_rn(b[4], b[1])), __fsub_rn(b[5], b[2]));
}
__device__ __forceinline__ float iou_rn(const float* a, float va, const float* b, float vb) {
  float z1 = fmaxf(a[0], b[0]), z2 = fminf(a[3], b[3]);
  float y1 = fmaxf(a[1], b[1]), y2 = fminf(a[4], b[4]);
  float x1 = fmaxf(a[2], b[2]), x2 = fminf(a[5], b[5]);
  float inter = __fmul_rn(__fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f)), fmaxf(__fsub_rn(z2, z1), 0.f));
  float uni = __fsub_rn(__fadd_rn(va, vb), inter);
  return __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));
}

// mask[i * nblk + jb] bit b : iou(i, jb*64+b) > thr  (only j > i matter)
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, int n, float thr, int nblk,
                                                      unsigned long long* __restrict__ mask) {
  const int ib = blockIdx.y, jb = blockIdx.x;
  if (jb < ib) return;
  __shared__ float sb[64][7];
  const int t = threadIdx.x;
  const int j = jb * 64 + t;
  if (j < n) {
#pragma unroll
    for (int c = 0; c < 6; ++c) sb[t][c] = boxes[(long long)j * 6 + c];
    sb[t][6] = vol_rn(sb[t]);
  }
  __syncthreads();
  const int i = ib * 64 + t;
  if (i >= n) return;
  float a[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) a[c] = boxes[(long long)i * 6 + c];
  const float va = vol_rn(a);
  unsigned long long bits = 0;
  const int lim = min(64, n - jb * 64);
  for (int b = 0; b < lim; ++b) {
    int jj = jb * 64 + b;
    if (jj <= i) continue;
    if (iou_rn(a, va, sb[b], sb[b][6]) > thr) bits |= 1ull << b;
  }
  mask[(long long)i * nblk + jb] = bits;
}

__global__ void __launch_bounds__(1024) nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int nblk,
                                                        int max_num, int* __restrict__ keep, int* __restrict__ count) {
  extern __shared__ unsigned long long removed[];  // [nblk]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long kept_bits;
  __shared__ int cnt;
  __shared__ int done;
  const int t = threadIdx.x;
  for (int w = t; w < nblk; w += blockDim.x) removed[w] = 0;
  if (t == 0) { cnt = 0; done = 0; }
  __syncthreads();
  for (int blk = 0; blk < nblk; ++blk) {
    const int base = blk * 64;
    const int lim = min(64, n - base);
    if (t < 64) diag[t] = t < lim ? mask[(long long)(base + t) * nblk + blk] : 0ull;
    __syncthreads();
    if (t == 0) {
      unsigned long long rem = removed[blk], kb = 0;
      int c = cnt;
      for (int b = 0; b < lim; ++b) {
        if (rem >> b & 1ull) continue;
        keep[c++] = base + b;
        kb |= 1ull << b;
        if (c >= max_num) { done = 1; break; }
        rem |= diag[b];
      }
      cnt = c;
      kept_bits = kb;
    }
    __syncthreads();
    if (done) break;
    const unsigned long long kb = kept_bits;
    for (int w = blk + 1 + t; w < nblk; w += blockDim.x) {
      unsigned long long acc = removed[w];
      unsigned long long bits = kb;
      while (bits) {
        int b = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= mask[(long long)(base + b) * nblk + w];
      }
      removed[w] = acc;
    }
    __syncthreads();
  }
  if (t == 0) *count = cnt;
}

__global__ void gather_boxes_kernel(const float* __restrict__ rows, const int* __restrict__ idx, const int* __restrict__ count,
                                    int max_rows, F6 div, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_rows) return;
  int c = *count;
#pragma unroll
  for (int j = 0; j < 6; ++j) out[(long long)i * 6 + j] = i < c ? __fdiv_rn(rows[(long long)idx[i] * 6 + j], div.v[j]) : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// overlaps / refinement / levels / mask targets
// ---------------------------------------------------------------------------------------------------------
// IoU of model.bbox_overlaps (model.py:374-410): no epsilon, 0/0 = NaN for two empty boxes
__device__ __forceinline__ float overlap_rn(const float* a, const float* b) {
  float z1 = fmaxf(a[0], b[0]), y1 = fmaxf(a[1], b[1]), x1 = fmaxf(a[2], b[2]);
  float z2 = fminf(a[3], b[3]), y2 = fminf(a[4], b[4]), x2 = fminf(a[5], b[5]);
  float inter = __fmul_rn(__fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f)), fmaxf(__fsub_rn(z2, z1), 0.f));
  float uni = __fsub_rn(__fadd_rn(vol_rn(a), vol_rn(b)), inter);
  return __fdiv_rn(inter, uni);
}

__global__ void overlaps_kernel(const float* __restrict__ b1, int n1, const float* __restrict__ b2, int n2, float* __restrict__ iou) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2) return;
  int i = t / n2, j = t % n2;
  float a[6], b[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { a[c] = b1[(long long)i * 6 + c]; b[c] = b2[(long long)j * 6 + c]; }
  iou[t] = overlap_rn(a, b);
}

// utils.compute_iou (utils.py:50-70): one box against many, with the +1e-6 of the NMS IoU
__global__ void iou_eps_kernel(const float* __restrict__ box, const float* __restrict__ boxes, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[6], b[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { a[c] = box[c]; b[c] = boxes[(long long)i * 6 + c]; }
  out[i] = iou_rn(a, vol_rn(a), b, vol_rn(b));
}

// utils.box_refinement (utils.py:101-128) of one box against its ground-truth box
__device__ __forceinline__ void refinement_rn(const float* box, const float* gt, const F6& sd, float* out) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float s = __fsub_rn(box[3 + j], box[j]);
    float c = __fadd_rn(box[j], __fmul_rn(0.5f, s));
    float gs = __fsub_rn(gt[3 + j], gt[j]);
    float gc = __fadd_rn(gt[j], __fmul_rn(0.5f, gs));
    out[j] = __fdiv_rn(__fdiv_rn(__fsub_rn(gc, c), s), sd.v[j]);
    out[3 + j] = __fdiv_rn(logf(__fdiv_rn(gs, s)), sd.v[3 + j]);
  }
}

__global__ void refinement_kernel(const float* __restrict__ box, const float* __restrict__ gt, int n, F6 sd, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  refinement_rn(box + i * 6, gt + i * 6, sd, out + i * 6);
}

// ---------------------------------------------------------------------------------------------------------
// detection targets without host round trips (model.detection_target_layer, model.py:414-563, split at its one
// data-dependent point: the two torch.randperm draws on the host generator need the candidate counts)
// ---------------------------------------------------------------------------------------------------------
// Part A, one block: normalised proposals (= gather_boxes), per-proposal max IoU / argmax over the ground-truth boxes
// (= bbox_overlaps + max), and the ascending index lists of the positive (IoU >= thr) and negative (IoU < thr) proposals
// (= the two torch.nonzero calls), with their lengths.  A NaN IoU is in neither list, as with the reference's comparisons.
__global__ void __launch_bounds__(1024) roi_candidates_kernel(const float* __restrict__ boxes, const int* __restrict__ keep,
                                                              const int* __restrict__ count, int max_rows, F6 div,
                                                              const float* __restrict__ gt, int G, float thr, float* __restrict__ rois,
                                                              float* __restrict__ iou_max, int* __restrict__ assign,
                                                              int* __restrict__ pos_list, int* __restrict__ neg_list,
                                                              int* __restrict__ counts) {
  __shared__ int s_warp[32];
  __shared__ int s_base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(max(*count, 0), max_rows);
  if (tid == 0) { s_base[0] = 0; s_base[1] = 0; }
  __syncthreads();
  for (int base = 0; base < max_rows; base += 1024) {
    const int i = base + tid;
    int flag = 0;
    if (i < max_rows) {
      float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float mx = 0.f;
      int arg = 0;
      if (i < n) {
        const int src = keep[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) r[j] = __fdiv_rn(boxes[(long long)src * 6 + j], div.v[j]);
        bool nan = false;
        mx = -INFINITY;
        for (int g = 0; g < G; ++g) {
          float b[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) b[j] = gt[g * 6 + j];
          const float v = overlap_rn(r, b);
          if (v != v) nan = true;
          if (v > mx) { mx = v; arg = g; }      // first occurrence of the maximum
        }
        if (nan) mx = __int_as_float(0x7fc00000);
        flag = (mx >= thr) ? 1 : ((mx < thr) ? 2 : 0);
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) rois[(long long)i * 6 + j] = r[j];
      iou_max[i] = mx;
      assign[i] = arg;
    }
    // block-wide exclusive scan of (positive flag | negative flag << 16)
    const int v = (flag == 1 ? 1 : 0) | (flag == 2 ? (1 << 16) : 0);
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;                     // inclusive over warps
    }
    __syncthreads();
    const int excl = incl - v + (warp ? s_warp[warp - 1] : 0);
    if (flag == 1) pos_list[s_base[0] + (excl & 0xffff)] = i;
    if (flag == 2) neg_list[s_base[1] + (excl >> 16)] = i;
    const int total = s_warp[31];
    __syncthreads();
    if (tid == 0) { s_base[0] += total & 0xffff; s_base[1] += total >> 16; }
    __syncthreads();
  }
  if (tid == 0) { counts[0] = n; counts[1] = s_base[0]; counts[2] = s_base[1]; }
}

// Part B: rows 0..P-1 = positives pos_list[perm[i]], rows P..R-1 = negatives neg_list[perm[i]]; class id and box deltas of the
// matched ground-truth box for the positives, zeros for the negatives (model.py:457-470, 545-560)
__global__ void roi_targets_kernel(const float* __restrict__ rois, const int* __restrict__ assign, const int* __restrict__ pos_list,
                                   const int* __restrict__ neg_list, const long long* __restrict__ perm, int P, int R,
                                   const float* __restrict__ gt, const int* __restrict__ gt_cls, F6 sd, float* __restrict__ out_rois,
                                   long long* __restrict__ cls, float* __restrict__ deltas) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const int src = i < P ? pos_list[perm[i]] : neg_list[perm[i]];
  float r[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) { r[j] = rois[(long long)src * 6 + j]; out_rois[i * 6 + j] = r[j]; }
  float d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  long long c = 0;
  if (i < P) {
    const int g = assign[src];
    float b[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) b[j] = gt[g * 6 + j];
    refinement_rn(r, b, sd, d);
    c = gt_cls[g];
  }
  cls[i] = c;
#pragma unroll
  for (int j = 0; j < 6; ++j) deltas[i * 6 + j] = d[j];
}

__global__ void roi_level_kernel(const float* __restrict__ boxes, int n, int* __restrict__ level) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = __fsub_rn(boxes[i * 6 + 3], boxes[i * 6 + 0]);
  float h = __fsub_rn(boxes[i * 6 + 4], boxes[i * 6 + 1]);
  float w = __fsub_rn(boxes[i * 6 + 5], boxes[i * 6 + 2]);
  float v = __fmul_rn(__fmul_rn(h, w), d);
  float l2 = __fdiv_rn(logf(v), logf(2.0f));
  float lv = __fadd_rn(4.f, __fmul_rn((float)(1. / 3.), l2));
  lv = rintf(lv);
  int li;
  if (lv != lv) li = INT_MIN;            // torch .int() of NaN / -inf lands on INT_MIN on x86
  else if (lv <= -2147483648.f) li = INT_MIN;
  else if (lv >= 2147483648.f) li = INT_MIN;
  else li = (int)lv;
  li = min(max(li, 2), 3);
  level[i] = li - 2;
}

__device__ __forceinline__ int py_slice_bound(long long v, int size) {
  if (v < 0) v += size;
  if (v < 0) v = 0;
  if (v > size) v = size;
  return (int)v;
}

__global__ void mask_target_kernel(const int* __restrict__ label, int D, int H, int W, const float* __restrict__ rois, int P,
                                   int ncls, int md, int mh, int mw, double* __restrict__ onehot,
                                   long long* __restrict__ cls_index) {
  const long long per = (long long)md * mh * mw;
  const long long total = (long long)P * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i / per);
    long long r = i % per;
    int ox = (int)(r % mw); r /= mw;
    int oy = (int)(r % mh);
    int oz = (int)(r / mh);
    const float* roi = rois + p * 6;
    // model.py:483-488: int(dim * coord), fp32 product, truncation
    int z1 = py_slice_bound((long long)__fmul_rn((float)D, roi[0]), D), z2 = py_slice_bound((long long)__fmul_rn((float)D, roi[3]), D);
    int y1 = py_slice_bound((long long)__fmul_rn((float)H, roi[1]), H), y2 = py_slice_bound((long long)__fmul_rn((float)H, roi[4]), H);
    int x1 = py_slice_bound((long long)__fmul_rn((float)W, roi[2]), W), x2 = py_slice_bound((long long)__fmul_rn((float)W, roi[5]), W);
    int cd = z2 - z1, ch = y2 - y1, cw = x2 - x1;
    int cls = -1;
    if (cd > 0 && ch > 0 && cw > 0) {
      // half-pixel-centre nearest neighbour (skimage>=0.19 order-0 resize == ndi.zoom grid_mode)
      int sz = (int)floor(((double)oz + 0.5) * ((double)cd / (double)md));
      int sy = (int)floor(((double)oy + 0.5) * ((double)ch / (double)mh));
      int sx = (int)floor(((double)ox + 0.5) * ((double)cw / (double)mw));
      sz = min(max(sz, 0), cd - 1); sy = min(max(sy, 0), ch - 1); sx = min(max(sx, 0), cw - 1);
      cls = label[((long long)(z1 + sz) * H + (y1 + sy)) * W + (x1 + sx)];
    }
    if (cls_index) cls_index[i] = (cls >= 0 && cls < ncls) ? cls : 0;
    if (onehot) {
      long long r2 = i % per;
      for (int c = 0; c < ncls; ++c) onehot[((long long)p * ncls + c) * per + r2] = (c == cls) ? 1.0 : 0.0;
    }
  }
}

}  // namespace cfun

using namespace cfun;

extern "C" size_t cfun_sort_workspace_size(int n) {
  if (n <= 0) return 256;
  size_t P = (size_t)next_pow2(n);
  return P * 8 + 512;
}

extern "C" int cfun_sort_desc(const float* scores, int n, int* order, void* ws, size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(scores && order && ws && n > 0);
  if (ws_bytes < cfun_sort_workspace_size(n)) { set_error("sort workspace too small"); return CFUN_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int P = next_pow2(n);
  float* keys = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  int* idx = reinterpret_cast<int*>(keys + P);
  sort_init_kernel<<<(unsigned)std::min<long long>(cdiv(P, 256), 1024), 256, 0, st>>>(scores, n, P, keys, idx);
  CFUN_LAUNCH_CHECK();
  const int nchunks = (P + SORT_CHUNK - 1) / SORT_CHUNK;
  sort_local_kernel<<<nchunks, 1024, 0, st>>>(keys, idx, P, 0, 0, 1);
  CFUN_LAUNCH_CHECK();
  for (int k = SORT_CHUNK * 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
      sort_global_kernel<<<(unsigned)std::min<long long>(cdiv(P / 2, 256), 2048), 256, 0, st>>>(keys, idx, P, k, j);
      CFUN_LAUNCH_CHECK();
    }
    sort_local_kernel<<<nchunks, 1024, 0, st>>>(keys, idx, P, k, SORT_CHUNK / 2, 0);
    CFUN_LAUNCH_CHECK();
  }
  copy_int_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(idx, order, n);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

static F6 to_f6(const float* p, float dflt) {
  F6 f;
  for (int i = 0; i < 6; ++i) f.v[i] = p ? p[i] : dflt;
  return f;
}

extern "C" int cfun_decode_clip(const float* anchors, const float* deltas, const float* scores, int sstride, int soff,
                                const int* order, int k, const float* std6_host, const float* window6_host,
                                float* boxes_out, float* scores_out, void* stream) {
  CFUN_CHECK_ARG(anchors && deltas && boxes_out && window6_host && k >= 0);
  CFUN_CHECK_ARG(!scores_out || scores);
  if (k == 0) return CFUN_OK;
  decode_clip_kernel<<<(unsigned)cdiv(k, 128), 128, 0, as_stream(stream)>>>(anchors, deltas, scores, sstride, soff, order, k,
                                                                            to_f6(std6_host, 1.f), to_f6(window6_host, 0.f),
                                                                            boxes_out, scores_out);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" size_t cfun_nms_workspace_size(int n) {
  if (n <= 0) return 256;
  size_t nblk = ((size_t)n + 63) / 64;
  return (size_t)n * nblk * 8 + 512;
}

extern "C" int cfun_nms3d(const float* boxes, int n, float threshold, int max_num, int* keep, int* count, void* ws,
                          size_t ws_bytes, void* stream) {
  CFUN_CHECK_ARG(keep && count && n >= 0 && max_num >= 1);
  cudaStream_t st = as_stream(stream);
  if (n == 0) { CFUN_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st)); return CFUN_OK; }
  CFUN_CHECK_ARG(boxes && ws);
  if (ws_bytes < cfun_nms_workspace_size(n)) { set_error("nms workspace too small"); return CFUN_ERR_WORKSPACE; }
  const int nblk = (n + 63) / 64;
  CFUN_CHECK_ARG((size_t)nblk * 8 <= 200 * 1024);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(align_up((size_t)ws, 256));
  nms_mask_kernel<<<dim3(nblk, nblk), 64, 0, st>>>(boxes, n, threshold, nblk, mask);
  CFUN_LAUNCH_CHECK();
  size_t smem = (size_t)nblk * 8;
  if (smem > 48 * 1024) CFUN_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<1, 1024, smem, st>>>(mask, n, nblk, max_num, keep, count);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_gather_boxes(const float* rows, const int* idx, const int* count, int max_rows,
                                 const float* div6_host, float* rows_out, void* stream) {
  CFUN_CHECK_ARG(rows && idx && count && rows_out && max_rows > 0);
  gather_boxes_kernel<<<(unsigned)cdiv(max_rows, 128), 128, 0, as_stream(stream)>>>(rows, idx, count, max_rows, to_f6(div6_host, 1.f), rows_out);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_bbox_overlaps3d(const float* boxes1, int n1, const float* boxes2, int n2, float* iou, void* stream) {
  CFUN_CHECK_ARG(n1 >= 0 && n2 >= 0);
  if (n1 == 0 || n2 == 0) return CFUN_OK;
  CFUN_CHECK_ARG(boxes1 && boxes2 && iou);
  overlaps_kernel<<<(unsigned)cdiv((long long)n1 * n2, 256), 256, 0, as_stream(stream)>>>(boxes1, n1, boxes2, n2, iou);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_roi_candidates(const float* boxes_sorted, const int* keep, const int* count, int max_rows, const float* div6_host,
                                   const float* gt_boxes, int n_gt, float iou_threshold, float* rois, float* iou_max, int* assign,
                                   int* pos_list, int* neg_list, int* counts3, void* stream) {
  CFUN_CHECK_ARG(boxes_sorted && keep && count && gt_boxes && rois && iou_max && assign && pos_list && neg_list && counts3);
  CFUN_CHECK_ARG(max_rows > 0 && max_rows <= (1 << 20) && n_gt > 0);
  roi_candidates_kernel<<<1, 1024, 0, as_stream(stream)>>>(boxes_sorted, keep, count, max_rows, to_f6(div6_host, 1.f), gt_boxes, n_gt,
                                                          iou_threshold, rois, iou_max, assign, pos_list, neg_list, counts3);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_roi_targets(const float* rois, const int* assign, const int* pos_list, const int* neg_list, const long long* perm,
                                int P, int R, const float* gt_boxes, const int* gt_class_ids, const float* std6_host, float* out_rois,
                                long long* class_ids, float* deltas, void* stream) {
  CFUN_CHECK_ARG(P >= 0 && R >= P);
  if (R == 0) return CFUN_OK;
  CFUN_CHECK_ARG(rois && assign && pos_list && neg_list && perm && gt_boxes && gt_class_ids && out_rois && class_ids && deltas);
  roi_targets_kernel<<<(unsigned)cdiv(R, 128), 128, 0, as_stream(stream)>>>(rois, assign, pos_list, neg_list, perm, P, R, gt_boxes,
                                                                          gt_class_ids, to_f6(std6_host, 1.f), out_rois, class_ids,
                                                                          deltas);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_iou3d_eps(const float* box, const float* boxes, int n, float* out, void* stream) {
  CFUN_CHECK_ARG(n >= 0);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(box && boxes && out);
  iou_eps_kernel<<<(unsigned)cdiv(n, 128), 128, 0, as_stream(stream)>>>(box, boxes, n, out);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_box_refinement(const float* box, const float* gt_box, int n, const float* std6_host, float* deltas,
                                   void* stream) {
  CFUN_CHECK_ARG(n >= 0);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(box && gt_box && deltas);
  refinement_kernel<<<(unsigned)cdiv(n, 128), 128, 0, as_stream(stream)>>>(box, gt_box, n, to_f6(std6_host, 1.f), deltas);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_roi_level(const float* boxes, int n, int* level, void* stream) {
  CFUN_CHECK_ARG(n >= 0);
  if (n == 0) return CFUN_OK;
  CFUN_CHECK_ARG(boxes && level);
  roi_level_kernel<<<(unsigned)cdiv(n, 128), 128, 0, as_stream(stream)>>>(boxes, n, level);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

extern "C" int cfun_mask_target_crop(const int* label, int D, int H, int W, const float* rois, int P, int ncls, int md,
                                     int mh, int mw, double* onehot, long long* cls_index, void* stream) {
  CFUN_CHECK_ARG(P >= 0 && D > 0 && H > 0 && W > 0 && ncls > 0 && md > 0 && mh > 0 && mw > 0);
  if (P == 0) return CFUN_OK;
  CFUN_CHECK_ARG(label && rois && (onehot || cls_index));
  long long total = (long long)P * md * mh * mw;
  mask_target_kernel<<<(unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()), 256, 0, as_stream(stream)>>>(
      label, D, H, W, rois, P, ncls, md, mh, mw, onehot, cls_index);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}
