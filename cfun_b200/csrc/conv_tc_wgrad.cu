// tcgen05 weight gradient of a stride-1 3-D convolution.
//
//   dW[tap][co][ci] = sum_{n,d,h,w} dY[n,d,h,w,co] * X[n, d+kd-pD, h+kh-pH, w+kw-pW, ci]
//
// The reduction runs over voxels, so both MMA operands need the voxel axis as K.  A pre-pass writes channel-major
// ("NCDHW") split-bf16 copies of dY and X; in that layout a run of Wk consecutive w-voxels of one (n,d,h) line is a
// contiguous K slab, and a [channels x Wk] TMA box lands in shared memory as a plain K-major, hardware-swizzled GEMM
// tile.  Per line-slab the CTA issues, for every tap of its tap group, D_tap[co, ci] += dY_tile[co, k] * X_tile_tap[ci, k]
// where X_tile_tap is the same box shifted by the tap offset -- out-of-bounds coordinates are zero-filled by the TMA
// unit (= the conv's zero padding).  Every tap owns Npad fp32 columns of TMEM (<= 512 columns per CTA, so the 27 taps are
// split into tap groups across blockIdx.y); the voxel range is split across blockIdx.x (split-K), and each CTA flushes its
// accumulators once at the end with fp32 atomics into dW.  M is always 128 rows (rows >= Cout are zero-filled by TMA):
// tcgen05 prices M = 64 and M = 128 identically.
// Operands are split-bf16 pairs, accumulated as hi*hi + lo*hi + hi*lo (see conv_tc.cu).
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int WG_THREADS = 192;

struct WgParams {
  int N, Do, Ho, Wo;             // dY spatial extents
  int kD, kH, kW, pD, pH, pW;
  int Cout, Cin, taps;           // Cin = input channels of this launch's slice
  int ci0, Cin_total;            // slice offset / full channel count (dW addressing)
  int Npad;                      // MMA N (Cin tile rounded up to 16)
  int Wk, swz;                   // K slab (voxels) and its byte width (Wk*2 = swizzle width)
  int kt_per_line;               // ceil(Wo / Wk)
  int T;                         // taps per CTA (tap group size)
  int nsplit, stages, tmem_cols;
  long long slabs_total;         // N*Do*Ho*kt_per_line
  long long slabs_per_cta;
  float* dw;                     // (Cout, Cin, taps) fp32, zero-initialised
  int debug_skip;                // bring-up aid (CFUN_WG_DEBUG): 1 no TMA, 2 no MMA, 4 no epilogue
};

// TMA requires the innermost box coordinate to be 16-byte aligned, so the +-1 voxel shifts along w (the K axis here)
// cannot be expressed as box coordinates: the pack pass materialises kW copies of X, copy kw pre-shifted by kw - pW along
// w, and the kernel picks the map of its tap's kw.  Shifts along h and d stay plain (outer-dimension) coordinates.
constexpr int WG_MAX_KW = 5;
struct alignas(64) XMaps {
  CUtensorMap hi[WG_MAX_KW];
  CUtensorMap lo[WG_MAX_KW];
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl,
                     const __grid_constant__ XMaps xmaps, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint8_t* ring = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int a_bytes = 128 * p.swz;                               // dY tile (one part)
  const int b_bytes = ((p.Npad * p.swz + 1023) / 1024) * 1024;       // X tile of one tap (one part)
  const int tap0 = blockIdx.y * p.T;
  const int ntap = min(p.T, p.taps - tap0);
  const int stage_bytes = parts * (a_bytes + p.T * b_bytes);
  const int m0 = blockIdx.z * 128;                               // Cout tile
  const long long s_beg = (long long)blockIdx.x * p.slabs_per_cta;
  const long long s_end = min(p.slabs_total, s_beg + p.slabs_per_cta);
  const int niter = (int)max(0LL, s_end - s_beg);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_yh);
    prefetch_tmap(&xmaps.hi[0]);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int slot = it % p.stages;
        const uint32_t ph = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&empty_bar[slot], ph ^ 1u, 110);
        long long s = s_beg + it;
        const int kt = (int)(s % p.kt_per_line); s /= p.kt_per_line;
        const int h = (int)(s % p.Ho); s /= p.Ho;
        const int d = (int)(s % p.Do);
        const int n = (int)(s / p.Do);
        const int w0 = kt * p.Wk;
        if (p.debug_skip & 1) { mbar_arrive_expect_tx(&full_bar[slot], 0); continue; }
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(parts * (a_bytes + ntap * p.Npad * p.swz)));
        uint8_t* sb = ring + (size_t)slot * stage_bytes;
        tma_load_5d(&map_yh, &full_bar[slot], sb, w0, h, d, n, m0);
        if (parts == 2) tma_load_5d(&map_yl, &full_bar[slot], sb + a_bytes, w0, h, d, n, m0);
        uint8_t* bb = sb + parts * a_bytes;
        const int khw = p.kH * p.kW;
        for (int t = 0; t < ntap; ++t) {
          const int tap = tap0 + t;
          const int kd = tap / khw, r = tap - kd * khw, kh = r / p.kW, kw = r - kh * p.kW;
          const int xh = h + kh - p.pH, xd = d + kd - p.pD;
          tma_load_5d(&xmaps.hi[kw], &full_bar[slot], bb + (size_t)(t * parts) * b_bytes, w0, xh, xd, n, 0);
          if (parts == 2) tma_load_5d(&xmaps.lo[kw], &full_bar[slot], bb + (size_t)(t * parts + 1) * b_bytes, w0, xh, xd, n, 0);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, elected lane issues (tc_ptx.cuh "issue-rate note")
    const uint32_t leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hiword = (uint32_t)(make_desc_kmajor(0, p.swz) >> 32);
    const uint32_t lbo1 = 1u << 16;
    const int ksteps = p.Wk / 16;
    for (int it = 0; it < niter; ++it) {
      const int slot = it % p.stages;
      const uint32_t ph = (uint32_t)((it / p.stages) & 1);
      mbar_wait(&full_bar[slot], ph, 120);
      tc_fence_after();
      const uint32_t sb = desc_addr(smem_u32(ring + (size_t)slot * stage_bytes)) | lbo1;
      const uint32_t bb = sb + (uint32_t)((parts * a_bytes) >> 4);
      const uint32_t acc = it > 0 ? 1u : 0u;
      for (int t = 0; t < ntap && !(p.debug_skip & 2); ++t) {
        const uint32_t dcol = tmem_base + (uint32_t)(t * p.Npad);
        const uint32_t bt = bb + (uint32_t)(((t * parts) * b_bytes) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ks < ksteps) {
            const uint32_t koff = (uint32_t)(ks * 2);   // 16 bf16 (32 B) along K inside the swizzled row
            const uint64_t a_hi = desc_join(hiword, sb + koff);
            const uint64_t b_hi = desc_join(hiword, bt + koff);
            if (leader) {
              if (ks == 0) umma_bf16(dcol, a_hi, b_hi, idesc, acc);
              else umma_bf16_acc(dcol, a_hi, b_hi, idesc);
              if (parts == 2) {
                const uint64_t a_lo = desc_join(hiword, sb + (uint32_t)(a_bytes >> 4) + koff);
                const uint64_t b_lo = desc_join(hiword, bt + (uint32_t)(b_bytes >> 4) + koff);
                umma_bf16_acc(dcol, a_lo, b_hi, idesc);
                umma_bf16_acc(dcol, a_hi, b_lo, idesc);
              }
            }
          }
        }
      }
      if (leader) umma_commit(&empty_bar[slot]);
      __syncwarp();
    }
    if (leader) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int co = m0 + quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0, 130);
    tc_fence_after();
    if (niter > 0 && !(p.debug_skip & 4)) {
      for (int t = 0; t < ntap; ++t) {
        const int tap = tap0 + t;
        for (int j = 0; j < p.Npad; j += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * p.Npad + j), r);
          tmem_ld_wait();
          if (co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = j + i;
              if (ci < p.Cin) atomicAdd(p.dw + ((long long)co * p.Cin_total + p.ci0 + ci) * p.taps + tap, __uint_as_float(r[i]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// fp32 NDHWC [rows, C] -> channel-major split bf16 [C][rows] (32x32 smem-tiled transpose)
__global__ void __launch_bounds__(256) pack_transpose_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, long long rows, int C,
                                                             long long pitch) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    long long r = r0 + j;
    int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < C) ? __ldg(x + r * C + c) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j;
    long long r = r0 + threadIdx.x;
    if (c < C && r < rows) {
      __nv_bfloat16 h, l;
      split_bf16(tile[threadIdx.x][j], h, l);
      hi[(long long)c * pitch + r] = h;
      if (lo) lo[(long long)c * pitch + r] = l;
    }
  }
}

// X [N*D*H rows of W voxels, C] fp32 -> kW channel-major copies [kw][C][N*D*H][Wp] with copy kw holding x[.., w + kw - pW, c]
// (zero outside [0, W)); Wp = output width rounded up to the TMA row granularity, so every copy is aligned to the dY lines.
__global__ void __launch_bounds__(256) pack_transpose_shift_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                                   __nv_bfloat16* __restrict__ lo, long long lines, int W, int Wo,
                                                                   int C, int kW, int pW, long long pitch) {
  // (Wo here is the padded line pitch Wp of the packed copies)
  __shared__ float tile[32][33];
  const long long line = blockIdx.x;                 // (n, d, h) line index of X
  const int wt = blockIdx.y;
  const int w0 = wt * 32 - pW;                       // first source voxel of copy kw = 0 (may be negative)
  const int c0 = blockIdx.z * 32;
  // stage 32 + (kW - 1) source voxels? keep it simple: one tile per kw
  for (int kw = 0; kw < kW; ++kw) {
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      int w = w0 + j + kw;
      int c = c0 + threadIdx.x;
      tile[j][threadIdx.x] = (w >= 0 && w < W && c < C) ? __ldg(x + (line * W + w) * C + c) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      int c = c0 + j;
      int wo = wt * 32 + threadIdx.x;
      if (c < C && wo < Wo) {
        __nv_bfloat16 h, l;
        split_bf16(tile[threadIdx.x][j], h, l);
        long long o = ((long long)kw * C + c) * pitch + line * Wo + wo;
        hi[o] = h;
        if (lo) lo[o] = l;
      }
    }
  }
}

struct WgPlan {
  int Wk, swz, T, ngroups, mtiles, Npad, stages, tmem_cols;
  int slices, Cs;               // input channels are processed in `slices` launches of Cs channels (MMA N <= 256)
  int Wp;                       // packed line pitch: Wout, or Wout rounded up to 16 when it is not a multiple of 8
  size_t off_yh, off_yl, off_xh, off_xl, total;
  long long rows_y, rows_x, lines_x;
  long long pitch_y, pitch_x;   // channel-plane pitch (elements): padded so that planes do not alias in L2
  size_t copy_elems;
};

static bool make_wg_plan(const cfun_conv3d_desc* d, WgPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  pl.Wp = (d->Wout & 7) ? (int)align_up((size_t)d->Wout, 16) : d->Wout;   // TMA row stride must be a multiple of 16 B
  if (d->kW > WG_MAX_KW) return false;
  if (d->Cin > 512) return false;
  pl.slices = d->Cin > 256 ? 2 : 1;
  if (d->Cin % pl.slices) return false;
  pl.Cs = d->Cin / pl.slices;
  pl.Wk = (pl.Wp % 64 == 0) ? 64 : ((pl.Wp % 32 == 0) ? 32 : 16);
  pl.swz = pl.Wk * 2;
  pl.Npad = (int)align_up((size_t)pl.Cs, 16);
  const int taps = d->kD * d->kH * d->kW;
  pl.T = std::min(taps, 512 / pl.Npad);
  // keep a stage (dY tile + T shifted X tiles, hi+lo) within ~100 KB so that two stages fit
  const int b_bytes = (int)align_up((size_t)pl.Npad * pl.swz, 1024);
  while (pl.T > 1 && 2 * (128 * pl.swz + pl.T * b_bytes) > 100 * 1024) --pl.T;
  pl.ngroups = (int)cdiv(taps, pl.T);
  pl.T = (int)cdiv(taps, pl.ngroups);
  pl.mtiles = (int)cdiv(d->Cout, 128);
  int cols = 32;
  while (cols < pl.T * pl.Npad) cols <<= 1;
  if (cols > 512) return false;
  pl.tmem_cols = cols;
  const size_t stage = 2 * (size_t)(128 * pl.swz + pl.T * b_bytes);
  pl.stages = (int)std::min<size_t>(4, (200 * 1024) / stage);
  if (pl.stages < 2) return false;
  pl.rows_y = (long long)d->N * d->Dout * d->Hout * pl.Wp;
  pl.rows_x = (long long)d->N * d->Din * d->Hin * d->Win;
  // The planes of consecutive channels are read by consecutive rows of every TMA box; an un-padded pitch is a multiple of
  // large powers of two for the usual extents (4 x 96^3 x 2 B = 27 x 256 KiB) and makes all rows of a box camp on the
  // same L2 slices.  17 x 128 B of padding per plane spreads them.
  pl.pitch_y = pl.rows_y + 1088;
  const size_t ybytes = align_up((size_t)pl.pitch_y * d->Cout * 2, 1024);
  pl.lines_x = (long long)d->N * d->Din * d->Hin;
  pl.pitch_x = pl.lines_x * pl.Wp + 1088;
  pl.copy_elems = (size_t)d->Cin * pl.pitch_x;                    // one pre-shifted copy of X (one part)
  const size_t xbytes = align_up(pl.copy_elems * 2 * d->kW, 1024);
  pl.off_yh = 0; pl.off_yl = ybytes; pl.off_xh = 2 * ybytes; pl.off_xl = 2 * ybytes + xbytes;
  pl.total = 2 * ybytes + 2 * xbytes + 2048;
  return true;
}

bool ds_supported(const cfun_conv3d_desc* d);     // conv_tc_wgrad_ds.cu (d-stacked, thin channels)
size_t ds_workspace(const cfun_conv3d_desc* d);
int ds_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st);

int s2d_conv(const cfun_conv3d_desc* d, int pass, const float* a, const float* b, const float* bias, float* out, float* dbias,
             int epi, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st);   // conv_s2d.cu

// capability of the generic channel-major kernel without the size policy (used by the space-to-depth path)
bool wg_capable(const cfun_conv3d_desc* d) {
  WgPlan pl;
  if (!make_wg_plan(d, pl)) return false;
  return d->Cin >= 16 && d->Cout >= 8;
}

bool tc_wgrad_supported(const cfun_conv3d_desc* d) {
  if (ds_supported(d)) return true;
  WgPlan pl;
  if (!make_wg_plan(d, pl)) return false;
  if (d->kD * d->kH * d->kW < 27) return false;
  if (d->Cin < 16 || d->Cout < 8) return false;
  return true;
}

size_t tc_wgrad_workspace(const cfun_conv3d_desc* d) {
  if (ds_supported(d)) return ds_workspace(d);
  WgPlan pl;
  if (!make_wg_plan(d, pl)) return 0;
  return pl.total;
}

static int encode_cmajor_map(CUtensorMap* m, void* base, int W, int H, int D, int N, int C, int Wk, int rows, int swz,
                             long long pitch) {
  cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)C};
  cuuint64_t strides[4] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)D * H * W * 2, (cuuint64_t)pitch * 2};
  cuuint32_t box[5] = {(cuuint32_t)Wk, 1, 1, 1, (cuuint32_t)rows};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = get_tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(channel-major) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  return CFUN_OK;
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

int tc_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  if (d->sD == 2) return s2d_conv(d, CFUN_PASS_BWD_WEIGHT, x, dy, nullptr, dw, dbias, 0, nsplit, ws, ws_bytes, st);
  if (ds_supported(d)) return ds_conv_bwd_weight(d, x, dy, dw, dbias, nsplit, ws, ws_bytes, st);
  WgPlan pl;
  CFUN_CHECK_ARG(make_wg_plan(d, pl));
  CFUN_CHECK_ARG(x && dy && dw && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d tc wgrad: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yh);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yl);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xh);
  __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xl);
  if (pl.Wp == d->Wout) {
    pack_transpose_kernel<<<dim3((unsigned)cdiv(pl.rows_y, 32), (unsigned)cdiv(d->Cout, 32)), dim3(32, 8), 0, st>>>(dy, yh, split ? yl : nullptr, pl.rows_y, d->Cout, pl.pitch_y);
  } else {   // padded lines: the line-aware kernel with one un-shifted copy writes zeros into [Wout, Wp)
    const long long lines_y = (long long)d->N * d->Dout * d->Hout;
    pack_transpose_shift_kernel<<<dim3((unsigned)lines_y, (unsigned)cdiv(pl.Wp, 32), (unsigned)cdiv(d->Cout, 32)), dim3(32, 8), 0, st>>>(
        dy, yh, split ? yl : nullptr, lines_y, d->Wout, pl.Wp, d->Cout, 1, 0, pl.pitch_y);
  }
  CFUN_LAUNCH_CHECK();
  pack_transpose_shift_kernel<<<dim3((unsigned)pl.lines_x, (unsigned)cdiv(pl.Wp, 32), (unsigned)cdiv(d->Cin, 32)), dim3(32, 8), 0, st>>>(
      x, xh, split ? xl : nullptr, pl.lines_x, d->Win, pl.Wp, d->Cin, d->kW, d->pW, pl.pitch_x);
  CFUN_LAUNCH_CHECK();
  const int taps = d->kD * d->kH * d->kW;
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * taps, st));

  CUtensorMap myh, myl;
  int rc;
  if ((rc = encode_cmajor_map(&myh, yh, pl.Wp, d->Hout, d->Dout, d->N, d->Cout, pl.Wk, 128, pl.swz, pl.pitch_y)) != CFUN_OK) return rc;
  if ((rc = encode_cmajor_map(&myl, split ? yl : yh, pl.Wp, d->Hout, d->Dout, d->N, d->Cout, pl.Wk, 128, pl.swz, pl.pitch_y)) != CFUN_OK) return rc;

  WgParams p;
  p.N = d->N; p.Do = d->Dout; p.Ho = d->Hout; p.Wo = d->Wout;
  p.kD = d->kD; p.kH = d->kH; p.kW = d->kW; p.pD = d->pD; p.pH = d->pH; p.pW = d->pW;
  p.Cout = d->Cout; p.Cin = pl.Cs; p.Cin_total = d->Cin; p.taps = taps;
  p.Npad = pl.Npad; p.Wk = pl.Wk; p.swz = pl.swz;
  p.kt_per_line = (int)cdiv(pl.Wp, pl.Wk);      // Wp = 8: one 16-wide slab per line, the upper half is TMA zero fill
  p.T = pl.T;
  p.nsplit = split ? 3 : 1;
  p.stages = pl.stages;
  p.tmem_cols = pl.tmem_cols;
  p.slabs_total = (long long)d->N * d->Dout * d->Hout * p.kt_per_line;
  // split-K over voxels: enough CTAs for ~2 waves, at least 32 slabs each
  long long ctas = std::max<long long>(1, (2LL * num_sms()) / ((long long)pl.ngroups * pl.mtiles));
  ctas = std::min<long long>(ctas, cdiv(p.slabs_total, 32));
  p.slabs_per_cta = cdiv(p.slabs_total, ctas);
  ctas = cdiv(p.slabs_total, p.slabs_per_cta);
  p.dw = dw;
  { const char* e = getenv("CFUN_WG_DEBUG"); p.debug_skip = e ? atoi(e) : 0; }
  const int parts = split ? 2 : 1;
  const int b_bytes = (int)align_up((size_t)pl.Npad * pl.swz, 1024);
  const size_t stage_bytes = (size_t)parts * (128 * pl.swz + pl.T * b_bytes);
  const size_t smem = 2048 + pl.stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)ctas, (unsigned)pl.ngroups, (unsigned)pl.mtiles);
  for (int sl = 0; sl < pl.slices; ++sl) {       // channel slices of X (MMA N <= 256): channel planes are contiguous in the pack
    XMaps xm;
    const size_t coff = (size_t)sl * pl.Cs * pl.pitch_x;
    for (int kw = 0; kw < WG_MAX_KW; ++kw) {
      const int k = kw < d->kW ? kw : 0;
      if ((rc = encode_cmajor_map(&xm.hi[kw], xh + (size_t)k * pl.copy_elems + coff, pl.Wp, d->Hin, d->Din, d->N, pl.Cs, pl.Wk, pl.Npad, pl.swz, pl.pitch_x)) != CFUN_OK) return rc;
      if ((rc = encode_cmajor_map(&xm.lo[kw], (split ? xl : xh) + (size_t)k * pl.copy_elems + coff, pl.Wp, d->Hin, d->Din, d->N, pl.Cs, pl.Wk, pl.Npad, pl.swz, pl.pitch_x)) != CFUN_OK) return rc;
    }
    p.ci0 = sl * pl.Cs;
    conv_tc_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(myh, myl, xm, p);
    CFUN_LAUNCH_CHECK();
  }
  if (dbias) return simt_bias_grad(dy, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}

int tc_debug_read_wgrad(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
