// tcgen05 weight gradient of a 3x3x3 / stride 1 / pad 1 convolution from halo-resident, group-planar operands.
//
//   dW[tap][co][ci] = sum_voxels dY[v, co] * X[v + tap - 1, ci]
//
// Same data layout as conv_tc_halo.cu: split-bf16 activations packed [C/8 groups][N*(D+2)][H][W][8 ch] (16-byte rows, one
// per voxel).  With the un-swizzled MN-major canonical layout ((1,n),(8,k)):((X,SBO),(1,LBO)) the voxel axis is K (8
// consecutive 16-B rows = one w-line of the 1 x 16 x 8 slab, LBO = distance to the next line) and the 8-channel groups are
// the MN axis (SBO = plane pitch).  Per slab a CTA loads the dY slab (16 lines x 128 B per group) and ONE plane of the X
// halo (18 x 10 voxels, the plane d + kd - 1 of its tap group kd), then issues for each of its taps (kh,kw) eight K=16
// MMAs whose B operand is the halo plane viewed at line offset kh and row offset kw -- the shift costs nothing.
// The nine taps of a kd-group own 9 x Npad fp32 TMEM columns; voxels are split across CTAs (split-K) and flushed once with
// fp32 atomics.  Compared with conv_tc_wgrad.cu (channel-major slabs re-loaded per tap, 64-byte TMA rows) this moves 3.5x
// fewer bytes per voxel and in rows of 128 / 160 bytes.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int HW_THREADS = 192;
constexpr int HW_HT = 16, HW_WT = 8, HW_HH = 18, HW_WH = 10;
constexpr int HW_YP = HW_HT * HW_WT * 16;                          // 2048 B: dY slab plane of one channel group
constexpr int HW_XP_DATA = HW_HH * HW_WH * 16;                     // 2880 B: one halo plane of one channel group
constexpr int HW_XP = (HW_XP_DATA + 127) / 128 * 128;              // 2944 B pitch

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G,
                       cudaStream_t st);   // conv_tc_halo.cu

__device__ __forceinline__ uint64_t make_desc_interleave_hw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tma_load_4d_hw(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct HwParams {
  int N, D, H, W, Cout, Cin;
  int Gy, Gx;                 // 8-channel groups of dY (this M tile) and X
  int Npad;                   // MMA N = Gx * 8
  int T9, ng9;                // (kh,kw) taps per CTA, groups of them
  int tilesH, tilesW;
  int nsplit, stages, tmem_cols;
  long long units_total, units_per_cta;
  float* dw;
};

__global__ void __launch_bounds__(HW_THREADS, 1)
conv_tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl,
                          const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, const HwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint8_t* ring = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int y_bytes = p.Gy * HW_YP, x_bytes = p.Gx * HW_XP;
  const int stage_bytes = parts * (y_bytes + x_bytes);
  const int kd = blockIdx.y / p.ng9;
  const int t9_0 = (blockIdx.y % p.ng9) * p.T9;
  const int nt9 = min(p.T9, 9 - t9_0);
  const int g0 = blockIdx.z * 16;                        // first dY channel group of this M tile
  const long long u_beg = (long long)blockIdx.x * p.units_per_cta;
  const long long u_end = min(p.units_total, u_beg + p.units_per_cta);
  const int niter = (int)max(0LL, u_end - u_beg);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_yh);
    prefetch_tmap(&map_xh);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int slot = it % p.stages;
        mbar_wait(&empty_bar[slot], (uint32_t)(((it / p.stages) & 1) ^ 1), 310);
        long long t = u_beg + it;
        const int wb = (int)(t % p.tilesW); t /= p.tilesW;
        const int hb = (int)(t % p.tilesH); t /= p.tilesH;
        const int d = (int)(t % p.D);
        const int n = (int)(t / p.D);
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(parts * (p.Gy * HW_YP + p.Gx * HW_XP_DATA)));
        uint8_t* sb = ring + (size_t)slot * stage_bytes;
        const int plane_y = n * (p.D + 2) + d + 1;
        const int plane_x = n * (p.D + 2) + d + kd;      // padded index of plane d + kd - 1
        for (int part = 0; part < parts; ++part) {
          const CUtensorMap* my = part == 0 ? &map_yh : &map_yl;
          const CUtensorMap* mx = part == 0 ? &map_xh : &map_xl;
          uint8_t* yb = sb + (size_t)part * y_bytes;
          uint8_t* xb = sb + (size_t)parts * y_bytes + (size_t)part * x_bytes;
          for (int g = 0; g < p.Gy; ++g) tma_load_4d_hw(my, &full_bar[slot], yb + g * HW_YP, wb * HW_WT * 8, hb * HW_HT, plane_y, g0 + g);
          for (int g = 0; g < p.Gx; ++g) tma_load_4d_hw(mx, &full_bar[slot], xb + g * HW_XP, (wb * HW_WT - 1) * 8, hb * HW_HT - 1, plane_x, g);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major (bits 15, 16), bf16 x bf16 -> fp32, M = 128, N = Npad
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
      for (int it = 0; it < niter; ++it) {
        const int slot = it % p.stages;
        mbar_wait(&full_bar[slot], (uint32_t)((it / p.stages) & 1), 320);
        tc_fence_after();
        const uint32_t sb = smem_u32(ring + (size_t)slot * stage_bytes);
        const uint32_t y_hi = sb, y_lo = sb + (uint32_t)y_bytes;
        const uint32_t x_hi = sb + (uint32_t)(parts * y_bytes), x_lo = x_hi + (uint32_t)x_bytes;
        const uint32_t first = it == 0 ? 0u : 1u;
#pragma unroll 1
        for (int t = 0; t < nt9; ++t) {
          const int t9 = t9_0 + t;
          const int kh = t9 / 3, kw = t9 - kh * 3;
          const uint32_t dcol = tmem_base + (uint32_t)(t * p.Npad);
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            const uint32_t ay = (uint32_t)(j * 2 * HW_WT * 16);                       // lines 2j, 2j+1 of the slab
            const uint32_t ax = (uint32_t)((((2 * j + kh) * HW_WH) + kw) * 16);       // same lines of the halo plane, shifted
            const uint64_t a_hi = make_desc_interleave_hw(y_hi + ay, HW_WT * 16, HW_YP);
            const uint64_t b_hi = make_desc_interleave_hw(x_hi + ax, HW_WH * 16, HW_XP);
            umma_bf16(dcol, a_hi, b_hi, idesc, first | (uint32_t)(j > 0));
            if (parts == 2) {
              const uint64_t a_lo = make_desc_interleave_hw(y_lo + ay, HW_WT * 16, HW_YP);
              const uint64_t b_lo = make_desc_interleave_hw(x_lo + ax, HW_WH * 16, HW_XP);
              umma_bf16(dcol, a_lo, b_hi, idesc, 1);
              umma_bf16(dcol, a_hi, b_lo, idesc, 1);
            }
          }
        }
        umma_commit(&empty_bar[slot]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int quad = warp & 3;
    const int co = g0 * 8 + quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0, 330);
    tc_fence_after();
    if (niter > 0) {
      for (int t = 0; t < nt9; ++t) {
        const int tap = kd * 9 + t9_0 + t;
        for (int j = 0; j < p.Npad; j += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * p.Npad + j), r);
          tmem_ld_wait();
          if (co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = j + i;
              if (ci < p.Cin) atomicAdd(p.dw + ((long long)co * p.Cin + ci) * 27 + tap, __uint_as_float(r[i]));
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

struct HwPlan {
  int Gy_total, Gx, Npad, T9, ng9, mtiles, stages, tmem_cols;
  size_t act_y, act_x, off_yh, off_yl, off_xh, off_xl, total, smem;
};

static bool make_hw_plan(const cfun_conv3d_desc* d, HwPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != 3 || d->kH != 3 || d->kW != 3 || d->pD != 1 || d->pH != 1 || d->pW != 1) return false;
  if (d->Hin < 8 || d->Win < 8) return false;
  pl.Gy_total = (int)cdiv(d->Cout, 8);
  pl.Gx = (int)align_up((size_t)d->Cin, 16) / 8;
  pl.Npad = pl.Gx * 8;
  if (pl.Npad > 256) return false;
  pl.T9 = std::min(9, 512 / pl.Npad);
  pl.ng9 = (int)cdiv(9, pl.T9);
  pl.T9 = (int)cdiv(9, pl.ng9);
  pl.mtiles = (int)cdiv(pl.Gy_total, 16);
  int cols = 32;
  while (cols < pl.T9 * pl.Npad) cols <<= 1;
  if (cols > 512) return false;
  pl.tmem_cols = cols;
  const int gy = std::min(pl.Gy_total, 16);
  const size_t stage = 2 * ((size_t)gy * HW_YP + (size_t)pl.Gx * HW_XP);
  // M = 128 makes the MMA read 16 channel-group planes from the dY tile base; the planes beyond the real ones only feed
  // accumulator rows that are never read back, but they must lie inside the allocation: 36 KB of slack after the ring
  const size_t slack = 36 * 1024;
  pl.stages = (int)std::min<size_t>(4, (224 * 1024 - 2048 - slack) / stage);
  if (pl.stages < 2) return false;
  pl.smem = 2048 + pl.stages * stage + slack;
  pl.act_y = align_up((size_t)pl.Gy_total * d->N * (d->Dout + 2) * d->Hout * d->Wout * 16, 1024);
  pl.act_x = align_up((size_t)pl.Gx * d->N * (d->Din + 2) * d->Hin * d->Win * 16, 1024);
  pl.off_yh = 0; pl.off_yl = pl.act_y; pl.off_xh = 2 * pl.act_y; pl.off_xl = 2 * pl.act_y + pl.act_x;
  pl.total = 2 * pl.act_y + 2 * pl.act_x + 2048;
  return true;
}

bool hw_supported(const cfun_conv3d_desc* d) {
  const char* e = getenv("CFUN_TC_HALO");
  if (e && e[0] == '0') return false;
  HwPlan pl;
  if (!make_hw_plan(d, pl)) return false;
  return d->Cin >= 16 && (d->Cin & 3) == 0 && d->Cout >= 8 && (d->Cout & 3) == 0;
}
size_t hw_workspace(const cfun_conv3d_desc* d) {
  HwPlan pl;
  return make_hw_plan(d, pl) ? pl.total : 0;
}

static int encode_gp_map(CUtensorMap* m, void* base, int W, int H, long long planes, int G, int box_w8, int box_h) {
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)G};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)box_w8, (cuuint32_t)box_h, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = get_tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(group-planar) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  return CFUN_OK;
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

int hw_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  HwPlan pl;
  CFUN_CHECK_ARG(make_hw_plan(d, pl));
  CFUN_CHECK_ARG(x && dy && dw && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d halo wgrad: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yh);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yl);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xh);
  __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xl);
  int prc;
  if ((prc = launch_pack_act_gp(dy, yh, split ? yl : nullptr, d->N, d->Dout, d->Hout, d->Wout, d->Cout, pl.Gy_total, st)) != CFUN_OK) return prc;
  if ((prc = launch_pack_act_gp(x, xh, split ? xl : nullptr, d->N, d->Din, d->Hin, d->Win, d->Cin, pl.Gx, st)) != CFUN_OK) return prc;
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * 27, st));
  CUtensorMap myh, myl, mxh, mxl;
  int rc;
  const long long planes = (long long)d->N * (d->Din + 2);
  if ((rc = encode_gp_map(&myh, yh, d->Wout, d->Hout, planes, pl.Gy_total, HW_WT * 8, HW_HT)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map(&myl, split ? yl : yh, d->Wout, d->Hout, planes, pl.Gy_total, HW_WT * 8, HW_HT)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map(&mxh, xh, d->Win, d->Hin, planes, pl.Gx, HW_WH * 8, HW_HH)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map(&mxl, split ? xl : xh, d->Win, d->Hin, planes, pl.Gx, HW_WH * 8, HW_HH)) != CFUN_OK) return rc;

  HwParams p;
  p.N = d->N; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cout = d->Cout; p.Cin = d->Cin;
  p.Gy = std::min(pl.Gy_total, 16);
  p.Gx = pl.Gx; p.Npad = pl.Npad; p.T9 = pl.T9; p.ng9 = pl.ng9;
  p.tilesH = (int)cdiv(d->Hin, HW_HT); p.tilesW = (int)cdiv(d->Win, HW_WT);
  p.nsplit = split ? 3 : 1;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
  p.units_total = (long long)d->N * d->Din * p.tilesH * p.tilesW;
  const int ygroups = 3 * pl.ng9 * pl.mtiles;
  long long ctas = std::max<long long>(1, (2LL * num_sms()) / ygroups);
  ctas = std::min<long long>(ctas, cdiv(p.units_total, 8));
  p.units_per_cta = cdiv(p.units_total, ctas);
  ctas = cdiv(p.units_total, p.units_per_cta);
  p.dw = dw;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // the last M tile may own fewer than 16 channel groups: the TMA loads only Gy real planes; rows beyond Cout are ignored
  if (pl.mtiles > 1 && (pl.Gy_total % 16) != 0) p.Gy = 16;   // full tiles load 16; the tail tile's extra planes are OOB zero fill
  dim3 grid((unsigned)ctas, (unsigned)(3 * pl.ng9), (unsigned)pl.mtiles);
  conv_tc_wgrad_halo_kernel<<<grid, HW_THREADS, pl.smem, st>>>(myh, myl, mxh, mxl, p);
  CFUN_LAUNCH_CHECK();
  if (dbias) return simt_bias_grad(dy, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}

int tc_debug_read_hw(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
