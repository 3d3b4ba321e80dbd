// tcgen05 weight gradient of a 3x3x3 / stride 1 / pad 1 convolution, "d-stacked" variant for thin-channel layers.
//
//   dW[kd][kh][kw][co][ci] = sum_{n,d,h,w} dY[n,d,h,w,co] * X[n, d+kd-1, h+kh-1, w+kw-1, ci]
//
// conv_tc_wgrad_halo.cu feeds M = 128 accumulator rows with Cout channels only (20..40 of 128 rows useful in the top
// levels of the U-Net, where 2/3 of the weight-gradient time goes) and splits kd over CTAs.  Here the work unit is one
// 18 x 10 halo plane p of X, and the A operand stacks the THREE dY planes that meet it (d = p-1, p, p+1 <-> kd = 2, 1, 0)
// along M: with the group-planar pack [C/8][N*(D+2)][H][W][8ch] one TMA box {8 w, 16 h, 3 planes, Gy groups} lands in
// shared memory as 3*Gy planes of 2 KB at a uniform pitch, which is exactly the MN-major un-swizzled canonical layout
// with SBO = 2 KB: accumulator row = (channel group, d-shift, channel) -> up to 5 groups x 3 shifts x 8 = 120 of 128 rows.
// The (kh,kw) taps stay free views of the X halo plane (descriptor start address), each tap owns Npad fp32 TMEM columns.
// Per unit: 9 taps x 8 K-steps x 3 split-bf16 MMAs instead of 27 x 8 x 3, two TMA instructions per operand part.
// Voxels are split across CTAs (split-K) and flushed once with fp32 atomics into dW (zero-initialised by the caller side).
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int DS_THREADS = 192;
constexpr int DS_HT = 16, DS_WT = 8, DS_HH = 18, DS_WH = 10;
constexpr int DS_YP = DS_HT * DS_WT * 16;        // 2048 B: one dY plane of one channel group
constexpr int DS_XP = DS_HH * DS_WH * 16;        // 2880 B: one X halo plane of one channel group (box-packed pitch)
constexpr int DS_MAX_GT = 5;                     // channel groups per M tile: 5 x 3 shifts = 15 of 16 row groups

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G,
                       cudaStream_t st);   // conv_tc_halo.cu
int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

__device__ __forceinline__ uint64_t make_desc_mn_ds(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tma_load_4d_ds(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct DsParams {
  int N, D, H, W, Cout, Cin;
  int Gt;                     // dY channel groups per M tile (<= 5)
  int Gx;                     // X channel groups of this launch's Cin slice (Npad / 8)
  int ci0;                    // first input channel of the slice (dW addressing); X groups start at ci0 / 8
  int Npad;                   // MMA N
  int T9, ng9;                // (kh,kw) taps per CTA, number of tap groups
  int ntap;                   // selected (kh,kw) taps (9 for a dense 3^3 kernel)
  int tap_list[9];            // their indices kh*3+kw; the space-to-depth weight gradient needs only {0,1,3,4}
  int kd_mask;                // bit kd set = this kd plane of dW is wanted
  int tilesH, tilesW;
  int nsplit, stages;
  int tmem_cols;
  int y_bytes, x_bytes, stage_bytes;
  long long units_total, units_per_cta;
  float* dw;
};

// LEAN (opt-in, CFUN_TC_LEAN=1, split mode; not yet validated): one leader region per unit instead of one per K step
// (see the note at conv_tc_halo_kernel)
template <bool LEAN>
__global__ void __launch_bounds__(DS_THREADS, 1)
conv_tc_wgrad_ds_kernel(const __grid_constant__ CUtensorMap map_yh, const __grid_constant__ CUtensorMap map_yl,
                        const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, const DsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint8_t* ring = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int t9_0 = blockIdx.y * p.T9;
  const int nt9 = min(p.T9, p.ntap - t9_0);
  const int g0 = blockIdx.z * p.Gt;                      // first dY channel group of this M tile
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
        mbar_wait(&empty_bar[slot], (uint32_t)(((it / p.stages) & 1) ^ 1), 410);
        long long t = u_beg + it;
        const int wb = (int)(t % p.tilesW); t /= p.tilesW;
        const int hb = (int)(t % p.tilesH); t /= p.tilesH;
        const int pl = (int)(t % p.D);                   // X plane of this unit
        const int n = (int)(t / p.D);
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(parts * (p.y_bytes + p.x_bytes)));
        uint8_t* sb = ring + (size_t)slot * p.stage_bytes;
        const int plane0 = n * (p.D + 2) + pl;           // padded index of dY plane pl-1 (and of X plane pl, minus one)
        for (int part = 0; part < parts; ++part) {
          tma_load_4d_ds(part == 0 ? &map_yh : &map_yl, &full_bar[slot], sb + (size_t)part * p.y_bytes, wb * DS_WT * 8,
                         hb * DS_HT, plane0, g0);
          tma_load_4d_ds(part == 0 ? &map_xh : &map_xl, &full_bar[slot], sb + (size_t)parts * p.y_bytes + (size_t)part * p.x_bytes,
                         (wb * DS_WT - 1) * 8, hb * DS_HT - 1, plane0 + 1, p.ci0 >> 3);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop (tc_ptx.cuh "issue-rate note"): descriptors = constant high word + 16-byte offsets
    const uint32_t leader = elect_one();
    // both operands MN-major (bits 15, 16), bf16 x bf16 -> fp32, M = 128, N = Npad
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiword = (uint32_t)(make_desc_mn_ds(0, DS_WT * 16, DS_YP) >> 32);
    const uint32_t b_hiword = (uint32_t)(make_desc_mn_ds(0, DS_WH * 16, DS_XP) >> 32);
    const uint32_t a_lbo = (uint32_t)((DS_WT * 16) >> 4) << 16, b_lbo = (uint32_t)((DS_WH * 16) >> 4) << 16;
    for (int it = 0; it < niter; ++it) {
      const int slot = it % p.stages;
      mbar_wait(&full_bar[slot], (uint32_t)((it / p.stages) & 1), 420);
      tc_fence_after();
      const uint32_t sb = smem_u32(ring + (size_t)slot * p.stage_bytes);
      const uint32_t y_hi = desc_addr(sb) | a_lbo, y_lo = desc_addr(sb + (uint32_t)p.y_bytes) | a_lbo;
      const uint32_t x_hi = desc_addr(sb + (uint32_t)(parts * p.y_bytes)) | b_lbo;
      const uint32_t x_lo = desc_addr(sb + (uint32_t)(parts * p.y_bytes + p.x_bytes)) | b_lbo;
      const uint32_t first = it == 0 ? 0u : 1u;
      if (LEAN) {
        if (leader) {
#pragma unroll 1
          for (int t = 0; t < nt9; ++t) {
            const int t9 = p.tap_list[t9_0 + t];
            const int kh = t9 / 3, kw = t9 - kh * 3;
            const uint32_t dcol = tmem_base + (uint32_t)(t * p.Npad);
            const uint32_t toff = (uint32_t)(kh * DS_WH + kw);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t a_hi = desc_join(a_hiword, y_hi + (uint32_t)(j * 2 * DS_WT));
              const uint64_t b_hi = desc_join(b_hiword, x_hi + (uint32_t)(j * 2 * DS_WH) + toff);
              if (j == 0) umma_bf16(dcol, a_hi, b_hi, idesc, first);
              else umma_bf16_acc(dcol, a_hi, b_hi, idesc);
              umma_bf16_acc(dcol, desc_join(a_hiword, y_lo + (uint32_t)(j * 2 * DS_WT)), b_hi, idesc);
              umma_bf16_acc(dcol, a_hi, desc_join(b_hiword, x_lo + (uint32_t)(j * 2 * DS_WH) + toff), idesc);
            }
          }
          umma_commit(&empty_bar[slot]);
        }
        __syncwarp();
        continue;
      }
#pragma unroll 1
      for (int t = 0; t < nt9; ++t) {
        const int t9 = p.tap_list[t9_0 + t];
        const int kh = t9 / 3, kw = t9 - kh * 3;
        const uint32_t dcol = tmem_base + (uint32_t)(t * p.Npad);
        const uint32_t toff = (uint32_t)(kh * DS_WH + kw);            // 16-byte rows: halo line kh, voxel kw
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // K step j = lines 2j, 2j+1 of every dY plane (2 x 8 rows of 16 B) and the same lines, shifted, of the halo plane
          const uint64_t a_hi = desc_join(a_hiword, y_hi + (uint32_t)(j * 2 * DS_WT));
          const uint64_t b_hi = desc_join(b_hiword, x_hi + (uint32_t)(j * 2 * DS_WH) + toff);
          if (leader) {
            if (j == 0) umma_bf16(dcol, a_hi, b_hi, idesc, first);
            else umma_bf16_acc(dcol, a_hi, b_hi, idesc);
            if (parts == 2) {
              const uint64_t a_lo = desc_join(a_hiword, y_lo + (uint32_t)(j * 2 * DS_WT));
              const uint64_t b_lo = desc_join(b_hiword, x_lo + (uint32_t)(j * 2 * DS_WH) + toff);
              umma_bf16_acc(dcol, a_lo, b_hi, idesc);
              umma_bf16_acc(dcol, a_hi, b_lo, idesc);
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
    // accumulator row = (channel group g, d-shift s, channel): row group rg = 3 g + s, s <-> dY plane p-1+s <-> kd = 2 - s
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int rg = row >> 3;
    const int g = rg / 3, s = rg - g * 3;
    const int co = (g0 + g) * 8 + (row & 7);
    const int kd = 2 - s;
    const bool row_ok = g < p.Gt && co < p.Cout && ((p.kd_mask >> kd) & 1);
    mbar_wait(tmem_full_bar, 0, 430);
    tc_fence_after();
    if (niter > 0) {
      for (int t = 0; t < nt9; ++t) {
        const int tap = kd * 9 + p.tap_list[t9_0 + t];
        for (int j = 0; j < p.Npad; j += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * p.Npad + j), r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ci = p.ci0 + j + i;
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

struct DsPlan {
  int Gy_total, Gt, Gx, Npad, T9, ng9, mtiles, stages, tmem_cols;
  int Gx_total, slices;        // input channels are processed in `slices` launches of Gx groups (MMA N = 8 Gx <= 128)
  int y_bytes, x_bytes, stage_bytes;
  size_t act_y, act_x, off_yh, off_yl, off_xh, off_xl, total, smem;
};

static bool make_ds_plan(const cfun_conv3d_desc* d, DsPlan& pl, int ntap = 9) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != 3 || d->kH != 3 || d->kW != 3 || d->pD != 1 || d->pH != 1 || d->pW != 1) return false;
  if (d->Hin < 8 || d->Win < 8) return false;
  pl.Gy_total = (int)cdiv(d->Cout, 8);
  pl.Gx_total = (int)align_up((size_t)d->Cin, 16) / 8;
  // wide inputs: slices of <= 80 channels (10 groups) keep two pipeline stages of X planes + 3 dY planes in shared memory
  pl.slices = pl.Gx_total <= 10 ? 1 : (int)cdiv(pl.Gx_total, 10);
  pl.Gx = (int)align_up((size_t)cdiv(pl.Gx_total, pl.slices), 2);
  pl.Npad = pl.Gx * 8;
  if (pl.Npad > 128 || pl.slices > 8) return false;
  pl.T9 = std::min(ntap, 512 / pl.Npad);
  pl.ng9 = (int)cdiv(ntap, pl.T9);
  pl.T9 = (int)cdiv(ntap, pl.ng9);
  int cols = 32;
  while (cols < pl.T9 * pl.Npad) cols <<= 1;
  if (cols > 512) return false;
  pl.tmem_cols = cols;
  pl.x_bytes = pl.Gx * DS_XP;
  // M = 128 always reads 16 row groups (32 KB) from the dY tile base; what lies behind the real 3*Gt planes only feeds
  // accumulator rows that are never read back, but it must be inside the allocation: slack after the ring
  const size_t budget = 227 * 1024 - 2048;
  pl.Gt = 0;
  for (int gt = std::min(pl.Gy_total, DS_MAX_GT); gt >= 1; --gt) {
    const size_t yb = (size_t)gt * 3 * DS_YP;
    const size_t stage = align_up(2 * (yb + (size_t)pl.x_bytes), 1024);
    const size_t slack = 32 * 1024;
    if (2 * stage + slack <= budget) {
      pl.Gt = gt;
      pl.y_bytes = (int)yb;
      pl.stage_bytes = (int)stage;
      pl.stages = (int)std::min<size_t>(4, (budget - slack) / stage);
      pl.smem = 2048 + pl.stages * stage + slack;
      break;
    }
  }
  if (pl.Gt == 0) return false;
  pl.mtiles = (int)cdiv(pl.Gy_total, pl.Gt);
  // balance the M tiles (e.g. 10 groups -> 2 x 5, 6 groups -> 2 x 3)
  pl.Gt = (int)cdiv(pl.Gy_total, pl.mtiles);
  pl.y_bytes = pl.Gt * 3 * DS_YP;
  pl.act_y = align_up((size_t)pl.Gy_total * d->N * (d->Dout + 2) * d->Hout * d->Wout * 16, 1024);
  pl.act_x = align_up((size_t)pl.Gx_total * d->N * (d->Din + 2) * d->Hin * d->Win * 16, 1024);
  pl.off_yh = 0; pl.off_yl = pl.act_y; pl.off_xh = 2 * pl.act_y; pl.off_xl = 2 * pl.act_y + pl.act_x;
  pl.total = 2 * pl.act_y + 2 * pl.act_x + 2048;
  return true;
}

bool ds_supported(const cfun_conv3d_desc* d) {
  const char* e = getenv("CFUN_TC_WGDS");        // "0" falls back to the kd-split halo kernel (A/B measurements)
  if (e && e[0] == '0') return false;
  const char* h = getenv("CFUN_TC_HALO");
  if (h && h[0] == '0') return false;
  DsPlan pl;
  if (!make_ds_plan(d, pl)) return false;
  return d->Cin >= 16 && (d->Cin & 3) == 0 && d->Cout >= 8 && (d->Cout & 3) == 0;
}
size_t ds_workspace(const cfun_conv3d_desc* d) {
  DsPlan pl;
  return make_ds_plan(d, pl) ? pl.total : 0;
}

static int encode_gp_map_ds(CUtensorMap* m, void* base, int W, int H, long long planes, int G, int box_w8, int box_h,
                            int box_p, int box_g) {
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)G};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)box_w8, (cuuint32_t)box_h, (cuuint32_t)box_p, (cuuint32_t)box_g};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = get_tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(group-planar, d-stacked) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  return CFUN_OK;
}

// yh/yl/xh/xl: group-planar split-bf16 packs (pack_act_gp) of dY (Gy_total groups) and X (Gx groups)
// gy_pack: channel groups the dY pack was written with (>= Gy_total; the fused backward shares the data gradient's pack,
// whose group count is rounded up to a multiple of 2 -- the extra group is zeros)
int ds_launch(const cfun_conv3d_desc* d, const DsPlan& pl, __nv_bfloat16* yh, __nv_bfloat16* yl, __nv_bfloat16* xh,
              __nv_bfloat16* xl, float* dw, bool split, int gy_pack, cudaStream_t st, int tap_mask = 0x1FF, int kd_mask = 7) {
  CUtensorMap myh, myl, mxh, mxl;
  int rc;
  const long long planes = (long long)d->N * (d->Din + 2);
  if ((rc = encode_gp_map_ds(&myh, yh, d->Wout, d->Hout, planes, gy_pack, DS_WT * 8, DS_HT, 3, pl.Gt)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&myl, split ? yl : yh, d->Wout, d->Hout, planes, gy_pack, DS_WT * 8, DS_HT, 3, pl.Gt)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&mxh, xh, d->Win, d->Hin, planes, pl.Gx_total, DS_WH * 8, DS_HH, 1, pl.Gx)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&mxl, split ? xl : xh, d->Win, d->Hin, planes, pl.Gx_total, DS_WH * 8, DS_HH, 1, pl.Gx)) != CFUN_OK) return rc;

  DsParams p;
  p.N = d->N; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cout = d->Cout; p.Cin = d->Cin;
  p.Gt = pl.Gt; p.Gx = pl.Gx; p.Npad = pl.Npad; p.T9 = pl.T9; p.ng9 = pl.ng9;
  p.ntap = 0;
  for (int t = 0; t < 9; ++t) { p.tap_list[t] = 0; if ((tap_mask >> t) & 1) p.tap_list[p.ntap++] = t; }
  p.kd_mask = kd_mask;
  p.tilesH = (int)cdiv(d->Hin, DS_HT); p.tilesW = (int)cdiv(d->Win, DS_WT);
  p.nsplit = split ? 3 : 1;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
  p.y_bytes = pl.y_bytes; p.x_bytes = pl.x_bytes; p.stage_bytes = pl.stage_bytes;
  p.units_total = (long long)d->N * d->Din * p.tilesH * p.tilesW;
  const int ygroups = pl.ng9 * pl.mtiles;
  long long ctas = std::max<long long>(1, num_sms() / ygroups);
  ctas = std::min<long long>(ctas, cdiv(p.units_total, 4));
  p.units_per_cta = cdiv(p.units_total, ctas);
  ctas = cdiv(p.units_total, p.units_per_cta);
  p.dw = dw;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_ds_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_ds_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)ctas, (unsigned)pl.ng9, (unsigned)pl.mtiles);
  const char* le = getenv("CFUN_TC_LEAN");
  const bool lean = split && le && le[0] == '1';
  for (int sl = 0; sl < pl.slices; ++sl) {     // groups beyond Gx_total in the last slice are TMA zero fill
    p.ci0 = sl * pl.Gx * 8;
    if (lean) conv_tc_wgrad_ds_kernel<true><<<grid, DS_THREADS, pl.smem, st>>>(myh, myl, mxh, mxl, p);
    else conv_tc_wgrad_ds_kernel<false><<<grid, DS_THREADS, pl.smem, st>>>(myh, myl, mxh, mxl, p);
    CFUN_LAUNCH_CHECK();
  }
  return CFUN_OK;
}

int ds_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, int nsplit,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  DsPlan pl;
  CFUN_CHECK_ARG(make_ds_plan(d, pl));
  CFUN_CHECK_ARG(x && dy && dw && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d d-stacked wgrad: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yh);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yl);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xh);
  __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xl);
  int prc;
  if ((prc = launch_pack_act_gp(dy, yh, split ? yl : nullptr, d->N, d->Dout, d->Hout, d->Wout, d->Cout, pl.Gy_total, st)) != CFUN_OK) return prc;
  if ((prc = launch_pack_act_gp(x, xh, split ? xl : nullptr, d->N, d->Din, d->Hin, d->Win, d->Cin, pl.Gx_total, st)) != CFUN_OK) return prc;
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * 27, st));
  int rc = ds_launch(d, pl, yh, yl, xh, xl, dw, split, pl.Gy_total, st);
  if (rc != CFUN_OK) return rc;
  if (dbias) return simt_bias_grad(dy, (long long)d->N * d->Dout * d->Hout * d->Wout, d->Cout, dbias, st);
  return CFUN_OK;
}

// weight gradient from ready-made packs (fused backward): xh/xl = forward's X pack (align16(Cin)/8 groups), yh/yl = the data
// gradient's dY pack (gy_pack groups)
int ds_bwd_weight_packed(const cfun_conv3d_desc* d, __nv_bfloat16* yh, __nv_bfloat16* yl, int gy_pack, __nv_bfloat16* xh,
                         __nv_bfloat16* xl, float* dw, cudaStream_t st) {
  DsPlan pl;
  CFUN_CHECK_ARG(make_ds_plan(d, pl));
  CFUN_CHECK_ARG(yh && yl && xh && xl && dw && gy_pack >= pl.Gy_total);
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * 27, st));
  return ds_launch(d, pl, yh, yl, xh, xl, dw, true, gy_pack, st);
}

// Sub-kernel weight gradient: only the (kh,kw) taps of tap_mask and the kd planes of kd_mask are computed and written
// (dw is zero elsewhere).  conv_s2d.cu uses it with {kh,kw in {0,1}} x {kd in {0,1}}: the 2x2x2 space-to-depth kernel is
// the corresponding corner of a 3x3x3 / pad-1 kernel.
bool ds_masked_supported(const cfun_conv3d_desc* d, int tap_mask) {
  DsPlan pl;
  return make_ds_plan(d, pl, __builtin_popcount(tap_mask & 0x1FF)) && d->Cin >= 16 && (d->Cin & 3) == 0 && d->Cout >= 8 && (d->Cout & 3) == 0;
}
size_t ds_masked_workspace(const cfun_conv3d_desc* d, int tap_mask) {
  DsPlan pl;
  return make_ds_plan(d, pl, __builtin_popcount(tap_mask & 0x1FF)) ? pl.total : 0;
}
int ds_conv_bwd_weight_masked(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, int tap_mask, int kd_mask,
                              void* ws, size_t ws_bytes, cudaStream_t st) {
  DsPlan pl;
  CFUN_CHECK_ARG(make_ds_plan(d, pl, __builtin_popcount(tap_mask & 0x1FF)));
  CFUN_CHECK_ARG(x && dy && dw && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d d-stacked wgrad (masked): workspace too small"); return CFUN_ERR_WORKSPACE; }
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yh);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_yl);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xh);
  __nv_bfloat16* xl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_xl);
  int rc;
  if ((rc = launch_pack_act_gp(dy, yh, yl, d->N, d->Dout, d->Hout, d->Wout, d->Cout, pl.Gy_total, st)) != CFUN_OK) return rc;
  if ((rc = launch_pack_act_gp(x, xh, xl, d->N, d->Din, d->Hin, d->Win, d->Cin, pl.Gx_total, st)) != CFUN_OK) return rc;
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * 27, st));
  return ds_launch(d, pl, yh, yl, xh, xl, dw, true, pl.Gy_total, st, tap_mask, kd_mask);
}

int tc_debug_read_ds(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
