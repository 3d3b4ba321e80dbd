// tcgen05 weight gradient of a 3x3x3 / stride 1 / pad 1 convolution: kd stacked along M, kw stacked along N.
//
//   dW[kd][kh][kw][co][ci] = sum_{n,d,h,w} dY[n,d,h,w,co] * X[n, d+kd-1, h+kh-1, w+kw-1, ci]
//
// Both operands come from the group-planar split-bf16 pack [C/8][N*(D+2)][H][W][8ch] (zero planes at d = -1, D) that the
// forward / data-gradient kernels use, and both are MN-major un-swizzled UMMA operands (16-byte rows = one voxel of one
// 8-channel group; 8 consecutive rows = 8 voxels of a line = one K atom; LBO = line pitch, SBO = plane pitch).
// The work unit is one HT x 8 tile of one X plane p:
//   * A (M = 128 accumulator rows): the THREE dY planes d = p-1, p, p+1 (kd = 2, 1, 0) of up to 5 channel groups, one TMA
//     box {8 w, HT h, 3 planes, Gt groups}: row = (channel group, kd, channel) -> 120 of 128 rows used at 40 channels.
//   * B (N = KW * 8 Gx columns): KW copies of the X tile, shifted by kw - 1 voxels along w (a TMA inner coordinate may be
//     any multiple of 16 bytes = one voxel), each copy Gx planes of (HT + 2) x 8 voxels: column = (kw, channel group,
//     channel).  The kh taps stay free views (start address + kh lines), each kh owns N fp32 TMEM columns.
// Why: these kernels are bound by shared-memory bandwidth, not by the tensor pipe -- every tcgen05.mma re-reads its 4 KB A
// tile and its N x 32 B B tile from shared memory at 128 B/clk, and the TMA fills share that port (measured: the previous
// kernel, 9 (kh,kw) taps x N = Cin <= 48, ran at 1.10 x that model: 11.9 k cycles per 128 voxels of a 40->40 layer).
// Stacking kw along N needs 3 x fewer MMAs for the same products: (3 x 8 x 3) x (4 KB + 4 KB) instead of (9 x 8 x 3) x
// (4 KB + 1.5 KB) per 128 voxels, at the price of writing the X tile three times.
// Voxels are split across CTAs (split-K) and flushed once with fp32 atomics into dW (zero-initialised by the caller side).
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int DS_THREADS = 192;
constexpr int DS_WT = 8;                         // tile width (voxels) = one K atom
constexpr int DS_LINE = DS_WT * 16;              // 128 B: one line of one channel group
constexpr int DS_MAX_GT = 5;                     // dY channel groups per M tile: 5 x 3 kd = 15 of 16 row groups (3 x 5 kd for 5^3 kernels)
constexpr int DS_MAX_STAGES = 6;

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G,
                       cudaStream_t st);   // conv_tc_halo.cu
int launch_pack_act_gp_pad(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, int P,
                           cudaStream_t st);
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
  int HT;                     // tile height (8 or 16 lines): K = HT * 8 voxels per unit, HT / 2 MMA K-steps
  int Gt;                     // dY channel groups per M tile (<= 5)
  int Gx;                     // X channel groups of this launch's Cin slice
  int slices;                 // Cin slices of Gx groups = gridDim.y (slice s: input channels from s * 8 Gx)
  int KS;                     // kernel size (3 or 5): KS dY planes are stacked along M, taps = KS^3, zero padding KS / 2
  int nkw, nkh;               // stacked kw copies / kh views
  int kw_list[5], kh_list[5];
  int kd_mask;                // bit kd set = this kd plane of dW is wanted
  int Ntot;                   // MMA N: 8 * Gx * nkw rounded up to 16 (the padding columns read whatever follows, never stored)
  int tilesH, tilesW;
  int nsplit, stages;
  int tmem_cols;
  int yp, xp;                 // plane pitches: HT * 128 (dY), (HT + 2) * 128 (X)
  int y_bytes, x_bytes, stage_bytes;
  long long units_total, units_per_cta;
  float* dw;
};

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
  const int g0 = blockIdx.z * p.Gt;                      // first dY channel group of this M tile
  const int ci0 = blockIdx.y * p.Gx * 8;                 // first input channel of this N slice
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
        const int pad = p.KS >> 1;
        const int plane0 = n * (p.D + 2 * pad) + pl;     // padded index of dY plane pl - pad (and of X plane pl, minus pad)
        for (int part = 0; part < parts; ++part) {
          tma_load_4d_ds(part == 0 ? &map_yh : &map_yl, &full_bar[slot], sb + (size_t)part * p.y_bytes, wb * DS_WT * 8,
                         hb * p.HT, plane0, g0);        // box: 8 w x HT h x KS planes x Gt groups
          uint8_t* xb = sb + (size_t)parts * p.y_bytes + (size_t)part * p.x_bytes;
          for (int k = 0; k < p.nkw; ++k)
            tma_load_4d_ds(part == 0 ? &map_xh : &map_xl, &full_bar[slot], xb + (size_t)k * p.Gx * p.xp,
                           (wb * DS_WT - pad + p.kw_list[k]) * 8, hb * p.HT - pad, plane0 + pad, ci0 >> 3);
        }
      }
    }
  } else if (warp == 1) {
    // descriptors = constant high word + 16-byte offsets in the low word; one leader region per unit
    const uint32_t leader = elect_one();
    // both operands MN-major (bits 15, 16), bf16 x bf16 -> fp32, M = 128, N = Ntot
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Ntot >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiword = (uint32_t)(make_desc_mn_ds(0, DS_LINE, (uint32_t)p.yp) >> 32);
    const uint32_t b_hiword = (uint32_t)(make_desc_mn_ds(0, DS_LINE, (uint32_t)p.xp) >> 32);
    const uint32_t lbo = (uint32_t)(DS_LINE >> 4) << 16;
    const int ksteps = p.HT >> 1;
    for (int it = 0; it < niter; ++it) {
      const int slot = it % p.stages;
      mbar_wait(&full_bar[slot], (uint32_t)((it / p.stages) & 1), 420);
      tc_fence_after();
      if (leader) {
        const uint32_t sb = smem_u32(ring + (size_t)slot * p.stage_bytes);
        const uint32_t y_hi = desc_addr(sb) | lbo, y_lo = desc_addr(sb + (uint32_t)p.y_bytes) | lbo;
        const uint32_t x_hi = desc_addr(sb + (uint32_t)(parts * p.y_bytes)) | lbo;
        const uint32_t x_lo = desc_addr(sb + (uint32_t)(parts * p.y_bytes + p.x_bytes)) | lbo;
        const uint32_t first = it == 0 ? 0u : 1u;
#pragma unroll 1
        for (int t = 0; t < p.nkh; ++t) {
          const uint32_t dcol = tmem_base + (uint32_t)(t * p.Ntot);
          const uint32_t toff = (uint32_t)(p.kh_list[t] * (DS_LINE >> 4));          // 16-byte rows: line kh of the X tile
#pragma unroll 2
          for (int j = 0; j < ksteps; ++j) {
            // K step j = lines 2j, 2j+1 of every dY plane and lines 2j+kh, 2j+kh+1 of every X plane (2 x 8 voxels)
            const uint32_t ko = (uint32_t)(j * 2 * (DS_LINE >> 4));
            const uint64_t a_hi = desc_join(a_hiword, y_hi + ko);
            const uint64_t b_hi = desc_join(b_hiword, x_hi + ko + toff);
            if (j == 0) umma_bf16(dcol, a_hi, b_hi, idesc, first);
            else umma_bf16_acc(dcol, a_hi, b_hi, idesc);
            if (parts == 2) {
              umma_bf16_acc(dcol, desc_join(a_hiword, y_lo + ko), b_hi, idesc);
              umma_bf16_acc(dcol, a_hi, desc_join(b_hiword, x_lo + ko + toff), idesc);
            }
          }
        }
        umma_commit(&empty_bar[slot]);
      }
      __syncwarp();
    }
    if (leader) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // accumulator row = (channel group g, d-shift s, channel): row group rg = KS g + s, s <-> dY plane p-pad+s <-> kd = KS-1-s
    // accumulator column (within the kh block) = (kw copy, channel group, channel)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int rg = row >> 3;
    const int g = rg / p.KS, s = rg - g * p.KS;
    const int co = (g0 + g) * 8 + (row & 7);
    const int kd = p.KS - 1 - s;
    const int T = p.KS * p.KS * p.KS;
    const bool row_ok = g < p.Gt && co < p.Cout && ((p.kd_mask >> kd) & 1);
    const int ncopy = p.Gx * 8;
    mbar_wait(tmem_full_bar, 0, 430);
    tc_fence_after();
    if (niter > 0) {
      for (int t = 0; t < p.nkh; ++t) {
        for (int j = 0; j < p.nkw * ncopy; j += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * p.Ntot + j), r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = j + i;
              const int k = col / ncopy;
              const int ci = ci0 + (col - k * ncopy);
              if (k < p.nkw && ci < p.Cin) {
                const int tap = (kd * p.KS + p.kh_list[t]) * p.KS + p.kw_list[k];
                atomicAdd(p.dw + ((long long)co * p.Cin + ci) * T + tap, __uint_as_float(r[i]));
              }
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
  int Gy_total, Gt, Gx, mtiles, stages, tmem_cols, HT, Ntot, nkw, nkh, KS, ctas_per_sm;
  int Gx_total, slices;        // input channels are processed in `slices` launches of Gx groups
  int yp, xp, y_bytes, x_bytes, stage_bytes;
  size_t act_y, act_x, off_yh, off_yl, off_xh, off_xl, total, smem;
};

static bool make_ds_plan(const cfun_conv3d_desc* d, DsPlan& pl, int nkh = 0, int nkw = 0) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != d->kH || d->kD != d->kW || (d->kD != 3 && d->kD != 5)) return false;
  if (d->pD != d->kD / 2 || d->pH != d->kD / 2 || d->pW != d->kD / 2) return false;
  if (d->Hin < 8 || d->Win < 8) return false;
  pl.KS = d->kD;
  if (nkh == 0) nkh = pl.KS;
  if (nkw == 0) nkw = pl.KS;
  if (nkh < 1 || nkh > pl.KS || nkw < 1 || nkw > pl.KS) return false;
  pl.nkh = nkh; pl.nkw = nkw;
  pl.Gy_total = (int)cdiv(d->Cout, 8);
  pl.Gx_total = (int)align_up((size_t)d->Cin, 16) / 8;    // groups in the X pack (shared with the forward: 16-channel chunks)
  const int gx_real = (int)cdiv(d->Cin, 8);
  // nkh accumulator blocks of N = 8 * nkw * Gx (rounded up to 16) columns must fit the 512 TMEM columns, N <= 256
  int gx_max = 0;
  for (int g = 1; g <= 32; ++g) {
    const int n = (int)align_up((size_t)8 * nkw * g, 16);
    if (n <= 256 && nkh * n <= 512) gx_max = g;
  }
  if (gx_max == 0) return false;
  pl.slices = (int)cdiv(gx_real, gx_max);
  if (pl.slices > 16) return false;
  pl.Gx = (int)cdiv(gx_real, pl.slices);
  pl.Ntot = (int)align_up((size_t)8 * nkw * pl.Gx, 16);
  int cols = 32;
  while (cols < nkh * pl.Ntot) cols <<= 1;
  pl.tmem_cols = cols;
  pl.mtiles = (int)cdiv(pl.Gy_total, 16 / pl.KS);           // KS stacked planes per group: 5 groups (3^3) or 3 groups (5^3) per M tile
  pl.Gt = (int)cdiv(pl.Gy_total, pl.mtiles);              // balanced M tiles (10 groups -> 2 x 5, 6 -> 2 x 3)
  // tile height: 16 lines when at least 3 pipeline stages fit, else 8
  const size_t budget = 227 * 1024 - 2048;
  pl.HT = 0;
  for (int ht = (d->Hin >= 16 ? 16 : 8); ht >= 8; ht -= 8) {
    const int yp = ht * DS_LINE, xp = (ht + pl.KS - 1) * DS_LINE;
    const size_t yb = (size_t)pl.Gt * pl.KS * yp, xb = (size_t)nkw * pl.Gx * xp;
    const size_t stage = align_up(2 * (yb + xb), 1024);
    // M = 128 always reads 16 dY planes and N = Ntot may read one X plane more than was loaded: keep that inside the allocation
    const size_t slack = (size_t)16 * yp + xp;
    const int st = (int)std::min<size_t>(DS_MAX_STAGES, (budget - slack) / stage);
    if (st >= 3 || (ht == 8 && st >= 2)) {
      pl.HT = ht; pl.yp = yp; pl.xp = xp; pl.y_bytes = (int)yb; pl.x_bytes = (int)xb; pl.stage_bytes = (int)stage;
      pl.stages = st;
      pl.smem = 2048 + (size_t)st * stage + slack;
      break;
    }
  }
  if (pl.HT == 0) return false;
  // Two co-resident CTAs per SM where two pipeline stages of an 8- or 16-line tile fit in half the shared memory and the
  // accumulators in half the TMEM (thin layers): the CTAs' barrier hand-shakes overlap each other's MMAs (see conv_tc_halo.cu).
  // CFUN_DS_CTAS=1 keeps one CTA per SM (A/B measurements).
  pl.ctas_per_sm = 1;
  {
    const char* c1 = getenv("CFUN_DS_CTAS");
    const size_t cap2 = 111 * 1024 - 2048;
    if (!(c1 && c1[0] == '1') && pl.tmem_cols <= 256) {
      for (int ht = (d->Hin >= 16 ? 16 : 8); ht >= 8; ht -= 8) {
        const int yp = ht * DS_LINE, xp = (ht + pl.KS - 1) * DS_LINE;
        const size_t yb = (size_t)pl.Gt * pl.KS * yp, xb = (size_t)nkw * pl.Gx * xp;
        const size_t stage = align_up(2 * (yb + xb), 1024);
        const size_t slack = (size_t)16 * yp + xp;
        if (slack + 2 * stage > cap2) continue;
        const int st = (int)std::min<size_t>(DS_MAX_STAGES, (cap2 - slack) / stage);
        pl.HT = ht; pl.yp = yp; pl.xp = xp; pl.y_bytes = (int)yb; pl.x_bytes = (int)xb; pl.stage_bytes = (int)stage;
        pl.stages = st;
        pl.smem = 2048 + (size_t)st * stage + slack;
        pl.ctas_per_sm = 2;
        break;
      }
    }
  }
  const int pad = pl.KS / 2;
  pl.act_y = align_up((size_t)pl.Gy_total * d->N * (d->Dout + 2 * pad) * d->Hout * d->Wout * 16, 1024);
  pl.act_x = align_up((size_t)pl.Gx_total * d->N * (d->Din + 2 * pad) * d->Hin * d->Win * 16, 1024);
  pl.off_yh = 0; pl.off_yl = pl.act_y; pl.off_xh = 2 * pl.act_y; pl.off_xl = 2 * pl.act_y + pl.act_x;
  pl.total = 2 * pl.act_y + 2 * pl.act_x + 2048;
  return true;
}

static bool ds_shape_ok(const cfun_conv3d_desc* d) {
  if (d->kD == 5) return d->Cin <= 16 && d->Cout >= 8 && (d->Cout & 3) == 0;      // out_upscale_conv (8 -> 8): the pack has a scalar path
  return d->Cin >= 16 && (d->Cin & 3) == 0 && d->Cout >= 8 && (d->Cout & 3) == 0;
}

bool ds_supported(const cfun_conv3d_desc* d) {
  const char* e = getenv("CFUN_TC_WGDS");        // "0" falls back to the channel-major kernel (A/B measurements)
  if (e && e[0] == '0') return false;
  DsPlan pl;
  if (!make_ds_plan(d, pl)) return false;
  return ds_shape_ok(d);
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

// tap_mask (bits kh*KS+kw) must be a product set {kh} x {kw}
static bool split_tap_mask(int tap_mask, int KS, int* kh_list, int& nkh, int* kw_list, int& nkw) {
  int khm = 0, kwm = 0;
  for (int t = 0; t < KS * KS; ++t) if ((tap_mask >> t) & 1) { khm |= 1 << (t / KS); kwm |= 1 << (t % KS); }
  nkh = nkw = 0;
  for (int i = 0; i < 5; ++i) { kh_list[i] = kw_list[i] = 0; }
  for (int i = 0; i < KS; ++i) { if ((khm >> i) & 1) kh_list[nkh++] = i; if ((kwm >> i) & 1) kw_list[nkw++] = i; }
  for (int a = 0; a < nkh; ++a) for (int b = 0; b < nkw; ++b) if (!((tap_mask >> (kh_list[a] * KS + kw_list[b])) & 1)) return false;
  return nkh > 0 && nkw > 0;
}

// yh/yl/xh/xl: group-planar split-bf16 packs (pack_act_gp) of dY (gy_pack groups) and X (Gx_total groups)
// gy_pack: channel groups the dY pack was written with (>= Gy_total; the fused backward shares the data gradient's pack,
// whose group count is rounded up to a multiple of 2 -- the extra group is zeros)
int ds_launch(const cfun_conv3d_desc* d, const DsPlan& pl, __nv_bfloat16* yh, __nv_bfloat16* yl, __nv_bfloat16* xh,
              __nv_bfloat16* xl, float* dw, bool split, int gy_pack, cudaStream_t st, int tap_mask = -1, int kd_mask = -1) {
  CUtensorMap myh, myl, mxh, mxl;
  int rc;
  if (tap_mask < 0) tap_mask = (1 << (pl.KS * pl.KS)) - 1;
  if (kd_mask < 0) kd_mask = (1 << pl.KS) - 1;
  const long long planes = (long long)d->N * (d->Din + 2 * (pl.KS / 2));
  if ((rc = encode_gp_map_ds(&myh, yh, d->Wout, d->Hout, planes, gy_pack, DS_WT * 8, pl.HT, pl.KS, pl.Gt)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&myl, split ? yl : yh, d->Wout, d->Hout, planes, gy_pack, DS_WT * 8, pl.HT, pl.KS, pl.Gt)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&mxh, xh, d->Win, d->Hin, planes, pl.Gx_total, DS_WT * 8, pl.HT + pl.KS - 1, 1, pl.Gx)) != CFUN_OK) return rc;
  if ((rc = encode_gp_map_ds(&mxl, split ? xl : xh, d->Win, d->Hin, planes, pl.Gx_total, DS_WT * 8, pl.HT + pl.KS - 1, 1, pl.Gx)) != CFUN_OK) return rc;

  DsParams p;
  p.N = d->N; p.D = d->Din; p.H = d->Hin; p.W = d->Win; p.Cout = d->Cout; p.Cin = d->Cin;
  p.HT = pl.HT; p.Gt = pl.Gt; p.Gx = pl.Gx; p.Ntot = pl.Ntot; p.KS = pl.KS;
  CFUN_CHECK_ARG(split_tap_mask(tap_mask, pl.KS, p.kh_list, p.nkh, p.kw_list, p.nkw));
  CFUN_CHECK_ARG(p.nkh == pl.nkh && p.nkw == pl.nkw);
  p.kd_mask = kd_mask;
  p.tilesH = (int)cdiv(d->Hin, pl.HT); p.tilesW = (int)cdiv(d->Win, DS_WT);
  p.nsplit = split ? 3 : 1;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
  p.yp = pl.yp; p.xp = pl.xp;
  p.y_bytes = pl.y_bytes; p.x_bytes = pl.x_bytes; p.stage_bytes = pl.stage_bytes;
  p.units_total = (long long)d->N * d->Din * p.tilesH * p.tilesW;
  p.slices = pl.slices;
  // split-K over CTAs: every (M tile, Cin slice) pair owns its own dW block, so the more of those there are the fewer
  // voxel splits (and atomic flushes of the same block) are needed to fill the SMs
  long long ctas = std::max<long long>(1, (long long)pl.ctas_per_sm * num_sms() / (pl.mtiles * pl.slices));
  ctas = std::min<long long>(ctas, cdiv(p.units_total, 4));
  p.units_per_cta = cdiv(p.units_total, ctas);
  ctas = cdiv(p.units_total, p.units_per_cta);
  p.dw = dw;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)ctas, (unsigned)pl.slices, (unsigned)pl.mtiles);   // groups beyond Gx_total in the last slice are TMA zero fill
  timing_begin(st);
  conv_tc_wgrad_ds_kernel<<<grid, DS_THREADS, pl.smem, st>>>(myh, myl, mxh, mxl, p);
  timing_end(st);
  CFUN_LAUNCH_CHECK();
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
  if ((prc = launch_pack_act_gp_pad(dy, yh, split ? yl : nullptr, d->N, d->Dout, d->Hout, d->Wout, d->Cout, pl.Gy_total, pl.KS / 2, st)) != CFUN_OK) return prc;
  if ((prc = launch_pack_act_gp_pad(x, xh, split ? xl : nullptr, d->N, d->Din, d->Hin, d->Win, d->Cin, pl.Gx_total, pl.KS / 2, st)) != CFUN_OK) return prc;
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * pl.KS * pl.KS * pl.KS, st));
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
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin * pl.KS * pl.KS * pl.KS, st));
  return ds_launch(d, pl, yh, yl, xh, xl, dw, true, gy_pack, st);
}

// Sub-kernel weight gradient: only the (kh,kw) taps of tap_mask (a product set) and the kd planes of kd_mask are computed
// and written (dw is zero elsewhere).  conv_s2d.cu uses it with {kh,kw in {0,1}} x {kd in {0,1}}: the 2x2x2 space-to-depth
// kernel is the corresponding corner of a 3x3x3 / pad-1 kernel.
static bool make_masked_plan(const cfun_conv3d_desc* d, int tap_mask, DsPlan& pl) {
  int khl[5], kwl[5], nkh, nkw;
  if (!d || d->kD != 3 || !split_tap_mask(tap_mask & 0x1FF, 3, khl, nkh, kwl, nkw)) return false;
  return make_ds_plan(d, pl, nkh, nkw);
}
bool ds_masked_supported(const cfun_conv3d_desc* d, int tap_mask) {
  DsPlan pl;
  return make_masked_plan(d, tap_mask, pl) && ds_shape_ok(d);
}
size_t ds_masked_workspace(const cfun_conv3d_desc* d, int tap_mask) {
  DsPlan pl;
  return make_masked_plan(d, tap_mask, pl) ? pl.total : 0;
}
int ds_conv_bwd_weight_masked(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, int tap_mask, int kd_mask,
                              void* ws, size_t ws_bytes, cudaStream_t st) {
  DsPlan pl;
  CFUN_CHECK_ARG(make_masked_plan(d, tap_mask, pl));
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
  return ds_launch(d, pl, yh, yl, xh, xl, dw, true, pl.Gy_total, st, tap_mask & 0x1FF, kd_mask);
}

int tc_debug_read_ds(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
