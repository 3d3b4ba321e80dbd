// EXPERIMENTAL (opt-in: CFUN_TC_COL=1), kept as a measured negative result.  Validated on B200 with the torch-free driver
// (`tools/prof_driver.bin unet 1 cmpcol`, profiles/r01_column_pass_ab.txt): forward and data gradient are BIT-IDENTICAL to
// conv_tc_halo.cu (same MMA order per tile), no pipeline time-outs -- but 40->40 @ 4x96^3 takes 1.47 ms against 1.33 ms.
// So the remaining gap of the halo kernel is not the shared-memory write traffic this variant removes; the default dispatch
// never selects it, and tests/test_gpu_ops.py::test_conv3d_tcgen05_column_pass runs only with the switch set.
//
// tcgen05 3x3x3 / stride 1 / pad 1 convolution for thin layers, "column pass" variant of conv_tc_halo.cu.
//
// Why: conv_tc_halo_kernel sits at its shared-memory ceiling: per 128-voxel tile the UMMA operand fetches
// need 8.1 k cycles of the 128 B/clk port and the TMA writes of the tile's halo (104 KB) plus the re-streamed weights
// (249 KB) another 2.8 k.  Here a CTA processes a COLUMN of T = 5..6 tiles stacked along d in one pass:
//   * T accumulator pairs [hi*hi | hi*lo] live in TMEM at once (T * 2 * Npad <= 512 columns);
//   * per K chunk the T+2 halo planes of the column are streamed once (plane p feeds tile p-kd for kd = 0..2), instead of
//     3 planes per tile: (T+2)/(3T) of the activation bytes;
//   * the three kd weight stages of a chunk stay resident for the whole plane sweep: the weights are read once per column,
//     1/T of the bytes.
// Shared-memory writes per tile drop from 353 KB to ~98 KB (40->40, T = 5).  Tile t of a column completes after the last
// chunk's plane t+2, so the epilogue of tile t overlaps the MMAs of tiles t+1.. and the next column starts with tile 0.
// Same operand format (group-planar split-bf16 packs, pack_w_halo layout), same MMA views as conv_tc_halo.cu.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int HC_THREADS = 224;      // warp 0: halo-plane producer, 1: MMA issuer, 2..5: epilogue, 6: weight producer
constexpr int HC_HT = 16, HC_WT = 8, HC_HH = 18, HC_WH = 10;
constexpr int HC_PLANE_DATA = HC_HH * HC_WH * 16;                   // 2880 B: one halo plane of one 8-channel group
constexpr int HC_PLANE = (HC_PLANE_DATA + 127) / 128 * 128;        // 2944
constexpr int HC_ASLOTS = 4;
constexpr int HC_BSTAGES = 6;        // two chunks' worth of (chunk, kd) weight stages
constexpr int HC_MAX_T = 6;

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G,
                       cudaStream_t st);                                                          // conv_tc_halo.cu
int launch_pack_w_halo(const float* w, __nv_bfloat16* out, int Cout, int Cin, int Npad, int CPC, int parts, int mode,
                       cudaStream_t st);                                                          // conv_tc_halo.cu

__device__ __forceinline__ uint64_t make_desc_hc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void hc_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void hc_tma_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void hc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct HcParams {
  int N, D, H, W, Cout;
  int CPC, Npad, T;            // K chunks, MMA N, tiles per column
  int tilesH, tilesW, dblocks;
  long long nunits;            // columns: N * dblocks * tilesH * tilesW
  int nsplit, epi;
  const float* bias;
  float* y;
  const uint8_t* wpack;        // [chunk][kd][tap9][kgroup2][part][Npad][8] bf16 (pack_w_halo_kernel)
};

__global__ void __launch_bounds__(HC_THREADS, 1)
conv_tc_hc_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l, const HcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);       // [HC_ASLOTS]
  uint64_t* a_empty = a_full + HC_ASLOTS;
  uint64_t* b_full = a_empty + HC_ASLOTS;                         // [HC_BSTAGES]
  uint64_t* b_empty = b_full + HC_BSTAGES;
  uint64_t* t_full = b_empty + HC_BSTAGES;                        // [HC_MAX_T]
  uint64_t* t_empty = t_full + HC_MAX_T;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + HC_MAX_T);
  uint8_t* base = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int a_slot_bytes = parts * 2 * HC_PLANE;                  // one halo plane of one K chunk: 2 groups x parts
  const int nrows = parts * p.Npad;
  const int b_stage_bytes = 9 * 2 * nrows * 16;
  uint8_t* a_ring = base;
  uint8_t* b_ring = base + (size_t)HC_ASLOTS * a_slot_bytes;
  const int T = p.T;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    if (parts == 2) prefetch_tmap(&map_l);
    for (int i = 0; i < HC_ASLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < HC_BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < HC_MAX_T; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== halo-plane producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        long long t = u;
        const int wb = (int)(t % p.tilesW); t /= p.tilesW;
        const int hb = (int)(t % p.tilesH); t /= p.tilesH;
        const int db = (int)(t % p.dblocks);
        const int n = (int)(t / p.dblocks);
        const int c_w = (wb * HC_WT - 1) * 8, c_h = hb * HC_HT - 1;
        const int plane0 = n * (p.D + 2) + db * T;              // padded index of plane d0 - 1
        for (int c = 0; c < p.CPC; ++c) {
          for (int pl = 0; pl < T + 2; ++pl) {
            mbar_wait(&a_empty[slot], ph ^ 1u, 610);
            mbar_arrive_expect_tx(&a_full[slot], (uint32_t)(parts * 2 * HC_PLANE_DATA));
            uint8_t* sl = a_ring + (size_t)slot * a_slot_bytes;
            for (int g = 0; g < 2; ++g) {
              hc_tma_4d(&map_h, &a_full[slot], sl + g * HC_PLANE, c_w, c_h, plane0 + pl, 2 * c + g);
              if (parts == 2) hc_tma_4d(&map_l, &a_full[slot], sl + (2 + g) * HC_PLANE, c_w, c_h, plane0 + pl, 2 * c + g);
            }
            if (++slot == HC_ASLOTS) { slot = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== weight producer: one stage per (chunk, kd), resident for the whole plane sweep =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        for (int q = 0; q < p.CPC * 3; ++q) {
          mbar_wait(&b_empty[st], ph ^ 1u, 620);
          mbar_arrive_expect_tx(&b_full[st], (uint32_t)b_stage_bytes);
          hc_bulk_load(b_ring + (size_t)st * b_stage_bytes, p.wpack + (size_t)q * b_stage_bytes, (uint32_t)b_stage_bytes, &b_full[st]);
          if (++st == HC_BSTAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    const uint32_t leader = elect_one();
    const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nrows >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
    const uint32_t a_hiword = (uint32_t)(make_desc_hc(0, HC_PLANE, HC_WH * 16) >> 32);
    const uint32_t b_hiword = (uint32_t)(make_desc_hc(0, (uint32_t)(nrows * 16), 128) >> 32);
    const uint32_t a_lbo = (uint32_t)(HC_PLANE >> 4) << 16, b_lbo = (uint32_t)nrows << 16;
    const uint32_t b_tap = (uint32_t)(2 * nrows);
    int slot = 0;
    uint32_t aph = 0;
    uint32_t cc = 0;                       // chunks consumed so far: stage of (chunk, kd) = (3 cc + kd) % HC_BSTAGES
    uint32_t pass = 0;
    for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x, ++pass) {
      uint32_t started = 0;                // accumulators that already hold a partial sum in this pass
      for (int c = 0; c < p.CPC; ++c, ++cc) {
        for (int pl = 0; pl < T + 2; ++pl) {
          mbar_wait(&a_full[slot], aph, 630);
          tc_fence_after();
          const uint32_t a_hi0 = desc_addr(a_base + (uint32_t)(slot * a_slot_bytes)) | a_lbo;
          const uint32_t a_lo0 = a_hi0 + (uint32_t)((2 * HC_PLANE) >> 4);
#pragma unroll 1
          for (int kd = 0; kd < 3; ++kd) {
            const int t = pl - kd;                               // the tile of the column this (plane, kd) pair feeds
            if (t < 0 || t >= T) continue;
            const uint32_t sidx = 3u * cc + (uint32_t)kd;
            const int st = (int)(sidx % HC_BSTAGES);
            if (t == 0) {                                        // first use of weight stage (chunk, kd)
              mbar_wait(&b_full[st], (sidx / HC_BSTAGES) & 1u, 640);
              tc_fence_after();
            }
            if (c == 0 && kd == 0) {                             // first MMA into accumulator t in this pass
              mbar_wait(&t_empty[t], (pass & 1u) ^ 1u, 650);
              tc_fence_after();
            }
            const uint32_t b0 = desc_addr(b_base + (uint32_t)(st * b_stage_bytes)) | b_lbo;
            const uint32_t dcol = tmem_base + (uint32_t)(t * nrows);
            const uint32_t acc = (started >> t) & 1u;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const uint32_t aoff = (uint32_t)((t9 / 3) * HC_WH + (t9 % 3));
              const uint64_t a_hi = desc_join(a_hiword, a_hi0 + aoff);
              const uint64_t b_all = desc_join(b_hiword, b0 + (uint32_t)t9 * b_tap);
              if (leader) {
                if (t9 == 0) umma_bf16(dcol, a_hi, b_all, idesc_2n, acc);
                else umma_bf16_acc(dcol, a_hi, b_all, idesc_2n);
                if (parts == 2) umma_bf16_acc(dcol, desc_join(a_hiword, a_lo0 + aoff), b_all, idesc_n);
              }
            }
            started |= 1u << t;
            if (leader) {
              if (t == T - 1) umma_commit(&b_empty[st]);                       // last plane that needs this weight stage
              if (c == p.CPC - 1 && kd == 2) umma_commit(&t_full[t]);          // tile t of the column is complete
            }
            __syncwarp();
          }
          if (leader) umma_commit(&a_empty[slot]);
          __syncwarp();
          if (++slot == HC_ASLOTS) { slot = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    const bool vec = (p.Cout & 3) == 0;
    uint32_t pass = 0;
    for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x, ++pass) {
      long long tt = u;
      const int wb = (int)(tt % p.tilesW); tt /= p.tilesW;
      const int hb = (int)(tt % p.tilesH); tt /= p.tilesH;
      const int db = (int)(tt % p.dblocks);
      const int n = (int)(tt / p.dblocks);
      const int oh = hb * HC_HT + lh, ow = wb * HC_WT + lw;
      for (int t = 0; t < T; ++t) {
        const int d = db * T + t;
        const bool ok = d < p.D && oh < p.H && ow < p.W;
        float* yrow = p.y + ((((long long)n * p.D + d) * p.H + oh) * p.W + ow) * (long long)p.Cout;
        mbar_wait(&t_full[t], pass & 1u, 660);
        tc_fence_after();
        for (int j = 0; j < p.Npad; j += 16) {
          uint32_t r[16], r2[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * nrows + j), r);
          if (parts == 2) tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * nrows + p.Npad + j), r2);
          tmem_ld_wait();
          if (ok) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float f = __uint_as_float(r[i]);
              if (parts == 2) f += __uint_as_float(r2[i]);
              if ((p.epi & CFUN_EPI_BIAS) && j + i < p.Cout) f += __ldg(p.bias + j + i);
              if (p.epi & CFUN_EPI_RELU) f = fmaxf(f, 0.f);
              v[i] = f;
            }
            if (vec && j + 16 <= p.Cout) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yrow + j + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (j + i < p.Cout) yrow[j + i] = v[i];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) hc_mbar_arrive(&t_empty[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

struct HcPlan {
  int Cs, Ct, N, D, H, W, Kp, G, CPC, Npad, T;
  size_t off_ah, off_al, off_w, total, act_bytes, w_bytes, smem;
};

static bool make_hc_plan(const cfun_conv3d_desc* d, int pass, HcPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != 3 || d->kH != 3 || d->kW != 3 || d->pD != 1 || d->pH != 1 || d->pW != 1) return false;
  if (pass == CFUN_PASS_FWD) { pl.Cs = d->Cin; pl.Ct = d->Cout; }
  else if (pass == CFUN_PASS_BWD_DATA) { pl.Cs = d->Cout; pl.Ct = d->Cin; }
  else return false;
  pl.N = d->N; pl.D = d->Din; pl.H = d->Hin; pl.W = d->Win;
  if (pl.H < 8 || pl.W < 8 || pl.D < 4) return false;
  pl.Kp = (int)align_up((size_t)pl.Cs, 16);
  pl.G = pl.Kp / 8;
  pl.CPC = pl.Kp / 16;
  if (pl.CPC > 4) return false;
  pl.Npad = (int)align_up((size_t)pl.Ct, 16);
  const int nrows = 2 * pl.Npad;
  if (nrows > 96) return false;                                  // six resident weight stages must fit beside the plane ring
  pl.T = std::min(HC_MAX_T, 512 / nrows);
  pl.T = std::min(pl.T, pl.D);
  const size_t a_bytes = (size_t)HC_ASLOTS * 2 * 2 * HC_PLANE;
  const size_t b_bytes = (size_t)HC_BSTAGES * 9 * 2 * nrows * 16;
  pl.smem = 2048 + a_bytes + b_bytes;
  if (pl.smem > 227 * 1024) return false;
  pl.act_bytes = align_up((size_t)pl.G * pl.N * (pl.D + 2) * pl.H * pl.W * 16, 1024);
  pl.w_bytes = align_up((size_t)pl.CPC * 3 * 2 * 9 * 2 * pl.Npad * 16, 1024);
  pl.off_ah = 0; pl.off_al = pl.act_bytes; pl.off_w = 2 * pl.act_bytes;
  pl.total = 2 * pl.act_bytes + pl.w_bytes + 2048;
  return true;
}

bool hc_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_TC_COL");          // opt-in only (experimental, see the header of this file)
  if (!(e && e[0] == '1')) return false;
  HcPlan pl;
  if (!make_hc_plan(d, pass, pl)) return false;
  return pl.Cs >= 16 && (pl.Cs & 3) == 0 && pl.Ct >= 8;
}
size_t hc_workspace(const cfun_conv3d_desc* d, int pass) {
  HcPlan pl;
  return make_hc_plan(d, pass, pl) ? pl.total : 0;
}

int hc_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
            int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  HcPlan pl;
  CFUN_CHECK_ARG(make_hc_plan(d, pass, pl));
  CFUN_CHECK_ARG(src && w && dst && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d hc: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  const int parts = split ? 2 : 1;
  __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(base + pl.off_ah);
  __nv_bfloat16* al = reinterpret_cast<__nv_bfloat16*>(base + pl.off_al);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(base + pl.off_w);
  int rc;
  if ((rc = launch_pack_act_gp(src, ah, split ? al : nullptr, pl.N, pl.D, pl.H, pl.W, pl.Cs, pl.G, st)) != CFUN_OK) return rc;
  if ((rc = launch_pack_w_halo(w, wp, d->Cout, d->Cin, pl.Npad, pl.CPC, parts, pass == CFUN_PASS_BWD_DATA ? 1 : 0, st)) != CFUN_OK) return rc;
  CUtensorMap mh, ml;
  for (int part = 0; part < 2; ++part) {
    void* b = part == 0 ? (void*)ah : (void*)(split ? al : ah);
    cuuint64_t dims[4] = {(cuuint64_t)pl.W * 8, (cuuint64_t)pl.H, (cuuint64_t)pl.N * (pl.D + 2), (cuuint64_t)pl.G};
    cuuint64_t strides[3] = {(cuuint64_t)pl.W * 16, (cuuint64_t)pl.H * pl.W * 16, (cuuint64_t)pl.N * (pl.D + 2) * pl.H * pl.W * 16};
    cuuint32_t box[4] = {HC_WH * 8, HC_HH, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_tensor_map_encoder()(part == 0 ? &mh : &ml, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(hc plane) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  }
  HcParams p;
  p.N = pl.N; p.D = pl.D; p.H = pl.H; p.W = pl.W; p.Cout = pl.Ct;
  p.CPC = pl.CPC; p.Npad = pl.Npad; p.T = pl.T;
  p.tilesH = (int)cdiv(pl.H, HC_HT); p.tilesW = (int)cdiv(pl.W, HC_WT); p.dblocks = (int)cdiv(pl.D, pl.T);
  p.nunits = (long long)pl.N * p.dblocks * p.tilesH * p.tilesW;
  p.nsplit = split ? 3 : 1;
  p.epi = epi; p.bias = bias; p.y = dst;
  p.wpack = reinterpret_cast<const uint8_t*>(wp);
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const unsigned grid = (unsigned)std::min<long long>(p.nunits, num_sms());
  conv_tc_hc_kernel<<<grid, HC_THREADS, pl.smem, st>>>(mh, ml, p);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int tc_debug_read_hc(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
