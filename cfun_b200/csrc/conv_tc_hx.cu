// tcgen05 3x3x3 / stride 1 / pad 1 convolution, halo-resident activations + CLUSTER-MULTICAST weights ("hx").
//
// Successor of conv_tc_halo.cu (same data layout, same MMA views: read that header first).  What the profiler said about
// the two older kernels: conv_tc_halo streams the whole weight tensor (249 KB for 40->40) through every CTA for every
// 128-voxel tile -- 10 GB of L2->SM traffic per 96^3 launch -- and conv_tc re-loads activations AND weights per tap, which
// makes the 128->256 RPN conv L2-bound at 2.6x its MMA time.  Here
//   * the activation halo of a 1 x 16 x 8 slab is loaded once per 16-channel K chunk into a 3-slot ring (any Cin);
//   * weight stages are loaded ONCE PER CLUSTER: each of the 2 / 4 CTAs of a thread-block cluster fetches 1/2 / 1/4 of a
//     stage with cp.async.bulk ... .multicast::cluster, the bytes land in all CTAs' rings and complete_tx on all their
//     `full` barriers; a stage is recycled when every CTA's MMA warp has committed on every CTA's `empty` barrier
//     (tcgen05.commit ... .multicast::cluster), so the barrier count is the cluster size;
//   * output channels are tiled by <= 128 (accumulator pairs [hi*hi | hi*lo], double buffered in 512 TMEM columns), the
//     N tiles are the outer loop so that the whole cluster works on the same weights at the same time; CTAs that run out
//     of tiles execute dummy iterations (same loads, same MMAs, no stores) to keep the cluster in lock-step;
//   * the MMA issue loop is warp-uniform with hoisted descriptors (tc_ptx.cuh).
// Forward and data gradient (flipped / transposed weights) share the kernel.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int HX_THREADS = 224;      // warp 0: halo producer, 1: MMA issuer, 2..5: epilogue, 6: weight producer
constexpr int HX_HT = 16, HX_WT = 8, HX_HH = 18, HX_WH = 10;
constexpr int HX_PLANE_DATA = 3 * HX_HH * HX_WH * 16;                 // 8640 B: one 8-channel group, one part
constexpr int HX_PLANE = (HX_PLANE_DATA + 127) / 128 * 128;          // 8704
constexpr int HX_ASLOTS = 3;
constexpr int HX_MAX_BSTAGES = 6;

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G,
                       cudaStream_t st);   // conv_tc_halo.cu

__device__ __forceinline__ uint64_t make_desc_il(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void hx_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void hx_bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void hx_tma_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void hx_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void hx_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct HxParams {
  int N, D, H, W, Cout;       // output extents == input extents (pad 1, stride 1)
  int CPC;                    // 16-channel K chunks
  int Npad, ntn;              // output channels per N tile (multiple of 16, <= 128), number of N tiles
  int tilesH, tilesW;
  long long ntiles;           // spatial tiles
  int iters;                  // spatial iterations per CTA (ceil(ntiles / grid)): identical for every CTA of a cluster
  int nsplit, tmem_cols, epi;
  int TPS, bstages, b_stage_bytes;   // taps per weight stage (9 or 3), ring depth, stage bytes
  int cluster;                // CTAs per cluster (1, 2, 4)
  int tapmask;                // bit t (0..26, kernel's own tap order) set = tap t is live (MASKED instantiations only)
  int debug;                  // bring-up (CFUN_HX_DEBUG): 1 = every CTA loads whole stages itself (no multicast loads)
  const float* bias;
  float* y;
  const uint8_t* wpack;       // [ntile][chunk][kd][tap9][kgroup2][part][Npad][8] bf16 (hi rows, then lo rows)
  double* stat_acc;           // optional [N][Cout][2]: per-(sample, channel) sum and sum of squares of y (InstanceNorm statistics)
};

// TPS: taps per weight stage, 9 (one kd plane) or 3 (one (kd, kh) row) -- compile-time so the issue loop is straight-line.
// MASKED: only the taps of p.tapmask are multiplied and weight stages without a live tap are neither loaded nor waited for
// (conv_s2d.cu: a 2x2x2 kernel embedded as one corner of the 3x3x3 stencil).
// LEAN (split mode, un-masked: the default there): one leader region per weight stage instead of one per tap -- straight-line
// MMA issue, 3.5 instead of 9.5 SASS instructions per MMA.  Bit-identical to the per-tap form; worth 0-4 % on B200, because
// the kernel is bound by operand fetch from shared memory, not by issue (profiles/r02_lean_validation.txt).
template <int TPS, bool MASKED, bool LEAN>
__global__ void __launch_bounds__(HX_THREADS, 1)
conv_tc_hx_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l, const HxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);       // [HX_ASLOTS]
  uint64_t* a_empty = a_full + HX_ASLOTS;
  uint64_t* b_full = a_empty + HX_ASLOTS;                         // [HX_MAX_BSTAGES]
  uint64_t* b_empty = b_full + HX_MAX_BSTAGES;
  uint64_t* t_full = b_empty + HX_MAX_BSTAGES;                    // [2]
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  uint8_t* base = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int a_slot_bytes = parts * 2 * HX_PLANE;                  // one K chunk: 2 channel groups x parts
  const int nrows = parts * p.Npad;                               // weight rows per K group: hi rows then lo rows
  uint8_t* a_ring = base;
  uint8_t* b_ring = base + (size_t)HX_ASLOTS * a_slot_bytes;
  constexpr int spc = 27 / TPS;                                   // weight stages per K chunk
  const uint32_t rank = p.cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
  const uint16_t self_mask = (uint16_t)(1u << rank);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    if (parts == 2) prefetch_tmap(&map_l);
    for (int i = 0; i < HX_ASLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.bstages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], (p.debug & 2) ? 1u : (uint32_t)p.cluster); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();      // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== halo producer (TMA tensor): runs up to HX_ASLOTS chunks ahead, independent of the weight ring =========
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int nt = 0; nt < p.ntn; ++nt) {
        for (int it = 0; it < p.iters; ++it) {
          long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
          if (t >= p.ntiles) t = p.ntiles - 1;                 // dummy iteration: keeps the cluster in lock-step
          const int wb = (int)(t % p.tilesW); t /= p.tilesW;
          const int hb = (int)(t % p.tilesH); t /= p.tilesH;
          const int d = (int)(t % p.D);
          const int n = (int)(t / p.D);
          const int c_w = (wb * HX_WT - 1) * 8;                 // inner coordinate in elements (multiple of 8 -> 16 B aligned)
          const int c_h = hb * HX_HT - 1;
          const int c_nd = n * (p.D + 2) + d;                   // padded plane index of d-1
          for (int c = 0; c < p.CPC; ++c) {
            mbar_wait(&a_empty[slot], ph ^ 1u, 510);
            mbar_arrive_expect_tx(&a_full[slot], (uint32_t)(parts * 2 * HX_PLANE_DATA));
            uint8_t* sl = a_ring + (size_t)slot * a_slot_bytes;
            for (int g = 0; g < 2; ++g) {
              hx_tma_4d(&map_h, &a_full[slot], sl + g * HX_PLANE, c_w, c_h, c_nd, 2 * c + g);
              if (parts == 2) hx_tma_4d(&map_l, &a_full[slot], sl + (2 + g) * HX_PLANE, c_w, c_h, c_nd, 2 * c + g);
            }
            if (++slot == HX_ASLOTS) { slot = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== weight producer (bulk copies, multicast across the cluster) =====================
    if (lane == 0) {
      const uint32_t slice = (uint32_t)(p.b_stage_bytes / p.cluster);
      int st = 0;
      uint32_t ph = 0;
      for (int nt = 0; nt < p.ntn; ++nt) {
        for (int it = 0; it < p.iters; ++it) {
          const uint8_t* src = p.wpack + (size_t)nt * p.CPC * spc * (size_t)p.b_stage_bytes + (size_t)rank * slice;
          for (int q = 0; q < p.CPC * spc; ++q, src += p.b_stage_bytes) {
            if (MASKED && !((p.tapmask >> ((q % spc) * TPS)) & ((1 << TPS) - 1))) continue;
            mbar_wait(&b_empty[st], ph ^ 1u, 520);
            mbar_arrive_expect_tx(&b_full[st], (uint32_t)p.b_stage_bytes);
            uint8_t* dst = b_ring + (size_t)st * p.b_stage_bytes + (size_t)rank * slice;
            if (p.cluster > 1 && !(p.debug & 1)) hx_bulk_load_mc(dst, src, slice, &b_full[st], cmask);
            else if (p.cluster > 1) hx_bulk_load(dst - (size_t)rank * slice, src - (size_t)rank * slice, (uint32_t)p.b_stage_bytes, &b_full[st]);
            else hx_bulk_load(dst, src, slice, &b_full[st]);
            if (++st == p.bstages) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    const uint32_t leader = elect_one();
    const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nrows >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
    const uint32_t a_hiword = (uint32_t)(make_desc_il(0, HX_PLANE, HX_WH * 16) >> 32);
    const uint32_t b_hiword = (uint32_t)(make_desc_il(0, (uint32_t)(nrows * 16), 128) >> 32);
    const uint32_t a_lbo = (uint32_t)(HX_PLANE >> 4) << 16, b_lbo = (uint32_t)nrows << 16;
    const uint32_t b_tap = (uint32_t)(2 * nrows);                 // 16-byte rows per tap in a weight stage
    int slot = 0, st = 0;
    uint32_t aph = 0, bph = 0;
    int local = 0;
    for (int nt = 0; nt < p.ntn; ++nt) {
      for (int it = 0; it < p.iters; ++it, ++local) {
        const int buf = local & 1;
        mbar_wait(&t_empty[buf], (uint32_t)(((local >> 1) & 1) ^ 1), 530);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(buf * nrows);
        uint32_t acc = 0;
        for (int c = 0; c < p.CPC; ++c) {
          mbar_wait(&a_full[slot], aph, 540);
          tc_fence_after();
          const uint32_t a_hi0 = desc_addr(a_base + (uint32_t)(slot * a_slot_bytes)) | a_lbo;
          const uint32_t a_lo0 = a_hi0 + (uint32_t)((2 * HX_PLANE) >> 4);
          for (int s = 0; s < spc; ++s) {
            if (MASKED && !((p.tapmask >> (s * TPS)) & ((1 << TPS) - 1))) continue;
            mbar_wait(&b_full[st], bph, 550);
            tc_fence_after();
            const uint32_t b0 = desc_addr(b_base + (uint32_t)(st * p.b_stage_bytes)) | b_lbo;
            const int tap0 = s * TPS;                             // first tap (0..26) of this stage
            const uint32_t a_s = (uint32_t)((tap0 / 9) * HX_HH * HX_WH + ((tap0 % 9) / 3) * HX_WH);   // halo row of (kd, kh0)
            if (LEAN) {
              if (leader) {
#pragma unroll
                for (int t = 0; t < TPS; ++t) {
                  const uint32_t aoff = a_s + (uint32_t)((t / 3) * HX_WH + (t % 3));
                  const uint64_t b_all = desc_join(b_hiword, b0 + (uint32_t)t * b_tap);
                  if (t == 0) umma_bf16(dcol, desc_join(a_hiword, a_hi0 + aoff), b_all, idesc_2n, acc);
                  else umma_bf16_acc(dcol, desc_join(a_hiword, a_hi0 + aoff), b_all, idesc_2n);
                  umma_bf16_acc(dcol, desc_join(a_hiword, a_lo0 + aoff), b_all, idesc_n);
                }
                if (p.cluster > 1) hx_commit_mc(&b_empty[st], cmask);
                else umma_commit(&b_empty[st]);
              }
              acc = 1;
              __syncwarp();
              if (++st == p.bstages) { st = 0; bph ^= 1u; }
              continue;
            }
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              if (MASKED) {
                if ((p.tapmask >> (tap0 + t)) & 1) {
                  const uint32_t aoff = a_s + (uint32_t)((t / 3) * HX_WH + (t % 3));
                  const uint64_t a_hi = desc_join(a_hiword, a_hi0 + aoff);
                  const uint64_t b_all = desc_join(b_hiword, b0 + (uint32_t)t * b_tap);
                  if (leader) {
                    umma_bf16(dcol, a_hi, b_all, idesc_2n, acc);
                    if (parts == 2) umma_bf16_acc(dcol, desc_join(a_hiword, a_lo0 + aoff), b_all, idesc_n);
                  }
                  acc = 1;
                }
              } else {
                const uint32_t aoff = a_s + (uint32_t)((t / 3) * HX_WH + (t % 3));
                const uint64_t a_hi = desc_join(a_hiword, a_hi0 + aoff);
                const uint64_t b_all = desc_join(b_hiword, b0 + (uint32_t)t * b_tap);
                if (leader) {
                  if (t == 0) umma_bf16(dcol, a_hi, b_all, idesc_2n, acc);   // [hi*hi | hi*lo] into columns [0,N) and [N,2N)
                  else umma_bf16_acc(dcol, a_hi, b_all, idesc_2n);
                  if (parts == 2) umma_bf16_acc(dcol, desc_join(a_hiword, a_lo0 + aoff), b_all, idesc_n);   // lo*hi
                }
              }
            }
            if (!MASKED) acc = 1;
            if (leader) {
              if (p.cluster > 1) hx_commit_mc(&b_empty[st], (p.debug & 2) ? self_mask : cmask);    // this CTA is done with its copy: tell every producer
              else umma_commit(&b_empty[st]);
            }
            __syncwarp();
            if (++st == p.bstages) { st = 0; bph ^= 1u; }
          }
          // inside a cluster the commits name their target CTA(s) by mask
          if (leader) { if (p.cluster > 1) hx_commit_mc(&a_empty[slot], self_mask); else umma_commit(&a_empty[slot]); }
          __syncwarp();
          if (++slot == HX_ASLOTS) { slot = 0; aph ^= 1u; }
        }
        if (leader) { if (p.cluster > 1) hx_commit_mc(&t_full[buf], self_mask); else umma_commit(&t_full[buf]); }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    const bool vec = (p.Cout & 3) == 0;
    int local = 0;
    for (int nt = 0; nt < p.ntn; ++nt) {
      const int ch0 = nt * p.Npad;
      for (int it = 0; it < p.iters; ++it, ++local) {
        long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
        const bool live = t < p.ntiles;
        if (!live) t = p.ntiles - 1;
        const int wb = (int)(t % p.tilesW); t /= p.tilesW;
        const int hb = (int)(t % p.tilesH); t /= p.tilesH;
        const int d = (int)(t % p.D);
        const int n = (int)(t / p.D);
        const int oh = hb * HX_HT + lh, ow = wb * HX_WT + lw;
        const bool ok = live && oh < p.H && ow < p.W;
        float* yrow = p.y + ((((long long)n * p.D + d) * p.H + oh) * p.W + ow) * (long long)p.Cout + ch0;
        const int buf = local & 1;
        mbar_wait(&t_full[buf], (uint32_t)((local >> 1) & 1), 560);
        tc_fence_after();
        for (int j = 0; j < p.Npad; j += 16) {
          if (ch0 + j >= p.Cout) break;
          uint32_t r[16], r2[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * nrows + j), r);
          if (parts == 2) tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * nrows + p.Npad + j), r2);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float f = __uint_as_float(r[i]);
            if (parts == 2) f += __uint_as_float(r2[i]);
            if ((p.epi & CFUN_EPI_BIAS) && ch0 + j + i < p.Cout) f += __ldg(p.bias + ch0 + j + i);
            if (p.epi & CFUN_EPI_RELU) f = fmaxf(f, 0.f);
            v[i] = ok ? f : 0.f;
          }
          if (ok) {
            if (vec && ch0 + j + 16 <= p.Cout) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yrow + j + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (ch0 + j + i < p.Cout) yrow[j + i] = v[i];
            }
          }
          if (p.stat_acc) warp_stats16(v, p.stat_acc + (long long)n * p.Cout * 2, ch0 + j, p.Cout, lane);
        }
        if ((p.debug & 8) && quad == 0 && live) {     // bring-up: raw operand words of this CTA's rings into channels 0..7 of voxel 0
          uint32_t r0[16];
          tmem_ld16(tmem_base + (uint32_t)(buf * nrows), r0);     // warp-collective
          tmem_ld_wait();
          if (lane == 0) {
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(a_ring);
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(b_ring);
            yrow[0] = (float)cluster_ctarank();
            yrow[1] = __uint_as_float(aw[(HX_HH * HX_WH + HX_WH + 1) * 4]);   // halo voxel (kd 1, kh 1, kw 1), channels 0..1
            yrow[2] = __uint_as_float(aw[(HX_HH * HX_WH + HX_WH + 1) * 4 + 1]);
            yrow[3] = __uint_as_float(bw[0]);
            yrow[4] = __uint_as_float(bw[1]);
            yrow[5] = __uint_as_float(r0[0]);
            yrow[6] = __uint_as_float(r0[1]);
            yrow[7] = (float)blockIdx.x;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) hx_mbar_arrive(&t_empty[buf]);     // 4 epilogue warps -> barrier count 4
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();      // no CTA leaves while a peer may still multicast / arrive into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// w (Cout, Cin, 27) fp32 -> [ntile][chunk][kd][tap9][kgroup2][part][Npad][8] bf16.  mode 1 = data gradient (rows = ci,
// k = co, taps mirrored).  One warp per (output row, group of 8 k): it reads the 8 x 27 source floats (one contiguous run
// of 216 in forward mode, 8 runs of 27 in data-gradient mode) coalesced into shared memory, then lanes 0..26 each split one
// tap's 8 values and store the 16-byte hi / lo rows.
__global__ void __launch_bounds__(256) pack_w_hx_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                                        int Cin, int Npad, int ntn, int CPC, int parts, int mode) {
  __shared__ float sm[8][8 * 27 + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kgroups = CPC * 2;
  const int units = ntn * Npad * kgroups;
  const int rows_src = mode == 0 ? Cout : Cin, k_src = mode == 0 ? Cin : Cout;
  float* buf = sm[warp];
  for (int u = blockIdx.x * 8 + warp; u < units; u += gridDim.x * 8) {
    const int nrow = u / kgroups, kgi = u - nrow * kgroups;
    const int k0 = kgi * 8;
    for (int t = lane; t < 216; t += 32) {
      const int j = t / 27, tap = t - j * 27;
      float v = 0.f;
      if (nrow < rows_src && k0 + j < k_src) {
        const long long idx = mode == 0 ? ((long long)nrow * Cin + (k0 + j)) * 27 + tap : ((long long)(k0 + j) * Cin + nrow) * 27 + tap;
        v = __ldg(w + idx);
      }
      buf[t] = v;
    }
    __syncwarp();
    if (lane < 27) {
      const int tap_o = mode == 0 ? lane : 26 - lane;
      const int kd = tap_o / 9, t9 = tap_o - kd * 9;
      const int nt = nrow / Npad, row = nrow - nt * Npad;
      const int c = kgi >> 1, kg = kgi & 1;
      __align__(16) __nv_bfloat16 h[8];
      __align__(16) __nv_bfloat16 l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(buf[j * 27 + lane], h[j], l[j]);
      const long long o = ((((((long long)nt * CPC + c) * 3 + kd) * 9 + t9) * 2 + kg) * parts) * Npad + row;     // 16-byte rows
      reinterpret_cast<uint4*>(out)[o] = *reinterpret_cast<const uint4*>(h);
      if (parts == 2) reinterpret_cast<uint4*>(out)[o + Npad] = *reinterpret_cast<const uint4*>(l);
    }
    __syncwarp();
  }
}

struct HxPlan {
  int Cs, Ct, N, D, H, W, Kp, G, CPC, Npad, ntn, tmem_cols, TPS, bstages, b_stage_bytes;
  size_t off_ah, off_al, off_w, total, act_bytes, w_bytes, smem;
};

static bool make_hx_plan(const cfun_conv3d_desc* d, int pass, HxPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != 3 || d->kH != 3 || d->kW != 3 || d->pD != 1 || d->pH != 1 || d->pW != 1) return false;
  if (pass == CFUN_PASS_FWD) { pl.Cs = d->Cin; pl.Ct = d->Cout; }
  else if (pass == CFUN_PASS_BWD_DATA) { pl.Cs = d->Cout; pl.Ct = d->Cin; }
  else return false;
  pl.N = d->N; pl.D = d->Din; pl.H = d->Hin; pl.W = d->Win;
  if (pl.H < 8 || pl.W < 8) return false;
  pl.Kp = (int)align_up((size_t)pl.Cs, 16);
  pl.G = pl.Kp / 8;
  pl.CPC = pl.Kp / 16;
  const int Np = (int)align_up((size_t)pl.Ct, 16);
  pl.ntn = (int)cdiv(Np, 128);
  pl.Npad = (int)align_up((size_t)cdiv(Np, pl.ntn), 16);
  int cols = 32;
  while (cols < 4 * pl.Npad) cols <<= 1;                     // [hi | lo] accumulator pairs, double buffered
  if (cols > 512) return false;
  pl.tmem_cols = cols;
  const int nrows = 2 * pl.Npad;
  pl.TPS = (9 * 2 * nrows * 16 <= 30 * 1024) ? 9 : 3;
  pl.b_stage_bytes = pl.TPS * 2 * nrows * 16;
  const size_t a_bytes = (size_t)HX_ASLOTS * 2 * 2 * HX_PLANE;
  const size_t budget = 227 * 1024 - 2048 - a_bytes;
  pl.bstages = (int)std::min<size_t>(HX_MAX_BSTAGES, budget / pl.b_stage_bytes);
  if (pl.bstages < 2) return false;
  pl.smem = 2048 + a_bytes + (size_t)pl.bstages * pl.b_stage_bytes;
  pl.act_bytes = align_up((size_t)pl.G * pl.N * (pl.D + 2) * pl.H * pl.W * 16, 1024);
  pl.w_bytes = align_up((size_t)pl.ntn * pl.CPC * 3 * 2 * 9 * 2 * pl.Npad * 16, 1024);
  pl.off_ah = 0; pl.off_al = pl.act_bytes; pl.off_w = 2 * pl.act_bytes;
  pl.total = 2 * pl.act_bytes + pl.w_bytes + 2048;
  return true;
}

bool hx_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_TC_HX");          // "0" falls back to conv_tc_halo / conv_tc (A/B measurements)
  if (e && e[0] == '0') return false;
  const char* h = getenv("CFUN_TC_HALO");
  if (h && h[0] == '0') return false;
  HxPlan pl;
  if (!make_hx_plan(d, pass, pl)) return false;
  return pl.Cs >= 16 && (pl.Cs & 3) == 0 && pl.Ct >= 8;
}
size_t hx_workspace(const cfun_conv3d_desc* d, int pass) {
  HxPlan pl;
  return make_hx_plan(d, pass, pl) ? pl.total : 0;
}

// cluster size and grid: the largest cluster (4, 2, 1) whose co-resident CTA count covers >= 90 % of the SMs
static void pick_cluster(size_t smem, long long ntiles, size_t w_pass_bytes, int& cluster, int& grid) {
  static int cached_cluster = 0, cached_grid = 0;
  static size_t cached_smem = 0;
  // Multicast halves / quarters the weight bytes pulled from L2.  Measured on B200 it is time-neutral at 2 CTAs and ~10 %
  // slower at 4 (fewer co-resident CTAs, lock-step coupling), so the default is pairs, and only where a pass over the
  // weights is large (>= 512 KB per tile); CFUN_HX_CLUSTER=1|2|4 forces a size.
  const char* e = getenv("CFUN_HX_CLUSTER");
  const int forced = e ? atoi(e) : (w_pass_bytes >= (512u << 10) ? 2 : 1);
  static int cached_forced = -1;
  if (cached_cluster == 0 || cached_smem != smem || cached_forced != forced) {
    cached_cluster = 1; cached_grid = num_sms(); cached_smem = smem; cached_forced = forced;
    for (int cl = 4; cl >= 2; cl >>= 1) {
      if (forced && cl != forced) continue;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(num_sms() / cl * cl));
      cfg.blockDim = dim3(HX_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, conv_tc_hx_kernel<9, false, false>, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
      const int g = std::min(nclusters * cl, num_sms() / cl * cl);
      if (g * 10 >= num_sms() * 9 || forced) { cached_cluster = cl; cached_grid = g; break; }
    }
    if (forced == 1) { cached_cluster = 1; cached_grid = num_sms(); }
  }
  cluster = cached_cluster;
  grid = cached_grid;
  if (ntiles < grid) grid = (int)(cdiv(ntiles, cluster) * cluster);
}

int hx_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               int tapmask = 0x7FFFFFF, double* stat_acc = nullptr);
int hx_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
            int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  return hx_conv_ex(d, pass, src, w, bias, dst, epi, nsplit, ws, ws_bytes, nullptr, nullptr, false, st);
}
size_t hx_pack_bytes(const cfun_conv3d_desc* d, int pass) {
  HxPlan pl;
  return make_hx_plan(d, pass, pl) ? pl.act_bytes : 0;
}
// ext_hi / ext_lo: see hl_conv_ex (conv_tc_halo.cu)
// tapmask: live taps of w (bit kd*9+kh*3+kw); masked taps are assumed to hold zero weights and are skipped
int hx_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               int tapmask, double* stat_acc) {
  HxPlan pl;
  CFUN_CHECK_ARG(make_hx_plan(d, pass, pl));
  CFUN_CHECK_ARG((src || ext_ready) && w && dst && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d hx: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  const int parts = split ? 2 : 1;
  __nv_bfloat16* ah = ext_hi ? ext_hi : reinterpret_cast<__nv_bfloat16*>(base + pl.off_ah);
  __nv_bfloat16* al = ext_hi ? ext_lo : reinterpret_cast<__nv_bfloat16*>(base + pl.off_al);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(base + pl.off_w);
  int rc;
  if (!(ext_hi && ext_ready))
    if ((rc = launch_pack_act_gp(src, ah, split ? al : nullptr, pl.N, pl.D, pl.H, pl.W, pl.Cs, pl.G, st)) != CFUN_OK) return rc;
  {
    const long long units = (long long)pl.ntn * pl.Npad * pl.CPC * 2;     // one warp each
    pack_w_hx_kernel<<<(unsigned)std::min<long long>(cdiv(units, 8), 16LL * num_sms()), 256, 0, st>>>(w, wp, d->Cout, d->Cin, pl.Npad, pl.ntn, pl.CPC, parts, pass == CFUN_PASS_BWD_DATA ? 1 : 0);
    CFUN_LAUNCH_CHECK();
  }
  static bool attr_set = false;
  if (!attr_set) {     // before the occupancy query of pick_cluster
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<9, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<9, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<9, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_hx_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  CUtensorMap mh, ml;
  for (int part = 0; part < 2; ++part) {
    void* b = part == 0 ? (void*)ah : (void*)(split ? al : ah);
    cuuint64_t dims[4] = {(cuuint64_t)pl.W * 8, (cuuint64_t)pl.H, (cuuint64_t)pl.N * (pl.D + 2), (cuuint64_t)pl.G};
    cuuint64_t strides[3] = {(cuuint64_t)pl.W * 16, (cuuint64_t)pl.H * pl.W * 16, (cuuint64_t)pl.N * (pl.D + 2) * pl.H * pl.W * 16};
    cuuint32_t box[4] = {HX_WH * 8, HX_HH, 3, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_tensor_map_encoder()(part == 0 ? &mh : &ml, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(hx halo) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  }
  HxParams p;
  p.N = pl.N; p.D = pl.D; p.H = pl.H; p.W = pl.W; p.Cout = pl.Ct;
  p.CPC = pl.CPC; p.Npad = pl.Npad; p.ntn = pl.ntn;
  p.tilesH = (int)cdiv(pl.H, HX_HT); p.tilesW = (int)cdiv(pl.W, HX_WT);
  p.ntiles = (long long)pl.N * pl.D * p.tilesH * p.tilesW;
  p.nsplit = split ? 3 : 1;
  p.tmem_cols = pl.tmem_cols;
  p.epi = epi; p.bias = bias; p.y = dst;
  p.wpack = reinterpret_cast<const uint8_t*>(wp);
  p.stat_acc = stat_acc;
  p.TPS = pl.TPS; p.bstages = pl.bstages;
  p.b_stage_bytes = split ? pl.b_stage_bytes : pl.b_stage_bytes / 2;
  int cluster, grid;
  pick_cluster(pl.smem, p.ntiles, (size_t)pl.CPC * 27 * 2 * (2 * pl.Npad) * 16, cluster, grid);
  p.cluster = cluster;
  { const char* e = getenv("CFUN_HX_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.iters = (int)cdiv(p.ntiles, grid);
  if (p.debug & 4) p.cluster = 1;          // bring-up: launched as a cluster, executed as independent CTAs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(HX_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const bool masked = (tapmask & 0x7FFFFFF) != 0x7FFFFFF;
  p.tapmask = 0;
  for (int t = 0; t < 27; ++t)      // the data gradient runs on mirrored taps (pack_w_hx_kernel mode 1)
    if ((tapmask >> t) & 1) p.tapmask |= 1 << (pass == CFUN_PASS_BWD_DATA ? 26 - t : t);
  const bool lean = split && !masked && !p.debug;   // validated bit-identical on B200 (profiles/r02_lean_validation.txt)
  timing_begin(st);
  if (pl.TPS == 9) {
    if (masked) CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<9, true, false>, mh, ml, p));
    else if (lean) CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<9, false, true>, mh, ml, p));
    else CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<9, false, false>, mh, ml, p));
  } else {
    if (masked) CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<3, true, false>, mh, ml, p));
    else if (lean) CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<3, false, true>, mh, ml, p));
    else CFUN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_hx_kernel<3, false, false>, mh, ml, p));
  }
  timing_end(st);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int tc_debug_read_hx(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
