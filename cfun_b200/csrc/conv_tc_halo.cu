// tcgen05 3x3x3 / stride 1 / pad 1 convolution with the input halo resident in shared memory (thin-channel layers).
//
// conv_tc.cu re-loads the shifted input box from L2 for every tap, which makes thin convs (Cin <= 64: the top levels of
// the U-Net mask branch, 92 % of the step's FLOPs) L2-bandwidth bound: 27 x 8 KB of activations per 128-voxel tile.
// Here one CTA tile is a 1 x 16 x 8 slab of output voxels; its 3 x 18 x 10 input halo is loaded ONCE per tile and the 27
// taps are 27 *views* of it, expressed purely through the UMMA shared-memory descriptor:
//   * activations are packed "group-planar": [Kp/8 groups][N*(D+2)][H][W][8 ch] split-bf16 (zero planes at d = -1, D), so a
//     TMA box (80 elements = 10 voxels x 8 ch, 18 rows, 3 planes) lands as one plane of 16-byte rows per 8-channel group;
//   * with the un-swizzled ("interleave") K-major canonical layout ((8,n),2):((1,SBO),LBO) a 128-row operand is 16 groups of 8
//     consecutive 16-byte rows; rows = w (contiguous), groups = h lines (SBO = 10*16 B), and tap (kd,kh,kw) of channel group g
//     is just start address = plane(g) + ((kd*18 + kh)*10 + kw)*16;
//   * every inner TMA coordinate is a multiple of 8 elements (16 B), the alignment rule the weight-gradient kernel ran into.
//
// Round 2 (what the measurements said): these kernels are bound by SHARED-MEMORY BANDWIDTH -- every tcgen05.mma re-reads
// its 4 KB A tile and N x 32 B of B from shared memory at 128 B/clk and the TMA fills share that port; the round-1 kernel ran
// at 1.1 x that model, and the issue-rate fix ("LEAN") was worth 0-4 %.  So this version removes bytes, not instructions:
//   * EXACT K.  The contraction index is (channel group, tap), 27 G entries of 8 channels; one K = 16 MMA step takes ANY two
//     entries, because the second half of a K-major operand sits at an arbitrary LBO from the first.  Entries are paired in
//     group-major order ((g,t),(g,t+1)), the pair straddling two groups has LBO = 2 planes - tap offset.  Cin = 20 (3
//     groups) needs 41 steps instead of 54 (K padded to 32), Cin = 40 (5 groups) 68 instead of 81 (K padded to 48).
//   * Only the G real channel groups are loaded (3 instead of 4 planes at 20 channels, 5 instead of 6 at 40), each with its own
//     full/empty barrier pair, so group g of the next tile streams in while groups g+1.. of this tile are being multiplied.
//   * RESIDENT WEIGHTS where they fit beside the halo (20->20, 40->20 dgrad: 84..126 KB): loaded once per CTA instead of
//     once per 128-voxel tile (110 KB per tile before); otherwise streamed through a ring of step-granular stages.
// Split-bf16 x3 arithmetic as in conv_tc.cu, issued as TWO MMAs per step: the weight tile holds the hi rows and the lo rows
// back to back, so  A_hi x [B_hi ; B_lo]  is one N = 2*Npad instruction (hi*hi and hi*lo land in adjacent accumulator
// column ranges, summed in the epilogue) and  A_lo x B_hi  a second N = Npad instruction.
// Persistent CTAs, TMEM accumulator double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// The data gradient is the same kernel with flipped / transposed weights.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace cfun {

constexpr int HL_THREADS = 224;                      // warp 0: halo producer, 1: MMA issuer, 2..5: epilogue, 6: weight producer
constexpr int HL_HT = 16, HL_WT = 8;                 // output slab 1 x 16 x 8
// kernel size KS = 3 (halo 3 x 18 x 10) or 5 (5 x 20 x 12: out_upscale_conv, mask_branch.py:87-88, at 192^3 in stage finetune)
__host__ __device__ constexpr int hl_hh(int KS) { return HL_HT + KS - 1; }
__host__ __device__ constexpr int hl_wh(int KS) { return HL_WT + KS - 1; }
__host__ __device__ constexpr int hl_plane_data(int KS) { return KS * hl_hh(KS) * hl_wh(KS) * 16; }      // one 8-channel group, one part
__host__ __device__ constexpr int hl_plane(int KS) { return (hl_plane_data(KS) + 127) / 128 * 128; }     // plane pitch (128 B aligned)
constexpr int HL_MAX_G = 8;                          // Cin <= 64
constexpr int HL_MAX_BSTAGES = 8;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, cudaStream_t st);

// ---- compile-time step table (G = real channel groups) -------------------------------------------------------------------
// entry e = (group g = e / 27, tap t = e % 27) lives at plane 2g (hi; lo = the next plane) + halo row of the tap; all in 16-byte units
__host__ __device__ constexpr uint32_t hl_entry_off16(int KS, int e) {
  return (uint32_t)(((2 * (e / (KS * KS * KS))) * hl_plane(KS) +
                     ((((e % (KS * KS * KS)) / (KS * KS)) * hl_hh(KS) + ((e % (KS * KS * KS)) % (KS * KS)) / KS) * hl_wh(KS) +
                      (e % (KS * KS * KS)) % KS) * 16) >> 4);
}
// A-descriptor low word of step s relative to the halo slot: start address >> 4 | LBO >> 4 << 16.  With an odd entry count the
// last step pairs the previous tap (multiplied by zero weights, see pack_w_halo_kernel) with the last real entry, so that both
// K halves read initialised shared memory.
__host__ __device__ constexpr uint32_t hl_a_word(int KS, int G, int s) {
  return 2 * s + 1 < KS * KS * KS * G
             ? (hl_entry_off16(KS, 2 * s) | ((hl_entry_off16(KS, 2 * s + 1) - hl_entry_off16(KS, 2 * s)) << 16))
             : ((hl_entry_off16(KS, 2 * s) - 1u) | (1u << 16));
}
// segment g = the steps whose first entry lies in group g: [seg_begin(g), seg_begin(g + 1)); 14, 13, 14, ... steps for T = 27
// taps (63, 62, ... for T = 125).  For even g (and g + 1 < G) the last step of the segment straddles into group g + 1.
// Weight stages are pieces of <= HL_HALF = 7 steps of a segment.
__host__ __device__ constexpr int hl_seg_begin(int T, int g) { return (T * g + 1) / 2; }
constexpr int HL_HALF = 7;

struct HlParams {
  int N, D, H, W, Cout;       // output extents == input extents (pad 1, stride 1)
  int Npad;                   // MMA N
  int tilesH, tilesW;
  long long ntiles;
  int nsplit;
  int tmem_cols;
  int epi;
  int resident;               // 1: the whole weight pack is loaded once per CTA; 0: ring of bstages stages of HL_HALF steps
  int bstages;
  const float* bias;
  float* y;
  const uint8_t* wpack;       // [step][khalf2][part][Npad][8] bf16 (hi rows, then lo rows)
  double* stat_acc;           // optional [N][Cout][2]: per-(sample, channel) sum and sum of squares of y (InstanceNorm statistics)
};

template <int G, bool SPLIT, int KS>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l, const HlParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);       // [HL_MAX_G]
  uint64_t* a_empty = a_full + HL_MAX_G;
  uint64_t* b_full = a_empty + HL_MAX_G;                          // [HL_MAX_BSTAGES]
  uint64_t* b_empty = b_full + HL_MAX_BSTAGES;
  uint64_t* t_full = b_empty + HL_MAX_BSTAGES;                    // [2]
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  __shared__ float s_stat[4][128];                                // per epilogue warp [2][64] statistics partials (p.stat_acc)
  uint8_t* base = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  constexpr int T = KS * KS * KS, PAD = KS / 2;
  constexpr int NSTEPS = (T * G + 1) / 2;
  constexpr int HL_PLANE = hl_plane(KS), HL_PLANE_DATA = hl_plane_data(KS), HL_WH = hl_wh(KS);
  constexpr int NH = (hl_seg_begin(T, 1) + HL_HALF - 1) / HL_HALF;        // stages per segment (the longest segment is the first)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int parts = SPLIT ? 2 : 1;
  const int nrows = parts * p.Npad;                               // weight rows per K half: hi rows then lo rows
  const int step_bytes = 2 * nrows * 16;
  uint8_t* a_slot = base;                                         // planes [g][part]
  uint8_t* b_ring = base + (size_t)G * 2 * HL_PLANE;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_h);
    if (parts == 2) prefetch_tmap(&map_l);
    for (int i = 0; i < G; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < HL_MAX_BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_stat[0][0])[i] = 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== halo producer: one TMA box per (group, part), one barrier pair per group =====================
    if (lane == 0) {
      int local = 0;
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++local) {
        long long t = tile;
        const int wb = (int)(t % p.tilesW); t /= p.tilesW;
        const int hb = (int)(t % p.tilesH); t /= p.tilesH;
        const int d = (int)(t % p.D);
        const int n = (int)(t / p.D);
        const int c_w = (wb * HL_WT - PAD) * 8;        // inner coordinate in elements (multiple of 8 -> 16 B aligned)
        const int c_h = hb * HL_HT - PAD;
        const int c_nd = n * (p.D + 2 * PAD) + d;      // padded plane index of d - PAD
        for (int g = 0; g < G; ++g) {
          mbar_wait(&a_empty[g], (uint32_t)((local & 1) ^ 1), 210);
          mbar_arrive_expect_tx(&a_full[g], (uint32_t)(parts * HL_PLANE_DATA));
          tma_load_4d(&map_h, &a_full[g], a_slot + (size_t)(2 * g) * HL_PLANE, c_w, c_h, c_nd, g);
          if (parts == 2) tma_load_4d(&map_l, &a_full[g], a_slot + (size_t)(2 * g + 1) * HL_PLANE, c_w, c_h, c_nd, g);
        }
      }
    }
  } else if (warp == 6) {
    // ===================== weight producer =====================
    if (lane == 0) {
      if (p.resident) {
        const uint32_t total = (uint32_t)(NSTEPS * step_bytes);
        mbar_arrive_expect_tx(&b_full[0], total);
        for (uint32_t off = 0; off < total; off += 32768u)
          bulk_load(b_ring + off, p.wpack + off, min(32768u, total - off), &b_full[0]);
      } else {
        int st = 0;
        uint32_t ph = 0;
        for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
          for (int q = 0; q < NH * G; ++q) {             // stage q = piece q % NH of segment q / NH
            const int g = q / NH;
            const int s0 = hl_seg_begin(T, g) + (q % NH) * HL_HALF;
            const int s1 = min(min(s0 + HL_HALF, hl_seg_begin(T, g + 1)), NSTEPS);
            if (s0 >= s1) continue;
            const uint32_t bytes = (uint32_t)((s1 - s0) * step_bytes);
            mbar_wait(&b_empty[st], ph ^ 1u, 220);
            mbar_arrive_expect_tx(&b_full[st], bytes);
            bulk_load(b_ring + (size_t)st * HL_HALF * step_bytes, p.wpack + (size_t)s0 * step_bytes, bytes, &b_full[st]);
            if (++st == p.bstages) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control, one straight-line leader region per weight stage ==========
    // A single warp retires a dependent instruction every ~4 cycles and the queue in front of the tensor pipe holds ~6 MMAs
    // (tools/umma_queue.cu), so the instruction count between MMAs IS the issue rate: the step table is compile-time
    // (template G) and every loop-carried quantity is a function of the uniform induction variable `it`, which keeps the
    // descriptors in uniform registers (no R2UR).
    const uint32_t leader = elect_one();
    const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nrows >> 3) << 17) | ((128u >> 4) << 24);
    // descriptor = constant high word (SBO, version) | low word (start address >> 4, LBO >> 4 in bits 16..29)
    constexpr uint32_t a_hiword = (uint32_t)((HL_WH * 16) >> 4) | (1u << 14);          // SBO = one halo line
    constexpr uint32_t b_hiword = (uint32_t)(128 >> 4) | (1u << 14);
    const uint32_t a_base = desc_addr(smem_u32(a_slot));
    const uint32_t b_base = desc_addr(smem_u32(b_ring)) | ((uint32_t)nrows << 16);
    const uint32_t stepw = (uint32_t)(step_bytes >> 4);
    const uint32_t stagew = stepw * HL_HALF;
    constexpr uint32_t lo_off = (uint32_t)(HL_PLANE >> 4);          // lo plane of a group follows its hi plane
    const uint32_t ring_n = (uint32_t)p.bstages;
    const uint32_t ring_inv = 0xFFFFFFFFu / ring_n + 1u;            // cnt / ring_n == umulhi(cnt, ring_inv) for cnt < 2^32 / ring_n
    const int iters = (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    if (p.resident) { mbar_wait(&b_full[0], 0, 245); tc_fence_after(); }
    for (int it = 0; it < iters; ++it) {
      const uint32_t buf = (uint32_t)it & 1u;
      mbar_wait(&t_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u, 230);
      const uint32_t dcol = tmem_base + buf * (uint32_t)nrows;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        mbar_wait(&a_full[g], buf, 240);
        if ((g & 1) == 0 && g + 1 < G) mbar_wait(&a_full[g + 1], buf, 241);     // the last step of an even segment straddles
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int s0 = hl_seg_begin(T, g) + h * HL_HALF;
          const int seg_end = hl_seg_begin(T, g + 1) < NSTEPS ? hl_seg_begin(T, g + 1) : NSTEPS;
          const int s1 = s0 + HL_HALF < seg_end ? s0 + HL_HALF : seg_end;
          if (s0 >= s1) continue;
          const bool last_piece = s1 == seg_end;
          // weight stage counter: stages of a tile are numbered in issue order (the producer skips the same empty pieces)
          int before = 0;
#pragma unroll
          for (int gg = 0; gg < G; ++gg) {
            const int sb = hl_seg_begin(T, gg), se = hl_seg_begin(T, gg + 1) < NSTEPS ? hl_seg_begin(T, gg + 1) : NSTEPS;
            const int pieces = (se - sb + HL_HALF - 1) / HL_HALF;
            if (gg < g) before += pieces;
          }
          constexpr int STAGES_PER_TILE = []() { int n = 0; for (int gg = 0; gg < G; ++gg) { const int sb = hl_seg_begin(T, gg); const int se = hl_seg_begin(T, gg + 1) < NSTEPS ? hl_seg_begin(T, gg + 1) : NSTEPS; n += (se - sb + HL_HALF - 1) / HL_HALF; } return n; }();
          const uint32_t cnt = (uint32_t)it * STAGES_PER_TILE + (uint32_t)(before + h);
          const uint32_t lap = ring_n == 1u ? cnt : __umulhi(cnt, ring_inv);       // (the reciprocal overflows for a one-stage ring)
          const uint32_t st = cnt - lap * ring_n;
          uint32_t bw;
          if (p.resident) bw = b_base + (uint32_t)s0 * stepw;
          else {
            mbar_wait(&b_full[st], lap & 1u, 250);
            bw = b_base + st * stagew;
          }
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int s = s0; s < s1; ++s) {
              if (s < NSTEPS) {
                const uint64_t b_all = desc_join(b_hiword, bw + (uint32_t)(s - s0) * stepw);
                const uint32_t aw = a_base + hl_a_word(KS, G, s);     // low 14 bits: address, bits 16..29: LBO (no carry between them)
                if (s == 0) umma_bf16(dcol, desc_join(a_hiword, aw), b_all, idesc_2n, 0u);     // [hi*hi | hi*lo]
                else umma_bf16_acc(dcol, desc_join(a_hiword, aw), b_all, idesc_2n);
                if (parts == 2) umma_bf16_acc(dcol, desc_join(a_hiword, aw + lo_off), b_all, idesc_n);   // lo*hi: first Npad rows only
              }
            }
            if (!p.resident) umma_commit(&b_empty[st]);
            if (last_piece) umma_commit(&a_empty[g]);
            if (last_piece && g == G - 1) umma_commit(&t_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int lh = row >> 3, lw = row & 7;
    const bool vec = (p.Cout & 3) == 0;
    int local = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++local) {
      long long t = tile;
      const int wb = (int)(t % p.tilesW); t /= p.tilesW;
      const int hb = (int)(t % p.tilesH); t /= p.tilesH;
      const int d = (int)(t % p.D);
      const int n = (int)(t / p.D);
      const int oh = hb * HL_HT + lh, ow = wb * HL_WT + lw;
      const bool ok = oh < p.H && ow < p.W;
      float* yrow = p.y + ((((long long)n * p.D + d) * p.H + oh) * p.W + ow) * (long long)p.Cout;
      const int buf = local & 1;
      mbar_wait(&t_full[buf], (uint32_t)((local >> 1) & 1), 260);
      tc_fence_after();
      for (int j = 0; j < p.Npad; j += 16) {
        uint32_t r[16], r2[16];
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * nrows + j), r);
        if (parts == 2) tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * nrows + p.Npad + j), r2);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float f = __uint_as_float(r[i]);
          if (parts == 2) f += __uint_as_float(r2[i]);
          if ((p.epi & CFUN_EPI_BIAS) && j + i < p.Cout) f += __ldg(p.bias + j + i);
          if (p.epi & CFUN_EPI_RELU) f = fmaxf(f, 0.f);
          v[i] = ok ? f : 0.f;
        }
        if (ok) {
          if (vec && j + 16 <= p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yrow + j + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (j + i < p.Cout) yrow[j + i] = v[i];
          }
        }
        if (p.stat_acc && j < p.Cout) {
          if (p.Cout <= 64) warp_stats16_shared(v, s_stat[quad], j, p.Cout, lane);
          else warp_stats16(v, p.stat_acc + (long long)n * p.Cout * 2, j, p.Cout, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);     // 4 epilogue warps -> barrier count 4
      if (p.stat_acc && p.Cout <= 64) {
        // flush the CTA's partial sums every 16 tiles, at a sample boundary and after the last tile (uniform over the 4 warps)
        const long long next = tile + gridDim.x;
        const bool flush = (local & 15) == 15 || next >= p.ntiles || next / ((long long)p.D * p.tilesH * p.tilesW) != n;
        if (flush) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int e = (int)threadIdx.x - 64;       // 0..127 over the epilogue warps
          const int st_ = e >> 6, c = e & 63;
          if (c < p.Cout) {
            const double t = ((double)s_stat[0][e] + (double)s_stat[1][e]) + ((double)s_stat[2][e] + (double)s_stat[3][e]);
            s_stat[0][e] = s_stat[1][e] = s_stat[2][e] = s_stat[3][e] = 0.f;
            if (t != 0.0) atomicAdd(p.stat_acc + ((long long)n * p.Cout + c) * 2 + st_, t);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
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

// x fp32 NDHWC [N,D,H,W,C] -> group-planar split bf16 [Kp/8][N*(D+2)][H][W][8]; d-planes 0 and D+1 of every sample are zero
__global__ void __launch_bounds__(256) pack_act_gp_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo, int N, int D, int H, int W, int C, int G, int P) {
  const long long HW = (long long)H * W;
  const long long vox_p = (long long)N * (D + 2 * P) * HW;  // padded voxel count per group (P zero planes on each side)
  const long long total = vox_p * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pv = i % vox_p;
    const int g = (int)(i / vox_p);
    const long long plane = pv / HW;                        // n*(D+2) + d'
    const int dp = (int)(plane % (D + 2 * P));
    const int n = (int)(plane / (D + 2 * P));
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
    if (dp < P || dp >= D + P) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j] = __float2bfloat16_rn(0.f); l[j] = h[j]; }
    } else {
      const long long src = (((long long)n * D + (dp - P)) * HW + (pv % HW)) * C + g * 8;
      float v[8];
      if (g * 8 + 8 <= C && (C & 3) == 0) {
        float4 a = __ldg(reinterpret_cast<const float4*>(x + src));
        float4 b = __ldg(reinterpret_cast<const float4*>(x + src + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (g * 8 + j < C) ? __ldg(x + src + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(v[j], h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + i * 8) = *reinterpret_cast<const uint4*>(l);
  }
}

// w (Cout, Cin, 27) fp32 -> [step][khalf2][part][Npad][8] bf16: K half `khalf` of step s is entry e = 2 s + khalf of the
// (channel group, tap) list, i.e. channels 8 (e / 27) .. + 7 at tap e % 27.  mode 1 = data gradient (rows = ci, k = co,
// taps mirrored).
__global__ void __launch_bounds__(256) pack_w_halo_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                                          int Cin, int Npad, int G, int nsteps, int parts, int mode, int T) {
  const long long total = (long long)nsteps * 2 * parts * Npad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int e8 = (int)(r % 8); r /= 8;
    const int row = (int)(r % Npad); r /= Npad;
    const int part = (int)(r % parts); r /= parts;
    const int kh = (int)(r % 2); r /= 2;
    int ent = 2 * (int)r + kh;
    if (((T * G) & 1) && (int)r == nsteps - 1) ent = kh ? T * G - 1 : T * G;      // odd entry count: (zero weights, last entry), see hl_a_word
    const int g = ent / T;
    const int k = g * 8 + e8;
    int tap = ent - g * T;
    int co, ci;
    if (mode == 0) { co = row; ci = k; }
    else { co = k; ci = row; tap = T - 1 - tap; }
    float v = 0.f;
    if (g < G && co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * T + tap];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    out[i] = part == 0 ? h : l;
  }
}

// Same pack through a shared-memory transpose.  Persistent blocks walk tiles of TV consecutive voxels: the tile (TV * C
// contiguous floats of x) is fetched with 16-byte cp.async into one of two shared-memory buffers while the previous tile
// is converted, so every block always has a tile of loads in flight; the conversion writes, per channel group, TV
// consecutive 16-byte rows.  Row pitch G*8 + 4 floats keeps the 16-byte shared-memory reads of a quarter warp on distinct
// banks; channels >= C of the last group read as zero.  Blocks [nwork, gridDim.x) zero the 2 P padding planes per sample.
// x2 != NULL: the source is the channel concatenation [x (C - C2 channels) | x2 (C2 channels)] of two tensors (the U-Net's
// torch.cat((up, skip), dim=1) feeding a conv: the concatenated fp32 tensor is never written).
// act != 0: the pack holds leaky_relu(x * act_scale[n][c], act_slope) (act_scale NULL = 1): the LeakyReLU / Dropout3d that
// precede a conv (mask_branch.py:127-131) applied on the way into its operand pack, bit-identical to affine_act_fwd.
__global__ void __launch_bounds__(256) pack_act_gp_tiled_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                                __nv_bfloat16* __restrict__ lo, int N, int D, int H, int W, int C, int G,
                                                                int TV, long long ntiles, int nwork, int P,
                                                                const float* __restrict__ x2, int C2, int act,
                                                                const float* __restrict__ act_scale, float act_slope) {
  extern __shared__ __align__(16) float tile[];
  const long long HW = (long long)H * W;
  const long long DHW = (long long)D * HW;
  const long long vox = (long long)N * DHW;                  // real voxels
  const long long vox_p = (long long)N * (D + 2 * P) * HW;   // padded voxels per group
  const int Cs = G * 8 + 4;
  if ((int)blockIdx.x >= nwork) {                            // zero planes: (g, n, 2 P planes, hw) rows of 16 bytes
    const long long total = (long long)G * N * 2 * P * HW;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = ((long long)blockIdx.x - nwork) * blockDim.x + threadIdx.x; i < total;
         i += (long long)(gridDim.x - nwork) * blockDim.x) {
      const long long hw = i % HW;
      long long r = i / HW;
      const int which = (int)(r % (2 * P)); r /= 2 * P;
      const int n = (int)(r % N);
      const int g = (int)(r / N);
      const long long pos = ((long long)n * (D + 2 * P) + (which < P ? which : D + which)) * HW + hw;
      reinterpret_cast<uint4*>(hi)[(long long)g * vox_p + pos] = z;
      if (lo) reinterpret_cast<uint4*>(lo)[(long long)g * vox_p + pos] = z;
    }
    return;
  }
  const int c4n = C >> 2;
  const int buf_floats = TV * Cs;
  auto fetch = [&](long long t, int b) {
    const long long v0 = t * TV;
    const int nv = (int)min((long long)TV, vox - v0);
    float* dst = tile + b * buf_floats;
    const int C1 = C - C2, c4a = C1 >> 2;
    for (int i = threadIdx.x; i < nv * c4n; i += blockDim.x) {
      const int v = i / c4n, c4 = i - v * c4n;
      const float* src = c4 < c4a ? x + (v0 + v) * (long long)C1 + 4 * c4 : x2 + (v0 + v) * (long long)C2 + 4 * (c4 - c4a);
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + v * Cs + 4 * c4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  long long t = blockIdx.x;
  int b = 0;
  if (t < ntiles) fetch(t, 0);
  for (; t < ntiles; t += nwork, b ^= 1) {
    if (t + nwork < ntiles) {
      fetch(t + nwork, b ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const long long v0 = t * TV;
    const int nv = (int)min((long long)TV, vox - v0);
    const float* cur = tile + b * buf_floats;
    for (int i = threadIdx.x; i < nv * G; i += blockDim.x) {
      const int g = i / nv, v = i - g * nv;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
      if (8 * g < C) a = *reinterpret_cast<const float4*>(cur + v * Cs + 8 * g);
      if (8 * g + 4 < C) bq = *reinterpret_cast<const float4*>(cur + v * Cs + 8 * g + 4);
      float f[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
      const long long gv = v0 + v;
      const long long n = gv / DHW, rem = gv - n * DHW;
      if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * g + j;
          const float m = (act_scale && c < C) ? __ldg(act_scale + n * C + c) : 1.f;
          const float t = fmaf(f[j], m, 0.f);
          f[j] = t > 0.f ? t : t * act_slope;
        }
      }
      __align__(16) __nv_bfloat16 h[8];
      __align__(16) __nv_bfloat16 l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(f[j], h[j], l[j]);
      const long long pos = (n * (D + 2 * P) + P) * HW + rem;
      reinterpret_cast<uint4*>(hi)[(long long)g * vox_p + pos] = *reinterpret_cast<const uint4*>(h);
      if (lo) reinterpret_cast<uint4*>(lo)[(long long)g * vox_p + pos] = *reinterpret_cast<const uint4*>(l);
    }
    __syncthreads();                                         // buffer b is refilled by the fetch of the next iteration
  }
}

// the 2 P zero planes per sample of a group-planar pack on their own (for packs whose data planes another kernel writes:
// cfun_instnorm_bwd_apply_pack)
__global__ void __launch_bounds__(256) pack_zero_planes_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int N,
                                                               int D, int H, int W, int G, int P) {
  const long long HW = (long long)H * W;
  const long long vox_p = (long long)N * (D + 2 * P) * HW;
  const long long total = (long long)G * N * 2 * P * HW;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    long long r = i / HW;
    const int which = (int)(r % (2 * P)); r /= 2 * P;
    const int n = (int)(r % N);
    const int g = (int)(r / N);
    const long long pos = ((long long)n * (D + 2 * P) + (which < P ? which : D + which)) * HW + hw;
    reinterpret_cast<uint4*>(hi)[(long long)g * vox_p + pos] = z;
    if (lo) reinterpret_cast<uint4*>(lo)[(long long)g * vox_p + pos] = z;
  }
}
int launch_pack_zero_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int G, int P, cudaStream_t st) {
  const long long total = (long long)G * N * 2 * P * H * W;
  pack_zero_planes_kernel<<<(unsigned)std::max<long long>(1, std::min<long long>(cdiv(total, 256), 2LL * num_sms())), 256, 0, st>>>(hi, lo, N, D, H, W, G, P);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

static int launch_pack_w_halo(const float* w, __nv_bfloat16* out, int Cout, int Cin, int Npad, int G, int nsteps, int parts, int mode,
                              int T, cudaStream_t st) {
  const long long wt = (long long)nsteps * 2 * parts * Npad * 8;
  pack_w_halo_kernel<<<(unsigned)std::min<long long>(cdiv(wt, 256), 4LL * num_sms()), 256, 0, st>>>(w, out, Cout, Cin, Npad, G, nsteps, parts, mode, T);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// host-side launcher (also used by conv_tc_hx.cu, conv_tc_wgrad_ds.cu, conv_fused.cu)
int launch_pack_act_gp_pad(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, int P, cudaStream_t st);
int launch_pack_act_gp(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, cudaStream_t st) {
  return launch_pack_act_gp_pad(x, hi, lo, N, D, H, W, C, G, 1, st);
}
// P zero planes before and after every sample (P = 1 for 3^3 kernels, 2 for 5^3)
int launch_pack_act_gp_pad(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H, int W, int C, int G, int P, cudaStream_t st) {
  const char* e = getenv("CFUN_PACK_TILED");          // "0": the one-thread-per-row kernel (A/B measurements)
  const int Cs = G * 8 + 4;
  const int TV = std::min(128, (6144 / Cs) / 32 * 32);  // two buffers of <= 24 KB: 4 blocks per SM
  if ((C & 3) == 0 && C > 16 && TV >= 32 && !(e && e[0] == '0')) {   // <= 16 channels: the row-per-thread kernel is faster (8 ch: 0.040 vs 0.064 ms)
    const long long vox = (long long)N * D * H * W;
    const long long ntile = cdiv(vox, TV);
    const long long zrows = (long long)G * N * 2 * P * H * W;
    const int nz = (int)std::max<long long>(1, std::min<long long>(cdiv(zrows, 1024), (long long)num_sms()));
    const int nwork = (int)std::min<long long>(ntile, 4LL * num_sms());
    pack_act_gp_tiled_kernel<<<(unsigned)(nwork + nz), 256, (size_t)2 * TV * Cs * sizeof(float), st>>>(x, hi, lo, N, D, H, W, C, G, TV, ntile, nwork, P, nullptr, 0, 0, nullptr, 0.f);
    CFUN_LAUNCH_CHECK();
    return CFUN_OK;
  }
  long long total = (long long)G * N * (D + 2 * P) * H * W;
  pack_act_gp_kernel<<<(unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()), 256, 0, st>>>(x, hi, lo, N, D, H, W, C, G, P);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// pack of the channel concatenation [a (C1) | b (C2)]; false = shape not handled by the tiled kernel (caller concatenates first)
bool pack_cat_supported(int C1, int C2, int G) {
  const int Cs = G * 8 + 4;
  const int TV = std::min(128, (6144 / Cs) / 32 * 32);
  return C1 > 0 && C2 > 0 && (C1 & 3) == 0 && (C2 & 3) == 0 && C1 + C2 > 16 && G * 8 >= C1 + C2 && TV >= 32;
}
int launch_pack_cat_gp_pad(const float* a, int C1, const float* b, int C2, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H,
                           int W, int G, int P, cudaStream_t st) {
  CFUN_CHECK_ARG(pack_cat_supported(C1, C2, G));
  const int Cs = G * 8 + 4;
  const int TV = std::min(128, (6144 / Cs) / 32 * 32);
  const long long vox = (long long)N * D * H * W;
  const long long ntile = cdiv(vox, TV);
  const long long zrows = (long long)G * N * 2 * P * H * W;
  const int nz = (int)std::max<long long>(1, std::min<long long>(cdiv(zrows, 1024), (long long)num_sms()));
  const int nwork = (int)std::min<long long>(ntile, 4LL * num_sms());
  pack_act_gp_tiled_kernel<<<(unsigned)(nwork + nz), 256, (size_t)2 * TV * Cs * sizeof(float), st>>>(a, hi, lo, N, D, H, W, C1 + C2, G, TV, ntile, nwork, P, b, C2, 0, nullptr, 0.f);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

// pack of leaky_relu(x * scale[n][c], slope); false from pack_preact_supported = shape not handled by the tiled kernel
bool pack_preact_supported(int C, int G) {
  const int Cs = G * 8 + 4;
  const int TV = std::min(128, (6144 / Cs) / 32 * 32);
  return (C & 3) == 0 && C > 16 && G * 8 >= C && TV >= 32;
}
int launch_pack_preact_gp_pad(const float* x, const float* scale, float slope, __nv_bfloat16* hi, __nv_bfloat16* lo, int N, int D, int H,
                              int W, int C, int G, int P, cudaStream_t st) {
  CFUN_CHECK_ARG(pack_preact_supported(C, G));
  const int Cs = G * 8 + 4;
  const int TV = std::min(128, (6144 / Cs) / 32 * 32);
  const long long vox = (long long)N * D * H * W;
  const long long ntile = cdiv(vox, TV);
  const long long zrows = (long long)G * N * 2 * P * H * W;
  const int nz = (int)std::max<long long>(1, std::min<long long>(cdiv(zrows, 1024), (long long)num_sms()));
  const int nwork = (int)std::min<long long>(ntile, 4LL * num_sms());
  pack_act_gp_tiled_kernel<<<(unsigned)(nwork + nz), 256, (size_t)2 * TV * Cs * sizeof(float), st>>>(x, hi, lo, N, D, H, W, C, G, TV, ntile, nwork, P, nullptr, 0, 1, scale, slope);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

}  // namespace cfun

// The operand pack on its own, for layout tests and bandwidth measurements (tools/ew_bench.py); the convolutions call
// launch_pack_act_gp_pad directly.  hi / lo: [G][N*(D+2P)][H][W][8] bf16 each, G >= ceil(C/8); lo may be NULL.
extern "C" int cfun_pack_act_gp(const float* x, void* hi, void* lo, int N, int D, int H, int W, int C, int G, int P, void* stream) {
  using namespace cfun;
  CFUN_CHECK_ARG(x && hi && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && G * 8 >= C && P >= 1 && P <= 2);
  return launch_pack_act_gp_pad(x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), N, D, H, W, C, G, P,
                                as_stream(stream));
}

namespace cfun {

struct HlPlan {
  int Cs, Ct, N, D, H, W, Kp, Gp, G, nsteps, Npad, tmem_cols, resident, sps, bstages, KS, ctas_per_sm;
  size_t off_ah, off_al, off_w, total, act_bytes, w_bytes, smem;
};

static bool make_hl_plan(const cfun_conv3d_desc* d, int pass, HlPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  if (d->kD != d->kH || d->kD != d->kW || (d->kD != 3 && d->kD != 5)) return false;
  if (d->pD != d->kD / 2 || d->pH != d->kD / 2 || d->pW != d->kD / 2) return false;
  pl.KS = d->kD;
  if (pass == CFUN_PASS_FWD) { pl.Cs = d->Cin; pl.Ct = d->Cout; }
  else if (pass == CFUN_PASS_BWD_DATA) { pl.Cs = d->Cout; pl.Ct = d->Cin; }
  else return false;
  pl.N = d->N; pl.D = d->Din; pl.H = d->Hin; pl.W = d->Win;
  if (pl.H < 8 || pl.W < 8) return false;
  const int T = pl.KS * pl.KS * pl.KS, HL_PLANE = hl_plane(pl.KS);
  if ((long long)pl.W * 8 > 0x7fffffffLL) return false;
  pl.Kp = (int)align_up((size_t)pl.Cs, 16);
  pl.Gp = pl.Kp / 8;                                       // groups in the pack (16-channel granularity, shared with hx / wgrad)
  pl.G = (int)cdiv(pl.Cs, 8);                              // groups that hold real channels: the only ones loaded and multiplied
  if (pl.G > (pl.KS == 3 ? HL_MAX_G : 1)) return false;     // 5^3: one channel group (the 8 -> 8 out_upscale_conv)
  pl.nsteps = (T * pl.G + 1) / 2;
  pl.Npad = (int)align_up((size_t)pl.Ct, 16);
  if (pl.Npad > 128) return false;                         // [hi | lo] accumulator pairs, double buffered: 4 * Npad <= 512
  int cols = 32;
  while (cols < 4 * pl.Npad) cols <<= 1;
  pl.tmem_cols = cols;
  const size_t a_bytes = (size_t)pl.G * 2 * HL_PLANE;
  const size_t step_bytes = (size_t)2 * 2 * pl.Npad * 16;
  const size_t budget = 227 * 1024 - 2048 - 2048 - a_bytes;     // 2 KB header + alignment slack, 2 KB static (s_stat)
  pl.w_bytes = align_up((size_t)pl.nsteps * step_bytes, 1024);
  const char* e = getenv("CFUN_HL_RESIDENT");              // "0": always stream the weights (A/B measurements)
  // Two co-resident CTAs per SM wherever they fit (thin layers: <= 32 input channels, and the 5^3 kernel): each CTA's barrier
  // hand-shakes overlap the other CTA's MMAs.  Needs <= 111 KB of dynamic shared memory per CTA and <= 256 TMEM columns.
  // Measured on B200, 20->20 @ 4x96^3: 0.54 -> 0.46 ms per kernel although the weights are no longer resident (a single
  // CTA with streamed weights: 0.59).  CFUN_HL_CTAS=1 keeps one CTA per SM (A/B measurements).
  pl.ctas_per_sm = 1;
  {
    const char* c1 = getenv("CFUN_HL_CTAS");
    const size_t cap2 = 111 * 1024;
    const size_t min_stages = (c1 && c1[0] == '2') ? 1 : 2;       // "2": also with a single weight stage (experiment)
    if (!(c1 && c1[0] == '1') && pl.tmem_cols <= 256 && 2048 + a_bytes + min_stages * HL_HALF * step_bytes <= cap2) {
      const size_t budget2 = cap2 - 2048 - a_bytes;
      pl.ctas_per_sm = 2;
      if (pl.w_bytes <= budget2 && !(e && e[0] == '0')) {
        pl.resident = 1; pl.sps = pl.nsteps; pl.bstages = 1;
        pl.smem = 2048 + a_bytes + pl.w_bytes;
      } else {
        pl.resident = 0;
        pl.sps = HL_HALF;
        pl.bstages = (int)std::min<size_t>(HL_MAX_BSTAGES, budget2 / (pl.sps * step_bytes));
        pl.smem = 2048 + a_bytes + (size_t)pl.bstages * pl.sps * step_bytes;
      }
    }
  }
  if (pl.ctas_per_sm == 2) {
  } else if (pl.w_bytes <= budget && !(e && e[0] == '0')) {
    pl.resident = 1; pl.sps = pl.nsteps; pl.bstages = 1;
    pl.smem = 2048 + a_bytes + pl.w_bytes;
  } else {
    pl.resident = 0;
    pl.sps = HL_HALF;
    pl.bstages = (int)std::min<size_t>(HL_MAX_BSTAGES, budget / (pl.sps * step_bytes));
    if (pl.bstages < 2) return false;
    pl.smem = 2048 + a_bytes + (size_t)pl.bstages * pl.sps * step_bytes;
  }
  pl.act_bytes = align_up((size_t)pl.Gp * pl.N * (pl.D + 2 * (pl.KS / 2)) * pl.H * pl.W * 16, 1024);
  pl.off_ah = 0; pl.off_al = pl.act_bytes; pl.off_w = 2 * pl.act_bytes;
  pl.total = 2 * pl.act_bytes + pl.w_bytes + 2048;
  return true;
}

bool hl_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_TC_HALO");       // "0" disables the halo kernels (A/B measurements)
  if (e && e[0] == '0') return false;
  const char* x = getenv("CFUN_TC_HX");         // "only": route every supported shape through conv_tc_hx.cu instead
  if (x && x[0] == 'o') return false;
  HlPlan pl;
  if (!make_hl_plan(d, pass, pl)) return false;
  if (pl.KS == 5) return true;            // any channel counts up to 8 (8 -> 8 heart, 3 -> 3 LiTS): the pack has a scalar path
  return pl.Cs >= 16 && (pl.Cs & 3) == 0 && pl.Ct >= 8;
}
size_t hl_workspace(const cfun_conv3d_desc* d, int pass) {
  HlPlan pl;
  return make_hl_plan(d, pass, pl) ? pl.total : 0;
}

// ext_hi / ext_lo (optional): caller-owned buffers for the split-bf16 activation pack (pl.act_bytes each, see
// hl_pack_bytes); ext_ready = the pack is already in them (fused backward, forward pack kept for the weight gradient).
int hl_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               double* stat_acc = nullptr);
int hl_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
            int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  return hl_conv_ex(d, pass, src, w, bias, dst, epi, nsplit, ws, ws_bytes, nullptr, nullptr, false, st);
}
size_t hl_pack_bytes(const cfun_conv3d_desc* d, int pass) {
  HlPlan pl;
  return make_hl_plan(d, pass, pl) ? pl.act_bytes : 0;
}
int hl_conv_ex(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
               int nsplit, void* ws, size_t ws_bytes, __nv_bfloat16* ext_hi, __nv_bfloat16* ext_lo, bool ext_ready, cudaStream_t st,
               double* stat_acc) {
  HlPlan pl;
  CFUN_CHECK_ARG(make_hl_plan(d, pass, pl));
  CFUN_CHECK_ARG((src || ext_ready) && w && dst && ws && get_tensor_map_encoder());
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d halo: workspace too small"); return CFUN_ERR_WORKSPACE; }
  const bool split = nsplit == 3;
  const int parts = split ? 2 : 1;
  __nv_bfloat16* ah = ext_hi ? ext_hi : reinterpret_cast<__nv_bfloat16*>(base + pl.off_ah);
  __nv_bfloat16* al = ext_hi ? ext_lo : reinterpret_cast<__nv_bfloat16*>(base + pl.off_al);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(base + pl.off_w);
  {
    if (!(ext_hi && ext_ready)) {
      int prc = launch_pack_act_gp_pad(src, ah, split ? al : nullptr, pl.N, pl.D, pl.H, pl.W, pl.Cs, pl.Gp, pl.KS / 2, st);
      if (prc != CFUN_OK) return prc;
    }
    int prc = launch_pack_w_halo(w, wp, d->Cout, d->Cin, pl.Npad, pl.G, pl.nsteps, parts, pass == CFUN_PASS_BWD_DATA ? 1 : 0,
                                 pl.KS * pl.KS * pl.KS, st);
    if (prc != CFUN_OK) return prc;
  }
  CUtensorMap mh, ml;
  for (int part = 0; part < 2; ++part) {
    void* b = part == 0 ? (void*)ah : (void*)(split ? al : ah);
    const cuuint64_t planes = (cuuint64_t)pl.N * (pl.D + 2 * (pl.KS / 2));
    cuuint64_t dims[4] = {(cuuint64_t)pl.W * 8, (cuuint64_t)pl.H, planes, (cuuint64_t)pl.Gp};
    cuuint64_t strides[3] = {(cuuint64_t)pl.W * 16, (cuuint64_t)pl.H * pl.W * 16, planes * pl.H * pl.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)hl_wh(pl.KS) * 8, (cuuint32_t)hl_hh(pl.KS), (cuuint32_t)pl.KS, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_tensor_map_encoder()(part == 0 ? &mh : &ml, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  }
  HlParams p;
  p.N = pl.N; p.D = pl.D; p.H = pl.H; p.W = pl.W; p.Cout = pl.Ct;
  p.Npad = pl.Npad;
  p.tilesH = (int)cdiv(pl.H, HL_HT); p.tilesW = (int)cdiv(pl.W, HL_WT);
  p.ntiles = (long long)pl.N * pl.D * p.tilesH * p.tilesW;
  p.nsplit = split ? 3 : 1;
  p.tmem_cols = pl.tmem_cols;
  p.epi = epi; p.bias = bias; p.y = dst;
  p.resident = pl.resident; p.bstages = pl.bstages;
  p.wpack = reinterpret_cast<const uint8_t*>(wp);
  p.stat_acc = stat_acc;
  const unsigned grid = (unsigned)std::min<long long>(p.ntiles, (long long)pl.ctas_per_sm * num_sms());
#define CFUN_HL_LAUNCH(GG, KK)                                                                                                  \
  case GG + 16 * KK: {                                                                                                          \
    static bool attr_set = false;                                                                                               \
    if (!attr_set) {                                                                                                            \
      CFUN_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<GG, true, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));  \
      CFUN_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<GG, false, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); \
      attr_set = true;                                                                                                          \
    }                                                                                                                           \
    timing_begin(st);                                                                                                           \
    if (split) conv_tc_halo_kernel<GG, true, KK><<<grid, HL_THREADS, pl.smem, st>>>(mh, ml, p);                                 \
    else conv_tc_halo_kernel<GG, false, KK><<<grid, HL_THREADS, pl.smem, st>>>(mh, ml, p);                                      \
    timing_end(st);                                                                                                             \
    break;                                                                                                                      \
  }
  switch (pl.G + 16 * pl.KS) {
    CFUN_HL_LAUNCH(2, 3) CFUN_HL_LAUNCH(3, 3) CFUN_HL_LAUNCH(4, 3) CFUN_HL_LAUNCH(5, 3) CFUN_HL_LAUNCH(6, 3) CFUN_HL_LAUNCH(7, 3)
    CFUN_HL_LAUNCH(8, 3) CFUN_HL_LAUNCH(1, 5)
    default: set_error("conv3d halo: unsupported channel group count %d for kernel size %d", pl.G, pl.KS); return CFUN_ERR_INVALID;
  }
#undef CFUN_HL_LAUNCH
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int tc_debug_read_halo(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}

}  // namespace cfun
