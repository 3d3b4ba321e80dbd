// tcgen05 implicit-GEMM 3-D convolution for sm_100a (stride 1, any odd/even kernel, arbitrary zero padding).
//
//   y[m, co] = sum_{tap, ci} x[m + tap - pad, ci] * w[co, ci, tap]        m = one output voxel
//
// Design (im2col-free, DESIGN.md "conv_tc"):
//   * operands are split-bf16 pairs: v = hi + lo with hi = bf16(v), lo = bf16(v - hi); the product is accumulated as
//     hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator (3 MMAs per K-chunk, relative error ~2^-16, inside the 1e-4
//     parity bar that a single bf16 or tf32 pass misses); `nsplit = 1` runs hi*hi only (fast mode, not parity grade);
//   * one CTA owns a 128-voxel output box (Dt x Ht x Wt) times BN <= 256 output channels; for every kernel tap and every
//     16-channel K-chunk the TMA loads the *shifted* input box (16 ch x Wt x Ht x Dt) straight out of the NDHWC tensor,
//     out-of-bounds coordinates are zero-filled by the TMA unit = the conv's zero padding, no halo code, no im2col buffer;
//     the 27 re-reads of the input tile hit L2, HBM sees each activation once;
//   * smem tiles are K-major, SWIZZLE_32B (row = 16 bf16 = 32 B = exactly one UMMA K step), descriptors per chunk;
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue
//     (tcgen05.ld 32x32b -> bias / ReLU -> coalesced fp32 NDHWC stores); full/empty mbarrier ring between TMA and MMA,
//     tcgen05.commit releases smem stages and publishes the accumulator.
// The data gradient of a stride-1 conv is the same kernel on dy with spatially flipped, (ci,co)-transposed weights and
// pad' = k-1-pad.
#include "tc_ptx.cuh"
#include <mutex>
#include <cstdlib>

namespace cfun {

// K-major, SWIZZLE_32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, SM100 "version 1"):
//   rows of 32 B, 8-row groups 256 B apart (SBO), LBO field = 1 (unused for a single 32 B K-slab)
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (>>4)
  d |= (uint64_t)(256 >> 4) << 32;   // stride byte offset (>>4)
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)6 << 61;            // layout type SWIZZLE_32B
  return d;
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
constexpr int TC_KCH = 4;            // 16-channel K chunks per pipeline stage
constexpr int TC_A_BYTES = 128 * 32; // one A operand part per chunk
constexpr int TC_THREADS = 192;

struct TcParams {
  int N, Do, Ho, Wo, Cout;
  int kD, kH, kW, pD, pH, pW;
  int CPC;           // K chunks per tap (= Cp / 16)
  int BN;            // output channels per CTA (multiple of 16, <= 256)
  int Nt, Dt, Ht, Wt;    // output box per CTA, Nt*Dt*Ht*Wt == 128 (Nt > 1 only for volumes smaller than a tile)
  int tilesD, tilesH, tilesW;
  int nsplit;        // 3: hi*hi + lo*hi + hi*lo ; 1: hi*hi
  int stages;
  int tmem_cols;
  int epi;
  int ksplit;        // > 1: blockIdx.z covers a slice of the K chunks and writes raw partial sums to part[z]
  const float* bias;
  float* y;
  float* part;       // [ksplit][N*Do*Ho*Wo][Cout]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: barriers first (small), then the stage ring aligned to 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint8_t* ring = smem_raw + 1024 + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned in the shared window

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_bytes = p.BN * 32;
  const int parts = p.nsplit == 3 ? 2 : 1;
  const int chunk_bytes = parts * (TC_A_BYTES + b_bytes);
  const int stage_bytes = TC_KCH * chunk_bytes;
  const int taps = p.kD * p.kH * p.kW;
  const int all_chunks = taps * p.CPC;
  const int q_begin = (int)(((long long)all_chunks * blockIdx.z) / p.ksplit);
  const int total_chunks = (int)(((long long)all_chunks * (blockIdx.z + 1)) / p.ksplit);     // end of this CTA's chunk range
  const int nstage_iters = (total_chunks - q_begin + TC_KCH - 1) / TC_KCH;

  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tilesW; t /= p.tilesW;
  const int th = t % p.tilesH; t /= p.tilesH;
  const int td = t % p.tilesD; t /= p.tilesD;
  const int n = t * p.Nt;            // first sample of the box
  const int d0 = td * p.Dt, h0 = th * p.Ht, w0 = tw * p.Wt;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_ah);
    prefetch_tmap(&map_bh);
    if (p.nsplit == 3) { prefetch_tmap(&map_al); prefetch_tmap(&map_bl); }
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      int q = q_begin;
      for (int it = 0; it < nstage_iters; ++it) {
        const int slot = it % p.stages;
        const uint32_t ph = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&empty_bar[slot], ph ^ 1u, 10);
        const int nch = min(TC_KCH, total_chunks - q);
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(nch * chunk_bytes));
        uint8_t* sbase = ring + (size_t)slot * stage_bytes;
        for (int c = 0; c < nch; ++c, ++q) {
          const int tap = q / p.CPC;
          const int cc = q - tap * p.CPC;
          const int khw = p.kH * p.kW;
          const int kd = tap / khw;
          const int r = tap - kd * khw;
          const int kh = r / p.kW;
          const int kw = r - kh * p.kW;
          uint8_t* cb = sbase + (size_t)c * chunk_bytes;
          const int xc = cc * 16, xw = w0 + kw - p.pW, xh = h0 + kh - p.pH, xd = d0 + kd - p.pD;
          tma_load_5d(&map_ah, &full_bar[slot], cb, xc, xw, xh, xd, n);
          tma_load_3d(&map_bh, &full_bar[slot], cb + parts * TC_A_BYTES, xc, n0, tap);
          if (parts == 2) {
            tma_load_5d(&map_al, &full_bar[slot], cb + TC_A_BYTES, xc, xw, xh, xd, n);
            tma_load_3d(&map_bl, &full_bar[slot], cb + 2 * TC_A_BYTES + b_bytes, xc, n0, tap);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues: tc_ptx.cuh "issue-rate note") ==========
    {
      const uint32_t leader = elect_one();
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t hiword = (uint32_t)(make_desc_sw32(0) >> 32);
      const uint32_t lbo1 = 1u << 16;
      int q = q_begin;
      uint32_t acc = 0;
      for (int it = 0; it < nstage_iters; ++it) {
        const int slot = it % p.stages;
        const uint32_t ph = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&full_bar[slot], ph, 20);
        tc_fence_after();
        const int nch = min(TC_KCH, total_chunks - q);
        const uint32_t sbase = smem_u32(ring + (size_t)slot * stage_bytes);
#pragma unroll
        for (int c = 0; c < TC_KCH; ++c) {
          if (c < nch) {
            const uint32_t cb = desc_addr(sbase + (uint32_t)(c * chunk_bytes)) | lbo1;
            const uint64_t a_hi = desc_join(hiword, cb);
            const uint64_t b_hi = desc_join(hiword, cb + (uint32_t)((parts * TC_A_BYTES) >> 4));
            if (leader) {
              umma_bf16(tmem_base, a_hi, b_hi, idesc, acc);
              if (parts == 2) {
                const uint64_t a_lo = desc_join(hiword, cb + (uint32_t)(TC_A_BYTES >> 4));
                const uint64_t b_lo = desc_join(hiword, cb + (uint32_t)((2 * TC_A_BYTES + b_bytes) >> 4));
                umma_bf16_acc(tmem_base, a_lo, b_hi, idesc);
                umma_bf16_acc(tmem_base, a_hi, b_lo, idesc);
              }
            }
            acc = 1;
          }
        }
        q += nch;
        if (leader) umma_commit(&empty_bar[slot]);   // frees the smem stage once the MMAs above have consumed it
        __syncwarp();
      }
      if (leader) umma_commit(tmem_full_bar);        // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // row of the 128-row tile == voxel inside the box
    const int lw = row % p.Wt;
    const int lh = (row / p.Wt) % p.Ht;
    const int ld = (row / (p.Wt * p.Ht)) % p.Dt;
    const int on = n + row / (p.Wt * p.Ht * p.Dt);
    const int od = d0 + ld, oh = h0 + lh, ow = w0 + lw;
    const bool vox_ok = on < p.N && od < p.Do && oh < p.Ho && ow < p.Wo;
    const bool raw = p.ksplit > 1;             // partial sums: bias / ReLU are applied by reduce_split_kernel
    float* yrow = (raw ? p.part + (long long)blockIdx.z * ((long long)p.N * p.Do * p.Ho * p.Wo * p.Cout) : p.y) +
                  ((((long long)on * p.Do + od) * p.Ho + oh) * p.Wo + ow) * (long long)p.Cout;
    mbar_wait(tmem_full_bar, 0, 30);
    tc_fence_after();
    const bool vec = (p.Cout & 3) == 0;
    for (int j = 0; j < p.BN; j += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)j, r);
      tmem_ld_wait();
      if (vox_ok) {
        const int c0 = n0 + j;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float f = __uint_as_float(r[i]);
          if (!raw && (p.epi & CFUN_EPI_BIAS) && c0 + i < p.Cout) f += __ldg(p.bias + c0 + i);
          if (!raw && (p.epi & CFUN_EPI_RELU)) f = fmaxf(f, 0.f);
          v[i] = f;
        }
        if (vec && c0 + 16 <= p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(yrow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.Cout) yrow[c0 + i] = v[i];
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

// y = sum over the K splits of conv_tc_kernel's partial tiles, in split order (deterministic), + bias, ReLU
__global__ void __launch_bounds__(256) reduce_split_kernel(const float* __restrict__ part, int ksplit, long long total4, int C4,
                                                           const float* __restrict__ bias, int epi, float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(part) + i);
    for (int z = 1; z < ksplit; ++z) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(part) + (long long)z * total4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (epi & CFUN_EPI_BIAS) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % C4));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (epi & CFUN_EPI_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    reinterpret_cast<float4*>(y)[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------
// operand packing
// ---------------------------------------------------------------------------------------------------------
// x [rows, C] fp32 -> hi/lo [rows, Cp] bf16 (channels >= C zero)
__global__ void __launch_bounds__(256) pack_act_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo, long long rows, int C, int Cp) {
  const int G = Cp / 8;
  const long long total = rows * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long r = i / G;
    const int c0 = g * 8;
    float v[8];
    if (c0 + 8 <= C && (C & 3) == 0) {
      float4 a = __ldg(reinterpret_cast<const float4*>(x + r * C + c0));
      float4 b = __ldg(reinterpret_cast<const float4*>(x + r * C + c0 + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? __ldg(x + r * C + c0 + j) : 0.f;
    }
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(v[j], h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + r * Cp + c0) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + r * Cp + c0) = *reinterpret_cast<const uint4*>(l);
  }
}

// w (Cout, Cin, taps) fp32 -> [tap][Np][Kp] hi/lo.  mode 0 (forward): row = co, k = ci, tap as is.
// mode 1 (data gradient): row = ci, k = co, tap mirrored (kD-1-kd, kH-1-kh, kW-1-kw).
__global__ void __launch_bounds__(256) pack_w_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ lo, int Cout, int Cin, int kD, int kH,
                                                        int kW, int Np, int Kp, int mode) {
  const int taps = kD * kH * kW;
  const long long total = (long long)taps * Np * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    long long r = i / Kp;
    const int row = (int)(r % Np);
    const int tap = (int)(r / Np);
    float v = 0.f;
    int co, ci, st = tap;
    if (mode == 0) { co = row; ci = k; }
    else {
      co = k; ci = row;
      int kd = tap / (kH * kW), rr = tap % (kH * kW), kh = rr / kW, kw = rr % kW;
      st = ((kD - 1 - kd) * kH + (kH - 1 - kh)) * kW + (kW - 1 - kw);
    }
    if (co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * taps + st];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
EncodeTiledFn get_tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TcPlan {
  // the conv this plan executes (already in "forward" form: source tensor, weights [tap][Np][Kp], target tensor)
  int N, Cs, Ds, Hs, Ws;     // source (K channels)
  int Ct, Dt_, Ht_, Wt_;     // target dims (output channels, spatial)
  int kD, kH, kW, pD, pH, pW;
  int Kp, Np, BN, ntiles_n;
  int bn, bd, bh, bw;        // output box (bn samples x bd x bh x bw voxels = 128 rows)
  int tilesN, tilesD, tilesH, tilesW;
  size_t off_ah, off_al, off_bh, off_bl, off_part, total;
  int ksplit;                // CTAs sharing one output tile along K (taps x channel chunks); > 1: partial sums + reduce_split_kernel
};

// 128-row output box = bn samples x bd x bh x bw voxels, every extent a power of two not larger than the tensor's (the
// TMA box never exceeds the tensor extent); minimises the padded volume, then prefers one sample per box and wide rows.
// bn > 1 lets volumes smaller than 128 voxels per sample (the 6^3 bottom of the U-Net) use the tensor-core path.
static void pick_box(int N, int Do, int Ho, int Wo, int& bn, int& bd, int& bh, int& bw) {
  long long best = -1;
  int best_pref = -1;
  bn = bd = bh = bw = 0;
  for (int n = 1; n <= 128 && n <= N; n <<= 1)
    for (int a = 1; n * a <= 128 && a <= Do; a <<= 1)
      for (int b = 1; n * a * b <= 128 && b <= Ho; b <<= 1) {
        const int c = 128 / (n * a * b);
        if (n * a * b * c != 128 || c > Wo) continue;
        const long long vol = cdiv(N, n) * n * cdiv(Do, a) * a * cdiv(Ho, b) * b * cdiv(Wo, c) * c;
        const int pref = (n == 1 ? 1000 : 0) + (c >= 8 ? 100 : 0) - abs(a - b) - abs(b - c);
        if (best < 0 || vol < best || (vol == best && pref > best_pref)) { best = vol; best_pref = pref; bn = n; bd = a; bh = b; bw = c; }
      }
}

static bool make_plan(const cfun_conv3d_desc* d, int pass, TcPlan& pl) {
  if (!d || d->sD != 1 || d->sH != 1 || d->sW != 1) return false;
  pl.N = d->N;
  pl.kD = d->kD; pl.kH = d->kH; pl.kW = d->kW;
  if (pass == CFUN_PASS_FWD) {
    pl.Cs = d->Cin; pl.Ds = d->Din; pl.Hs = d->Hin; pl.Ws = d->Win;
    pl.Ct = d->Cout; pl.Dt_ = d->Dout; pl.Ht_ = d->Hout; pl.Wt_ = d->Wout;
    pl.pD = d->pD; pl.pH = d->pH; pl.pW = d->pW;
  } else if (pass == CFUN_PASS_BWD_DATA) {
    pl.Cs = d->Cout; pl.Ds = d->Dout; pl.Hs = d->Hout; pl.Ws = d->Wout;
    pl.Ct = d->Cin; pl.Dt_ = d->Din; pl.Ht_ = d->Hin; pl.Wt_ = d->Win;
    pl.pD = d->kD - 1 - d->pD; pl.pH = d->kH - 1 - d->pH; pl.pW = d->kW - 1 - d->pW;
    if (pl.pD < 0 || pl.pH < 0 || pl.pW < 0) return false;
  } else {
    return false;
  }
  pl.Kp = (int)align_up((size_t)pl.Cs, 16);
  pl.Np = (int)align_up((size_t)pl.Ct, 16);
  pl.ntiles_n = (int)cdiv(pl.Np, 256);
  pl.BN = (int)align_up((size_t)cdiv(pl.Np, pl.ntiles_n), 16);
  pick_box(pl.N, std::min(pl.Dt_, pl.Ds), std::min(pl.Ht_, pl.Hs), std::min(pl.Wt_, pl.Ws), pl.bn, pl.bd, pl.bh, pl.bw);
  if (pl.bd == 0) return false;
  pl.tilesN = (int)cdiv(pl.N, pl.bn);
  pl.tilesD = (int)cdiv(pl.Dt_, pl.bd); pl.tilesH = (int)cdiv(pl.Ht_, pl.bh); pl.tilesW = (int)cdiv(pl.Wt_, pl.bw);
  const size_t rows = (size_t)pl.N * pl.Ds * pl.Hs * pl.Ws;
  const size_t act = align_up(rows * pl.Kp * 2, 1024);
  const size_t wgt = align_up((size_t)pl.kD * pl.kH * pl.kW * (size_t)(pl.BN * pl.ntiles_n) * pl.Kp * 2, 1024);
  pl.off_ah = 0; pl.off_al = act; pl.off_bh = 2 * act; pl.off_bl = 2 * act + wgt; pl.total = 2 * act + 2 * wgt + 2048;
  // Few output tiles and a long contraction (the 6^3 x 320-channel bottom of the U-Net: ~20 CTAs each streaming 27 taps x 20
  // chunks): split K over blockIdx.z so the launch fills the SMs; the partial tiles are summed in a fixed order afterwards.
  const long long base_ctas = (long long)pl.tilesN * pl.tilesD * pl.tilesH * pl.tilesW * pl.ntiles_n;
  const int chunks = pl.kD * pl.kH * pl.kW * (pl.Kp / 16);
  pl.ksplit = 1;
  const char* e = getenv("CFUN_TC_KSPLIT");                // "0": never split (A/B measurements)
  if (2 * base_ctas <= num_sms() && !(e && e[0] == '0'))
    pl.ksplit = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(num_sms() / base_ctas, chunks / 16), 16));
  pl.off_part = pl.total;
  if (pl.ksplit > 1) pl.total += align_up((size_t)pl.ksplit * pl.N * pl.Dt_ * pl.Ht_ * pl.Wt_ * pl.Ct * sizeof(float), 1024);
  return true;
}

bool tc_wgrad_supported(const cfun_conv3d_desc* d);
size_t tc_wgrad_workspace(const cfun_conv3d_desc* d);
bool hx_supported(const cfun_conv3d_desc* d, int pass);     // conv_tc_hx.cu: halo-resident, cluster-multicast weights
size_t hx_workspace(const cfun_conv3d_desc* d, int pass);
int hx_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
            int nsplit, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_debug_read_hx(int* out8);
bool hl_supported(const cfun_conv3d_desc* d, int pass);
size_t hl_workspace(const cfun_conv3d_desc* d, int pass);
int hl_conv(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
            int nsplit, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_debug_read_halo(int* out8);
int tc_debug_read_ds(int* out8);

bool wg_capable(const cfun_conv3d_desc* d);   // conv_tc_wgrad.cu
bool s2d_supported(const cfun_conv3d_desc* d, int pass);   // conv_s2d.cu: stride-2 3^3 convs by space-to-depth
size_t s2d_workspace(const cfun_conv3d_desc* d, int pass);
int s2d_conv(const cfun_conv3d_desc* d, int pass, const float* a, const float* b, const float* bias, float* out, float* dbias,
             int epi, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st);

// geometry / alignment capability of the generic tcgen05 kernels, without the "is it worth it" policy of tc_supported
bool tc_capable(const cfun_conv3d_desc* d, int pass) {
  if (pass == CFUN_PASS_BWD_WEIGHT) return wg_capable(d);
  TcPlan pl;
  if (!make_plan(d, pass, pl)) return false;
  if (pl.Cs < 16 || (pl.Cs & 3) || pl.Ct < 8) return false;
  if (pl.ntiles_n > 8) return false;
  const size_t stage_bytes = (size_t)TC_KCH * 2 * (TC_A_BYTES + pl.BN * 32);
  return (200 * 1024) / stage_bytes >= 2;
}

bool tc_supported(const cfun_conv3d_desc* d, int pass) {
  static int sm100 = -1;
  if (sm100 < 0) sm100 = cfun_device_is_sm100();
  if (!sm100 || !get_tensor_map_encoder()) return false;
  if (d && d->sD == 2 && s2d_supported(d, pass)) return true;
  if (pass == CFUN_PASS_BWD_WEIGHT) {
    const char* e = getenv("CFUN_TC_WGRAD");     // "0" keeps the weight gradient on CUDA cores (A/B measurements)
    if (e && e[0] == '0') return false;
    return tc_wgrad_supported(d);
  }
  if (hl_supported(d, pass) || hx_supported(d, pass)) return true;
  TcPlan pl;
  if (!make_plan(d, pass, pl)) return false;
  const int taps = pl.kD * pl.kH * pl.kW;
  if (taps < 27) return false;                       // pointwise / P3D factorised convs stay on CUDA cores
  if (pl.Cs < 16 || (pl.Cs & 3) || pl.Ct < 8) return false;
  if (pl.ntiles_n > 8) return false;
  return true;
}

// heuristic used by CFUN_CONV_ALGO_AUTO: tensor cores only where the launch is big enough to pay for the operand packing
bool tc_preferred(const cfun_conv3d_desc* d, int pass) {
  if (!tc_supported(d, pass)) return false;
  const long long vox = (long long)d->N * d->Dout * d->Hout * d->Wout;
  const double flop = 2.0 * (double)vox * d->Cin * d->Cout * d->kD * d->kH * d->kW;
  return vox >= 2048 || flop >= 5e8;      // the 6^3 x 320-channel bottom of the U-Net is tiny in voxels, not in work
}

size_t tc_workspace(const cfun_conv3d_desc* d, int pass) {
  if (d && d->sD == 2) return s2d_workspace(d, pass);
  if (pass == CFUN_PASS_BWD_WEIGHT) return tc_wgrad_workspace(d);
  if (hl_supported(d, pass)) return hl_workspace(d, pass);
  if (hx_supported(d, pass)) return hx_workspace(d, pass);
  TcPlan pl;
  if (!make_plan(d, pass, pl)) return 0;
  return pl.total;
}

static int encode_act_map(CUtensorMap* m, void* base, const TcPlan& pl) {
  cuuint64_t dims[5] = {(cuuint64_t)pl.Kp, (cuuint64_t)pl.Ws, (cuuint64_t)pl.Hs, (cuuint64_t)pl.Ds, (cuuint64_t)pl.N};
  cuuint64_t strides[4] = {(cuuint64_t)pl.Kp * 2, (cuuint64_t)pl.Ws * pl.Kp * 2, (cuuint64_t)pl.Hs * pl.Ws * pl.Kp * 2,
                           (cuuint64_t)pl.Ds * pl.Hs * pl.Ws * pl.Kp * 2};
  cuuint32_t box[5] = {16, (cuuint32_t)pl.bw, (cuuint32_t)pl.bh, (cuuint32_t)pl.bd, (cuuint32_t)pl.bn};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = get_tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  return CFUN_OK;
}

static int encode_w_map(CUtensorMap* m, void* base, const TcPlan& pl) {
  const int rows = pl.BN * pl.ntiles_n;
  cuuint64_t dims[3] = {(cuuint64_t)pl.Kp, (cuuint64_t)rows, (cuuint64_t)(pl.kD * pl.kH * pl.kW)};
  cuuint64_t strides[2] = {(cuuint64_t)pl.Kp * 2, (cuuint64_t)rows * pl.Kp * 2};
  cuuint32_t box[3] = {16, (cuuint32_t)pl.BN, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = get_tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return CFUN_ERR_CUDA; }
  return CFUN_OK;
}

static int run_tc(const cfun_conv3d_desc* d, int pass, const float* src, const float* w, const float* bias, float* dst, int epi,
                  int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcPlan pl;
  CFUN_CHECK_ARG(make_plan(d, pass, pl));
  CFUN_CHECK_ARG(src && w && dst && ws);
  CFUN_CHECK_ARG(get_tensor_map_encoder() != nullptr);
  const size_t base = align_up((size_t)ws, 1024);
  if (ws_bytes < pl.total || base + pl.total - 2048 > (size_t)ws + ws_bytes) { set_error("conv3d tc: workspace too small"); return CFUN_ERR_WORKSPACE; }
  __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(base + pl.off_ah);
  __nv_bfloat16* al = reinterpret_cast<__nv_bfloat16*>(base + pl.off_al);
  __nv_bfloat16* bh = reinterpret_cast<__nv_bfloat16*>(base + pl.off_bh);
  __nv_bfloat16* bl = reinterpret_cast<__nv_bfloat16*>(base + pl.off_bl);
  const bool split = nsplit == 3;
  const long long rows = (long long)pl.N * pl.Ds * pl.Hs * pl.Ws;
  {
    long long total = rows * (pl.Kp / 8);
    pack_act_kernel<<<(unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()), 256, 0, st>>>(src, ah, split ? al : nullptr, rows, pl.Cs, pl.Kp);
    CFUN_LAUNCH_CHECK();
    const int Nrows = pl.BN * pl.ntiles_n;
    long long wt = (long long)pl.kD * pl.kH * pl.kW * Nrows * pl.Kp;
    pack_w_tc_kernel<<<(unsigned)std::min<long long>(cdiv(wt, 256), 8LL * num_sms()), 256, 0, st>>>(
        w, bh, split ? bl : nullptr, d->Cout, d->Cin, d->kD, d->kH, d->kW, Nrows, pl.Kp, pass == CFUN_PASS_BWD_DATA ? 1 : 0);
    CFUN_LAUNCH_CHECK();
  }
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = encode_act_map(&mah, ah, pl)) != CFUN_OK) return rc;
  if ((rc = encode_act_map(&mal, split ? al : ah, pl)) != CFUN_OK) return rc;
  if ((rc = encode_w_map(&mbh, bh, pl)) != CFUN_OK) return rc;
  if ((rc = encode_w_map(&mbl, split ? bl : bh, pl)) != CFUN_OK) return rc;

  TcParams p;
  p.N = pl.N; p.Do = pl.Dt_; p.Ho = pl.Ht_; p.Wo = pl.Wt_; p.Cout = pl.Ct;
  p.kD = pl.kD; p.kH = pl.kH; p.kW = pl.kW; p.pD = pl.pD; p.pH = pl.pH; p.pW = pl.pW;
  p.CPC = pl.Kp / 16;
  p.BN = pl.BN;
  p.Nt = pl.bn; p.Dt = pl.bd; p.Ht = pl.bh; p.Wt = pl.bw;
  p.tilesD = pl.tilesD; p.tilesH = pl.tilesH; p.tilesW = pl.tilesW;
  p.nsplit = split ? 3 : 1;
  const int parts = split ? 2 : 1;
  const size_t stage_bytes = (size_t)TC_KCH * parts * (TC_A_BYTES + pl.BN * 32);
  int stages = (int)std::min<size_t>(8, (200 * 1024) / stage_bytes);
  CFUN_CHECK_ARG(stages >= 2);
  p.stages = stages;
  int cols = 32;
  while (cols < pl.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.epi = epi;
  p.bias = bias;
  p.y = dst;
  const size_t smem = 1024 + stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int ksplit = pl.ksplit;
  if ((pl.Ct & 3) || (bias && ((size_t)bias & 15))) ksplit = 1;      // reduce_split_kernel works on float4
  p.ksplit = ksplit;
  p.part = reinterpret_cast<float*>(base + pl.off_part);
  dim3 grid((unsigned)((long long)pl.tilesN * pl.tilesD * pl.tilesH * pl.tilesW), (unsigned)pl.ntiles_n, (unsigned)ksplit);
  conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(mah, mal, mbh, mbl, p);
  CFUN_LAUNCH_CHECK();
  if (ksplit > 1) {
    const long long total4 = (long long)pl.N * pl.Dt_ * pl.Ht_ * pl.Wt_ * pl.Ct / 4;
    reduce_split_kernel<<<(unsigned)std::min<long long>(cdiv(total4, 256), 4LL * num_sms()), 256, 0, st>>>(p.part, ksplit, total4, pl.Ct / 4,
                                                                                                       bias, epi, dst);
    CFUN_LAUNCH_CHECK();
  }
  return CFUN_OK;
}

int tc_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, int nsplit,
                void* ws, size_t ws_bytes, cudaStream_t st) {
  CFUN_CHECK_ARG(!(epi & CFUN_EPI_BIAS) || bias);
  if (d->sD == 2) return s2d_conv(d, CFUN_PASS_FWD, x, w, bias, y, nullptr, epi, nsplit, ws, ws_bytes, st);
  // thin layers (Cin <= 64): conv_tc_halo.cu, measured 7-10 % faster there; everything else 3^3: conv_tc_hx.cu
  if (hl_supported(d, CFUN_PASS_FWD)) return hl_conv(d, CFUN_PASS_FWD, x, w, bias, y, epi, nsplit, ws, ws_bytes, st);
  if (hx_supported(d, CFUN_PASS_FWD)) return hx_conv(d, CFUN_PASS_FWD, x, w, bias, y, epi, nsplit, ws, ws_bytes, st);
  return run_tc(d, CFUN_PASS_FWD, x, w, bias, y, epi, nsplit, ws, ws_bytes, st);
}

int tc_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, int nsplit, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  if (d->sD == 2) return s2d_conv(d, CFUN_PASS_BWD_DATA, dy, w, nullptr, dx, nullptr, 0, nsplit, ws, ws_bytes, st);
  if (hl_supported(d, CFUN_PASS_BWD_DATA)) return hl_conv(d, CFUN_PASS_BWD_DATA, dy, w, nullptr, dx, 0, nsplit, ws, ws_bytes, st);
  if (hx_supported(d, CFUN_PASS_BWD_DATA)) return hx_conv(d, CFUN_PASS_BWD_DATA, dy, w, nullptr, dx, 0, nsplit, ws, ws_bytes, st);
  return run_tc(d, CFUN_PASS_BWD_DATA, dy, w, nullptr, dx, 0, nsplit, ws, ws_bytes, st);
}

}  // namespace cfun

extern "C" int cfun_pack_split_bf16(const float* x, void* hi, void* lo, long long rows, int C, int Cpad, void* stream) {
  using namespace cfun;
  CFUN_CHECK_ARG(x && hi && rows >= 0 && C > 0 && Cpad >= C && Cpad % 8 == 0);
  if (rows == 0) return CFUN_OK;
  long long total = rows * (Cpad / 8);
  pack_act_kernel<<<(unsigned)std::min<long long>(cdiv(total, 256), 32LL * num_sms()), 256, 0, as_stream(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), rows, C, Cpad);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

namespace cfun {
int tc_debug_read_wgrad(int* out8);
int tc_debug_read_conv(int* out8) {
  int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CFUN_CUDA(cudaMemcpyFromSymbol(out8, g_tc_debug, sizeof(z)));
  CFUN_CUDA(cudaMemcpyToSymbol(g_tc_debug, z, sizeof(z)));
  return CFUN_OK;
}
}  // namespace cfun

// debugging aid: {site+1 (0 = no time-out), blockIdx.x, blockIdx.y, threadIdx.x, parity, spins}; reading resets it.
extern "C" int cfun_tc_debug_status(int* out8_host) {
  using namespace cfun;
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("cfun_tc_debug_status: device in error state"); return CFUN_ERR_CUDA; }
  int a[8], b[8], c[8];
  int rc = tc_debug_read_conv(a);
  if (rc != CFUN_OK) return rc;
  rc = tc_debug_read_wgrad(b);
  if (rc != CFUN_OK) return rc;
  rc = tc_debug_read_halo(c);
  if (rc != CFUN_OK) return rc;
  int f[8], g[8];
  rc = tc_debug_read_ds(f);
  if (rc != CFUN_OK) return rc;
  rc = tc_debug_read_hx(g);
  if (rc != CFUN_OK) return rc;
  for (int i = 0; i < 8; ++i)
    out8_host[i] = a[0] ? a[i] : (b[0] ? b[i] : (c[0] ? c[i] : (f[0] ? f[i] : g[i])));
  return CFUN_OK;
}
