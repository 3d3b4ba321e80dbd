// Pointwise (1x1x1, stride 1) convolutions with at most 8 output channels: the segmentation / deep-supervision heads of
// the mask U-Net (mask_branch.py:72,86-87: conv3d_l4 40->8 @ 96^3, ds2_1x1_conv3d, ds3_1x1_conv3d) and the RPN class /
// bbox heads (model.py:714-715).  They move 0.7 GB per pass at 96^3 and do 0.3 GFLOP: HBM-bound streaming kernels, which
// the generic implicit GEMM ran 3-6x off the bandwidth roofline (its tiles assume Cout >= 16).
//   forward        y[v][co]  = sum_ci x[v][ci] w[co][ci] (+ bias, ReLU)   one thread per voxel pair, weights in shared memory
//   data gradient  dx[v][ci] = sum_co dy[v][co] w[co][ci]                 same
//   weight gradient dw[co][ci] = sum_v dy[v][co] x[v][ci]                 4 x 2 register tiles over shared-memory staged chunks,
//                                                                         persistent blocks, one atomic flush
// Arithmetic is plain fp32 FMA (the exact-mode CUDA-core path).
#include "common.cuh"

namespace cfun {

constexpr int PW_MAX_CIN = 320;
constexpr int PW_THREADS = 256;

// weights staged as wsm[ci][8] (output channels beyond Cout are zero)
__global__ void __launch_bounds__(PW_THREADS) pw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, long long M, int Cin,
                                                            int Cout, int epi) {
  __shared__ __align__(16) float wsm[PW_MAX_CIN * 8];
  for (int i = threadIdx.x; i < Cin * 8; i += PW_THREADS) {
    const int ci = i >> 3, co = i & 7;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int c4n = Cin >> 2;
  for (long long v = (long long)blockIdx.x * PW_THREADS + threadIdx.x; v < M; v += (long long)gridDim.x * PW_THREADS) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(x + v * Cin);
#pragma unroll 2
    for (int c4 = 0; c4 < c4n; ++c4) {
      const float4 xv = __ldg(xr + c4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8 + 4);
        acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
        acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
        acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
        acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((epi & CFUN_EPI_BIAS) && j < Cout) acc[j] += __ldg(bias + j);
      if (epi & CFUN_EPI_RELU) acc[j] = fmaxf(acc[j], 0.f);
    }
    float* yo = y + v * Cout;
    if (Cout == 8) {
      reinterpret_cast<float4*>(yo)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(yo)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < Cout) yo[j] = acc[j];
    }
  }
}

// weights staged as wsm[co][Cin] (rows beyond Cout are zero)
__global__ void __launch_bounds__(PW_THREADS) pw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, long long M, int Cin, int Cout) {
  __shared__ __align__(16) float wsm[8 * PW_MAX_CIN];
  for (int i = threadIdx.x; i < 8 * Cin; i += PW_THREADS) {
    const int co = i / Cin, ci = i - co * Cin;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int c4n = Cin >> 2;
  for (long long v = (long long)blockIdx.x * PW_THREADS + threadIdx.x; v < M; v += (long long)gridDim.x * PW_THREADS) {
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = j < Cout ? __ldg(dy + v * Cout + j) : 0.f;
    float4* xo = reinterpret_cast<float4*>(dx + v * Cin);
#pragma unroll 2
    for (int c4 = 0; c4 < c4n; ++c4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(wsm + j * Cin + c4 * 4);
        o.x = fmaf(g[j], wv.x, o.x); o.y = fmaf(g[j], wv.y, o.y); o.z = fmaf(g[j], wv.z, o.z); o.w = fmaf(g[j], wv.w, o.w);
      }
      xo[c4] = o;
    }
  }
}

// Streaming forms of the two kernels above for Cin <= 80 (the 96^3 / 48^3 heads, where a thread's 160..320-byte row walk made
// every warp load touch 32 sectors for 16 useful bytes each): persistent blocks of TV threads move tiles of TV voxels through
// shared memory with coalesced 16-byte transfers -- cp.async double-buffered on the way in (forward), one staged tile on the
// way out (data gradient) -- and thread v works on row v of the tile.  A row pitch with an odd number of 16-byte words keeps
// the row accesses of a quarter warp on distinct banks.  Same arithmetic order as pw_fwd_kernel / pw_dgrad_kernel.
__device__ __forceinline__ void pw_cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void pw_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pw_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TV>
__global__ void __launch_bounds__(TV) pw_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, long long M, int Cin, int Cout,
                                                     int epi, int pitch) {
  extern __shared__ __align__(16) float sm[];            // wsm[Cin][8] | tile[2][TV][pitch]
  float* wsm = sm;
  float* tile = sm + Cin * 8;
  for (int i = threadIdx.x; i < Cin * 8; i += TV) {
    const int ci = i >> 3, co = i & 7;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  const int c4n = Cin >> 2;
  const long long ntiles = (M + TV - 1) / TV;
  auto fetch = [&](long long t, int b) {
    const long long v0 = t * TV;
    const int nv = (int)min((long long)TV, M - v0);
    const float* src = x + v0 * Cin;
    float* dst = tile + (size_t)b * TV * pitch;
    for (int i = threadIdx.x; i < nv * c4n; i += TV) {
      const int v = i / c4n, c4 = i - v * c4n;
      pw_cp_async16(dst + v * pitch + 4 * c4, src + 4 * (long long)i);
    }
    pw_cp_commit();
  };
  long long t = blockIdx.x;
  int b = 0;
  if (t < ntiles) fetch(t, 0);
  for (; t < ntiles; t += gridDim.x, b ^= 1) {
    if (t + gridDim.x < ntiles) { fetch(t + gridDim.x, b ^ 1); pw_cp_wait<1>(); }
    else pw_cp_wait<0>();
    __syncthreads();
    const long long v = t * TV + threadIdx.x;
    if (v < M) {
      const float* row = tile + (size_t)b * TV * pitch + threadIdx.x * pitch;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 2
      for (int c4 = 0; c4 < c4n; ++c4) {
        const float4 xv = *reinterpret_cast<const float4*>(row + 4 * c4);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wsm + (c4 * 4 + j) * 8 + 4);
          acc[0] = fmaf(xs[j], w0.x, acc[0]); acc[1] = fmaf(xs[j], w0.y, acc[1]);
          acc[2] = fmaf(xs[j], w0.z, acc[2]); acc[3] = fmaf(xs[j], w0.w, acc[3]);
          acc[4] = fmaf(xs[j], w1.x, acc[4]); acc[5] = fmaf(xs[j], w1.y, acc[5]);
          acc[6] = fmaf(xs[j], w1.z, acc[6]); acc[7] = fmaf(xs[j], w1.w, acc[7]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if ((epi & CFUN_EPI_BIAS) && j < Cout) acc[j] += __ldg(bias + j);
        if (epi & CFUN_EPI_RELU) acc[j] = fmaxf(acc[j], 0.f);
      }
      float* yo = y + v * Cout;
      if (Cout == 8) {
        reinterpret_cast<float4*>(yo)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(yo)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < Cout) yo[j] = acc[j];
      }
    }
    __syncthreads();                       // buffer b is refilled by the next iteration's fetch
  }
}

template <int TV>
__global__ void __launch_bounds__(TV) pw_dgrad2_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                       long long M, int Cin, int Cout, int pitch) {
  extern __shared__ __align__(16) float sm[];            // wsm[8][Cin] | tile[TV][pitch]
  float* wsm = sm;
  float* tile = sm + 8 * Cin;
  for (int i = threadIdx.x; i < 8 * Cin; i += TV) {
    const int co = i / Cin, ci = i - co * Cin;
    wsm[i] = co < Cout ? __ldg(w + (long long)co * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int c4n = Cin >> 2;
  const long long ntiles = (M + TV - 1) / TV;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long v0 = t * TV;
    const int nv = (int)min((long long)TV, M - v0);
    if ((int)threadIdx.x < nv) {
      const long long v = v0 + threadIdx.x;
      float g[8];
      if (Cout == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dy + v * 8)), c = __ldg(reinterpret_cast<const float4*>(dy + v * 8) + 1);
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = j < Cout ? __ldg(dy + v * Cout + j) : 0.f;
      }
      float* row = tile + threadIdx.x * pitch;
#pragma unroll 2
      for (int c4 = 0; c4 < c4n; ++c4) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(wsm + j * Cin + c4 * 4);
          o.x = fmaf(g[j], wv.x, o.x); o.y = fmaf(g[j], wv.y, o.y); o.z = fmaf(g[j], wv.z, o.z); o.w = fmaf(g[j], wv.w, o.w);
        }
        *reinterpret_cast<float4*>(row + 4 * c4) = o;
      }
    }
    __syncthreads();
    float4* out = reinterpret_cast<float4*>(dx + v0 * Cin);
    for (int i = threadIdx.x; i < nv * c4n; i += TV) {
      const int v = i / c4n, c4 = i - v * c4n;
      out[i] = *reinterpret_cast<const float4*>(tile + v * pitch + 4 * c4);
    }
    __syncthreads();
  }
}

// Round 2: register-tiled weight gradient.  A thread owns a 4 (ci) x 2 (co) tile of dW; a block stages PW2_V voxels of x and dy
// in shared memory (coalesced float4 / float2 rows) and its thread groups -- one group = (Cin / 4) x (Cout_p / 2) threads --
// walk disjoint voxel subsets of the chunk: one LDS.128 + one LDS.64 per 8 FMAs, so the kernel is bound by the HBM read of x
// and dy (0.68 GB at 40 -> 8 @ 4 x 96^3) instead of by shared-memory instruction issue.  One atomic flush per block.
constexpr int PW2_V = 128;
// nbuf = 2: the next chunk is fetched with cp.async while this one is multiplied (chunks up to 50 KB); 1: fetch, wait, multiply
__global__ void __launch_bounds__(PW_THREADS) pw_wgrad2_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, long long M, int Cin, int Cout, int Cop, int nbuf) {
  extern __shared__ __align__(16) float sm[];            // nbuf x { xs[PW2_V][Cin], gs[PW2_V][Cop] }
  const int chunk_floats = PW2_V * (Cin + Cop);
  const int tci = Cin >> 2, tco = Cop >> 1;              // thread tiles along ci / co
  const int gsize = tci * tco;
  const int ngroups = PW_THREADS / gsize;
  const int grp = threadIdx.x / gsize, tin = threadIdx.x - grp * gsize;
  const bool active = grp < ngroups;
  const int c4 = tin % tci, o2 = tin / tci;
  const bool async_g = Cop == Cout && (Cout & 3) == 0;   // dy rows are whole 16-byte pieces
  float acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long nchunks = (M + PW2_V - 1) / PW2_V;
  auto fetch = [&](long long ch, int b) {
    float* xs = sm + (size_t)b * chunk_floats;
    float* gs = xs + (size_t)PW2_V * Cin;
    const long long v0 = ch * PW2_V;
    const int nv = (int)min((long long)PW2_V, M - v0);
    for (int i = threadIdx.x; i < nv * tci; i += PW_THREADS) pw_cp_async16(xs + 4 * i, x + v0 * Cin + 4 * (long long)i);
    if (async_g) {
      for (int i = threadIdx.x; i < nv * (Cout >> 2); i += PW_THREADS) pw_cp_async16(gs + 4 * i, dy + v0 * Cout + 4 * (long long)i);
    } else {
      for (int i = threadIdx.x; i < nv * Cop; i += PW_THREADS) {
        const int v = i / Cop, c = i - v * Cop;
        gs[i] = c < Cout ? __ldg(dy + (v0 + v) * Cout + c) : 0.f;
      }
    }
    pw_cp_commit();
  };
  long long ch = blockIdx.x;
  int b = 0;
  if (ch < nchunks) fetch(ch, 0);
  for (; ch < nchunks; ch += gridDim.x) {
    const long long next = ch + gridDim.x;
    if (nbuf == 2 && next < nchunks) { fetch(next, b ^ 1); pw_cp_wait<1>(); }
    else pw_cp_wait<0>();
    __syncthreads();
    const int nv = (int)min((long long)PW2_V, M - ch * PW2_V);
    const float* xs = sm + (size_t)b * chunk_floats;
    const float* gs = xs + (size_t)PW2_V * Cin;
    if (active) {
      for (int v = grp; v < nv; v += ngroups) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)v * Cin + 4 * c4);
        const float2 gv = *reinterpret_cast<const float2*>(gs + (size_t)v * Cop + 2 * o2);
        acc[0][0] = fmaf(gv.x, xv.x, acc[0][0]); acc[0][1] = fmaf(gv.x, xv.y, acc[0][1]);
        acc[0][2] = fmaf(gv.x, xv.z, acc[0][2]); acc[0][3] = fmaf(gv.x, xv.w, acc[0][3]);
        acc[1][0] = fmaf(gv.y, xv.x, acc[1][0]); acc[1][1] = fmaf(gv.y, xv.y, acc[1][1]);
        acc[1][2] = fmaf(gv.y, xv.z, acc[1][2]); acc[1][3] = fmaf(gv.y, xv.w, acc[1][3]);
      }
    }
    __syncthreads();
    if (nbuf == 2) b ^= 1;
    else if (next < nchunks) fetch(next, 0);
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int co = 2 * o2 + a;
      if (co < Cout)
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) atomicAdd(dw + (long long)co * Cin + 4 * c4 + b2, acc[a][b2]);
    }
  }
}

int simt_bias_grad(const float* dy, long long M, int C, float* dbias, cudaStream_t st);

bool pw_supported(const cfun_conv3d_desc* d, int pass) {
  const char* e = getenv("CFUN_CONV_PW");          // "0": generic implicit GEMM (A/B measurements)
  if (e && e[0] == '0') return false;
  if (!d || d->kD != 1 || d->kH != 1 || d->kW != 1 || d->sD != 1 || d->sH != 1 || d->sW != 1 || d->pD || d->pH || d->pW) return false;
  if (d->Cout > 8 || d->Cout < 1 || (d->Cin & 3) || d->Cin > PW_MAX_CIN || d->Cin < 4) return false;
  // measured on B200 (tools/conv_cases.py pw): the forward wins everywhere (40->8 @ 4x96^3: 0.34 -> 0.18 ms), the data
  // gradient only for narrow inputs (0.46 -> 0.27 ms at 40 channels, slower from 80 up), the weight gradient nowhere
  // (0.69 -> 1.0 ms with the round-1 kernel; the register-tiled pw_wgrad2_kernel of round 2 takes it)
  if (pass == CFUN_PASS_BWD_DATA && d->Cin > 80) return false;       // streaming pw_dgrad2_kernel up to 80 channels
  if (pass == CFUN_PASS_BWD_WEIGHT) {
    const int cop = (d->Cout + 1) & ~1;
    if ((d->Cin >> 2) * (cop >> 1) > PW_THREADS) return false;
    if ((size_t)PW2_V * (d->Cin + cop) * sizeof(float) > 200 * 1024) return false;
  }
  return true;
}

// pitch (floats) of a staged [voxel][Cin] tile: an odd number of 16-byte words per row
static inline int pw_pitch(int Cin) { return ((Cin >> 2) & 1) ? Cin : Cin + 4; }
static inline long long pw_rows(const cfun_conv3d_desc* d) { return (long long)d->N * d->Dout * d->Hout * d->Wout; }

int pw_conv_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y, int epi, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_FWD) && x && w && y);
  CFUN_CHECK_ARG(!(epi & CFUN_EPI_BIAS) || bias);
  const long long M = pw_rows(d);
  const char* e = getenv("CFUN_PW_STREAM");         // "0": the row-per-thread kernels (A/B measurements)
  if (d->Cin <= 80 && M >= 4096 && !(e && e[0] == '0')) {
    const int pitch = pw_pitch(d->Cin);
    static bool attr_set = false;
    if (!attr_set) {
      CFUN_CUDA(cudaFuncSetAttribute(pw_fwd2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CFUN_CUDA(cudaFuncSetAttribute(pw_fwd2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    if (d->Cin <= 40) {
      const size_t smem = sizeof(float) * ((size_t)d->Cin * 8 + 2 * 256 * pitch);
      const unsigned grid = (unsigned)std::min<long long>(cdiv(M, 256), 2LL * num_sms());
      pw_fwd2_kernel<256><<<grid, 256, smem, st>>>(x, w, bias, y, M, d->Cin, d->Cout, epi, pitch);
    } else {
      const size_t smem = sizeof(float) * ((size_t)d->Cin * 8 + 2 * 128 * pitch);
      const unsigned grid = (unsigned)std::min<long long>(cdiv(M, 128), 2LL * num_sms());
      pw_fwd2_kernel<128><<<grid, 128, smem, st>>>(x, w, bias, y, M, d->Cin, d->Cout, epi, pitch);
    }
    CFUN_LAUNCH_CHECK();
    return CFUN_OK;
  }
  const unsigned grid = (unsigned)std::min<long long>(cdiv(M, PW_THREADS), 16LL * num_sms());
  pw_fwd_kernel<<<grid, PW_THREADS, 0, st>>>(x, w, bias, y, M, d->Cin, d->Cout, epi);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int pw_conv_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_BWD_DATA) && dy && w && dx);
  const long long M = pw_rows(d);
  const char* e = getenv("CFUN_PW_STREAM");
  if (d->Cin <= 80 && M >= 4096 && !(e && e[0] == '0')) {
    const int pitch = pw_pitch(d->Cin);
    static bool attr_set = false;
    if (!attr_set) {
      CFUN_CUDA(cudaFuncSetAttribute(pw_dgrad2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    const size_t smem = sizeof(float) * ((size_t)d->Cin * 8 + 256 * pitch);
    const unsigned grid = (unsigned)std::min<long long>(cdiv(M, 256), 2LL * num_sms());
    pw_dgrad2_kernel<256><<<grid, 256, smem, st>>>(dy, w, dx, M, d->Cin, d->Cout, pitch);
    CFUN_LAUNCH_CHECK();
    return CFUN_OK;
  }
  const unsigned grid = (unsigned)std::min<long long>(cdiv(M, PW_THREADS), 16LL * num_sms());
  pw_dgrad_kernel<<<grid, PW_THREADS, 0, st>>>(dy, w, dx, M, d->Cin, d->Cout);
  CFUN_LAUNCH_CHECK();
  return CFUN_OK;
}

int pw_conv_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t st) {
  CFUN_CHECK_ARG(pw_supported(d, CFUN_PASS_BWD_WEIGHT) && x && dy && dw);
  const long long M = pw_rows(d);
  CFUN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->Cin, st));
  const int cop = (d->Cout + 1) & ~1;
  const size_t chunk = (size_t)PW2_V * (d->Cin + cop) * sizeof(float);
  const int nbuf = chunk <= 50 * 1024 ? 2 : 1;
  const size_t smem = nbuf * chunk;
  static bool attr_set = false;
  if (!attr_set) {
    CFUN_CUDA(cudaFuncSetAttribute(pw_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const long long blocks = std::max<long long>(1, std::min<long long>(cdiv(M, PW2_V), (smem > 64 * 1024 ? 1LL : 3LL) * num_sms()));
  pw_wgrad2_kernel<<<(unsigned)blocks, PW_THREADS, smem, st>>>(x, dy, dw, M, d->Cin, d->Cout, cop, nbuf);
  CFUN_LAUNCH_CHECK();
  if (dbias) return simt_bias_grad(dy, M, d->Cout, dbias, st);
  return CFUN_OK;
}

}  // namespace cfun
