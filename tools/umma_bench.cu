// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N and of the shared-memory operand
// layout (K-major / MN-major, swizzle mode, SBO / LBO, start-address alignment).  Operands are whatever the (zeroed)
// shared memory holds: only the issue / operand-fetch cost is measured.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu && tools/umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  uint32_t idesc;
  uint32_t a_off, b_off;        // start offsets into smem
  uint32_t a_lbo, a_sbo, a_layout;
  uint32_t b_lbo, b_sbo, b_layout;
  uint32_t a_step, b_step;      // start-address increment between consecutive MMAs (cycled over 8 steps)
  int nmma;
  int reps;
  int nissue;    // issuing warps (each its own accumulator columns and commit barrier)
  uint32_t idesc2;   // alternation experiment (round 2): instruction descriptor of the "other" MMA shape, 0 = off
  int block;         // consecutive MMAs per shape (1 = hi, lo, hi, lo, ...; 8 = 8 x hi then 8 x lo)
  uint32_t a2_off;   // A start offset of the second shape
  int commit_every;  // commit experiment: a tcgen05.commit (to a barrier nobody waits on) after every commit_every-th group of 16 MMAs, 0 = off
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int BLOCK>
__global__ void __launch_bounds__(128, 1) bench_kernel_t(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint64_t dummy_bar;
  __shared__ long long tmax[4];
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot + (uint32_t)warp * 128u;
  uint64_t& bar = bars[warp];
  if ((threadIdx.x & 31) == 0) tmax[warp] = 0;
  uint32_t leader = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
  if (warp < c.nissue) {
    const uint32_t base = smem_u32(smem);
    long long best = 1ll << 60;
    for (int r = 0; r < c.reps; ++r) {
      // descriptors hoisted: only the low word (start address) changes, the issue loop is 1 add + 1 MMA per instruction
      const uint64_t da0 = mk_desc(base + c.a_off, c.a_lbo, c.a_sbo, c.a_layout);
      const uint64_t db0 = mk_desc(base + c.b_off, c.b_lbo, c.b_sbo, c.b_layout);
      const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
      const uint32_t a_lo = (uint32_t)da0, b_lo = (uint32_t)db0;
      const uint32_t as = c.a_step >> 4, bs = c.b_step >> 4;
      long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < c.nmma; i += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(u & 7) * as);
          const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)(u & 7) * bs);
          const bool second = BLOCK > 0 && ((u / (BLOCK > 0 ? BLOCK : 1)) & 1);
          const uint64_t da2 = da + (uint64_t)(c.a2_off >> 4);
          if (leader)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                "l"(second ? da2 : da), "l"(db), "r"(second ? c.idesc2 : c.idesc));
        }
        if (c.commit_every && leader && ((i >> 4) % c.commit_every) == c.commit_every - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy_bar)) : "memory");
      }
      if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(&bar)), "r"((uint32_t)(r & 1))
            : "memory");
      }
      long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (leader) tmax[warp] = best;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = max(max(tmax[0], tmax[1]), max(tmax[2], tmax[3]));
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
  }
}

static void launch_bench(const Cfg& c, long long* out) {
  switch (c.idesc2 ? c.block : 0) {
    case 1: bench_kernel_t<1><<<148, 128, 200 * 1024>>>(c, out); break;
    case 2: bench_kernel_t<2><<<148, 128, 200 * 1024>>>(c, out); break;
    case 4: bench_kernel_t<4><<<148, 128, 200 * 1024>>>(c, out); break;
    case 8: bench_kernel_t<8><<<148, 128, 200 * 1024>>>(c, out); break;
    default: bench_kernel_t<0><<<148, 128, 200 * 1024>>>(c, out); break;
  }
}
static uint32_t idesc(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) |
         ((128u >> 4) << 24);
}

int main(int argc, char** argv) {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench_kernel_t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_kernel_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_kernel_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_kernel_t<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench_kernel_t<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Row { const char* name; bool a_mn, b_mn; uint32_t a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout, a_step, b_step, a_off, b_off; };
  const uint32_t B0 = 100 * 1024;
  Row rows[] = {
      // name                                  a_mn  b_mn   a_lbo a_sbo a_lay  b_lbo b_sbo b_lay  a_step b_step a_off b_off
      {"K-major SW128 canonical", false, false, 16, 1024, 2, 16, 1024, 2, 32, 32, 0, B0},
      {"K-major SW32 (conv_tc)", false, false, 16, 256, 6, 16, 256, 6, 0, 0, 0, B0},
      {"K-major noswz sbo128 lbo2048 (dense)", false, false, 2048, 128, 0, 2048, 128, 0, 0, 0, 0, B0},
      {"K-major noswz sbo160 lbo8704 (halo A), B dense", false, false, 8704, 160, 0, 2048, 128, 0, 16, 0, 0, B0},
      {"K-major noswz sbo160 lbo8704 aligned start", false, false, 8704, 160, 0, 2048, 128, 0, 0, 0, 0, B0},
      {"K-major noswz sbo256 lbo8704", false, false, 8704, 256, 0, 2048, 128, 0, 0, 0, 0, B0},
      {"K-major noswz sbo144 lbo8704", false, false, 8704, 144, 0, 2048, 128, 0, 0, 0, 0, B0},
      {"MN-major noswz A sbo2048 lbo128, B sbo2880 lbo160 (ds)", true, true, 128, 2048, 0, 160, 2880, 0, 256, 320, 0, B0},
      {"MN-major noswz A sbo2048 lbo128, B dense sbo128? (A only)", true, false, 128, 2048, 0, 2048, 128, 0, 256, 0, 0, B0},
      {"MN-major noswz A sbo2064 lbo128, B kmaj dense", true, false, 128, 2064, 0, 2048, 128, 0, 256, 0, 0, B0},
      {"MN-major noswz A sbo2176 lbo128, B kmaj dense", true, false, 128, 2176, 0, 2048, 128, 0, 256, 0, 0, B0},
      {"MN-major noswz A sbo128 lbo2048 (dense), B kmaj dense", true, false, 2048, 128, 0, 2048, 128, 0, 0, 0, 0, B0},
      {"K-major dense A, MN-major B sbo2880 lbo160", false, true, 2048, 128, 0, 160, 2880, 0, 0, 320, 0, B0},
      {"K-major dense A, MN-major B sbo128 lbo2048 dense", false, true, 2048, 128, 0, 2048, 128, 0, 0, 0, 0, B0},
      // round 2: exact-K pairing in conv_tc_halo.cu puts the second K half at an arbitrary (small) LBO
      {"K-major noswz sbo160 lbo16 (tap kw+1), B dense", false, false, 16, 160, 0, 2048, 128, 0, 16, 0, 0, B0},
      {"K-major noswz sbo160 lbo128, B dense", false, false, 128, 160, 0, 2048, 128, 0, 16, 0, 0, B0},
      {"K-major noswz sbo160 lbo2528 (kd wrap), B dense", false, false, 2528, 160, 0, 2048, 128, 0, 16, 0, 0, B0},
      {"K-major noswz sbo160 lbo11296 (group straddle), B dense", false, false, 11296, 160, 0, 2048, 128, 0, 16, 0, 0, B0},
      {"K-major noswz sbo160 lbo8704, B dense, A stepping 32 B", false, false, 8704, 160, 0, 2048, 128, 0, 32, 0, 0, B0},
  };
  if (argc > 1 && !strcmp(argv[1], "alt")) {
    // does switching the MMA shape (N = 2 Npad for A_hi x [B_hi;B_lo], N = Npad for A_lo x B_hi) between consecutive
    // instructions cost anything?  cycles per PAIR of MMAs for block sizes 1, 2, 4, 8 and for two same-shape MMAs
    printf("alternation: cycles per (N1, N2) pair, halo-A layout; block = consecutive MMAs of one shape\n");
    const int pairs[][2] = {{96, 48}, {64, 32}, {96, 96}, {48, 48}, {64, 64}, {32, 32}, {256, 128}};
    for (auto& pr : pairs) {
      printf("N1=%3d N2=%3d :", pr[0], pr[1]);
      for (int block : {1, 2, 4, 8}) {
        Cfg c;
        c.idesc = idesc(pr[0], false, false); c.idesc2 = idesc(pr[1], false, false);
        c.a_off = 0; c.b_off = B0; c.a_lbo = 8704; c.a_sbo = 160; c.a_layout = 0; c.b_lbo = 2048; c.b_sbo = 128; c.b_layout = 0;
        c.a_step = 16; c.b_step = 0; c.nmma = 512; c.reps = 4; c.nissue = 1; c.block = block; c.a2_off = 17408; c.commit_every = 0;
        launch_bench(c, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); break; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  block %d: %6.1f", block, 2.0 * (double)mx / c.nmma);
      }
      printf("\n");
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "commit")) {
    printf("commit cadence: cycles per (N1, N2) pair with a tcgen05.commit after every k x 16 MMAs (k = 0: none)\n");
    const int pairs[][2] = {{96, 48}, {64, 32}, {128, 128}};
    for (auto& pr : pairs) {
      printf("N1=%3d N2=%3d :", pr[0], pr[1]);
      for (int k : {0, 4, 2, 1}) {
        Cfg c;
        c.idesc = idesc(pr[0], false, false); c.idesc2 = idesc(pr[1], false, false);
        c.a_off = 0; c.b_off = B0; c.a_lbo = 8704; c.a_sbo = 160; c.a_layout = 0; c.b_lbo = 2048; c.b_sbo = 128; c.b_layout = 0;
        c.a_step = 16; c.b_step = 0; c.nmma = 512; c.reps = 4; c.nissue = 1; c.block = 1; c.a2_off = 17408; c.commit_every = k;
        launch_bench(c, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); break; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  every %2d MMAs: %6.1f", k * 16, 2.0 * (double)mx / c.nmma);
      }
      printf("\n");
    }
    return 0;
  }
  const int Ns[] = {32, 48, 64, 96, 128};
  printf("%-62s", "layout \\ N: cycles per MMA (M=128,K=16)");
  for (int N : Ns) printf(" %6d", N);
  printf("\n");
  int ri = 0;
  for (const Row& r : rows) for (int nissue = 1; nissue <= 4; ++nissue) {
    if (nissue == 1) ++ri;
    if (nissue != 1 || (ri != 1 && ri != 4 && ri < 15)) continue;
    printf("%-50s issuers=%d ", r.name, nissue);
    for (int N : Ns) {
      Cfg c;
      c.idesc = idesc(N, r.a_mn, r.b_mn);
      c.a_off = r.a_off; c.b_off = r.b_off;
      c.a_lbo = r.a_lbo; c.a_sbo = r.a_sbo; c.a_layout = r.a_layout;
      c.b_lbo = r.b_lbo; c.b_sbo = r.b_sbo; c.b_layout = r.b_layout;
      c.a_step = r.a_step; c.b_step = r.b_step;
      c.nmma = 512; c.reps = 4; c.nissue = nissue; c.idesc2 = 0; c.block = 1; c.a2_off = 0; c.commit_every = 0;
      launch_bench(c, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); break; }
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf(" %6.1f", (double)mx / c.nmma / nissue);
    }
    printf("\n");
    fflush(stdout);
  }
  return 0;
}
