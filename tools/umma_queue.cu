// How deep is the queue between an issuing thread and the tensor pipe?  Time for the elected lane to ISSUE k back-to-back
// tcgen05.mma (M=128, N=96, K=16; ~56 cycles each in the pipe) without waiting for completion, k = 1..40.  While the queue has
// room an issue costs a few cycles; once it is full every further issue waits for an MMA to retire (~56 cycles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_queue.bin tools/umma_queue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int K>
__device__ __forceinline__ long long issue_k(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t leader) {
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < K; ++i)
    if (leader)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da),
                   "l"(db), "r"(idesc));
  long long t1 = clock64();
  return t1 - t0;
}
__global__ void __launch_bounds__(128, 1) k(long long* out, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (warp == 0) {
    uint32_t leader = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    const uint32_t base = smem_u32(smem);
    auto mk = [](uint32_t a, uint32_t lbo, uint32_t sbo) { return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46); };
    const uint64_t da = mk(base, 8704, 160), db = mk(base + 32768, 2048, 128);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long res[12];
    int r = 0;
    uint32_t ph = 0;
#define RUN(K)                                                                                                          \
    {                                                                                                                   \
      long long best = 1ll << 60;                                                                                       \
      for (int rep = 0; rep < 3; ++rep) {                                                                               \
        long long t = issue_k<K>(tslot, da, db, idesc, leader);                                                         \
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory"); \
        uint32_t ok = 0;                                                                                                \
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(ph) : "memory"); \
        ph ^= 1;                                                                                                        \
        if (t < best) best = t;                                                                                         \
      }                                                                                                                 \
      res[r++] = best;                                                                                                  \
    }
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(6) RUN(8) RUN(12) RUN(16) RUN(24) RUN(32) RUN(48) RUN(64)
    if (leader && blockIdx.x == 0) for (int i = 0; i < 12; ++i) out[i] = res[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
  }
}
int main() {
  long long* out;
  cudaMalloc(&out, 12 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int ks[12] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
  for (int N : {96, 32, 256}) {
    k<<<148, 128, 100 * 1024>>>(out, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
    long long h[12];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d issue cycles for k MMAs:", N);
    for (int i = 0; i < 12; ++i) printf("  k=%d:%lld", ks[i], h[i]);
    printf("\n");
  }
  return 0;
}
