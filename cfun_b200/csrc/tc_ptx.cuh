// Inline-PTX wrappers for the Blackwell async machinery used by the tcgen05 kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05.{alloc,mma,commit,ld,fence}.  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace cfun {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a reported error, never as a hung GPU and never as silently wrong
// results.  On time-out the first offender records (site, block, thread, parity) in g_tc_debug and TRAPS: the launch
// fails, the CUDA context reports the error at the next synchronising call and every later library call returns
// CFUN_ERR_CUDA.  Bring-up builds (CFUN_NVCC_FLAGS=-DCFUN_TC_NOTRAP) fall through instead, so that the kernel drains and
// the host can read the record with cfun_tc_debug_status().
static __device__ int g_tc_debug[8];   // one copy per translation unit (no -rdc); each TU exports a reader
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int site = 0) {
  // The spin itself must be tight: the latency from a barrier's completion to the waiter's next instruction is on the critical
  // path of every producer/consumer hand-off (a system-scope flag load per failed try cost ~1 us of wake-up latency and ~15 %
  // of the conv kernels' time).  The debug flag is polled once per 256 failed tries.
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if ((spin & 255u) != 255u) continue;
    if (spin > (1u << 22) || *((volatile int*)&g_tc_debug[0]) != 0) {
      if (atomicCAS(&g_tc_debug[0], 0, site + 1) == 0) {
        g_tc_debug[1] = (int)blockIdx.x; g_tc_debug[2] = (int)blockIdx.y; g_tc_debug[3] = (int)threadIdx.x;
        g_tc_debug[4] = (int)parity; g_tc_debug[5] = (int)spin;
      }
#ifndef CFUN_TC_NOTRAP
      __trap();
#endif
      return;
    }
  }
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Issue-rate note (tools/umma_bench.cu, B200): one tcgen05.mma M=128 K=16 costs max(N/2, 32 + N/4) cycles in the tensor
// pipe (A fetch 4 KB + B fetch N*32 B at 128 B/clk from shared memory, independent of layout / swizzle).  A lone thread
// that rebuilds 64-bit descriptors per instruction issues one MMA per ~65-100 cycles and starves the pipe for N < 256:
// issuer loops therefore run warp-uniform (all 32 lanes execute the address arithmetic, so it lives in uniform
// registers), keep the descriptor high words constant, add 16-byte offsets to the low word, and only the elected lane
// executes the tcgen05 instructions.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(p));
  return p;
}
// start-address field of a shared-memory matrix descriptor: (address >> 4) in 14 bits.  The mask matters inside a
// thread-block cluster, where the shared-window address of CTA rank r carries r in its upper bits (they would spill into
// the LBO field of a hoisted low word).
__device__ __forceinline__ uint32_t desc_addr(uint32_t saddr) { return (saddr & 0x3FFFFu) >> 4; }
__device__ __forceinline__ uint64_t desc_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }
// accumulate = compile-time true: no predicate set-up instruction
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// generic K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, SM100 "version 1").
//   swizzle_bytes in {32, 64, 128}: rows are swizzle_bytes long, 8-row groups are 8*swizzle_bytes apart (SBO);
//   the LBO field is unused for K-major swizzled operands (set to 1 like CUTLASS does).
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr, int swizzle_bytes) {
  const uint64_t layout = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * swizzle_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_tensor_map_encoder();   // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no libcuda link)

// InstanceNorm statistics from a conv epilogue: v[16] = 16 channels of this lane's output voxel (zeros where the lane holds
// no voxel).  A butterfly transpose-reduce (31 shuffles for the 32 values v, v^2) leaves lane L with the warp total of
// channel L & 15 (L < 16: sum, L >= 16: sum of squares).
__device__ __forceinline__ float warp_stats16_reduce(const float* v, int lane) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = v[i]; a[16 + i] = v[i] * v[i]; }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? a[k] : a[k + off];
      const float keep = upper ? a[k + off] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}
// ... straight into acc_n[(ch0 + c) * 2 + stat], the layout cfun_instnorm_stats accumulates into (one double atomic per lane)
__device__ __forceinline__ void warp_stats16(const float* v, double* __restrict__ acc_n, int ch0, int Cout, int lane) {
  const float t = warp_stats16_reduce(v, lane);
  const int c = ch0 + (lane & 15);
  if (c < Cout) atomicAdd(acc_n + (long long)c * 2 + (lane >> 4), (double)t);
}
// ... or into this warp's private shared-memory accumulator s_warp[2][64] (plain read-modify-write: lane L is the only
// writer of its slot, so the sums are run-to-run deterministic), which conv_tc_halo.cu flushes every few tiles -- every CTA
// works on the same sample at the same time, so per-tile global atomics would all land on the same 2 Cout doubles
__device__ __forceinline__ void warp_stats16_shared(const float* v, float* s_warp, int ch0, int Cout, int lane) {
  const float t = warp_stats16_reduce(v, lane);
  const int c = ch0 + (lane & 15);
  if (c < Cout) s_warp[(lane >> 4) * 64 + c] += t;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

}  // namespace cfun
