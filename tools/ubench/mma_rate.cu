// tcgen05.mma issue/execute rate: one elected thread issues REPS back-to-back MMAs on (uninitialised) smem operands and
// waits for the commit; clock64 around it.  Variants: cta_group::1 (M=128) with A from smem or from TMEM, cta_group::2
// (M=256).  Every SM runs the same loop so that chip-level effects (power) are included.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 mma_rate.cu -o mma_rate
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../../mofo_b200/csrc/common.cuh"
using namespace mofo;
namespace mofo { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } }

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MODE 0: 1-CTA, A smem.  MODE 1: 1-CTA, A TMEM.  MODE 2: 2-CTA (cluster of 2), A smem.
// MODE 3: 1-CTA, both operands MN-major (wgrad layout).  MODE 4: 2-CTA, both MN-major.
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  constexpr bool TWO = MODE == 2 || MODE == 4;
  const bool leader = !TWO || ctarank() == 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else tmem_alloc(slot, 512);
  }
  tc_fence_before();
  if (TWO) csync(); else __syncthreads();
  tc_fence_after();
  uint32_t tb; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tb) : "r"(slot));
  tb = __shfl_sync(0xffffffffu, tb, 0);
  long long t0 = 0, t1 = 0;
  if (warp == 0 && leader && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(TWO ? 256 : 128, N, MODE >= 3, MODE >= 3);
    const uint64_t ad = MODE >= 3 ? umma_desc_mnmajor(sA, 8192) : umma_desc_kmajor(sA), bd = MODE >= 3 ? umma_desc_mnmajor(sB, 8192) : umma_desc_kmajor(sB);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 0) umma_bf16(tb, ad + 2 * k, bd + 2 * k, idesc, 1u);
        if (MODE == 1) umma_ts(tb, tb + 256 + 8 * k, bd + 2 * k, idesc);
        if (MODE == 2) umma_2sm(tb, ad + 2 * k, bd + 2 * k, idesc);
        if (MODE == 3) umma_bf16(tb, ad + 128 * k, bd + 128 * k, idesc, 1u);
        if (MODE == 4) umma_2sm(tb, ad + 128 * k, bd + 128 * k, idesc);
      }
    }
    if (TWO) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)1) : "memory");
    else tc_commit(bar);
    mbar_wait(bar, 0);
    t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before();
  if (TWO) csync(); else __syncthreads();
  if (warp == 0) {
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    else tmem_dealloc(tb, 512);
  }
}

template <int MODE, int N>
void run(const char* name, int reps, long long* d) {
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (MODE == 2 || MODE == 4) ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) {
    cudaLaunchKernelEx(&cfg, rate_kernel<MODE, N>, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (reps * 4);
  const double flop = 2.0 * (MODE == 2 ? 128 : 128) * N * 16;   // per SM per instruction
  printf("%-28s N=%3d: %7.1f clk / MMA  -> %6.0f flop/clk/SM\n", name, N, per, flop / per);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int reps = 512;
  run<0, 64>("1cta A=smem", reps, d);  run<0, 128>("1cta A=smem", reps, d);  run<0, 256>("1cta A=smem", reps, d);
  run<1, 64>("1cta A=tmem", reps, d);  run<1, 128>("1cta A=tmem", reps, d);  run<1, 256>("1cta A=tmem", reps, d);
  run<3, 128>("1cta MN-major", reps, d); run<3, 192>("1cta MN-major", reps, d); run<3, 256>("1cta MN-major", reps, d);
  run<4, 128>("2cta MN-major", reps, d); run<4, 256>("2cta MN-major", reps, d);
  run<2, 64>("2cta A=smem (per-SM N/2 of B)", reps, d); run<2, 128>("2cta A=smem", reps, d); run<2, 256>("2cta A=smem", reps, d);
  return 0;
}
