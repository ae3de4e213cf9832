// Bring-up test for a 2-CTA (cta_group::2) tcgen05 GEMM: C[M,N] = A[M,K] · B[N,K]^T, bf16 in, f32 out.
// Cluster of 2 CTAs per 256 x BN output tile: CTA r loads A rows [r*128, +128) and B rows [r*BN/2, +BN/2) of the tile;
// the leader CTA issues tcgen05.mma.cta_group::2 (M = 256); each CTA drains its own 128 TMEM lanes.
// Standalone so that protocol bugs cannot touch the library: nvcc -arch=sm_100a gemm2cta.cu -lcuda -o gemm2cta
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mofo_b200/csrc/common.cuh"
using namespace mofo;
namespace mofo { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } }

constexpr int BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = 128 * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_leader, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {   // arrive on a (possibly remote) barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, float* C) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES), tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  auto smem_a = [&](int s) { return base + s * STAGE_BYTES; };
  auto smem_b = [&](int s) { return base + s * STAGE_BYTES + A_BYTES; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_n = N / BN;
  const int tile = blockIdx.x >> 1;
  const int m_blk = tile / num_n, n_blk = tile % num_n;
  const int kblocks = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    if (lane == 0) {   // producer (both CTAs): own halves, completion bytes land on the LEADER's full barrier
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t lbar = full_bar(stage) & PEER_MASK;
        if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
        else mbar_arrive_cluster(lbar);
        tma_load_2d_2sm(smem_a(stage), &tmA, lbar, kb * BK, m_blk * 256 + rank * 128);
        tma_load_2d_2sm(smem_b(stage), &tmB, lbar, kb * BK, n_blk * BN + rank * (BN / 2));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (leader && lane == 0) {   // MMA issuer: leader CTA only
      const uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t ad = umma_desc_kmajor(smem_a(stage)), bd = umma_desc_kmajor(smem_b(stage));
        for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        commit_2sm(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      commit_2sm(tfull_bar);
    }
  } else {   // epilogue: each CTA drains its own 128 lanes
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int row = m_blk * 256 + rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tc_wait_ld();
      if (row < M)
        for (int e = 0; e < 32; ++e) C[static_cast<size_t>(row) * N + n_blk * BN + c0 + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

static CUtensorMap make_map(void* p, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 512, N = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 256;
  cudaFree(0);
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((float)(rand() % 7 - 3));
  for (auto& v : hB) v = __float2bfloat16((float)(rand() % 5 - 2));
  __nv_bfloat16 *dA, *dB; float* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, (size_t)M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, (size_t)M * N * 4);
  CUtensorMap tA = make_map(dA, M, K, 128), tB = make_map(dB, N, K, BN / 2);
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  cudaFuncSetAttribute(gemm2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = (M / 256) * (N / BN);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gemm2cta_kernel<<<tiles * 2, 192, smem>>>(tA, tB, M, N, K, dC);
  cudaError_t err = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 2;
  std::vector<float> hC((size_t)M * N);
  cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; long bad = 0;
  if ((long)M * N * K <= 2e9) {
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += __bfloat162float(hA[(size_t)i * K + k]) * __bfloat162float(hB[(size_t)j * K + k]);
      double d = fabs(s - hC[(size_t)i * N + j]);
      if (!(d <= 1e-3)) { if (bad < 5) printf("mismatch (%d,%d): ref %f got %f\n", i, j, s, hC[(size_t)i * N + j]); ++bad; }
      if (d > maxerr) maxerr = d;
    }
    printf("M=%d N=%d K=%d  max abs err %.3g  mismatches %ld / %ld\n", M, N, K, maxerr, bad, (long)M * N);
  }
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) gemm2cta_kernel<<<tiles * 2, 192, smem>>>(tA, tB, M, N, K, dC);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("time %.1f us  %.0f TFLOP/s\n", ms * 100, 2.0 * M * N * K / (ms / 10 * 1e-3) / 1e12);
  return bad ? 1 : 0;
}
