// Device helpers shared by the attention kernels (attention.cu: streaming kernels for any S; attention_small.cu:
// single-pass kernels for S <= 192): tile constants, tcgen05 wrappers for score / probability tiles.
#pragma once
#include "common.cuh"

namespace mofo {

constexpr int AT = 128;                       // rows per CTA tile (q rows in fwd/dq, kv rows in dkv)
constexpr int BT = 64;                        // inner (streamed) tile: kv rows in fwd/dq, q rows in dkv
constexpr int TILE_BYTES = AT * 64 * 2;       // 16 KB: [128 x 64] bf16
constexpr int HTILE_BYTES = BT * 64 * 2;      //  8 KB: [ 64 x 64] bf16
constexpr int ATT_THREADS = 256;              // 2 threads per tile row: each owns 32 of the 64 inner columns

__device__ __forceinline__ void check_align(uint32_t base) {
  if (base & 1023u) {
    if (threadIdx.x == 0) printf("mofo: dynamic shared memory base not 1024-B aligned (0x%x)\n", base);
    __trap();
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {   // single FMNMX3 on sm_100
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// D[128 x 64] = A[128 x 64] · B[64 x 64]^T, both K-major SW128 tiles straight from TMA (4 K-steps of 16)
__device__ __forceinline__ void mma_ab_t(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  const uint64_t a0 = umma_desc_kmajor(a_tile), b0 = umma_desc_kmajor(b_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, k != 0);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// D[128 x 64] (+)= P[128 x 64 bf16, in TMEM: lane = row, column k/2 holds elements (k, k+1)] · T[64 x 64] (MN-major smem)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The bf16 operand is written by the threads over their OWN fp32 accumulator columns: the thread pair of a row owns
// fp32 columns [0,32) and [32,64) of a 64-column region and stores its 32 bf16 values (16 packed columns) at region
// columns [0,16) resp. [32,48).  K-steps of 16 elements therefore start at columns {0, 8, 32, 40}.
__device__ __forceinline__ void mma_ptmem_t(uint32_t d_tmem, uint32_t p_tmem, uint32_t t_tile, bool accumulate) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
  const uint64_t b0 = umma_desc_mnmajor(t_tile, 8192);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16_ts(d_tmem, p_tmem + (k >> 1) * 32 + (k & 1) * 8, b0 + 128 * k, idesc, (accumulate || k != 0) ? 1u : 0u);
}
__device__ __forceinline__ void tmem_ld16a(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8_async(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16_async(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// same as tmem_store_bf16_row but WITHOUT the trailing tcgen05.wait::st: the caller waits once before handing over
__device__ __forceinline__ void tmem_store_bf16_row_async(uint32_t taddr, const float (&v)[32]) {
  uint32_t r[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) r[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_store_bf16_row(uint32_t taddr, const float (&v)[32]) {
  uint32_t pk[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
  tmem_st16(taddr, pk);
}

int get_tmap(CUtensorMap* out, const void* p, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);  // gemm.cu

// single-pass kernels (attention_small.cu); return MOFO_OK or an error status
constexpr int ATTN_SMALL_MAX_S = 192;
int attn_small_fwd(const void* qkv, int B, int S, int H, float scale, void* out, float* lse, cudaStream_t stream);
int attn_small_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int S, int H, float scale,
                   void* dqkv, cudaStream_t stream);

}  // namespace mofo
