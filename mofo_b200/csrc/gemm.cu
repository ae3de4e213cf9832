// tcgen05 / TMEM / TMA GEMM kernels for the dense layers of the MOFO pretraining step.
//
//   gemm_tn_kernel    C[M,N] = epi(A[M,K] · B[N,K]^T)   persistent, warp-specialised:
//                       warp 16 = TMA producer, warp 17 = tcgen05.mma issuer (+TMEM owner),
//                       warps 0-15 = epilogue (TMEM -> registers -> fused epilogue -> smem staging -> coalesced global;
//                       the epilogue's second input is prefetched one chunk ahead).
//                     128 x BN output tiles (BN = 128/192/256), BLOCK_K = 64 (one 128-byte swizzle row),
//                     multi-stage smem ring, two TMEM accumulators so the epilogue of tile i overlaps the
//                     main loop of tile i+1.
//   gemm_tn2_kernel   the same GEMM and epilogues on CTA PAIRS (cluster of 2, tcgen05 cta_group::2): 256 x BN tiles
//                     (BN = 128/192/224/256), each CTA loads its 128 A rows and half of the B tile; used whenever the
//                     cost model in pick_tiles() says so (these layers are bound by L2 -> SM operand traffic).
//   gemm_wgrad_kernel dW[N,K] += dY[M,N]^T · X[M,K]     both operands MN-major (reduction over rows),
//                     split over M across CTAs, 16-byte vector red.add into the gradient arena; the bias
//                     gradient (column sums of dY) rides along as one extra N=16 MMA against a ones tile.
//                     (A CTA-pair variant of wgrad was measured 3-7% SLOWER on every layer shape - wgrad is not
//                     bound by L2 -> SM operand traffic, its tensor pipe is ~55-65% busy - and was dropped.)
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 16;                      // warps 0-15: epilogue; warp 16: TMA producer; warp 17: MMA issuer
constexpr int GEMM_THREADS = (EPI_WARPS + 2) * 32;
constexpr int A_TILE_BYTES = BM * BK * 2;          // 16 KB
constexpr int STAGING_BYTES = EPI_WARPS * 2048;    // per epilogue warp: 32 rows x 64 B
constexpr int WG_EPI_WARPS = 8;                    // wgrad kernel: warps 0-7 epilogue, 8 producer, 9 MMA
constexpr int WG_THREADS = (WG_EPI_WARPS + 2) * 32;
constexpr int WG_STAGING_BYTES = WG_EPI_WARPS * 4096;

struct EpiParams {
  const float* bias;
  const float* resid;
  int ldr;
  const __nv_bfloat16* aux;
  int ldaux;
  const float* pos;
  const int32_t* row_idx;
  int group_rows, out_group_rows;
  void* out0;
  int ldo0;
  void* out1;
  int ldo1;
  const float* row_scale;    // BIAS_RESID_F32 only: out = resid + row_scale[m / group_rows] * (acc + bias)   (DropPath)
};

// erf-form GELU (nn.GELU default).  erf by Abramowitz-Stegun 7.1.25 (3-term, |abs err| <= 2.5e-5, below bf16
// resolution of the outputs it feeds): one MUFU.RCP + one MUFU.EX2 + a handful of FMA-class ops instead of
// libdevice erff's branchy ~25.  The epilogue is instruction-bound, so every instruction per element counts.
// e = exp(-x^2/2) is shared with the pdf term of the derivative.
__device__ __forceinline__ void erf_parts(float x, float& erf_v, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.47047f, z, 1.0f));
  const float poly = t * fmaf(t, fmaf(t, 0.7478556f, -0.0958798f), 0.3480242f);
  e = exp2f(-1.4426950408889634f * z * z);
  erf_v = copysignf(fmaf(-poly, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float er, e;
  erf_parts(x, er, e);
  const float hx = 0.5f * x;
  return fmaf(hx, er, hx);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float er, e;
  erf_parts(x, er, e);
  return fmaf(x * 0.3989422804014327f, e, fmaf(0.5f, er, 0.5f));
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// wgrad staging tile: 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t stg_addr(uint32_t stg, int row, int chunk) {
  return stg + static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}
// ---- gemm_tn per-warp staging tile: 32 rows x 64 B (32 bf16 or 16 f32 output columns per row) ------------------
// 16-byte chunks XOR-swizzled by ((row >> 1) & 3): conflict-free both for "own row" accesses (thread `lane` touches
// row `lane`) and for "cooperative" accesses (instruction i touches rows 8i..8i+7, 4 lanes per row), so every global
// instruction of the epilogue moves 8 full 64-byte row segments.
__device__ __forceinline__ uint32_t stg64(uint32_t stg, int row, int chunk) {
  return stg + static_cast<uint32_t>(row) * 64u + (static_cast<uint32_t>(chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// cooperative 64-byte-row access: 4 instructions cover 32 rows; lane -> (row 8i + lane/4, chunk lane%4)
template <typename RowPtr>
__device__ __forceinline__ void coop_fetch(uint4 (&pre)[4], int lane, int row_base, int M, int valid_chunks, RowPtr rowptr) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int grow = row_base + 8 * i + (lane >> 2), chunk = lane & 3;
    pre[i] = make_uint4(0, 0, 0, 0);
    if (grow < M && chunk < valid_chunks) pre[i] = __ldg(reinterpret_cast<const uint4*>(rowptr(grow)) + chunk);
  }
}
__device__ __forceinline__ void coop_stage(uint32_t stg, int lane, const uint4 (&pre)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts128(stg64(stg, 8 * i + (lane >> 2), lane & 3), pre[i]);
}
template <typename RowPtr>
__device__ __forceinline__ void coop_store(uint32_t stg, int lane, int row_base, int M, int valid_chunks, RowPtr rowptr) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + (lane >> 2), chunk = lane & 3;
    const int grow = row_base + row;
    if (grow < M && chunk < valid_chunks) *(reinterpret_cast<uint4*>(rowptr(grow)) + chunk) = lds128(stg64(stg, row, chunk));
  }
}

// ---- bias + erf-GELU epilogue of one chunk (32 rows x 32 columns per warp), MOFO_GELU_V2 -------------------------------
// Same arithmetic as erf_parts() with the constants folded (t = 1/(1 + p|u|/sqrt2), e = exp2(-log2e/2 * u^2),
// Phi = 0.5 + copysign(0.5, u) * (1 - poly(t) e), gelu = u Phi, gelu' = Phi + u pdf(u)): 14 FMA-class + 2 MUFU per element.
// What changed the kernel's speed is not the count but the SHAPE of the code: the round-2 ncu profile showed the epilogue
// warps issuing 0.54 instructions per cycle and scheduler with ~4 warps each, and the SASS showed why - with the whole
// chunk's bias slice (32 registers) held next to the 32 accumulators ptxas had serialised the elements (MUFU.RCP / MUFU.EX2
// alternating one element at a time, a ~100-cycle dependent chain per element, two in flight).  Here eight elements go
// through every stage together (eight independent chains per warp), the bias arrives eight columns at a time one batch
// ahead, and each batch's gelu' words leave for the staging tile as soon as they exist.
#ifndef MOFO_GELU_V2
#define MOFO_GELU_V2 1
#endif
#ifndef MOFO_EPI_PREFETCH
#define MOFO_EPI_PREFETCH 1      // chunks of the epilogue's second input kept in flight per warp; 2 measured 2-4 % SLOWER (below)
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// Pulls the bias slices of this warp's chunks of the coming tile into L1 while the tile's main loop is still running (lane i
// takes the warp's i-th chunk): the epilogue's broadcast bias loads then hit L1 (~40 clk) instead of L2 (~700 clk).  ncu,
// round 2: the bias add was the top stall of the GELU epilogue (long scoreboard, 12 % of all samples).
template <int COLS, int NCHUNK>
__device__ __forceinline__ void prefetch_bias(const float* bias, int lane, int grp, int n_tile0, int N) {
  const int ch = grp + lane * (EPI_WARPS / 4);
  const int n0 = n_tile0 + ch * COLS;
  if (bias != nullptr && ch < NCHUNK && n0 < N) {
    prefetch_l1(bias + n0);
    prefetch_l1(bias + min(n0 + COLS, N) - 1);
  }
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool BIAS>
__device__ __forceinline__ void gelu_chunk(const EpiParams& ep, uint32_t stg, int lane, int row_base, int M, int n0, int N,
                                           uint32_t taddr) {
  constexpr float kP = 0.47047f * 0.70710678118654752f;          // A-S 7.1.25's p, for z = |u| / sqrt(2)
  constexpr float kE = -0.5f * 1.4426950408889634f;              // exp(-u^2 / 2) = exp2(kE * u^2)
  constexpr float kD = 0.3989422804014327f;                      // 1 / sqrt(2 pi)
  const int rem = N - n0;
  const int valid_chunks = rem >= 32 ? 4 : rem / 8;
  // bias columns past N are never stored: their loads are redirected to the last valid float4 instead of being predicated
  // (a predicated load needs its zero default materialised: 28 CS2R per chunk in the first version of this function)
  const float4* bp = reinterpret_cast<const float4*>(ep.bias + n0);
  const int qmax = (rem >= 32 ? 32 : rem) / 4 - 1;
  auto bias4 = [&](int q) { return BIAS ? __ldg(bp + min(q, qmax)) : make_float4(0.f, 0.f, 0.f, 0.f); };
  uint32_t r[32];
  tmem_ld32(taddr, r);
  float4 b0 = bias4(0), b1 = bias4(1);           // requested before the accumulator wait: the two latencies overlap
  tc_wait_ld();
  uint32_t packed[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 nb0 = b0, nb1 = b1;
    if (c < 3) { nb0 = bias4(2 * c + 2); nb1 = bias4(2 * c + 3); }
    float u[8], t[8], e[8];
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = __uint_as_float(r[c * 8 + k]) + bb[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {          // GELU acts on the bf16-rounded pre-activation (what F.linear returns under autocast)
      const uint32_t ub = pack_bf16(u[2 * k], u[2 * k + 1]);
      u[2 * k] = bf16_lo(ub); u[2 * k + 1] = bf16_hi(ub);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = rcp_approx(fmaf(kP, fabsf(u[k]), 1.0f));
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = ex2_approx(kE * (u[k] * u[k]));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float poly = t[k] * fmaf(t[k], fmaf(t[k], 0.7478556f, -0.0958798f), 0.3480242f);
      t[k] = fmaf(copysignf(0.5f, u[k]), fmaf(-poly, e[k], 1.0f), 0.5f);            // Phi(u)
    }
    uint32_t dp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      packed[c * 4 + k] = pack_bf16(u[2 * k] * t[2 * k], u[2 * k + 1] * t[2 * k + 1]);
      dp[k] = pack_bf16(fmaf(u[2 * k] * kD, e[2 * k], t[2 * k]), fmaf(u[2 * k + 1] * kD, e[2 * k + 1], t[2 * k + 1]));
    }
    sts128(stg64(stg, lane, c), make_uint4(dp[0], dp[1], dp[2], dp[3]));
    b0 = nb0; b1 = nb1;
  }
  __syncwarp();
  coop_store(stg, lane, row_base, M, valid_chunks,
             [&](int g) { return reinterpret_cast<__nv_bfloat16*>(ep.out0) + static_cast<size_t>(g) * ep.ldo0 + n0; });
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c)
    sts128(stg64(stg, lane, c), make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]));
  __syncwarp();
  coop_store(stg, lane, row_base, M, valid_chunks,
             [&](int g) { return reinterpret_cast<__nv_bfloat16*>(ep.out1) + static_cast<size_t>(g) * ep.ldo1 + n0; });
  __syncwarp();
}

template <int EPI>
struct EpiTraits {
  static constexpr bool F32_OUT = (EPI == MOFO_EPI_BIAS_RESID_F32 || EPI == MOFO_EPI_BIAS_POS_F32);
  static constexpr int COLS = F32_OUT ? 16 : 32;                 // 64 output bytes per row per chunk
  static constexpr bool HAS_BIAS = (EPI == MOFO_EPI_BIAS_BF16 || EPI == MOFO_EPI_BIAS_GELU_BF16 ||
                                    EPI == MOFO_EPI_BIAS_RESID_F32 || EPI == MOFO_EPI_BIAS_POS_F32);
  static constexpr bool HAS_OPERAND = (EPI == MOFO_EPI_BIAS_RESID_F32 || EPI == MOFO_EPI_BIAS_POS_F32 ||
                                       EPI == MOFO_EPI_GELU_BWD_BF16);   // a second [M,N] input read by the epilogue
};

// global -> registers prefetch of the epilogue's additive / multiplicative operand for one chunk
template <int EPI>
__device__ __forceinline__ void operand_fetch(const EpiParams& ep, uint4 (&pre)[4], int lane, int row_base, int M, int n0, int N) {
  using T = EpiTraits<EPI>;
  const int rem = N - n0;
  const int valid_chunks = rem >= T::COLS ? 4 : (T::F32_OUT ? rem / 4 : rem / 8);
  if (EPI == MOFO_EPI_BIAS_RESID_F32)
    coop_fetch(pre, lane, row_base, M, valid_chunks, [&](int g) { return ep.resid + static_cast<size_t>(g) * ep.ldr + n0; });
  else if (EPI == MOFO_EPI_BIAS_POS_F32)
    coop_fetch(pre, lane, row_base, M, valid_chunks, [&](int g) { return ep.pos + static_cast<size_t>(ep.row_idx[g]) * N + n0; });
  else if (EPI == MOFO_EPI_GELU_BWD_BF16)
    coop_fetch(pre, lane, row_base, M, valid_chunks, [&](int g) { return ep.aux + static_cast<size_t>(g) * ep.ldaux + n0; });
}

// One warp, one chunk (64 output bytes per row) starting at column n0; `taddr` = this warp's TMEM lanes + chunk column.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, uint32_t stg, int lane, int row_base, int M, int n0,
                                               int N, uint32_t taddr, const uint4 (&pre)[4]) {
  using T = EpiTraits<EPI>;
  constexpr int COLS = T::COLS;
#if MOFO_GELU_V2
  if constexpr (EPI == MOFO_EPI_BIAS_GELU_BF16) {
    if (ep.bias != nullptr) gelu_chunk<true>(ep, stg, lane, row_base, M, n0, N, taddr);
    else gelu_chunk<false>(ep, stg, lane, row_base, M, n0, N, taddr);
    return;
  }
#endif
  const int rem = N - n0;                                   // > 0
  const int valid_chunks = rem >= COLS ? 4 : (T::F32_OUT ? rem / 4 : rem / 8);
  float v[COLS];
  // the chunk's bias slice is requested BEFORE the accumulator read is waited for, so the two latencies overlap (the
  // bias add used to sit behind a long-scoreboard stall in every chunk: ncu source view, round 2)
  float4 bv[COLS / 4];
  const bool has_bias = T::HAS_BIAS && ep.bias != nullptr;
  if (has_bias) {
#pragma unroll
    for (int c = 0; c < COLS / 4; ++c)
      bv[c] = n0 + c * 4 < N ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (COLS == 32) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e % COLS] = __uint_as_float(r[e]);
  } else {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e % COLS] = __uint_as_float(r[e]);
  }
  if (has_bias) {
#pragma unroll
    for (int c = 0; c < COLS / 4; ++c) {
      v[c * 4 + 0] += bv[c].x; v[c * 4 + 1] += bv[c].y; v[c * 4 + 2] += bv[c].z; v[c * 4 + 3] += bv[c].w;
    }
  }
  if (EPI == MOFO_EPI_BIAS_RESID_F32 && ep.row_scale != nullptr) {
    // stochastic depth (modeling_finetune.py:28, 218-219): the residual BRANCH of sample b is scaled by mask_b / keep_prob
    const int grow = min(row_base + lane, M - 1);
    const float sc = __ldg(ep.row_scale + grow / ep.group_rows);
#pragma unroll
    for (int e = 0; e < COLS; ++e) v[e] *= sc;
  }
  auto out_row = [&](int grow) -> size_t {
    return EPI == MOFO_EPI_BIAS_POS_F32
               ? static_cast<size_t>(grow / ep.group_rows) * ep.out_group_rows + (grow % ep.group_rows)
               : static_cast<size_t>(grow);
  };
  if (T::HAS_OPERAND) {            // transpose the prefetched operand rows to "own row" through the staging tile
    coop_stage(stg, lane, pre);
    __syncwarp();
  }
  if (T::F32_OUT) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 q = lds128(stg64(stg, lane, c));
      v[c * 4 + 0] += __uint_as_float(q.x); v[c * 4 + 1] += __uint_as_float(q.y);
      v[c * 4 + 2] += __uint_as_float(q.z); v[c * 4 + 3] += __uint_as_float(q.w);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(stg64(stg, lane, c), make_uint4(__float_as_uint(v[c * 4]), __float_as_uint(v[c * 4 + 1]),
                                              __float_as_uint(v[c * 4 + 2]), __float_as_uint(v[c * 4 + 3])));
    __syncwarp();
    coop_store(stg, lane, row_base, M, valid_chunks,
               [&](int g) { return reinterpret_cast<float*>(ep.out0) + out_row(g) * ep.ldo0 + n0; });
    __syncwarp();
  } else {
    if (EPI == MOFO_EPI_GELU_BWD_BF16) {      // aux holds gelu'(u) saved by the forward epilogue: one multiply per element
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = lds128(stg64(stg, lane, c));
        v[c * 8 + 0] *= bf16_lo(a.x); v[c * 8 + 1] *= bf16_hi(a.x);
        v[c * 8 + 2] *= bf16_lo(a.y); v[c * 8 + 3] *= bf16_hi(a.y);
        v[c * 8 + 4] *= bf16_lo(a.z); v[c * 8 + 5] *= bf16_hi(a.z);
        v[c * 8 + 6] *= bf16_lo(a.w); v[c * 8 + 7] *= bf16_hi(a.w);
      }
      __syncwarp();
    }
    uint32_t packed[COLS / 2];
    if (EPI == MOFO_EPI_BIAS_GELU_BF16) {
      // GELU acts on the bf16-rounded pre-activation u (what F.linear returns under autocast).  One erf/exp evaluation
      // yields both gelu(u) (out1, feeds fc2) and gelu'(u) (out0, kept for backward instead of u itself).
      uint32_t dpacked[COLS / 2];
#pragma unroll
      for (int e = 0; e < COLS / 2; ++e) {
        const uint32_t ub = pack_bf16(v[2 * e], v[2 * e + 1]);
        const float u0 = bf16_lo(ub), u1 = bf16_hi(ub);
        float er0, e0, er1, e1;
        erf_parts(u0, er0, e0);
        erf_parts(u1, er1, e1);
        const float h0 = 0.5f * u0, h1 = 0.5f * u1;
        packed[e] = pack_bf16(fmaf(h0, er0, h0), fmaf(h1, er1, h1));
        dpacked[e] = pack_bf16(fmaf(u0 * 0.3989422804014327f, e0, fmaf(0.5f, er0, 0.5f)),
                               fmaf(u1 * 0.3989422804014327f, e1, fmaf(0.5f, er1, 0.5f)));
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts128(stg64(stg, lane, c), make_uint4(dpacked[c * 4], dpacked[c * 4 + 1], dpacked[c * 4 + 2], dpacked[c * 4 + 3]));
      __syncwarp();
      coop_store(stg, lane, row_base, M, valid_chunks,
                 [&](int g) { return reinterpret_cast<__nv_bfloat16*>(ep.out0) + static_cast<size_t>(g) * ep.ldo0 + n0; });
      __syncwarp();
    } else {
#pragma unroll
      for (int e = 0; e < COLS / 2; ++e) packed[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(stg64(stg, lane, c), make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]));
    __syncwarp();
    {
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(EPI == MOFO_EPI_BIAS_GELU_BF16 ? ep.out1 : ep.out0);
      const int ldo = EPI == MOFO_EPI_BIAS_GELU_BF16 ? ep.ldo1 : ep.ldo0;
      coop_store(stg, lane, row_base, M, valid_chunks, [&](int g) { return outp + static_cast<size_t>(g) * ldo + n0; });
    }
    __syncwarp();
  }
}

template <int BN>
struct TnCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BN == 128 ? 6 : 4;
  static constexpr int TMEM_COLS = BN == 128 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               EpiParams ep) {
  using Cfg = TnCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto smem_a = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return base + s * Cfg::STAGE_BYTES + A_TILE_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform role
  const int num_n = (N + BN - 1) / BN, num_m = (M + BM - 1) / BM;
  const int tiles = num_m * num_n;
  const int kblocks = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == EPI_WARPS + 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // uniform register -> tcgen05.mma issues without a per-lane loop
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_trigger();

  if (warp == EPI_WARPS) {
    if (elect_one()) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          tma_load_2d(smem_a(stage), &tmA, full_bar(stage), kb * BK, m_blk * BM);
          tma_load_2d(smem_b(stage), &tmB, full_bar(stage), kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    if (elect_one()) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_kmajor(smem_a(stage)), bdesc = umma_desc_kmajor(smem_b(stage));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {  // ===== epilogue warps: quarter = TMEM lane group, grp = which chunks of the tile =====
    using T = EpiTraits<EPI>;
    constexpr int COLS = T::COLS;
    constexpr int NCHUNK = BN / COLS;
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t stg = stg_base + warp * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row_base = m_blk * BM + quarter * 32;
      // the epilogue's second input (residual / gelu' / pos rows) for this warp's first TWO chunks is requested here, while
      // the tile's main loop still runs, and stays two chunks ahead afterwards: one chunk of work (~200 instructions) does
      // not cover an HBM round trip (ncu, round 2: 50-60 % of these kernels' stall samples were long-scoreboard waits on
      // exactly these registers, at 23-29 % issue utilisation)
      constexpr int CSTEP = EPI_WARPS / 4;
      uint4 pre[4], pre2[4];
      if (T::HAS_OPERAND) {
        if (n_blk * BN + grp * COLS < N) operand_fetch<EPI>(ep, pre, lane, row_base, M, n_blk * BN + grp * COLS, N);
        if (MOFO_EPI_PREFETCH > 1 && grp + CSTEP < NCHUNK && n_blk * BN + (grp + CSTEP) * COLS < N)
          operand_fetch<EPI>(ep, pre2, lane, row_base, M, n_blk * BN + (grp + CSTEP) * COLS, N);
      }
      if (T::HAS_BIAS) prefetch_bias<COLS, NCHUNK>(ep.bias, lane, grp, n_blk * BN, N);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = grp; ch < NCHUNK; ch += CSTEP) {
        const int n0 = n_blk * BN + ch * COLS;
        if (n0 >= N) break;
        uint4 cur[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = pre[i];
        if (T::HAS_OPERAND) {
          if (MOFO_EPI_PREFETCH > 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[i] = pre2[i];
            const int n_fetch = n0 + 2 * CSTEP * COLS;
            if (ch + 2 * CSTEP < NCHUNK && n_fetch < N) operand_fetch<EPI>(ep, pre2, lane, row_base, M, n_fetch, N);
          } else {
            const int n_next = n0 + CSTEP * COLS;
            if (ch + CSTEP < NCHUNK && n_next < N) operand_fetch<EPI>(ep, pre, lane, row_base, M, n_next, N);
          }
        }
        epilogue_chunk<EPI>(ep, stg, lane, row_base, M, n0, N, taddr + ch * COLS, cur);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// gemm_tn2_kernel: the same GEMM on CTA PAIRS (thread-block cluster of 2, tcgen05 cta_group::2).
// A pair owns a 256 x BN output tile: CTA r loads A rows [r*128, +128) and B rows [r*BN/2, +BN/2) of the tile, the
// leader CTA issues one M=256 tcgen05.mma that reads both CTAs' shared memory, and each CTA drains its own 128 TMEM
// lanes through the same fused epilogues.  Per SM and per k-block this moves 16 KB + BN*64 B from L2 instead of
// 16 KB + BN*128 B: these GEMMs (K = 384..3072, ~85 flop/B per 128 x 256 tile) are bound by L2 -> SM bandwidth
// (~43 B/clk/SM), not by the tensor pipe, so halving the B traffic is what buys time.
// Barrier protocol (all barriers live at the same smem offset in both CTAs):
//   full[s]   leader's only, count 2: leader's expect_tx(2 x stage bytes) + the peer producer's remote arrive;
//             both producers' TMA bytes complete on it
//   empty[s]  one per CTA, count 1: tcgen05.commit multicast to both CTAs when the MMAs that read the stage finish
//   tfull[a]  one per CTA, count 1: commit multicast when the pair tile's accumulator is complete
//   tempty[a] leader's only, count 2 x EPI_WARPS: every epilogue warp of both CTAs arrives after draining
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;          // clears the CTA-rank bit of a shared::cluster address -> rank 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
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
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");   // NOT .release.cluster: that costs a MEMBAR.ALL.GPU per arrive
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {        // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}

template <int BN>
struct Tn2Cfg {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;          // per CTA
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = BN == 128 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                EpiParams ep) {
  using Cfg = Tn2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto smem_a = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return base + s * Cfg::STAGE_BYTES + A_TILE_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_n = (N + BN - 1) / BN, num_m2 = (M + 2 * BM - 1) / (2 * BM);
  const int tiles = num_m2 * num_n;
  const int kblocks = (K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(static_cast<uint32_t>(Cfg::TMEM_COLS)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  pdl_wait();
  pdl_trigger();

  if (warp == EPI_WARPS) {
    if (elect_one()) {  // ===== TMA producer (both CTAs): own A rows and own half of the B rows =====
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < tiles; tile += n_pairs) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t lbar = full_bar(stage) & PEER_MASK;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          else mbar_arrive_cluster(lbar);
          tma_load_2d_2sm(smem_a(stage), &tmA, lbar, kb * BK, m_blk * 2 * BM + static_cast<int>(rank) * BM);
          tma_load_2d_2sm(smem_b(stage), &tmB, lbar, kb * BK, n_blk * BN + static_cast<int>(rank) * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    if (leader && elect_one()) {  // ===== MMA issuer: leader CTA only =====
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < tiles; tile += n_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_kmajor(smem_a(stage)), bdesc = umma_desc_kmajor(smem_b(stage));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit_2sm(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {  // ===== epilogue warps (both CTAs): own 128 rows of the pair tile =====
    using T = EpiTraits<EPI>;
    constexpr int COLS = T::COLS;
    constexpr int NCHUNK = BN / COLS;
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t stg = stg_base + warp * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < tiles; tile += n_pairs) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row_base = m_blk * 2 * BM + static_cast<int>(rank) * BM + quarter * 32;
      // operand rows two chunks ahead, as in gemm_tn_kernel
      constexpr int CSTEP = EPI_WARPS / 4;
      uint4 pre[4], pre2[4];
      if (T::HAS_OPERAND) {
        if (n_blk * BN + grp * COLS < N) operand_fetch<EPI>(ep, pre, lane, row_base, M, n_blk * BN + grp * COLS, N);
        if (MOFO_EPI_PREFETCH > 1 && grp + CSTEP < NCHUNK && n_blk * BN + (grp + CSTEP) * COLS < N)
          operand_fetch<EPI>(ep, pre2, lane, row_base, M, n_blk * BN + (grp + CSTEP) * COLS, N);
      }
      if (T::HAS_BIAS) prefetch_bias<COLS, NCHUNK>(ep.bias, lane, grp, n_blk * BN, N);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = grp; ch < NCHUNK; ch += CSTEP) {
        const int n0 = n_blk * BN + ch * COLS;
        if (n0 >= N) break;
        uint4 cur[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = pre[i];
        if (T::HAS_OPERAND) {
          if (MOFO_EPI_PREFETCH > 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[i] = pre2[i];
            const int n_fetch = n0 + 2 * CSTEP * COLS;
            if (ch + 2 * CSTEP < NCHUNK && n_fetch < N) operand_fetch<EPI>(ep, pre2, lane, row_base, M, n_fetch, N);
          } else {
            const int n_next = n0 + CSTEP * COLS;
            if (ch + CSTEP < NCHUNK && n_next < N) operand_fetch<EPI>(ep, pre, lane, row_base, M, n_next, N);
          }
        }
        epilogue_chunk<EPI>(ep, stg, lane, row_base, M, n0, N, taddr + ch * COLS, cur);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc) & PEER_MASK);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync();      // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers / TMEM
  if (warp == EPI_WARPS + 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(static_cast<uint32_t>(Cfg::TMEM_COLS)) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// wgrad: dW[n, k] += sum_m dY[m, n] * X[m, k]
// ---------------------------------------------------------------------------------------------------
// NT = 128-row n-tiles per CTA: the CTA owns an (NT*128) x BNW block of dW, i.e. NT accumulators that share every X tile
// (one tcgen05.mma per n-tile and K-step against the same B operand).  These GEMMs are bound by L2 -> SM operand traffic
// (a 128 x 256 tile is 85 flop per operand byte, the SM would need ~96 B/clk at the tensor rate and gets ~45-60), so
// what NT buys is operand bytes per flop: 2 x 192 -> 112 flop/B, 3 x 128 -> 98 flop/B.  TMEM: NT*BNW accumulator columns
// + NT*16 for the bias-gradient MMAs.
template <int BNW, int NT>
struct WgCfg {
  static constexpr int A_BYTES = NT * 2 * 64 * 128;       // per n-tile: two 64-wide n panels x 64 m rows
  static constexpr int B_BYTES = (BNW / 64) * 64 * 128;   // BNW/64 panels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = NT > 1 ? 3 : (BNW == 128 ? 6 : 4);
  static constexpr int ACC_COLS = NT * BNW + NT * 16;
  static constexpr int TMEM_COLS = ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators do not fit tensor memory");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + WG_STAGING_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BNW, int NT>
__device__ __forceinline__ void wgrad_body(const CUtensorMap& tmY, const CUtensorMap& tmX, int M, int N, int K,
                                           float* __restrict__ dW, int ldw, int kb_per_split, float* __restrict__ dbias,
                                           int skip_lo, int skip_hi, int tile_idx, int split_idx) {
  using Cfg = WgCfg<BNW, NT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BIAS_COL = NT * BNW;                    // first TMEM column of the bias-gradient accumulators
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + WG_STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  auto smem_a = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform role
  const int num_k = (K + BNW - 1) / BNW;
  const int n_blk = tile_idx / num_k, k_blk = tile_idx % num_k;          // n_blk counts (NT*128)-row blocks
  const int kblocks_total = (M + BK - 1) / BK;
  const int kb0 = split_idx * kb_per_split;
  const int kb1 = min(kblocks_total, kb0 + kb_per_split);
  const int nkb = kb1 - kb0;   // >= 1 by construction of the grid
  // bias gradient = column sums of dY = dY^T · 1: one extra N=16 MMA per K-step against an all-ones tile, issued by
  // the CTAs of the first k-tile only.  The ones tile (16 reduction rows x 128 B) aliases the epilogue staging area,
  // which is not touched before the last MMA has completed.
  const bool do_bias = dbias != nullptr && k_blk == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (do_bias && threadIdx.x < 128) {
    sts128(stg_base + threadIdx.x * 16, make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u));
    fence_proxy_async_smem();
  }
  if (warp == WG_EPI_WARPS && lane == 0) { tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmX); }
  if (warp == WG_EPI_WARPS + 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // uniform register -> tcgen05.mma issues without a per-lane loop
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_trigger();

  if (warp == WG_EPI_WARPS) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        const int m0 = (kb0 + i) * BK;
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
#pragma unroll
        for (int p = 0; p < 2 * NT; ++p) tma_load_2d(smem_a(stage) + p * 8192, &tmY, full_bar(stage), n_blk * 128 * NT + p * 64, m0);
#pragma unroll
        for (int p = 0; p < BNW / 64; ++p) tma_load_2d(smem_b(stage) + p * 8192, &tmX, full_bar(stage), k_blk * BNW + p * 64, m0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == WG_EPI_WARPS + 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BNW, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t bdesc = umma_desc_mnmajor(smem_b(stage) + k * 2048, 8192);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint64_t adesc = umma_desc_mnmajor(smem_a(stage) + nt * 16384 + k * 2048, 8192);
            umma_bf16(tmem_base + nt * BNW, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            if (do_bias)
              umma_bf16(tmem_base + BIAS_COL + nt * 16, adesc, umma_desc_mnmajor(stg_base, 8192), umma_idesc_bf16(128, 16, 1, 1),
                        (i | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(tfull_bar);
    }
  } else {
    // epilogue: TMEM -> staging tile -> coalesced fp32 vector red.add into dW
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t stg = stg_base + warp * 4096;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int nt = 0; nt < NT; ++nt) {
    const int n_base = (n_blk * NT + nt) * 128 + quarter * 32;
    if (n_base - quarter * 32 >= N) break;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + nt * BNW;
    if (do_bias && grp == 0) {
      uint32_t r[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + BIAS_COL + nt * 16, r);
      tc_wait_ld();
      const int n = n_base + lane;
      if (n < N && !(n >= skip_lo && n < skip_hi)) atomicAdd(dbias + n, __uint_as_float(r[0]));
    }
#pragma unroll 1
    for (int ch = grp; ch < BNW / 32; ch += 2) {
      const int k0 = k_blk * BNW + ch * 32;
      if (k0 >= K) break;
      uint32_t r[32];
      tmem_ld32(taddr + ch * 32, r);
      tc_wait_ld();
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(stg_addr(stg, lane, c), make_uint4(r[c * 4], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]));
      __syncwarp();
      {   // 16-byte vector reductions (red.global.add.v4.f32): one instruction covers 4 rows x 128 B
        const int c4 = lane & 7, kk = k0 + 4 * c4;
        if (kk < K) {                                      // K % 8 == 0: a 4-column group is valid as a whole
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + (lane >> 3);
            const int n = n_base + row;
            if (n < N) {
              const uint4 q = lds128(stg_addr(stg, row, c4));
              atomicAdd(reinterpret_cast<float4*>(dW + static_cast<size_t>(n) * ldw + kk),
                        make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)));
            }
          }
        }
      }
      __syncwarp();
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_EPI_WARPS + 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BNW, int NT>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, int M, int N, int K,
                  float* __restrict__ dW, int ldw, int kb_per_split, float* __restrict__ dbias, int skip_lo, int skip_hi) {
  wgrad_body<BNW, NT>(tmY, tmX, M, N, K, dW, ldw, kb_per_split, dbias, skip_lo, skip_hi, blockIdx.x, blockIdx.y);
}

// Grouped launch: up to WG_MAX_GROUP independent weight-gradient problems with the same reduction length M in ONE grid
// (blockIdx.x walks the concatenated tile lists, blockIdx.y the common M split).  The four weight gradients of a transformer
// block are independent of each other and off the backward critical path; launched one by one at M = 5120 (the encoder)
// each pays its own ramp-up / drain (~7 us of a ~25 us kernel) on a half-filled machine.
constexpr int WG_MAX_GROUP = 4;
struct WgProblem {
  CUtensorMap tmY, tmX;
  float* dW;
  float* dbias;
  int N, K, ldw, skip_lo, skip_hi, tile0;       // tile0: first blockIdx.x of this problem
};
struct WgGroup {
  WgProblem p[WG_MAX_GROUP];
  int n, M, kb_per_split;
};

template <int BNW>
__global__ void __launch_bounds__(WG_THREADS, 1) gemm_wgrad_grouped_kernel(const __grid_constant__ WgGroup g) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < WG_MAX_GROUP; ++i)
    if (i < g.n && static_cast<int>(blockIdx.x) >= g.p[i].tile0) pi = i;
  const WgProblem& q = g.p[pi];
  wgrad_body<BNW, 1>(q.tmY, q.tmX, g.M, q.N, q.K, q.dW, q.ldw, g.kb_per_split, q.dbias, q.skip_lo, q.skip_hi,
                     static_cast<int>(blockIdx.x) - q.tile0, blockIdx.y);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct TmapKey {
  const void* p; uint64_t rows, cols, ld; uint32_t box_rows;
  bool operator<(const TmapKey& o) const {
    return std::tie(p, rows, cols, ld, box_rows) < std::tie(o.p, o.rows, o.cols, o.ld, o.box_rows);
  }
};
static std::map<TmapKey, CUtensorMap> g_tmaps;
static std::mutex g_tmap_mu;

int get_tmap(CUtensorMap* out, const void* p, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  TmapKey k{p, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(k);
    if (it != g_tmaps.end()) { *out = it->second; return MOFO_OK; }
  }
  int rc = make_tmap_bf16_2d(out, p, rows, cols, ld, box_rows);
  if (rc != MOFO_OK) return rc;
  std::lock_guard<std::mutex> g(g_tmap_mu);
  if (g_tmaps.size() > 4096) g_tmaps.clear();
  g_tmaps[k] = *out;
  return MOFO_OK;
}

template <int BN, int EPI>
static int launch_tn(const CUtensorMap& tA, const CUtensorMap& tB, int M, int N, int K, const EpiParams& ep, cudaStream_t s) {
  using Cfg = TnCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  MOFO_CUDA(launch_pdl(gemm_tn_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, tA, tB, M, N, K, ep));
  return MOFO_OK;
}

template <int BN, int EPI>
static int launch_tn2(const CUtensorMap& tA, const CUtensorMap& tB, int M, int N, int K, const EpiParams& ep, cudaStream_t s) {
  using Cfg = Tn2Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(gemm_tn2_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  // persistent grid = the number of CTA pairs that are co-resident (a pair that had to wait for a free TPC would
  // start its share of the tiles only after the others finished theirs)
  static int max_pairs = 0;
  if (max_pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() & ~1); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    MOFO_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tn2_kernel<BN, EPI>, &cfg));
    if (n < 1) { set_error("gemm_tn2: no co-resident CTA pair fits"); return MOFO_ERR_CUDA; }
    max_pairs = n < sm_count() / 2 ? n : sm_count() / 2;
    if (getenv("MOFO_GEMM_VERBOSE")) fprintf(stderr, "mofo: gemm_tn2<%d,%d> max active CTA pairs = %d\n", BN, EPI, n);
  }
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int grid = 2 * (tiles < max_pairs ? tiles : max_pairs);
  MOFO_CUDA(launch_pdl(gemm_tn2_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, tA, tB, M, N, K, ep));
  return MOFO_OK;
}

// Tile configuration = (CTA pairs or single CTAs) x BN.  Estimated cost = waves x per-tile time, the per-tile time
// being the slowest of the tensor pipe (bn/2 clk per k16 step), the L2 -> SM operand traffic (~43 B/clk/SM measured
// chip-wide; a pair member loads only half of the B tile) and the epilogue, plus a fixed per-tile overhead.
// Calibrated on B200 against tools/prof_gemm_shapes.py; MOFO_GEMM_2CTA=0/1 and MOFO_FORCE_BN override (tuning aids).
struct TileChoice { bool pairs; int bn; };
static TileChoice pick_tiles(int M, int N, int K) {
  static const int forced_bn = [] { const char* e = getenv("MOFO_FORCE_BN"); return e ? atoi(e) : 0; }();
  static const int forced_pairs = [] { const char* e = getenv("MOFO_GEMM_2CTA"); return e ? atoi(e) : -1; }();
  const int sms = sm_count();
  TileChoice best{false, 128};
  double best_cost = 1e30;
  for (int pairs = 0; pairs < 2; ++pairs) {
    if (pairs && (M < 2 * BM || (sms & 1))) continue;
    if (forced_pairs >= 0 && pairs != (forced_pairs != 0) && !(pairs == 0 && M < 2 * BM)) continue;
    const int cands[4] = {256, 224, 192, 128};                     // 224: pairs only (3 full waves for N = 2304 / 3072 at M = 5120)
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (bn > 128 && N < bn) continue;
      if (bn == 224 && !pairs) continue;
      if (forced_bn && bn != forced_bn && !(forced_bn > N && bn == 128)) continue;
      const int rows = pairs ? 2 * BM : BM;
      const long tiles = static_cast<long>((M + rows - 1) / rows) * ((N + bn - 1) / bn);
      const long slots = pairs ? sms / 2 : sms;
      const double waves = static_cast<double>((tiles + slots - 1) / slots);
      const double t_mma = (K / 16.0) * (bn / 2.0);
      const double t_l2 = (BM + (pairs ? bn / 2 : bn)) * 2.0 * K / 43.0 * (pairs ? 1.0 : 0.85);
      const double t_epi = bn * 14.0;
      const double t = (t_mma > t_l2 ? (t_mma > t_epi ? t_mma : t_epi) : (t_l2 > t_epi ? t_l2 : t_epi)) + 700.0;
      const double cost = waves * t;
      if (cost < best_cost - 1e-9) { best_cost = cost; best = TileChoice{pairs != 0, bn}; }
    }
  }
  return best;
}

template <int EPI>
static int dispatch_bn(int bn, const CUtensorMap& tA, const CUtensorMap& tB, int M, int N, int K, const EpiParams& ep, cudaStream_t s) {
  switch (bn) {
    case 256: return launch_tn<256, EPI>(tA, tB, M, N, K, ep, s);
    case 192: return launch_tn<192, EPI>(tA, tB, M, N, K, ep, s);
    default:  return launch_tn<128, EPI>(tA, tB, M, N, K, ep, s);
  }
}

template <int EPI>
static int dispatch_bn2(int bn, const CUtensorMap& tA, const CUtensorMap& tB, int M, int N, int K, const EpiParams& ep, cudaStream_t s) {
  switch (bn) {
    case 256: return launch_tn2<256, EPI>(tA, tB, M, N, K, ep, s);
    case 224: return launch_tn2<224, EPI>(tA, tB, M, N, K, ep, s);
    case 192: return launch_tn2<192, EPI>(tA, tB, M, N, K, ep, s);
    default:  return launch_tn2<128, EPI>(tA, tB, M, N, K, ep, s);
  }
}

template <int BNW, int NT>
static int launch_wgrad(const CUtensorMap& tY, const CUtensorMap& tX, int M, int N, int K, float* dW, int ldw, float* dbias,
                        int skip_lo, int skip_hi, cudaStream_t s) {
  using Cfg = WgCfg<BNW, NT>;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute((gemm_wgrad_kernel<BNW, NT>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((N + 128 * NT - 1) / (128 * NT)) * ((K + BNW - 1) / BNW);
  const int kblocks = (M + BK - 1) / BK;
  int splits = sm_count() / tiles;
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  const int kb_per_split = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per_split - 1) / kb_per_split;
  dim3 grid(tiles, splits);
  MOFO_CUDA(launch_pdl((gemm_wgrad_kernel<BNW, NT>), grid, dim3(WG_THREADS), Cfg::SMEM_BYTES, s, tY, tX, M, N, K, dW, ldw, kb_per_split,
                       dbias, skip_lo, skip_hi));
  return MOFO_OK;
}

}  // namespace mofo

using namespace mofo;

template <int BNW>
static int launch_wgrad_grouped(int n, const mofo_bf16* const* dY, const int* ldy, const mofo_bf16* const* X, const int* ldx, int M,
                                const int* N, const int* K, float* const* dW, const int* ldw, float* const* dbias,
                                const int* dbias_skip_lo, const int* dbias_skip_hi, void* stream) {
  static WgGroup g;                      // ~1.3 KB: filled on the host, passed by value as the kernel parameter
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    MOFO_CHECK_ARG(dY[i] && X[i] && dW[i] && N[i] > 0 && K[i] > 0 && N[i] % 8 == 0 && K[i] % BNW == 0,
                   "gemm_wgrad_grouped: problem %d: N=%d K=%d (need N%%8==0, K%%%d==0)", i, N[i], K[i], BNW);
    MOFO_CHECK_ARG(ldy[i] >= N[i] && ldx[i] >= K[i] && ldy[i] % 8 == 0 && ldx[i] % 8 == 0 && ldw[i] >= K[i] && ldw[i] % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(dW[i]) & 15) == 0, "gemm_wgrad_grouped: problem %d: bad leading dimension / alignment", i);
    int rc = get_tmap(&g.p[i].tmY, dY[i], M, N[i], ldy[i], 64);
    if (rc) return rc;
    rc = get_tmap(&g.p[i].tmX, X[i], M, K[i], ldx[i], 64);
    if (rc) return rc;
    g.p[i].dW = dW[i]; g.p[i].dbias = dbias ? dbias[i] : nullptr;
    g.p[i].N = N[i]; g.p[i].K = K[i]; g.p[i].ldw = ldw[i];
    g.p[i].skip_lo = dbias_skip_lo ? dbias_skip_lo[i] : 0; g.p[i].skip_hi = dbias_skip_hi ? dbias_skip_hi[i] : 0;
    g.p[i].tile0 = tiles;
    tiles += ((N[i] + 127) / 128) * (K[i] / BNW);
  }
  g.n = n; g.M = M;
  const int kblocks = (M + BK - 1) / BK;
  int splits = (2 * sm_count()) / tiles;                    // ~2 waves of equal-sized work items
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  g.kb_per_split = (kblocks + splits - 1) / splits;
  splits = (kblocks + g.kb_per_split - 1) / g.kb_per_split;
  using Cfg = WgCfg<BNW, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(gemm_wgrad_grouped_kernel<BNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  MOFO_CUDA(launch_pdl(gemm_wgrad_grouped_kernel<BNW>, dim3(tiles, splits), dim3(WG_THREADS), Cfg::SMEM_BYTES,
                       static_cast<cudaStream_t>(stream), g));
  return MOFO_OK;
}

extern "C" {

int mofo_gemm_tn(const mofo_bf16* A, int lda, const mofo_bf16* B, int ldb, int M, int N, int K, int epilogue,
                 const float* bias, const float* resid, int ldr, const mofo_bf16* aux_bf16, int ldaux, const float* pos,
                 const int32_t* row_idx, int group_rows, int out_group_rows, void* out0, int ldo0, void* out1, int ldo1,
                 const float* row_scale, void* stream) {
  MOFO_CHECK_ARG(A && B && out0, "gemm_tn: null pointer");
  MOFO_CHECK_ARG(!row_scale || (epilogue == MOFO_EPI_BIAS_RESID_F32 && group_rows > 0), "gemm_tn: row_scale needs BIAS_RESID_F32 and group_rows");
  MOFO_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, "gemm_tn: M=%d N=%d K=%d (need K%%8==0, N%%8==0)", M, N, K);
  MOFO_CHECK_ARG(lda >= K && ldb >= K && lda % 8 == 0 && ldb % 8 == 0, "gemm_tn: bad leading dimensions lda=%d ldb=%d", lda, ldb);
  MOFO_CHECK_ARG(ldo0 % 8 == 0 && (reinterpret_cast<uintptr_t>(out0) & 15) == 0, "gemm_tn: out0 must be 16-B aligned with ld%%8==0");
  EpiParams ep{bias, resid, ldr, reinterpret_cast<const __nv_bfloat16*>(aux_bf16), ldaux, pos, row_idx,
               group_rows > 0 ? group_rows : M, out_group_rows > 0 ? out_group_rows : M, out0, ldo0, out1, ldo1, row_scale};
  switch (epilogue) {
    case MOFO_EPI_BIAS_GELU_BF16:
      MOFO_CHECK_ARG(out1 && ldo1 % 8 == 0, "gemm_tn: BIAS_GELU needs out1"); break;
    case MOFO_EPI_BIAS_RESID_F32:
      MOFO_CHECK_ARG(resid && ldr % 4 == 0, "gemm_tn: BIAS_RESID needs resid"); break;
    case MOFO_EPI_GELU_BWD_BF16:
      MOFO_CHECK_ARG(aux_bf16 && ldaux % 8 == 0, "gemm_tn: GELU_BWD needs aux_bf16"); break;
    case MOFO_EPI_BIAS_POS_F32:
      MOFO_CHECK_ARG(pos && row_idx, "gemm_tn: BIAS_POS needs pos and row_idx"); break;
    case MOFO_EPI_BIAS_BF16: case MOFO_EPI_PLAIN_BF16: break;
    default: MOFO_CHECK_ARG(false, "gemm_tn: unknown epilogue %d", epilogue);
  }
  const TileChoice tc = pick_tiles(M, N, K);
  const bool pairs = tc.pairs;
  const int bn = tc.bn;
  CUtensorMap tA, tB;
  int rc = get_tmap(&tA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = get_tmap(&tB, B, N, K, ldb, pairs ? bn / 2 : bn);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pairs) {
    switch (epilogue) {
      case MOFO_EPI_BIAS_BF16:      return dispatch_bn2<MOFO_EPI_BIAS_BF16>(bn, tA, tB, M, N, K, ep, s);
      case MOFO_EPI_BIAS_GELU_BF16: return dispatch_bn2<MOFO_EPI_BIAS_GELU_BF16>(bn, tA, tB, M, N, K, ep, s);
      case MOFO_EPI_BIAS_RESID_F32: return dispatch_bn2<MOFO_EPI_BIAS_RESID_F32>(bn, tA, tB, M, N, K, ep, s);
      case MOFO_EPI_PLAIN_BF16:     return dispatch_bn2<MOFO_EPI_PLAIN_BF16>(bn, tA, tB, M, N, K, ep, s);
      case MOFO_EPI_GELU_BWD_BF16:  return dispatch_bn2<MOFO_EPI_GELU_BWD_BF16>(bn, tA, tB, M, N, K, ep, s);
      default:                      return dispatch_bn2<MOFO_EPI_BIAS_POS_F32>(bn, tA, tB, M, N, K, ep, s);
    }
  }
  switch (epilogue) {
    case MOFO_EPI_BIAS_BF16:      return dispatch_bn<MOFO_EPI_BIAS_BF16>(bn, tA, tB, M, N, K, ep, s);
    case MOFO_EPI_BIAS_GELU_BF16: return dispatch_bn<MOFO_EPI_BIAS_GELU_BF16>(bn, tA, tB, M, N, K, ep, s);
    case MOFO_EPI_BIAS_RESID_F32: return dispatch_bn<MOFO_EPI_BIAS_RESID_F32>(bn, tA, tB, M, N, K, ep, s);
    case MOFO_EPI_PLAIN_BF16:     return dispatch_bn<MOFO_EPI_PLAIN_BF16>(bn, tA, tB, M, N, K, ep, s);
    case MOFO_EPI_GELU_BWD_BF16:  return dispatch_bn<MOFO_EPI_GELU_BWD_BF16>(bn, tA, tB, M, N, K, ep, s);
    default:                      return dispatch_bn<MOFO_EPI_BIAS_POS_F32>(bn, tA, tB, M, N, K, ep, s);
  }
}

int mofo_gemm_wgrad(const mofo_bf16* dY, int ldy, const mofo_bf16* X, int ldx, int M, int N, int K, float* dW, int ldw,
                    float* dbias, int dbias_skip_lo, int dbias_skip_hi, void* stream) {
  MOFO_CHECK_ARG(dY && X && dW, "gemm_wgrad: null pointer");
  MOFO_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "gemm_wgrad: M=%d N=%d K=%d (need N%%8==0, K%%8==0)", M, N, K);
  MOFO_CHECK_ARG(ldy >= N && ldx >= K && ldy % 8 == 0 && ldx % 8 == 0 && ldw >= K, "gemm_wgrad: bad leading dimensions");
  MOFO_CHECK_ARG(ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0, "gemm_wgrad: dW must be 16-byte aligned with ldw%%4==0");
  CUtensorMap tY, tX;
  int rc = get_tmap(&tY, dY, M, N, ldy, 64);
  if (rc) return rc;
  rc = get_tmap(&tX, X, M, K, ldx, 64);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Default: one n-tile x 256 / 192 / 128.  The multi-accumulator tiles (MOFO_WGRAD_NT=2 or 3: 2 n-tiles x 192, 3 x 128) cut
  // the operand bytes per flop by 25-30 % but measured NO faster on B200 (decoder fc1 62.2 vs 62.9 us, encoder fc1 30.4 vs
  // 27.5 us, step 13.22 vs 13.13 ms): wgrad is not limited by L2 -> SM operand bytes.  Kept opt-in for experiments.
  static const int max_nt = [] { const char* e = getenv("MOFO_WGRAD_NT"); return e ? atoi(e) : 1; }();
  if (max_nt >= 2 && N % 256 == 0 && K % 192 == 0) return launch_wgrad<192, 2>(tY, tX, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, s);
  if (max_nt >= 3 && N % 384 == 0 && K % 128 == 0) return launch_wgrad<128, 3>(tY, tX, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, s);
  if (K % 256 == 0) return launch_wgrad<256, 1>(tY, tX, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, s);
  if (K % 192 == 0) return launch_wgrad<192, 1>(tY, tX, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, s);
  return launch_wgrad<128, 1>(tY, tX, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, s);
}

int mofo_gemm_wgrad_grouped(int n, const mofo_bf16* const* dY, const int* ldy, const mofo_bf16* const* X, const int* ldx, int M,
                            const int* N, const int* K, float* const* dW, const int* ldw, float* const* dbias,
                            const int* dbias_skip_lo, const int* dbias_skip_hi, void* stream) {
  MOFO_CHECK_ARG(n >= 1 && n <= WG_MAX_GROUP && dY && X && N && K && dW && ldy && ldx && ldw && M > 0, "gemm_wgrad_grouped: bad argument (1..%d problems)", WG_MAX_GROUP);
  // one k-tile width for the whole group: 192 when every K divides by it (ViT-S / ViT-B: 384, 768, 1536, 3072), else 256
  // (ViT-L: 1024, 4096; decoder width 512)
  bool all192 = true, all256 = true;
  for (int i = 0; i < n; ++i) { all192 = all192 && K[i] > 0 && K[i] % 192 == 0; all256 = all256 && K[i] > 0 && K[i] % 256 == 0; }
  MOFO_CHECK_ARG(all192 || all256, "gemm_wgrad_grouped: every K must be a multiple of 192, or every K a multiple of 256");
  if (all192) return launch_wgrad_grouped<192>(n, dY, ldy, X, ldx, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, stream);
  return launch_wgrad_grouped<256>(n, dY, ldy, X, ldx, M, N, K, dW, ldw, dbias, dbias_skip_lo, dbias_skip_hi, stream);
}

}  // extern "C"
