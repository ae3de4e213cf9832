// Fused multi-head attention (head_dim 64) on tcgen05 tensor cores, forward and backward.
// Replaces modeling_finetune.py:85-95 (Attention.forward between the qkv and proj linears).
//
// Layout: qkv bf16 [B*S, 3*H*64] (q | k | v), out / dout bf16 [B*S, H*64], lse / delta f32 [B,H,S].
// All three kernels use 256 threads: each row of the 128-row tile is owned by two threads (32 of the 64 streamed
// columns each) that read their accumulator slice straight from TMEM, so the softmax needs one bf16 exchange per row
// and no shuffles.  One elected lane of warp 0 additionally issues the TMA loads and tcgen05.mma - from a branch
// the compiler can prove warp-uniform (shfl-broadcast warp id + elect.sync), otherwise every UTCHMMA / UTMALDG gets
// wrapped in a per-lane ELECT / R2UR / BRA.U.ANY loop that costs ~80 clk per instruction.
// Occupancy is what these kernels live on (the per-tile chain MMA -> tcgen05.ld -> softmax -> tcgen05.st -> MMA is
// ~2500 clk): forward 4 CTAs/SM (64 registers, 128 TMEM columns), dQ 3 (S and dP share one TMEM region), dK/dV 2.
//   S = Q K^T and friends:  both operands K-major SW128 tiles [128 rows x 64 d] straight from TMA.
//   P V / dS K / P^T dO ...: A = bf16 operand written by the threads straight into TENSOR MEMORY (tcgen05.st over their
//                            own fp32 accumulator columns; tcgen05.mma with A in TMEM), so it never touches shared memory;
//                            B = the same TMA tile re-read as an MN-major operand (rows = reduction index).
// Everything is kept in the log2 domain: t = s * scale * log2(e), p = exp2(t - lse2).
#include <stdlib.h>

#include "../../include/mofo_b200.h"
#include "attn_helpers.cuh"

namespace mofo {

#ifndef MOFO_ATTN_PAIRBAR
#define MOFO_ATTN_PAIRBAR 0      // 1: measured neutral (0.229 vs 0.230 ms forward at B = 32), kept as a switch
#endif
#ifndef MOFO_ATTN_PAD
#define MOFO_ATTN_PAD 0          // tuning aid: extra dynamic smem per CTA to force lower occupancy in variant builds
#endif
#ifdef MOFO_ATTN_TRACE
__device__ long long g_trace[64 * 8 * 2];
#define TRACE_AT(it, slot) do { if (trace_on && (tid == 0 || tid == 255)) g_trace[((it) * 8 + (slot)) * 2 + (tid != 0)] = clock64(); } while (0)
#define TRACE(slot) TRACE_AT(i, slot)
#else
#define TRACE_AT(it, slot) do { } while (0)
#define TRACE(slot) do { } while (0)
#endif

// =================================================================================================
// forward: CTA = 128 q rows of one (clip, head); streams 64-row K/V tiles (double buffered).
// TMEM: S [0,64) (P aliases it) | O [64,128).  smem 48.6 KB, 64 registers, 4 CTAs / SM.
// O accumulates in TMEM across the kv tiles (tcgen05.mma accumulate), so the threads never read it back inside the
// loop.  The exponent reference m_ref is updated lazily: only when the running row maximum exceeds it by more than
// 2^8 is the O row (and the row sum) rescaled in TMEM (tcgen05.ld / .st), which happens in the first tile(s) only;
// otherwise p = exp2(t - m_ref) may exceed 1 (<= 256), harmless in bf16 / fp32.  The final lse is exact.
// Per iteration there is ONE tensor-pipe round trip: thread 0 issues P(j)V(j) and S(j+1) back to back and the
// single commit of S(j+1) also covers P(j)V(j) (the tensor pipe executes in order).
// =================================================================================================
constexpr int FWD_SMEM = TILE_BYTES + 4 * HTILE_BYTES + 512 + 64;   // Q, K0,V0,K1,V1, max xchg, barriers

__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, int S, int H,
                float c /*scale*log2e*/, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo,
                float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  check_align(base);
  const uint32_t sQ = base;
  auto sK = [&](int b) { return base + TILE_BYTES + (2 * b) * HTILE_BYTES; };
  auto sV = [&](int b) { return base + TILE_BYTES + (2 * b + 1) * HTILE_BYTES; };
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(smem_raw + TILE_BYTES + 4 * HTILE_BYTES);   // [2][128]
  float* xsum = reinterpret_cast<float*>(smem_raw + TILE_BYTES);                                    // aliases K/V at the end
  const uint32_t bars = base + TILE_BYTES + 4 * HTILE_BYTES + 512;
  const uint32_t bar_q = bars, bar_s = bars + 8, bar_fin = bars + 16, tmem_slot = bars + 40;
  auto bar_kv = [&](int b) { return bars + 24 + 8 * b; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const int q0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * S;
  const int n_kv = (S + BT - 1) / BT;

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_fin, 1); mbar_init(bar_kv(0), 1); mbar_init(bar_kv(1), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // provably warp-uniform -> MMA operands stay in uniform registers
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_wait();
  pdl_trigger();
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_off + half * 32, tO = tmem_base + lane_off + 64 + half * 32;

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_q, TILE_BYTES);
    tma_load_2d(sQ, &tm_q, bar_q, h * 64, row0 + q0);
    mbar_expect_tx(bar_kv(0), 2 * HTILE_BYTES);
    tma_load_2d(sK(0), &tm_kv, bar_kv(0), (H + h) * 64, row0);
    tma_load_2d(sV(0), &tm_kv, bar_kv(0), (2 * H + h) * 64, row0);
    if (n_kv > 1) {
      mbar_expect_tx(bar_kv(1), 2 * HTILE_BYTES);
      tma_load_2d(sK(1), &tm_kv, bar_kv(1), (H + h) * 64, row0 + BT);
      tma_load_2d(sV(1), &tm_kv, bar_kv(1), (2 * H + h) * 64, row0 + BT);
    }
    mbar_wait(bar_q, 0);
    mbar_wait(bar_kv(0), 0);
    tc_fence_after();
    mma_ab_t(tmem_base, sQ, sK(0));                 // S(0)
    tc_commit(bar_s);
  }

  float m_ref = -INFINITY, l_run = 0.f;             // l_run: partial row sum over this thread's columns
#ifdef MOFO_ATTN_TRACE
  const bool trace_on = MOFO_ATTN_TRACE == 1 && blockIdx.x == 6 && blockIdx.y == 3 && blockIdx.z == 10;
#endif

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j & 1;
    TRACE_AT(j, 0);
    mbar_wait(bar_s, j & 1);                        // S(j) ready; also: P(j-1)V(j-1) done -> P tile and K/V buffer buf^1 free
    tc_fence_after();
    if (warp_u == 0 && j >= 1 && j + 1 < n_kv && elect_one()) {       // refill the buffer tile j-1 used with tile j+1
      mbar_expect_tx(bar_kv(buf ^ 1), 2 * HTILE_BYTES);
      tma_load_2d(sK(buf ^ 1), &tm_kv, bar_kv(buf ^ 1), (H + h) * 64, row0 + (j + 1) * BT);
      tma_load_2d(sV(buf ^ 1), &tm_kv, bar_kv(buf ^ 1), (2 * H + h) * 64, row0 + (j + 1) * BT);
    }
    const int kv_valid = S - j * BT - half * 32;    // own columns >= kv_valid are padding (last tile only)
    // 64 registers -> 4 CTAs / SM: the score row is read from TMEM twice - once for the row maximum, then again in
    // two 16-column halves for exp2 / pack / store - instead of being held in 32 registers (TMEM reads are cheap,
    // a fourth resident CTA is worth 6 % at S = 1568).
    float mx = -INFINITY;
    {
      uint32_t r[32];
      TRACE_AT(j, 1);
      tmem_ld32(tS, r);
      tc_wait_ld();
      TRACE_AT(j, 2);
      if (kv_valid >= 32) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
#pragma unroll
          for (int k = 0; k < 4; ++k) m4[k] = fmax3(m4[k], __uint_as_float(r[e + 2 * k]), __uint_as_float(r[e + 2 * k + 1]));
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) if (e < kv_valid) mx = fmaxf(mx, __uint_as_float(r[e]));
      }
    }
    xch[half * 128 + row] = __float2bfloat16_ru(mx);
    TRACE_AT(j, 3);
#if MOFO_ATTN_PAIRBAR
    // the row maximum is exchanged between the two warps that share these 32 rows only (named barrier, 64 threads): nothing
    // else in the CTA depends on this point, and a CTA-wide barrier made every warp wait for the slowest pair (ncu, round 2:
    // 7.7 % of the kernel's stall samples sat on the instruction behind it).  xch is rewritten only after the CTA-wide
    // barrier below, so the partner's read of this iteration is always complete.
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
#else
    __syncthreads();
#endif
    TRACE_AT(j, 4);
    mx = fmaxf(__bfloat162float(xch[row]), __bfloat162float(xch[128 + row])) * c;
    const bool bump = mx > m_ref + 8.0f;
    if (__any_sync(0xffffffffu, bump) && j > 0) {
      const float alpha = bump ? exp2f(m_ref - mx) : 1.0f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[16];
        tmem_ld16a(tO + hh * 16, o);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st16_async(tO + hh * 16, o);
      }
      tc_wait_st();
      l_run *= alpha;
    }
    if (bump) m_ref = mx;
    float rs = 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t r[16];
      tmem_ld16a(tS + hh * 16, r);
      tc_wait_ld();
      uint32_t pk[8];
      float s2[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        float p0 = exp2f(__uint_as_float(r[e]) * c - m_ref), p1 = exp2f(__uint_as_float(r[e + 1]) * c - m_ref);
        if (kv_valid < 32) {
          if (hh * 16 + e >= kv_valid) p0 = 0.f;
          if (hh * 16 + e + 1 >= kv_valid) p1 = 0.f;
        }
        s2[0] += p0; s2[1] += p1;
        pk[e >> 1] = pack_bf16(p0, p1);
      }
      rs += s2[0] + s2[1];
      tmem_st8_async(tS + hh * 8, pk);          // P (bf16) over this thread's own, already consumed, S columns
    }
    tc_wait_st();
    l_run += rs;
    tc_fence_before();
    TRACE_AT(j, 5);
    __syncthreads();
    TRACE_AT(j, 6);
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      mma_ptmem_t(tmem_base + 64, tmem_base, sV(buf), j != 0);   // O += P(j) V(j), P read from TMEM
      if (j + 1 < n_kv) {
        mbar_wait(bar_kv(buf ^ 1), ((j + 1) >> 1) & 1);
        tc_fence_after();
        mma_ab_t(tmem_base, sQ, sK(buf ^ 1));                   // S(j+1)
        tc_commit(bar_s);
      } else {
        tc_commit(bar_fin);
      }
    }
    TRACE_AT(j, 7);
  }

  mbar_wait(bar_fin, 0);                // every MMA has completed: O is final, K/V smem is free
  tc_fence_after();
  xsum[half * 128 + row] = l_run;
  __syncthreads();
  const float l_tot = xsum[row] + xsum[128 + row];
  uint32_t o[32];
  tmem_ld32(tO, o);
  tc_wait_ld();
  const int q = q0 + row;
  if (q < S) {
    const float inv = 1.0f / l_tot;
    const size_t off = (static_cast<size_t>(row0 + q) * H + h) * 64 + half * 32;
    __nv_bfloat16* dst = out + off;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * inv);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 v;
      v.x = pack_bf16(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
      v.y = pack_bf16(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
      v.z = pack_bf16(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
      v.w = pack_bf16(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
      reinterpret_cast<uint4*>(dst)[g] = v;
      if (out_lo) {
        // second bf16 word of O (o - bf16(o)): backward's delta = dO . O is then accurate to 2^-17 instead of 2^-9; an
        // error e_i in delta leaks e_i * (P-weighted mean key) into dQ (see attention_small.cu), which is what made
        // the decoder's q_bias gradients (batch-wide sums of cancelling terms) 20-70x noisier than the reference's
        uint4 w;
        w.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) - bf16_lo(v.x), __uint_as_float(o[g * 8 + 1]) - bf16_hi(v.x));
        w.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) - bf16_lo(v.y), __uint_as_float(o[g * 8 + 3]) - bf16_hi(v.y));
        w.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) - bf16_lo(v.z), __uint_as_float(o[g * 8 + 5]) - bf16_hi(v.z));
        w.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) - bf16_lo(v.w), __uint_as_float(o[g * 8 + 7]) - bf16_hi(v.w));
        reinterpret_cast<uint4*>(out_lo + off)[g] = w;
      }
    }
    if (half == 0) lse[(static_cast<size_t>(b) * H + h) * S + q] = m_ref + log2f(l_tot);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// =================================================================================================
// backward, part 0: delta[b,h,q] = sum_d dO * O
// =================================================================================================
// 8 lanes share one (row, head): lane j reads the j-th 16-byte chunk of the 128-byte O and dO rows, so every warp
// load covers 4 full rows (fully used sectors); 4 rows per thread in flight; 3 shuffles finish the dot product.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out,
                                                         const __nv_bfloat16* __restrict__ out_lo,
                                                         const __nv_bfloat16* __restrict__ dout, int rows, int S, int H,
                                                         float* __restrict__ delta) {
  pdl_wait();
  pdl_trigger();
  const int total = rows * H;                                       // (row, head) pairs, head fastest
  const int j = threadIdx.x & 7;
  const int g0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // first pair of this 8-lane group
  const int stride = (gridDim.x * blockDim.x) >> 3;
  for (int base = g0; base < total; base += 4 * stride) {
    uint4 a[4], bq[4], al[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * stride;
      al[u] = make_uint4(0, 0, 0, 0);
      if (i < total) {
        a[u] = __ldg(reinterpret_cast<const uint4*>(out + static_cast<size_t>(i) * 64) + j);
        bq[u] = __ldg(reinterpret_cast<const uint4*>(dout + static_cast<size_t>(i) * 64) + j);
        if (out_lo) al[u] = __ldg(reinterpret_cast<const uint4*>(out_lo + static_cast<size_t>(i) * 64) + j);
      } else {
        a[u] = make_uint4(0, 0, 0, 0); bq[u] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float s = (bf16_lo(a[u].x) + bf16_lo(al[u].x)) * bf16_lo(bq[u].x) + (bf16_hi(a[u].x) + bf16_hi(al[u].x)) * bf16_hi(bq[u].x) +
                (bf16_lo(a[u].y) + bf16_lo(al[u].y)) * bf16_lo(bq[u].y) + (bf16_hi(a[u].y) + bf16_hi(al[u].y)) * bf16_hi(bq[u].y) +
                (bf16_lo(a[u].z) + bf16_lo(al[u].z)) * bf16_lo(bq[u].z) + (bf16_hi(a[u].z) + bf16_hi(al[u].z)) * bf16_hi(bq[u].z) +
                (bf16_lo(a[u].w) + bf16_lo(al[u].w)) * bf16_lo(bq[u].w) + (bf16_hi(a[u].w) + bf16_hi(al[u].w)) * bf16_hi(bq[u].w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int i = base + u * stride;
      if (j == 0 && i < total) {
        const int row = i / H, h = i % H;
        const int b = row / S, q = row % S;
        delta[(static_cast<size_t>(b) * H + h) * S + q] = s;
      }
    }
  }
}

// =================================================================================================
// backward, part 1: dQ.  CTA = 128 q rows of one (clip, head); streams 64-row K/V tiles (2-stage TMA ring).
// TMEM: [0,64) holds, in turn, S = Q K^T, then dP = dO V^T, then dS (bf16) | dQ [64,128) -> 128 columns.
// S and dP are TIME-MULTIPLEXED through one 64-column region (the threads turn S into P, kept in registers, before dP
// is issued): a second tensor-pipe round trip per tile, but the CTA needs 128 instead of 256 TMEM columns, 64 KB of
// shared memory and <= 85 registers, so 3 CTAs fit per SM instead of 2 - the per-tile dependency chain, not the
// tensor pipe, is what bounds these kernels.
// Rows q >= S and kv >= S: garbage rows only pollute their own (never stored) output rows, so only the
// reduction (column) index is masked, and only in the last tile.
// =================================================================================================
constexpr int DQ_NST = 2;                                                       // K/V ring depth
constexpr int DQ_SMEM = 2 * TILE_BYTES + DQ_NST * 2 * HTILE_BYTES + 128;   // Q, dO, K/V ring, barriers

__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                   const __grid_constant__ CUtensorMap tm_do, int S, int H, float c, float scale,
                   const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  check_align(base);
  const uint32_t sQ = base, sdO = base + TILE_BYTES;
  auto sK = [&](int b) { return base + 2 * TILE_BYTES + (2 * b) * HTILE_BYTES; };
  auto sV = [&](int b) { return base + 2 * TILE_BYTES + (2 * b + 1) * HTILE_BYTES; };
  const uint32_t bars = base + 2 * TILE_BYTES + DQ_NST * 2 * HTILE_BYTES;
  const uint32_t bar_q = bars, bar_s = bars + 8, bar_dp = bars + 16, bar_fin = bars + 24, tmem_slot = bars + 56;
  auto bar_kv = [&](int b) { return bars + 32 + 8 * b; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const int q0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * S;
  const int n_kv = (S + BT - 1) / BT;
  auto load_kv = [&](int t) {      // elected lane: tile t -> ring stage t % DQ_NST
    const int st = t % DQ_NST;
    mbar_expect_tx(bar_kv(st), 2 * HTILE_BYTES);
    tma_load_2d(sK(st), &tm_kv, bar_kv(st), (H + h) * 64, row0 + t * BT);
    tma_load_2d(sV(st), &tm_kv, bar_kv(st), (2 * H + h) * 64, row0 + t * BT);
  };

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_fin, 1);
    for (int st = 0; st < DQ_NST; ++st) mbar_init(bar_kv(st), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_do);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // provably warp-uniform -> MMA operands stay in uniform registers
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_wait();
  pdl_trigger();
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_off + half * 32;            // S, then dP, then dS (bf16) in this thread's columns
  const uint32_t tdQ = tmem_base + lane_off + 64 + half * 32;

  const int q = q0 + row;
  const bool q_ok = q < S;
  const float my_lse = q_ok ? lse[(static_cast<size_t>(b) * H + h) * S + q] : 0.f;
  const float my_delta = q_ok ? delta[(static_cast<size_t>(b) * H + h) * S + q] : 0.f;

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_q, 2 * TILE_BYTES);
    tma_load_2d(sQ, &tm_q, bar_q, h * 64, row0 + q0);
    tma_load_2d(sdO, &tm_do, bar_q, h * 64, row0 + q0);
    for (int t = 0; t < DQ_NST && t < n_kv; ++t) load_kv(t);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_kv(0), 0);
    tc_fence_after();
    mma_ab_t(tmem_base, sQ, sK(0));            // S(0) = Q K^T
    tc_commit(bar_s);
  }

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j % DQ_NST, nbuf = (j + 1) % DQ_NST;
    mbar_wait(bar_s, j & 1);    // S(j) ready; also covers the dQ MMA of iteration j-1 -> stage (j-1) % NST is free
    tc_fence_after();
    if (warp_u == 0 && j >= 1 && j + 1 < n_kv && elect_one()) load_kv(j + 1);
    const int kv_valid = S - j * BT - half * 32;
    float p[32];
    {
      uint32_t rs[32];
      tmem_ld32(tS, rs);
      tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) p[e] = exp2f(__uint_as_float(rs[e]) * c - my_lse);
      if (kv_valid < 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) p[e] = (e < kv_valid) ? p[e] : 0.f;
      }
    }
    tc_fence_before();
    __syncthreads();            // every thread has consumed S(j): the region is free for dP(j)
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      mma_ab_t(tmem_base, sdO, sV(buf));       // dP(j) = dO V^T over S(j)
      tc_commit(bar_dp);
    }
    mbar_wait(bar_dp, j & 1);
    tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rp[16], pk[8];
      tmem_ld16a(tS + hh * 16, rp);
      tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; e += 2)
        pk[e >> 1] = pack_bf16(p[hh * 16 + e] * (__uint_as_float(rp[e]) - my_delta),
                               p[hh * 16 + e + 1] * (__uint_as_float(rp[e + 1]) - my_delta));
      tmem_st8_async(tS + hh * 8, pk);         // dS (bf16) over this thread's own, already consumed, dP columns
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      mma_ptmem_t(tmem_base + 64, tmem_base, sK(buf), j != 0);  // dQ += dS K, dS read from TMEM
      if (j + 1 < n_kv) {
        mbar_wait(bar_kv(nbuf), ((j + 1) / DQ_NST) & 1);
        tc_fence_after();
        mma_ab_t(tmem_base, sQ, sK(nbuf));     // S(j+1) over dS(j), which the dQ MMA has consumed (in-order pipe)
        tc_commit(bar_s);
      } else {
        tc_commit(bar_fin);
      }
    }
  }
  mbar_wait(bar_fin, 0);
  tc_fence_after();
  {
    __nv_bfloat16* dst = dqkv + static_cast<size_t>(row0 + q) * (3 * H * 64) + h * 64 + half * 32;
    uint32_t r[32];
    tmem_ld32(tdQ, r);
    tc_wait_ld();
    if (q_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * scale, __uint_as_float(r[g * 8 + 1]) * scale);
        o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * scale, __uint_as_float(r[g * 8 + 3]) * scale);
        o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * scale, __uint_as_float(r[g * 8 + 5]) * scale);
        o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * scale, __uint_as_float(r[g * 8 + 7]) * scale);
        reinterpret_cast<uint4*>(dst)[g] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// =================================================================================================
// backward, part 2: dK, dV.  CTA = 128 kv rows of one (clip, head); streams 64-row Q/dO tiles.
// TMEM: S^T [0,64) | dP^T [64,128) | dV [128,192) | dK [192,256) -> 2 CTAs / SM.  smem 96.6 KB (4-stage Q/dO ring); P^T / dS^T alias S^T / dP^T in TMEM.
// =================================================================================================
constexpr int DKV_NST = 4;                                                            // Q/dO ring depth (prefetch distance 3)
constexpr int DKV_SMEM = 2 * TILE_BYTES + DKV_NST * 2 * HTILE_BYTES + 1024 + 128;   // K,V, Q/dO ring, stats x2, barriers

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                    const __grid_constant__ CUtensorMap tm_do, int S, int H, float c, float scale,
                    const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  check_align(base);
  const uint32_t sK = base, sV = base + TILE_BYTES;
  auto sQ = [&](int b) { return base + 2 * TILE_BYTES + (2 * b) * HTILE_BYTES; };
  auto sdO = [&](int b) { return base + 2 * TILE_BYTES + (2 * b + 1) * HTILE_BYTES; };
  float* vec = reinterpret_cast<float*>(smem_raw + 2 * TILE_BYTES + DKV_NST * 2 * HTILE_BYTES);   // 2 x [lse 64 | delta 64]
  const uint32_t bars = base + 2 * TILE_BYTES + DKV_NST * 2 * HTILE_BYTES + 1024;
  const uint32_t bar_kv = bars, bar_s = bars + 8, bar_fin = bars + 16, tmem_slot = bars + 56, bar_dp = bars + 64;
  auto bar_q = [&](int b) { return bars + 24 + 8 * b; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const int kv0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * S;
  const int n_q = (S + BT - 1) / BT;
  auto load_q = [&](int t) {       // thread 0: q tile t -> ring stage t % DKV_NST
    const int st = t % DKV_NST;
    mbar_expect_tx(bar_q(st), 2 * HTILE_BYTES);
    tma_load_2d(sQ(st), &tm_q, bar_q(st), h * 64, row0 + t * BT);
    tma_load_2d(sdO(st), &tm_do, bar_q(st), h * 64, row0 + t * BT);
  };

  if (tid == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_fin, 1);
    for (int st = 0; st < DKV_NST; ++st) mbar_init(bar_q(st), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_do);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);      // provably warp-uniform -> MMA operands stay in uniform registers
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_wait();
  pdl_trigger();
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_off + half * 32, tdP = tmem_base + lane_off + 64 + half * 32;

  const int kv = kv0 + row;
  const bool kv_ok = kv < S;
#ifdef MOFO_ATTN_TRACE
  const bool trace_on = MOFO_ATTN_TRACE == 3 && blockIdx.x == 6 && blockIdx.y == 3 && blockIdx.z == 10;
#endif

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_kv, 2 * TILE_BYTES);
    tma_load_2d(sK, &tm_kv, bar_kv, (H + h) * 64, row0 + kv0);
    tma_load_2d(sV, &tm_kv, bar_kv, (2 * H + h) * 64, row0 + kv0);
    for (int t = 0; t < DKV_NST - 1 && t < n_q; ++t) load_q(t);
    mbar_wait(bar_kv, 0);
    mbar_wait(bar_q(0), 0);
    tc_fence_after();
    mma_ab_t(tmem_base, sK, sQ(0));             // S^T  = K Q^T
    tc_commit(bar_s);
    mma_ab_t(tmem_base + 64, sV, sdO(0));       // dP^T = V dO^T
    tc_commit(bar_dp);
  }

  // per-column (q) statistics: threads 0-63 fetch lse, 64-127 fetch delta, two tiles AHEAD into a register and one
  // tile ahead into the double-buffered vec[], so neither the global-load latency nor a block barrier sits between
  // the S^T hand-off and the softmax recompute.  vec[(i+1)&1] is written during iteration i (before its block
  // barrier) and read during iteration i+1 (after it); it was last read during iteration i-1.
  const float* stat_src = (tid < 64 ? lse : delta) + (static_cast<size_t>(b) * H + h) * S;
  float stat_next = 0.f;
  if (tid < 128) {
    vec[tid] = (tid & 63) < S ? stat_src[tid & 63] : 0.f;
    stat_next = BT + (tid & 63) < S ? stat_src[BT + (tid & 63)] : 0.f;
  }
  __syncthreads();
  for (int i = 0; i < n_q; ++i) {
    const int buf = i % DKV_NST, nbuf = (i + 1) % DKV_NST;
    if (tid < 128) {
      vec[((i + 1) & 1) * 128 + tid] = stat_next;
      const int qq = (i + 2) * BT + (tid & 63);
      stat_next = qq < S ? stat_src[qq] : 0.f;
    }
    TRACE(0);
    mbar_wait(bar_s, i & 1);     // S^T(i) ready (it was issued between the dV and dK MMAs of iteration i-1)
    tc_fence_after();
    TRACE(1);
    const int q_valid = S - i * BT - half * 32;
    const float* lse_s = vec + (i & 1) * 128 + half * 32;
    const float* del_s = lse_s + 64;
    uint32_t rs[32], rp[32];
    tmem_ld32(tS, rs);
    tc_wait_ld();
    TRACE(2);
    mbar_wait(bar_dp, i & 1);    // dP^T(i) ready; also covers dV/dK of iteration i-1 -> ring stage (i-1)%NST is free
    tc_fence_after();
    if (warp_u == 0 && i + DKV_NST - 1 < n_q && elect_one()) load_q(i + DKV_NST - 1);
    tmem_ld32(tdP, rp);
    TRACE(3);
    float p[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 l = reinterpret_cast<const float4*>(lse_s)[g];
      p[4 * g + 0] = exp2f(__uint_as_float(rs[4 * g + 0]) * c - l.x);
      p[4 * g + 1] = exp2f(__uint_as_float(rs[4 * g + 1]) * c - l.y);
      p[4 * g + 2] = exp2f(__uint_as_float(rs[4 * g + 2]) * c - l.z);
      p[4 * g + 3] = exp2f(__uint_as_float(rs[4 * g + 3]) * c - l.w);
    }
    if (q_valid < 32) {                                        // ragged last q tile: padding columns contribute nothing
#pragma unroll
      for (int e = 0; e < 32; ++e) p[e] = (e < q_valid) ? p[e] : 0.f;
    }
    tmem_store_bf16_row_async(tS, p);                          // P^T over this thread's own S^T columns
    TRACE(4);
    tc_wait_ld();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 d = reinterpret_cast<const float4*>(del_s)[g];
      p[4 * g + 0] *= __uint_as_float(rp[4 * g + 0]) - d.x;
      p[4 * g + 1] *= __uint_as_float(rp[4 * g + 1]) - d.y;
      p[4 * g + 2] *= __uint_as_float(rp[4 * g + 2]) - d.z;
      p[4 * g + 3] *= __uint_as_float(rp[4 * g + 3]) - d.w;
    }
    tmem_store_bf16_row_async(tdP, p);                         // dS^T over this thread's own dP^T columns
    tc_wait_st();
    tc_fence_before();
    TRACE(5);
    __syncthreads();
    TRACE(6);
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      // order on the (in-order) tensor pipe: dV(i), S^T(i+1), dK(i), dP^T(i+1) - the next score tile, which starts
      // the next iteration's dependency chain, does not queue behind dK
      mma_ptmem_t(tmem_base + 128, tmem_base, sdO(buf), i != 0);       // dV += P^T dO   (A from TMEM)
      if (i + 1 < n_q) {
        mbar_wait(bar_q(nbuf), ((i + 1) / DKV_NST) & 1);
        tc_fence_after();
        mma_ab_t(tmem_base, sK, sQ(nbuf));                             // S^T(i+1) over P^T(i), which dV has consumed
        tc_commit(bar_s);
        mma_ptmem_t(tmem_base + 192, tmem_base + 64, sQ(buf), i != 0); // dK += dS^T Q   (A from TMEM)
        mma_ab_t(tmem_base + 64, sV, sdO(nbuf));                       // dP^T(i+1) over dS^T(i)
        tc_commit(bar_dp);
      } else {
        mma_ptmem_t(tmem_base + 192, tmem_base + 64, sQ(buf), i != 0);
        tc_commit(bar_fin);
      }
    }
    TRACE(7);
  }
  mbar_wait(bar_fin, 0);
  tc_fence_after();
  {   // half 0 writes dV, half 1 writes dK (64 columns each)
    __nv_bfloat16* dst = dqkv + static_cast<size_t>(row0 + kv) * (3 * H * 64) + ((half == 0 ? 2 * H : H) + h) * 64;
    const float sc = half == 0 ? 1.0f : scale;
    const uint32_t tsrc = tmem_base + lane_off + (half == 0 ? 128 : 192);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tsrc + c0, r);
      tc_wait_ld();
      if (kv_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * sc, __uint_as_float(r[g * 8 + 1]) * sc);
          o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * sc, __uint_as_float(r[g * 8 + 3]) * sc);
          o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * sc, __uint_as_float(r[g * 8 + 5]) * sc);
          o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * sc, __uint_as_float(r[g * 8 + 7]) * sc);
          reinterpret_cast<uint4*>(dst + c0)[g] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace mofo

using namespace mofo;

// MOFO_ATTN_SMALL=0 routes short sequences through the streaming kernels too (A/B measurements, tests of both paths)
static bool use_small_attention() {
  const char* e = getenv("MOFO_ATTN_SMALL");
  return !(e && e[0] == '0');
}

extern "C" {

#ifdef MOFO_ATTN_TRACE
int mofo_debug_read_trace(long long* host, int n) {   // tuning aid, only in -DMOFO_ATTN_TRACE builds
  return cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif

int mofo_attn_fwd(const mofo_bf16* qkv, int B, int S, int H, float scale, mofo_bf16* out, mofo_bf16* out_lo, float* lse,
                  void* stream) {
  MOFO_CHECK_ARG(qkv && out && lse, "attn_fwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "attn_fwd: bad shape B=%d S=%d H=%d", B, S, H);
  if (S <= ATTN_SMALL_MAX_S && use_small_attention())            // whole score row in TMEM: single-pass kernels
    return attn_small_fwd(qkv, B, S, H, scale, out, lse, static_cast<cudaStream_t>(stream));
  CUtensorMap tq, tkv;
  int rc = get_tmap(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * H * 64, 3ull * H * 64, AT);
  if (rc) return rc;
  rc = get_tmap(&tkv, qkv, static_cast<uint64_t>(B) * S, 3ull * H * 64, 3ull * H * 64, BT);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM + MOFO_ATTN_PAD));
    attr_set = true;
  }
  dim3 grid((S + AT - 1) / AT, H, B);
  MOFO_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(ATT_THREADS), FWD_SMEM + MOFO_ATTN_PAD, static_cast<cudaStream_t>(stream), tq, tkv, S, H,
                       scale * 1.4426950408889634f, reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<__nv_bfloat16*>(out_lo), lse));
  return MOFO_OK;
}

int mofo_attn_bwd(const mofo_bf16* qkv, const mofo_bf16* out, const mofo_bf16* out_lo, const mofo_bf16* dout, const float* lse,
                  int B, int S, int H, float scale, mofo_bf16* dqkv, float* delta, void* stream) {
  MOFO_CHECK_ARG(qkv && out && dout && lse && dqkv && delta, "attn_bwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "attn_bwd: bad shape B=%d S=%d H=%d", B, S, H);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (S <= ATTN_SMALL_MAX_S && use_small_attention())
    return attn_small_bwd(qkv, out, dout, lse, B, S, H, scale, dqkv, s);
  const uint64_t rows = static_cast<uint64_t>(B) * S;
  CUtensorMap tq128, tq64, td128, td64;
  int rc = get_tmap(&tq128, qkv, rows, 3ull * H * 64, 3ull * H * 64, AT);
  if (rc) return rc;
  rc = get_tmap(&tq64, qkv, rows, 3ull * H * 64, 3ull * H * 64, BT);
  if (rc) return rc;
  rc = get_tmap(&td128, dout, rows, 1ull * H * 64, 1ull * H * 64, AT);
  if (rc) return rc;
  rc = get_tmap(&td64, dout, rows, 1ull * H * 64, 1ull * H * 64, BT);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM + MOFO_ATTN_PAD));
    MOFO_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM + MOFO_ATTN_PAD));
    attr_set = true;
  }
  {
    const long pairs = static_cast<long>(rows) * H;                 // 8 lanes per pair, 4 pairs per thread
    long blocks = (pairs * 8 / 4 + 255) / 256;
    const long cap = 8L * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    MOFO_CUDA(launch_pdl(attn_delta_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s,
                         reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(out_lo),
                         reinterpret_cast<const __nv_bfloat16*>(dout), static_cast<int>(rows), S, H, delta));
  }
  dim3 grid((S + AT - 1) / AT, H, B);
  const float c = scale * 1.4426950408889634f;
  MOFO_CUDA(launch_pdl(attn_bwd_dq_kernel, grid, dim3(ATT_THREADS), DQ_SMEM + MOFO_ATTN_PAD, s, tq128, tq64, td128, S, H, c, scale, lse,
                       static_cast<const float*>(delta), reinterpret_cast<__nv_bfloat16*>(dqkv)));
  MOFO_CUDA(launch_pdl(attn_bwd_dkv_kernel, grid, dim3(ATT_THREADS), DKV_SMEM + MOFO_ATTN_PAD, s, tq128, tq64, td64, S, H, c, scale, lse,
                       static_cast<const float*>(delta), reinterpret_cast<__nv_bfloat16*>(dqkv)));
  return MOFO_OK;
}

}  // extern "C"
