// Single-pass attention for short sequences (S <= 192, head_dim 64): the encoder of the MOFO pretraining step attends
// over the 160 visible tokens of a clip (modeling_pretrain.py:90 -> modeling_finetune.py:85-95), 12 layers x (clip, head)
// units of 160 x 160 scores.  The streaming kernels of attention.cu spend most of such a launch on per-CTA prologue and
// on a 128-row tile that is 3/4 empty.  Here ONE CTA owns a whole (clip, head):
//   * all operands (Q, K, V and, in backward, dO) are fetched once by TMA as 128-row SW128 tiles;
//   * the complete score row lives in tensor memory: S = Q K^T is one tcgen05.mma group with N = the padded kv count
//     (multiple of 32), so there is no online softmax, no running maximum, no rescaling of O;
//   * P / dS are written back to TMEM as bf16 over the thread's own fp32 columns and consumed as the A operand of the
//     following MMA (tcgen05.mma with A in TMEM), as in attention.cu;
//   * backward is ONE kernel: q-major items (S, dP -> dS -> dQ) then kv-major items (S^T, dP^T -> P^T, dS^T -> dV, dK)
//     over the same shared-memory operands; delta_i = sum_j P_ij dP_ij comes out of the q-major items themselves (no
//     separate kernel, and consistent with the recomputed P to fp32 rounding).
// 256 threads; thread (row, half) owns N/2 score columns of its row.  Rows of the second 128-row tile beyond S are never
// computed (their warps skip the softmax work) and never stored.
#include "../../include/mofo_b200.h"
#include "attn_helpers.cuh"

namespace mofo {

constexpr int SM_THREADS = 256;

#ifdef MOFO_ATTN_TRACE            // tuning aid (variant builds): clock64 stamps of one mid-grid CTA, thread 0 and thread 255
__device__ long long g_strace[32 * 2];
#define STRACE(slot) do { if (strace_on && (tid == 0 || tid == 255)) g_strace[(slot) * 2 + (tid != 0)] = clock64(); } while (0)
#else
#define STRACE(slot) do { } while (0)
#endif

// D[128 x N] = A[128 x 64] * B[N x 64]^T, both K-major SW128 tiles (B: N rows, contiguous)
template <int N>
__device__ __forceinline__ void mma_abt_n(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
  const uint64_t a0 = umma_desc_kmajor(a_tile), b0 = umma_desc_kmajor(b_tile);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, k != 0);
}
// D[128 x 64] = P[128 x N bf16 in TMEM] * T[N x 64] (T: MN-major SW128 tile, rows = reduction index).
// P layout: the two threads of a row own fp32 columns [0, N/2) and [N/2, N) of the region and write their N/2 bf16
// values (N/4 packed columns) at the START of their own range, so K-step k (16 elements) sits at column
// (k / (N/32)) * (N/2) + (k % (N/32)) * 8.
template <int N>
__device__ __forceinline__ void mma_pt_n(uint32_t d_tmem, uint32_t p_tmem, uint32_t t_tile) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
  constexpr int KH = N / 32;                     // K-steps per half
  const uint64_t b0 = umma_desc_mnmajor(t_tile, 8192);
#pragma unroll
  for (int k = 0; k < N / 16; ++k)
    umma_bf16_ts(d_tmem, p_tmem + (k / KH) * (N / 2) + (k % KH) * 8, b0 + 128 * k, idesc, k != 0 ? 1u : 0u);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int N> struct SmallCfg {
  static constexpr int NH = N / 2;                                   // columns per thread
  static constexpr int TILE2 = 2 * TILE_BYTES;                       // 256 rows x 128 B
  static constexpr int FWD_SMEM = 3 * TILE2 + 2 * 2 * 128 * 4 + 128;       // Q, K, V | max / sum exchange | barriers
  static constexpr int BWD_SMEM = 4 * TILE2 + 3 * 256 * 4 + 128;           // Q, K, V, dO | lse, delta, partials | barriers
};

// =================================================================================================
// forward.  TMEM (256 columns): S [0, N) with P over it | O(0) [192, 256) | O(1) [128, 192) - the second q tile's output
// lands in columns of the score region that are free once P(1) is written (P occupies [0, N/4) and [N/2, 3N/4)), so the
// first tile's epilogue (TMEM -> global) runs behind the second tile's softmax -> P V chain instead of in front of it.
// 2 CTAs / SM.
// =================================================================================================
template <int N>
__global__ void __launch_bounds__(SM_THREADS, 2)
attn_small_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, int S, int H, float c /*scale*log2e*/,
                      __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  using Cfg = SmallCfg<N>;
  constexpr int NH = Cfg::NH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  check_align(base);
  const uint32_t sQ = base, sK = base + Cfg::TILE2, sV = base + 2 * Cfg::TILE2;
  float* xmax = reinterpret_cast<float*>(smem_raw + 3 * Cfg::TILE2);          // [2][128]
  float* xsum = xmax + 256;                                                  // [2][128]
  const uint32_t bars = base + 3 * Cfg::TILE2 + 2048;
  const uint32_t bar_ld = bars, bar_s = bars + 8, tmem_slot = bars + 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
  const int h = blockIdx.x, b = blockIdx.y;
  const int row0 = b * S;
  const int n_qt = S > 128 ? 2 : 1;
#ifdef MOFO_ATTN_TRACE
  const bool strace_on = MOFO_ATTN_TRACE == 4 && blockIdx.x == 3 && blockIdx.y == 10;
#endif
  STRACE(0);

  if (tid == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_s, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_qkv);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_wait();
  pdl_trigger();
  const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t tS = tmem_base + lane_off + half * NH;          // this thread's score columns
  constexpr bool SPLIT_O = 3 * N / 4 <= 128;     // O(1) columns [128, 192) clear of the packed P of the second tile (N <= 160)
  constexpr int O1_SHIFT = SPLIT_O ? 64 : 0;     // otherwise O(1) reuses O(0)'s columns and the epilogues stay in order
  const uint32_t tO = tmem_base + lane_off + 192 + half * 32;    // this thread's 32 output columns of O(0); O(1) is 64 lower
  STRACE(1);

  if (warp_u == 0 && elect_one()) {
    const int nt = N > 128 ? 2 : 1;                              // 128-row boxes per K / V tile
    mbar_expect_tx(bar_ld, (n_qt + 2 * nt) * TILE_BYTES);
    for (int t = 0; t < n_qt; ++t) tma_load_2d(sQ + t * TILE_BYTES, &tm_qkv, bar_ld, h * 64, row0 + t * 128);
    for (int t = 0; t < nt; ++t) {
      tma_load_2d(sK + t * TILE_BYTES, &tm_qkv, bar_ld, (H + h) * 64, row0 + t * 128);
      tma_load_2d(sV + t * TILE_BYTES, &tm_qkv, bar_ld, (2 * H + h) * 64, row0 + t * 128);
    }
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    STRACE(2);
    mma_abt_n<N>(tmem_base, sQ, sK);                             // S(0)
    tc_commit(bar_s);
  }

  auto epilogue = [&](int t, float m_row, float l_tot) {      // O(t) is final in TMEM
    const int q = t * 128 + row;
    uint32_t o[32];
    tmem_ld32(tO - O1_SHIFT * t, o);
    tc_wait_ld();
    if (q < S) {
      const float inv = 1.0f / l_tot;
      __nv_bfloat16* dst = out + (static_cast<size_t>(row0 + q) * H + h) * 64 + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
        v.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
        v.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
        v.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[g] = v;
      }
      if (half == 0) lse[(static_cast<size_t>(b) * H + h) * S + q] = m_row + log2f(l_tot);
    }
  };

  float m_prev = 0.f, l_prev = 1.f, m_last = 0.f;
  for (int t = 0; t < n_qt; ++t) {
    mbar_wait(bar_s, t & 1);                     // S(t) ready; for t = 1 also: O(0) = P(0) V complete
    tc_fence_after();
    STRACE(3 + 4 * t);
    if (t == 1) {
      l_prev = xsum[row] + xsum[128 + row];                       // tile 0's row sum, before xsum is reused
      if (!SPLIT_O) epilogue(0, m_prev, l_prev);
    }
    STRACE(4 + 4 * t);
    const bool active = t * 128 + quad * 32 < S;                  // warp-uniform: this warp's rows hold real queries
    float m = 0.f;
    if (active) {
      // the thread's whole score slice (NH columns) is held in registers: one TMEM round trip, and the max / exp2 /
      // sum / pack chains below are NH-way independent, so latency is hidden by instruction-level parallelism (only two
      // CTAs fit per SM - tensor memory - so there are few warps to hide it otherwise)
      const int col0 = half * NH;
      uint32_t r[NH];
#pragma unroll
      for (int ch = 0; ch < NH / 16; ++ch) tmem_ld16a(tS + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(&r[ch * 16]));
      tc_wait_ld();
      const bool ragged = col0 + NH > S;                          // this slice holds padding columns (kv >= S)
      if (ragged) {
#pragma unroll
        for (int e = 0; e < NH; ++e) if (col0 + e >= S) r[e] = 0xff800000u;   // -inf
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < NH; e += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = fmax3(m4[k], __uint_as_float(r[e + 2 * k]), __uint_as_float(r[e + 2 * k + 1]));
      }
      xmax[half * 128 + row] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      named_bar_sync(1 + quad, 64);                               // the two warps that share these 32 rows
      m = fmaxf(xmax[row], xmax[128 + row]) * c;
      float rs[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[NH / 2];
#pragma unroll
      for (int e = 0; e < NH; e += 2) {
        const float p0 = exp2f(fmaf(__uint_as_float(r[e]), c, -m)), p1 = exp2f(fmaf(__uint_as_float(r[e + 1]), c, -m));
        rs[(e >> 1) & 3] += p0 + p1;
        pk[e >> 1] = pack_bf16(p0, p1);
      }
#pragma unroll
      for (int ch = 0; ch < NH / 16; ++ch)                        // P over this thread's own, already consumed, columns
        tmem_st8_async(tS + ch * 8, *reinterpret_cast<uint32_t(*)[8]>(&pk[ch * 8]));
      tc_wait_st();
      named_bar_sync(1 + quad, 64);                               // partner has read xsum (tile 0's row sum) / xmax
      xsum[half * 128 + row] = (rs[0] + rs[1]) + (rs[2] + rs[3]);
    }
    if (t == 0) m_prev = m;
    m_last = m;
    tc_fence_before();
    STRACE(5 + 4 * t);
    __syncthreads();
    STRACE(6 + 4 * t);
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      mma_pt_n<N>(tmem_base + 192 - O1_SHIFT * t, tmem_base, sV); // O(t) = P(t) V
      if (t + 1 < n_qt) mma_abt_n<N>(tmem_base, sQ + TILE_BYTES, sK);   // S(1) over P(0), which the PV MMA has consumed
      tc_commit(bar_s);
    }
    if (t == 1 && SPLIT_O) epilogue(0, m_prev, l_prev);     // overlaps the P(1) V MMA (different TMEM columns)
  }
  mbar_wait(bar_s, n_qt & 1);
  tc_fence_after();
  STRACE(11);
  epilogue(n_qt - 1, m_last, xsum[row] + xsum[128 + row]);
  STRACE(12);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
  STRACE(13);
}

// =================================================================================================
// backward (dQ, dK, dV in one kernel).  TMEM (512 columns): R1 [0, N) | R2 [N, 2N) | acc0 [384, 448) | acc1 [448, 512).
// Items, in order: A0, A1 (q-major, 128 q rows each): R1 = S = Q_t K^T, R2 = dP = dO_t V^T; threads: dS = P (dP - delta)
// -> bf16 over R2; acc0 = dQ_t = dS K.   B0, B1 (kv-major, 128 kv rows each): R1 = S^T = K_u Q^T, R2 = dP^T = V_u dO^T;
// threads: P^T -> bf16 over R1, dS^T -> bf16 over R2; acc0 = dV_u = P^T dO, acc1 = dK_u = dS^T Q.
// The tensor pipe executes in order, so the input MMAs of item i+1 are issued right behind the output MMAs of item i
// and one commit covers both; the accumulators of item i are stored while item i+1's inputs are being computed.
// =================================================================================================
template <int N>
__global__ void __launch_bounds__(SM_THREADS, 1)
attn_small_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, int S, int H,
                      float c, float scale, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv) {
  using Cfg = SmallCfg<N>;
  constexpr int NH = Cfg::NH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  check_align(base);
  const uint32_t sQ = base, sK = base + Cfg::TILE2, sV = base + 2 * Cfg::TILE2, sdO = base + 3 * Cfg::TILE2;
  float* lse_s = reinterpret_cast<float*>(smem_raw + 4 * Cfg::TILE2);        // [256]: +inf beyond S (p = 0 there)
  float* del_s = lse_s + 256;                                                // [256]
  float* xdel = del_s + 256;                                                 // [2][128] partial row sums of P dP
  const uint32_t bars = base + 4 * Cfg::TILE2 + 3072;
  const uint32_t bar_ld = bars, bar_in = bars + 8, tmem_slot = bars + 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
  const int h = blockIdx.x, b = blockIdx.y;
  const int row0 = b * S;
  const int n_t = S > 128 ? 2 : 1;               // 128-row tiles (q tiles == kv tiles)
  const int n_items = 2 * n_t;

  if (tid == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_in, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_wait();
  pdl_trigger();
  const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t tR1 = tmem_base + lane_off + half * NH, tR2 = tmem_base + lane_off + N + half * NH;
  const uint32_t tA0 = tmem_base + 384, tA1 = tmem_base + 448;

  auto issue_in = [&](int item) {                // elected lane: the two input MMAs of an item
    const int t = item % n_t;
    if (item < n_t) {                            // q-major
      mma_abt_n<N>(tmem_base, sQ + t * TILE_BYTES, sK);           // S   = Q_t K^T
      mma_abt_n<N>(tmem_base + N, sdO + t * TILE_BYTES, sV);      // dP  = dO_t V^T
    } else {                                     // kv-major
      mma_abt_n<N>(tmem_base, sK + t * TILE_BYTES, sQ);           // S^T  = K_u Q^T
      mma_abt_n<N>(tmem_base + N, sV + t * TILE_BYTES, sdO);      // dP^T = V_u dO^T
    }
  };

  if (warp_u == 0 && elect_one()) {
    mbar_expect_tx(bar_ld, 8 * TILE_BYTES);
    for (int t = 0; t < 2; ++t) {
      tma_load_2d(sQ + t * TILE_BYTES, &tm_qkv, bar_ld, h * 64, row0 + t * 128);
      tma_load_2d(sK + t * TILE_BYTES, &tm_qkv, bar_ld, (H + h) * 64, row0 + t * 128);
      tma_load_2d(sV + t * TILE_BYTES, &tm_qkv, bar_ld, (2 * H + h) * 64, row0 + t * 128);
      tma_load_2d(sdO + t * TILE_BYTES, &tm_do, bar_ld, h * 64, row0 + t * 128);
    }
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    issue_in(0);
    tc_commit(bar_in);
  }
  // per-row statistics: lse (log2 domain, from forward); delta is produced by the q-major items
  lse_s[tid] = tid < S ? lse[(static_cast<size_t>(b) * H + h) * S + tid] : INFINITY;
  del_s[tid] = 0.f;
  __syncthreads();

  auto store_outputs = [&](int item) {           // accumulators of a finished item -> global (bf16)
    const int t = item % n_t;
    const int r = t * 128 + row;                 // q row (items A) or kv row (items B)
    if (item < n_t) {                            // dQ_t: 64 columns in acc0, 32 per thread
      uint32_t v[32];
      tmem_ld32(tA0 + lane_off + half * 32, v);
      tc_wait_ld();
      if (r < S) {
        __nv_bfloat16* dst = dqkv + static_cast<size_t>(row0 + r) * (3 * H * 64) + h * 64 + half * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * scale, __uint_as_float(v[g * 8 + 1]) * scale);
          o.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * scale, __uint_as_float(v[g * 8 + 3]) * scale);
          o.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * scale, __uint_as_float(v[g * 8 + 5]) * scale);
          o.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * scale, __uint_as_float(v[g * 8 + 7]) * scale);
          reinterpret_cast<uint4*>(dst)[g] = o;
        }
      }
    } else {                                     // half 0 stores dV_u (acc0), half 1 stores dK_u (acc1, x scale)
      __nv_bfloat16* dst = dqkv + static_cast<size_t>(row0 + r) * (3 * H * 64) + ((half == 0 ? 2 * H : H) + h) * 64;
      const float sc = half == 0 ? 1.0f : scale;
      const uint32_t tsrc = (half == 0 ? tA0 : tA1) + lane_off;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tsrc + c0, v);
        tc_wait_ld();
        if (r < S) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * sc, __uint_as_float(v[g * 8 + 1]) * sc);
            o.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * sc, __uint_as_float(v[g * 8 + 3]) * sc);
            o.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * sc, __uint_as_float(v[g * 8 + 5]) * sc);
            o.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * sc, __uint_as_float(v[g * 8 + 7]) * sc);
            reinterpret_cast<uint4*>(dst + c0)[g] = o;
          }
        }
      }
    }
  };

  for (int item = 0; item < n_items; ++item) {
    mbar_wait(bar_in, item & 1);                 // inputs of this item ready; outputs of the previous item complete
    tc_fence_after();
    if (item > 0) store_outputs(item - 1);
    const int t = item % n_t;
    const bool active = t * 128 + quad * 32 < S;                  // warp-uniform
    if (active) {
      const int col0 = half * NH;
      // both slices (S and dP, NH columns each) are held in registers: one TMEM round trip per item, NH-way independent
      // arithmetic (one CTA per SM, 255 registers per thread available)
      uint32_t rs[NH], rp[NH];
#pragma unroll
      for (int ch = 0; ch < NH / 16; ++ch) {
        tmem_ld16a(tR1 + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(&rs[ch * 16]));
        tmem_ld16a(tR2 + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(&rp[ch * 16]));
      }
      uint32_t pk[NH / 2];
      if (item < n_t) {                          // q-major: row statistics, columns = kv (mask kv >= S)
        // delta_i = sum_j P_ij dP_ij is formed HERE from the recomputed fp32 P and the fp32 dP accumulators (not from
        // the bf16 O of the forward pass), so that sum_j dS_ij = 0 holds to fp32 rounding exactly as in the reference's
        // softmax backward: any error e_i in delta leaks e_i * (P-weighted mean key) into dQ, which the q_bias gradient
        // (a sum over all tokens of a clip batch) accumulates.
        const float my_lse = lse_s[t * 128 + row];
        tc_wait_ld();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const bool ragged = col0 + NH > S;
#pragma unroll
        for (int e = 0; e < NH; ++e) {
          float p = exp2f(fmaf(__uint_as_float(rs[e]), c, -my_lse));
          if (ragged && col0 + e >= S) p = 0.f;
          acc[e & 3] = fmaf(p, __uint_as_float(rp[e]), acc[e & 3]);
          rs[e] = __float_as_uint(p);
        }
        xdel[half * 128 + row] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        named_bar_sync(1 + quad, 64);                             // the two warps that share these 32 rows
        const float my_del = xdel[row] + xdel[128 + row];
        if (half == 0) del_s[t * 128 + row] = my_del;             // column statistic of the kv-major items
#pragma unroll
        for (int e = 0; e < NH; e += 2)
          pk[e >> 1] = pack_bf16(__uint_as_float(rs[e]) * (__uint_as_float(rp[e]) - my_del),
                                 __uint_as_float(rs[e + 1]) * (__uint_as_float(rp[e + 1]) - my_del));
#pragma unroll
        for (int ch = 0; ch < NH / 16; ++ch)     // dS over this thread's own, already consumed, dP columns
          tmem_st8_async(tR2 + ch * 8, *reinterpret_cast<uint32_t(*)[8]>(&pk[ch * 8]));
        named_bar_sync(1 + quad, 64);                             // xdel may be rewritten by the next q-major item
      } else {                                   // kv-major: column statistics (q); lse_s = +inf beyond S gives p = 0
        uint32_t dk[NH / 2];
        const float4* l4 = reinterpret_cast<const float4*>(lse_s + col0);
        const float4* d4 = reinterpret_cast<const float4*>(del_s + col0);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < NH / 4; ++g) {
          const float4 l = l4[g], d = d4[g];
          const float p0 = exp2f(fmaf(__uint_as_float(rs[4 * g + 0]), c, -l.x)), p1 = exp2f(fmaf(__uint_as_float(rs[4 * g + 1]), c, -l.y));
          const float p2 = exp2f(fmaf(__uint_as_float(rs[4 * g + 2]), c, -l.z)), p3 = exp2f(fmaf(__uint_as_float(rs[4 * g + 3]), c, -l.w));
          pk[2 * g] = pack_bf16(p0, p1); pk[2 * g + 1] = pack_bf16(p2, p3);
          dk[2 * g] = pack_bf16(p0 * (__uint_as_float(rp[4 * g + 0]) - d.x), p1 * (__uint_as_float(rp[4 * g + 1]) - d.y));
          dk[2 * g + 1] = pack_bf16(p2 * (__uint_as_float(rp[4 * g + 2]) - d.z), p3 * (__uint_as_float(rp[4 * g + 3]) - d.w));
        }
#pragma unroll
        for (int ch = 0; ch < NH / 16; ++ch) {
          tmem_st8_async(tR1 + ch * 8, *reinterpret_cast<uint32_t(*)[8]>(&pk[ch * 8]));      // P^T
          tmem_st8_async(tR2 + ch * 8, *reinterpret_cast<uint32_t(*)[8]>(&dk[ch * 8]));      // dS^T
        }
      }
      tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
      if (item < n_t) {
        mma_pt_n<N>(tA0, tmem_base + N, sK);                      // dQ_t = dS K
      } else {
        mma_pt_n<N>(tA0, tmem_base, sdO);                         // dV_u = P^T dO
        mma_pt_n<N>(tA1, tmem_base + N, sQ);                      // dK_u = dS^T Q
      }
      if (item + 1 < n_items) issue_in(item + 1);
      tc_commit(bar_in);
    }
  }
  mbar_wait(bar_in, n_items & 1);
  tc_fence_after();
  store_outputs(n_items - 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// -------------------------------------------------------------------------------------------------
template <int N>
static int launch_small_fwd(const CUtensorMap& tq, int B, int S, int H, float scale, void* out, float* lse, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallCfg<N>::FWD_SMEM));
    attr_set = true;
  }
  MOFO_CUDA(launch_pdl(attn_small_fwd_kernel<N>, dim3(H, B), dim3(SM_THREADS), SmallCfg<N>::FWD_SMEM, s, tq, S, H,
                       scale * 1.4426950408889634f, reinterpret_cast<__nv_bfloat16*>(out), lse));
  return MOFO_OK;
}

template <int N>
static int launch_small_bwd(const CUtensorMap& tq, const CUtensorMap& td, int B, int S, int H, float scale, const void* out,
                            const void* dout, const float* lse, void* dqkv, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    MOFO_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallCfg<N>::BWD_SMEM));
    attr_set = true;
  }
  MOFO_CUDA(launch_pdl(attn_small_bwd_kernel<N>, dim3(H, B), dim3(SM_THREADS), SmallCfg<N>::BWD_SMEM, s, tq, td, S, H,
                       scale * 1.4426950408889634f, scale, lse, reinterpret_cast<__nv_bfloat16*>(dqkv)));
  return MOFO_OK;
}

#ifdef MOFO_ATTN_TRACE
extern "C" int mofo_debug_read_strace(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_strace, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif

int attn_small_fwd(const void* qkv, int B, int S, int H, float scale, void* out, float* lse, cudaStream_t stream) {
  CUtensorMap tq;
  int rc = get_tmap(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * H * 64, 3ull * H * 64, 128);
  if (rc) return rc;
  if (S <= 64) return launch_small_fwd<64>(tq, B, S, H, scale, out, lse, stream);
  if (S <= 128) return launch_small_fwd<128>(tq, B, S, H, scale, out, lse, stream);
  if (S <= 160) return launch_small_fwd<160>(tq, B, S, H, scale, out, lse, stream);
  return launch_small_fwd<192>(tq, B, S, H, scale, out, lse, stream);
}

int attn_small_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int S, int H, float scale,
                   void* dqkv, cudaStream_t stream) {
  CUtensorMap tq, td;
  int rc = get_tmap(&tq, qkv, static_cast<uint64_t>(B) * S, 3ull * H * 64, 3ull * H * 64, 128);
  if (rc) return rc;
  rc = get_tmap(&td, dout, static_cast<uint64_t>(B) * S, 1ull * H * 64, 1ull * H * 64, 128);
  if (rc) return rc;
  if (S <= 64) return launch_small_bwd<64>(tq, td, B, S, H, scale, out, dout, lse, dqkv, stream);
  if (S <= 128) return launch_small_bwd<128>(tq, td, B, S, H, scale, out, dout, lse, dqkv, stream);
  if (S <= 160) return launch_small_bwd<160>(tq, td, B, S, H, scale, out, dout, lse, dqkv, stream);
  return launch_small_bwd<192>(tq, td, B, S, H, scale, out, dout, lse, dqkv, stream);
}

}  // namespace mofo
