// HBM-/latency-bound kernels of the MOFO pretraining step: tube masking, tubelet gather, LayerNorm fwd/bwd,
// decoder-input assembly, target + MSE, weight casts, bias-gradient column sums, gradient norm.
// All of them are integer/byte/elementwise work: coalesced 16-byte accesses, no tensor cores.
#include <stdlib.h>

#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

// =================================================================================================
// (1) tube masking  — masking_generator.py:17-24, 43-85
// =================================================================================================
struct WordStream {
  const uint32_t* ws;   // first n_s words, staged in shared memory
  const uint32_t* wg;   // the whole stream in global memory (words beyond the staged prefix)
  int n_s;
  int n;
  int pos;
  bool ok;
  // numpy legacy random_interval (masked rejection on 32-bit MT19937 outputs)
  __device__ int interval(int mx) {
    if (mx == 0) return 0;
    uint32_t mask = static_cast<uint32_t>(mx);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while (true) {
      if (pos >= n) { ok = false; return 0; }
      uint32_t v = (pos < n_s ? ws[pos] : wg[pos]) & mask;
      ++pos;
      if (v <= static_cast<uint32_t>(mx)) return static_cast<int>(v);
    }
  }
};

template <typename T>
__device__ void legacy_shuffle(T* x, int n, WordStream& ws) {
  for (int i = n - 1; i > 0; --i) {
    int j = ws.interval(i);
    T tmp = x[i]; x[i] = x[j]; x[j] = tmp;
  }
}

// one warp per clip; lane 0 runs the (inherently sequential) Fisher-Yates passes in shared memory,
// the whole warp does the predicate, the compactions and the writes.
template <bool BB>
__global__ void tube_mask_kernel(const double* __restrict__ bb_first, const uint32_t* __restrict__ rng_words, int B,
                                 int W, int Wsm, int T, int H, int Wd, int nmask, double ratio_bb,
                                 uint8_t* __restrict__ mask, int32_t* __restrict__ vis_idx,
                                 int32_t* __restrict__ msk_idx, int32_t* __restrict__ words_used) {
  extern __shared__ int16_t sm[];
  const int npf = H * Wd;
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * warps + wid;
  int16_t* index = sm + wid * (3 * npf);     // in-box ids, later shuffled
  int16_t* remaining = index + npf;          // candidates 0..nmask-1 not selected
  int16_t* f = remaining + npf;              // per-frame mask (0/1)
  if (b >= B) return;
  // the clip's word stream is staged in shared memory by the whole warp (coalesced): lane 0's sequential draws then
  // cost a shared-memory load each instead of a dependent global load (~270 draws per clip)
  uint32_t* wsm = reinterpret_cast<uint32_t*>(sm + ((warps * 3 * npf + 1) & ~1)) + static_cast<size_t>(wid) * Wsm;
  const int nstage = W < Wsm ? W : Wsm;
  for (int i = lane; i < nstage; i += 32) wsm[i] = rng_words[static_cast<size_t>(b) * W + i];
  __syncwarp();
  WordStream ws{wsm, rng_words + static_cast<size_t>(b) * W, nstage, W, 0, true};

  for (int i = lane; i < npf; i += 32) f[i] = 0;
  __syncwarp();
  if (BB) {
    const double b0 = bb_first[b * 4 + 0], b1 = bb_first[b * 4 + 1], b2 = bb_first[b * 4 + 2], b3 = bb_first[b * 4 + 3];
    int n_in = 0;
    for (int base = 0; base < npf; base += 32) {
      int id = base + lane;
      bool in = false;
      if (id < npf) {
        int j = id / Wd, k = id % Wd;
        double x1t = j * 16, x2t = j * 16 + 16, y1t = k * 16, y2t = k * 16 + 16;
        in = !((b0 > x2t || b2 < x1t) && (b1 > y2t || b3 < y1t));       // :55  (cross predicate, x vs row j)
      }
      unsigned bal = __ballot_sync(0xffffffffu, in);
      if (in) index[n_in + __popc(bal & ((1u << lane) - 1))] = static_cast<int16_t>(id);
      n_in += __popc(bal);
    }
    __syncwarp();
    int cap = 0;
    if (lane == 0) {
      legacy_shuffle(index, n_in, ws);                                   // :62
      cap = static_cast<int>(static_cast<double>(n_in) * ratio_bb);     // int(len(index)*ratio)  :64
      cap = cap < nmask ? cap : nmask;
      for (int i = 0; i < cap; ++i) f[index[i]] = 1;                     // :67-68
    }
    cap = __shfl_sync(0xffffffffu, cap, 0);
    __syncwarp();
    int n_rem = 0;                                                       // setdiff1d(arange(nmask), selected)  :72
    for (int base = 0; base < nmask; base += 32) {
      int id = base + lane;
      bool keep = id < nmask && f[id] == 0;
      unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) remaining[n_rem + __popc(bal & ((1u << lane) - 1))] = static_cast<int16_t>(id);
      n_rem += __popc(bal);
    }
    __syncwarp();
    if (lane == 0) {
      legacy_shuffle(remaining, n_rem, ws);                              // :75
      int need = nmask - cap;                                            // :71
      for (int i = 0; i < need; ++i) f[remaining[i]] = 1;                // :76-77
    }
  } else {
    for (int i = lane; i < npf; i += 32) f[i] = (i >= npf - nmask) ? 1 : 0;   // hstack([zeros, ones])  :18-21
    __syncwarp();
    if (lane == 0) legacy_shuffle(f, npf, ws);                                // :22
  }
  __syncwarp();
  if (lane == 0) words_used[b] = ws.ok ? ws.pos : -1;

  // tile over T slabs (:84) + ascending index lists
  const int nvis = npf - nmask;
  int cv = 0, cm = 0;
  for (int base = 0; base < npf; base += 32) {
    int id = base + lane;
    bool valid = id < npf;
    bool m = valid && f[id] != 0;
    unsigned balm = __ballot_sync(0xffffffffu, m);
    unsigned balv = __ballot_sync(0xffffffffu, valid && !m);
    int pm = cm + __popc(balm & ((1u << lane) - 1));
    int pv = cv + __popc(balv & ((1u << lane) - 1));
    if (valid) {
      for (int t = 0; t < T; ++t) {
        mask[(static_cast<size_t>(b) * T + t) * npf + id] = m ? 1 : 0;
        if (m) msk_idx[(static_cast<size_t>(b) * T + t) * nmask + pm] = t * npf + id;
        else   vis_idx[(static_cast<size_t>(b) * T + t) * nvis + pv] = t * npf + id;
      }
    }
    cm += __popc(balm);
    cv += __popc(balv);
  }
}

// index lists of an arbitrary boolean mask: one warp per clip, ballot/popc scan (ascending order)
__global__ void mask_indices_kernel(const uint8_t* __restrict__ mask, int B, int N, int n_msk,
                                    int32_t* __restrict__ vis_idx, int32_t* __restrict__ msk_idx,
                                    int32_t* __restrict__ bad_rows) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n_vis = N - n_msk;
  int cv = 0, cm = 0;
  for (int base = 0; base < N; base += 32) {
    const int id = base + lane;
    const bool valid = id < N;
    const bool m = valid && mask[static_cast<size_t>(b) * N + id] != 0;
    const unsigned balm = __ballot_sync(0xffffffffu, m);
    const unsigned balv = __ballot_sync(0xffffffffu, valid && !m);
    const int pm = cm + __popc(balm & ((1u << lane) - 1));
    const int pv = cv + __popc(balv & ((1u << lane) - 1));
    if (m && pm < n_msk) msk_idx[static_cast<size_t>(b) * n_msk + pm] = id;
    if (valid && !m && pv < n_vis) vis_idx[static_cast<size_t>(b) * n_vis + pv] = id;
    cm += __popc(balm);
    cv += __popc(balv);
  }
  if (cm != n_msk) {        // malformed row: every index entry stays a valid token id (0), and the row is reported
    for (int i = min(cm, n_msk) + lane; i < n_msk; i += 32) msk_idx[static_cast<size_t>(b) * n_msk + i] = 0;
    for (int i = min(cv, n_vis) + lane; i < n_vis; i += 32) vis_idx[static_cast<size_t>(b) * n_vis + i] = 0;
    if (lane == 0) atomicAdd(bad_rows, 1);
  }
}

// =================================================================================================
// (2) tubelet gather (im2col of the visible tubes)  — modeling_finetune.py:238-248 + modeling_pretrain.py:90
// =================================================================================================
// 192 threads per tube: thread -> (segment = (c,p0,p1) row of 16 px, half of 8 px).
__global__ void __launch_bounds__(192) gather_tubes_kernel(const float* __restrict__ video,
                                                           const int32_t* __restrict__ idx, int n_idx, int frames,
                                                           int size, __nv_bfloat16* __restrict__ A) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x;
  const int b = row / n_idx;
  const int hw = size >> 4;
  const int tok = min(max(idx[row], 0), (frames >> 1) * hw * hw - 1);     // never index outside the clip
  const int t = tok / (hw * hw), h = (tok / hw) % hw, w = tok % hw;
  const int seg = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int c = seg >> 5, p0 = (seg >> 4) & 1, p1 = seg & 15;
  const float* src = video + ((static_cast<size_t>(b) * 3 + c) * frames + 2 * t + p0) * size * size +
                     static_cast<size_t>(16 * h + p1) * size + 16 * w + half * 8;
  float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
  float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
  uint4 o;
  o.x = pack_bf16(v0.x, v0.y); o.y = pack_bf16(v0.z, v0.w);
  o.z = pack_bf16(v1.x, v1.y); o.w = pack_bf16(v1.z, v1.w);
  *reinterpret_cast<uint4*>(A + static_cast<size_t>(row) * 1536 + seg * 16 + half * 8) = o;
}

// =================================================================================================
// (5) LayerNorm
// =================================================================================================
constexpr int LN_MAX_CHUNKS = 8;  // D <= 1024 (float4 chunk per lane per iteration)

__device__ __forceinline__ size_t map_row(int m, int group_rows, int in_group_rows, int in_row_offset) {
  return static_cast<size_t>(m / group_rows) * in_group_rows + in_row_offset + (m % group_rows);
}

template <int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int M, int D, float eps,
                                                            int group_rows, int in_group_rows, int in_row_offset,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int nch = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + map_row(m, group_rows, in_group_rows, in_row_offset) * D);
  float4 v[NCH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int ch = lane + i * 32;
    if (ch < nch) { v[i] = xr[ch]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int ch = lane + i * 32;
    if (ch < nch) {
      float a = v[i].x - mu, b2 = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b2 * b2) + (c * c + d * d);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { mean[m] = mu; rstd[m] = rs; }
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(m) * D);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int ch = lane + i * 32;
    if (ch < nch) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
      float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + ch);
      uint2 o;
      o.x = pack_bf16((v[i].x - mu) * rs * g.x + bt.x, (v[i].y - mu) * rs * g.y + bt.y);
      o.y = pack_bf16((v[i].z - mu) * rs * g.z + bt.z, (v[i].w - mu) * rs * g.w + bt.w);
      yr[ch] = o;
    }
  }
}

// SPLIT warps per row (1 for D <= 384, 2 above: each warp owns a contiguous half of the columns and the two exchange
// their partial row sums through shared memory and a 64-thread named barrier), grid-stride over rows; per-lane
// register partials of dgamma/dbeta, reduced through shared memory per CTA and accumulated with one 16-byte vector
// reduction per 4 columns per CTA.
template <int NCH, int SPLIT>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres, int M, int D,
    int group_rows, int in_group_rows, int in_row_offset, float* __restrict__ dx_f32,
    __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
    const float* __restrict__ bf16_row_scale) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float red[];  // [warps / SPLIT][2*D]
  __shared__ float2 xs[2][8];     // SPLIT == 2: partial (s1, s2) of every warp, double buffered over the row loop
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, warps = (blockDim.x >> 5) / SPLIT;   // warps: row slots per CTA
  const int slot = wid / SPLIT, hw = wid % SPLIT;
  const int nchw = (D >> 2) / SPLIT;                 // float4 chunks per warp
  const int ch0 = hw * nchw;                         // first chunk of this warp's column range
  const int nch = ch0 + nchw;                        // one past its last chunk
  float4 dg[NCH], db[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { dg[i] = make_float4(0, 0, 0, 0); db[i] = make_float4(0, 0, 0, 0); }

  // Rows are software-pipelined when a lane owns <= 3 chunks (D <= 384, or D = 768 split over two warps): the loads
  // of the warp's NEXT row are issued before the current row is reduced and stored, so a warp always has one row of
  // x / dy / dres in flight (more chunks per lane would not fit the 128 registers available at 2 CTAs / SM).
  constexpr bool PIPE = NCH <= 3;
  const int stride = gridDim.x * warps;
  float4 xv[PIPE ? NCH : 1], rres[NCH];
  uint2 d2[PIPE ? NCH : 1];
  float mu = 0.f, rs = 0.f;
  auto fetch = [&](int m, float4 (&xo)[NCH], uint2 (&dyo)[NCH], float4 (&ro)[NCH], float& muo, float& rso) {
    const size_t xrow = map_row(m, group_rows, in_group_rows, in_row_offset);
    const float4* xr = reinterpret_cast<const float4*>(x + xrow * D);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(m) * D);
    muo = mean[m]; rso = rstd[m];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = ch0 + lane + i * 32;
      if (ch < nch) {
        xo[i] = xr[ch];
        dyo[i] = dyr[ch];
        if (dres) ro[i] = reinterpret_cast<const float4*>(dres + xrow * D)[ch];
      }
    }
  };
  int m = blockIdx.x * warps + slot;
  int it = 0;
  if constexpr (PIPE) {
    if (m < M) fetch(m, xv, d2, rres, mu, rs);
  }
  for (; m < M; m += stride) {
    const size_t xrow = map_row(m, group_rows, in_group_rows, in_row_offset);
    float4 xn[PIPE ? NCH : 1], rn[PIPE ? NCH : 1];
    uint2 dn[PIPE ? NCH : 1];
    float mun = 0.f, rsn = 0.f;
    if constexpr (PIPE) {
      if (m + stride < M) fetch(m + stride, xn, dn, rn, mun, rsn);
    }
    float4 xh[NCH], g[NCH];
    float s1 = 0.f, s2 = 0.f;
    auto accum = [&](int i, int ch, const float4& xq, const uint2& dq) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
      const float4 d = make_float4(bf16_lo(dq.x), bf16_hi(dq.x), bf16_lo(dq.y), bf16_hi(dq.y));
      xh[i] = make_float4((xq.x - mu) * rs, (xq.y - mu) * rs, (xq.z - mu) * rs, (xq.w - mu) * rs);
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
    };
    if constexpr (PIPE) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = ch0 + lane + i * 32;
        if (ch < nch) accum(i, ch, xv[i], d2[i]);
      }
    } else {                 // D > 384: x / dy are consumed chunk by chunk as they arrive (registers)
      const float4* xr = reinterpret_cast<const float4*>(x + xrow * D);
      const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(m) * D);
      mu = mean[m]; rs = rstd[m];
      if (dres) {            // issued together with x / dy: one global-latency phase per row instead of two
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int ch = ch0 + lane + i * 32;
          if (ch < nch) rres[i] = reinterpret_cast<const float4*>(dres + xrow * D)[ch];
        }
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = ch0 + lane + i * 32;
        if (ch < nch) accum(i, ch, xr[ch], dyr[ch]);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if constexpr (SPLIT == 2) {
      if (lane == 0) xs[it & 1][wid] = make_float2(s1, s2);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
      const float2 o = xs[it & 1][wid ^ 1];
      s1 += o.x; s2 += o.y;
      ++it;
    }
    s1 /= D;
    s2 /= D;
    const float bsc = bf16_row_scale ? __ldg(bf16_row_scale + m / group_rows) : 1.0f;   // DropPath scale of the consuming branch
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      int ch = ch0 + lane + i * 32;
      if (ch < nch) {
        float4 o = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                               rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
        if (dres) { o.x += rres[i].x; o.y += rres[i].y; o.z += rres[i].z; o.w += rres[i].w; }
        if (dx_f32) reinterpret_cast<float4*>(dx_f32 + xrow * D)[ch] = o;
        if (dx_bf16) {
          uint2 p; p.x = pack_bf16(o.x * bsc, o.y * bsc); p.y = pack_bf16(o.z * bsc, o.w * bsc);
          reinterpret_cast<uint2*>(dx_bf16 + xrow * D)[ch] = p;
        }
      }
    }
    if constexpr (PIPE) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) { xv[i] = xn[i]; d2[i] = dn[i]; rres[i] = rn[i]; }
      mu = mun; rs = rsn;
    }
  }
#ifdef MOFO_LN_SKIP_REDUCE      // tuning aid (variant builds only): bounds the cost of the parameter-gradient reduction
  if (dg[0].x != 12345.678f) return;
#endif
  // CTA reduction of the parameter-gradient partials
  float* mine = red + static_cast<size_t>(slot) * 2 * D;
  const int Dq = D >> 2;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int ch = ch0 + lane + i * 32;
    if (ch < nch) {
      reinterpret_cast<float4*>(mine)[ch] = dg[i];
      reinterpret_cast<float4*>(mine)[Dq + ch] = db[i];
    }
  }
  __syncthreads();
  // one 16-byte vector reduction per 4 columns per CTA straight into dgamma / dbeta (red.global.add.v4.f32): no
  // workspace pass, no ticket, no fence - the kernel boundary publishes the sums
  for (int c4 = threadIdx.x; c4 < (2 * D) >> 2; c4 += blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w2 = 0; w2 < warps; ++w2) {
      const float4 v = reinterpret_cast<const float4*>(red + static_cast<size_t>(w2) * 2 * D)[c4];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int c = c4 << 2;
    atomicAdd(reinterpret_cast<float4*>(c < D ? dgamma + c : dbeta + (c - D)), s);
  }
}

// =================================================================================================
// (6) decoder input assembly  — modeling_pretrain.py:260-263
// =================================================================================================
__global__ void __launch_bounds__(256) assemble_fwd_kernel(const float* __restrict__ mask_token, const float* __restrict__ pos,
                                    const int32_t* __restrict__ msk_idx, int rows, int n_vis, int n_msk, int Dd,
                                    float* __restrict__ x_full) {
  pdl_wait();
  pdl_trigger();
  // warp per masked row (b*n_msk + j), 8 rows per CTA, grid-stride: a lane moves Dd/128 float4 per row
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const float4* mt = reinterpret_cast<const float4*>(mask_token);
  for (int r = blockIdx.x * wpc + wid; r < rows; r += gridDim.x * wpc) {
    const int b = r / n_msk, j = r % n_msk;
    const float4* p = reinterpret_cast<const float4*>(pos + static_cast<size_t>(msk_idx[r]) * Dd);
    float4* o = reinterpret_cast<float4*>(x_full + (static_cast<size_t>(b) * (n_vis + n_msk) + n_vis + j) * Dd);
    for (int c = lane; c < (Dd >> 2); c += 32) {
      float4 a = __ldg(mt + c), q = __ldg(p + c);
      o[c] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
  }
}

// grid = (row chunks, B).  Threads own float4 column chunks; rows of the chunk are streamed.
constexpr int ASM_RG = 4;          // row groups per CTA (blockDim.y)
__global__ void assemble_bwd_kernel(const float* __restrict__ dx_full, int n_vis, int n_msk, int Dd, int rows_per_cta,
                                    float* __restrict__ dmask_token, __nv_bfloat16* __restrict__ dvis) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float4 part[];                      // [ASM_RG][Dd / 4]
  const int b = blockIdx.y;
  const int N = n_vis + n_msk;
  const int C4 = Dd >> 2;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  const int c = threadIdx.x, ry = threadIdx.y;
  float4 acc = make_float4(0, 0, 0, 0);
  if (c < C4) {
    const float4* src = reinterpret_cast<const float4*>(dx_full + static_cast<size_t>(b) * N * Dd) + c;
    for (int r = r0 + ry; r < r1; r += 4 * ASM_RG) {    // 4 independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * ASM_RG;
        v[u] = rr < r1 ? src[static_cast<size_t>(rr) * C4] : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * ASM_RG;
        if (rr >= r1) continue;
        if (rr < n_vis) {
          uint2 p; p.x = pack_bf16(v[u].x, v[u].y); p.y = pack_bf16(v[u].z, v[u].w);
          reinterpret_cast<uint2*>(dvis + (static_cast<size_t>(b) * n_vis + rr) * Dd)[c] = p;
        } else {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        }
      }
    }
    part[ry * C4 + c] = acc;
  }
  __syncthreads();
  if (ry == 0 && c < C4 && r1 > n_vis) {               // this CTA saw masked rows: one 16-byte reduction per 4 columns
#pragma unroll
    for (int g = 1; g < ASM_RG; ++g) {
      const float4 o = part[g * C4 + c];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(reinterpret_cast<float4*>(dmask_token) + c, acc);
  }
}

// zero the first `nz` rows of every group of `group_rows` rows of an f32 and / or a bf16 [groups*group_rows, D] matrix
__global__ void __launch_bounds__(256) zero_rows_kernel(float* __restrict__ f32, __nv_bfloat16* __restrict__ b16, int group_rows,
                                                        int nz, int D, int64_t total4) {
  pdl_wait();
  pdl_trigger();
  const int C4 = D >> 2;
  const int64_t per_group = static_cast<int64_t>(nz) * C4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t g = i / per_group, r = i % per_group;
    const int64_t dst = g * group_rows * C4 + r;
    if (f32) reinterpret_cast<float4*>(f32)[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b16) reinterpret_cast<uint2*>(b16)[dst] = make_uint2(0u, 0u);
  }
}

// =================================================================================================
// (6d) key-masked softmax of the box-focused classifier's cross attention  - modeling_finetune.py:125-160 (CrossAttention:
// queries = tokens in the box, keys / values = tokens outside it; 3 heads of 256, so the score rows do not fit the 64-wide
// attention kernels and the products run as tcgen05 GEMMs with these two kernels between them)
// =================================================================================================
// One CTA of 128 threads per score row; a thread owns up to MS_CHUNKS float4 column chunks (Nk <= 2048, Nk % 4 == 0), so a
// row is read once, reduced in registers / two block reductions, and written once.
constexpr int MS_THREADS = 128, MS_CHUNKS = 4;

template <bool MAX>
__device__ __forceinline__ float ms_block_reduce(float v, float* red) {
  v = MAX ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();                                   // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < MS_THREADS / 32; ++i) r = MAX ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// P[row, k] = softmax_k(scale * S[row, k]) over the keys with allowed[b, k] != 0, exactly 0 elsewhere; rows = B * rows_per_b
__global__ void __launch_bounds__(MS_THREADS) masked_softmax_fwd_kernel(const float* __restrict__ S, const uint8_t* __restrict__ allowed,
                                                                        int rows_per_b, int Nk, float scale,
                                                                        __nv_bfloat16* __restrict__ P) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[MS_THREADS / 32];
  const size_t row = blockIdx.x;
  const int b = static_cast<int>(row / rows_per_b);
  const float4* src = reinterpret_cast<const float4*>(S + row * Nk);
  const uchar4* al = reinterpret_cast<const uchar4*>(allowed + static_cast<size_t>(b) * Nk);
  const int C4 = Nk >> 2;
  float4 v[MS_CHUNKS];
  const float ninf = __int_as_float(0xff800000);
  float mx = ninf;
#pragma unroll
  for (int i = 0; i < MS_CHUNKS; ++i) {
    const int c = threadIdx.x + i * MS_THREADS;
    v[i] = make_float4(ninf, ninf, ninf, ninf);
    if (c < C4) {
      const float4 x = src[c];
      const uchar4 a = __ldg(al + c);
      v[i].x = a.x ? x.x * scale : ninf; v[i].y = a.y ? x.y * scale : ninf;
      v[i].z = a.z ? x.z * scale : ninf; v[i].w = a.w ? x.w * scale : ninf;
    }
    mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
  }
  mx = ms_block_reduce<true>(mx, red);
  if (mx == ninf) mx = 0.f;                          // no key allowed (the host never passes such a row): all-zero row, no NaN
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MS_CHUNKS; ++i) {
    v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  sum = ms_block_reduce<false>(sum, red);
  const float inv = sum > 0.f ? __fdividef(1.0f, sum) : 0.f;
  uint2* dst = reinterpret_cast<uint2*>(P + row * Nk);
#pragma unroll
  for (int i = 0; i < MS_CHUNKS; ++i) {
    const int c = threadIdx.x + i * MS_THREADS;
    if (c < C4) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x * inv, v[i].y * inv), hi = __floats2bfloat162_rn(v[i].z * inv, v[i].w * inv);
      dst[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
}

// dS[row, k] = scale * P * (dP - sum_k P * dP)   (softmax backward with the 1/sqrt(d) of `q * scale` folded in)
__global__ void __launch_bounds__(MS_THREADS) masked_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP,
                                                                        int Nk, float scale, __nv_bfloat16* __restrict__ dS) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[MS_THREADS / 32];
  const size_t row = blockIdx.x;
  const uint2* ps = reinterpret_cast<const uint2*>(P + row * Nk);
  const float4* ds = reinterpret_cast<const float4*>(dP + row * Nk);
  const int C4 = Nk >> 2;
  float4 p[MS_CHUNKS], g[MS_CHUNKS];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < MS_CHUNKS; ++i) {
    const int c = threadIdx.x + i * MS_THREADS;
    p[i] = make_float4(0, 0, 0, 0); g[i] = make_float4(0, 0, 0, 0);
    if (c < C4) {
      const uint2 w = ps[c];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
      const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
      p[i] = make_float4(a.x, a.y, bb.x, bb.y);
      g[i] = ds[c];
    }
    dot += (p[i].x * g[i].x + p[i].y * g[i].y) + (p[i].z * g[i].z + p[i].w * g[i].w);
  }
  dot = ms_block_reduce<false>(dot, red);
  uint2* dst = reinterpret_cast<uint2*>(dS + row * Nk);
#pragma unroll
  for (int i = 0; i < MS_CHUNKS; ++i) {
    const int c = threadIdx.x + i * MS_THREADS;
    if (c < C4) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(scale * p[i].x * (g[i].x - dot), scale * p[i].y * (g[i].y - dot));
      const __nv_bfloat162 hi = __floats2bfloat162_rn(scale * p[i].z * (g[i].z - dot), scale * p[i].w * (g[i].w - dot));
      dst[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
}

// dst bf16 [M, N] (row stride ldd) = src f32 [M, N] (row stride lds); 4 elements per thread
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, int lds, int64_t total4, int N4,
                                                            __nv_bfloat16* __restrict__ dst, int ldd) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = blockIdx.x * 256L + threadIdx.x; i < total4; i += gridDim.x * 256L) {
    const int64_t r = i / N4;
    const int c = static_cast<int>(i - r * N4) * 4;
    const float4 x = *reinterpret_cast<const float4*>(src + r * lds + c);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
    *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}

// =================================================================================================
// (6b) token mean pooling (finetuning classifier)  — modeling_finetune.py:400-401  fc_norm(x.mean(1))
// =================================================================================================
// grid = (column blocks of 128, B), block (32, 8): lane = one float4 column chunk, 8 row groups stream the N rows with 4
// independent 16-byte loads in flight each; the row groups are combined in a FIXED order through shared memory, so the
// result is deterministic (no atomics: a mean that feeds a LayerNorm amplifies last-bit noise).
constexpr int TM_RG = 8;
__global__ void __launch_bounds__(256) token_mean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wts, int N,
                                                             int D, float inv_n, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float4 part[TM_RG][32];
  const int b = blockIdx.y;
  const int C4 = D >> 2;
  const int c = blockIdx.x * 32 + threadIdx.x, ry = threadIdx.y;
  float4 acc = make_float4(0, 0, 0, 0);
  if (c < C4) {
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * N * D) + c;
    const float* wb = wts ? wts + static_cast<size_t>(b) * N : nullptr;     // per-token weights (box-focused pooling) or plain mean
    for (int r = ry; r < N; r += 4 * TM_RG) {
      float4 v[4];
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * TM_RG;
        v[u] = rr < N ? src[static_cast<size_t>(rr) * C4] : make_float4(0, 0, 0, 0);
        wv[u] = rr < N ? (wb ? __ldg(wb + rr) : inv_n) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(wv[u], v[u].x, acc.x); acc.y = fmaf(wv[u], v[u].y, acc.y);
        acc.z = fmaf(wv[u], v[u].z, acc.z); acc.w = fmaf(wv[u], v[u].w, acc.w);
      }
    }
  }
  part[ry][threadIdx.x] = acc;
  __syncthreads();
  if (ry == 0 && c < C4) {
#pragma unroll
    for (int g = 1; g < TM_RG; ++g) {
      const float4 o = part[g][threadIdx.x];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    reinterpret_cast<float4*>(out + static_cast<size_t>(b) * D)[c] = acc;
  }
}

// dx[b, n, :] = dpooled[b, :] / N for every token n (f32 and / or bf16 copy), 16 bytes per thread and iteration
__global__ void __launch_bounds__(256) token_mean_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ wts,
                                                             int N, int D, float inv_n, int64_t total4, float* __restrict__ dx_f32,
                                                             __nv_bfloat16* __restrict__ dx_bf16,
                                                             const float* __restrict__ bf16_row_scale) {
  pdl_wait();
  pdl_trigger();
  const int C4 = D >> 2;
  const int64_t per_clip = static_cast<int64_t>(N) * C4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_clip), c = static_cast<int>(i % C4);
    float4 v = __ldg(reinterpret_cast<const float4*>(dpooled + static_cast<size_t>(b) * D) + c);
    const float wn = wts ? __ldg(wts + i / C4) : inv_n;          // i / C4 = b * N + n
    v.x *= wn; v.y *= wn; v.z *= wn; v.w *= wn;
    if (dx_f32) reinterpret_cast<float4*>(dx_f32)[i] = v;
    if (dx_bf16) {
      const float sc = bf16_row_scale ? __ldg(bf16_row_scale + b) : 1.0f;
      uint2 p; p.x = pack_bf16(v.x * sc, v.y * sc); p.y = pack_bf16(v.z * sc, v.w * sc);
      reinterpret_cast<uint2*>(dx_bf16)[i] = p;
    }
  }
}

// =================================================================================================
// (6c) tokens inside the motion box (box-focused classifier)  — modeling_finetune.py:589-630, 555-585
// =================================================================================================
// The reference paints the box of every frame into an all-zero clip, runs an all-ones Conv3d (patch_yab) over it, averages
// the (identical) channels, clamps to [0,1] and casts to bool: a tube is "in the box" iff ANY pixel of either of its two
// frames lies inside that frame's box.  Closed form: per frame j, [16h, 16h+16) x [16w, 16w+16) intersects the slice
// [y1:y2) x [x1:x2) (Python slice semantics: negative bounds count from the end, all bounds clamp to [0, size]).
// weights (fusing 'weighted_mean', mode 1, :571-572): (mean_in * 1 + mean_out * 0.5) / 2 -> 0.5 / n_in for tokens in the box,
// 0.25 / n_out for the others; no token in the box -> plain mean (:560-562); fusing 'org' (mode 0): plain mean.  n_out = 0
// gives NaN in the reference (mean of an empty selection) and here.
// mode 2 = fusing 'soft_attn' (:573-574) AS WRITTEN: SoftAttention.forward (:282-303) multiplies x [n, c] by a [n, 1, 1], which
// broadcasts to [n, n, c]; summing dim 1 and then taking .mean(0) leaves (sum_i a_i) * mean_j x_j, and the a_i are
// normalised to sum to 1 - so the module returns the plain mean of its token set and the fused feature is
// mean_in + mean_out (weights 1 / n_in and 1 / n_out); its own parameters receive a mathematically zero gradient.
__device__ __forceinline__ int py_slice_bound(long long v, int size) {
  if (v < 0) { v += size; if (v < 0) v = 0; }
  if (v > size) v = size;
  return static_cast<int>(v);
}
__global__ void __launch_bounds__(256) box_tokens_kernel(const long long* __restrict__ boxes, int frames, int size, int mode,
                                                         uint8_t* __restrict__ inbox, float* __restrict__ weights) {
  __shared__ int cnt_s[8];
  __shared__ int n_in_s;
  const int b = blockIdx.x;
  const int hw = size >> 4, N = (frames >> 1) * hw * hw;
  int cnt = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int t = n / (hw * hw), h = (n / hw) % hw, w = n % hw;
    bool in = false;
#pragma unroll
    for (int p0 = 0; p0 < 2; ++p0) {
      const long long* bb = boxes + (static_cast<size_t>(b) * frames + 2 * t + p0) * 4;
      const int x1 = py_slice_bound(bb[0], size), y1 = py_slice_bound(bb[1], size);
      const int x2 = py_slice_bound(bb[2], size), y2 = py_slice_bound(bb[3], size);
      in = in || (max(16 * h, y1) < min(16 * h + 16, y2) && max(16 * w, x1) < min(16 * w + 16, x2));
    }
    inbox[static_cast<size_t>(b) * N + n] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) cnt_s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { int s = 0; for (int i = 0; i < (blockDim.x >> 5); ++i) s += cnt_s[i]; n_in_s = s; }
  __syncthreads();
  if (weights == nullptr) return;
  const int n_in = n_in_s, n_out = N - n_in;
  const float num_in = mode == 2 ? 1.0f : 0.5f, num_out = mode == 2 ? 1.0f : 0.25f;
  const float w_in = (mode == 0 || n_in == 0) ? 1.0f / N : __fdiv_rn(num_in, static_cast<float>(n_in));
  const float w_out = (mode == 0 || n_in == 0) ? 1.0f / N : (n_out == 0 ? __int_as_float(0x7fc00000) : __fdiv_rn(num_out, static_cast<float>(n_out)));
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const bool in = inbox[static_cast<size_t>(b) * N + n] != 0;
    weights[static_cast<size_t>(b) * N + n] = (n_out == 0 && mode != 0 && n_in != 0) ? __int_as_float(0x7fc00000) : (in ? w_in : w_out);
  }
}

// =================================================================================================
// (7) target + MSE  — engine_for_pretraining.py:258-304
// =================================================================================================
__device__ __forceinline__ float block_sum_128(float v, float* sh) {  // 128 threads; result broadcast
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// one CTA (128 threads) per masked tube.  thread -> (p0, p1, quad of 4 px) for each of the 3 channels.  The 12 label
// values of a thread (4 pixels x 3 channels) are CONTIGUOUS in the reference's feature order f = p*3 + c, so the thread
// reads its 24 bytes of the prediction row and writes its 24 bytes of dpred directly (a warp covers 768 contiguous
// bytes): no shared-memory interleave, two block barriers per tube (the mean and the variance of the three channels).
__global__ void __launch_bounds__(128) target_mse_kernel(const float* __restrict__ video,
                                                         const int32_t* __restrict__ msk_idx,
                                                         const __nv_bfloat16* __restrict__ pred, int n_msk, int frames,
                                                         int size, int normalize_target, float gscale,
                                                         float* __restrict__ loss_partials,
                                                         __nv_bfloat16* __restrict__ dpred,
                                                         float* __restrict__ labels_out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[4];
  __shared__ float sh3[2][4][3];
  const int row = blockIdx.x;
  const int b = row / n_msk;
  const int hw = size >> 4;
  const int tok = min(max(msk_idx[row], 0), (frames >> 1) * hw * hw - 1);   // never index outside the clip
  const int t = tok / (hw * hw), h = (tok / hw) % hw, w = tok % hw;
  const int tid = threadIdx.x;
  const int p0 = tid >> 6, p1 = (tid >> 2) & 15, q = tid & 3;
  const int pbase = p0 * 256 + p1 * 16 + q * 4;         // pixel index p = p0*256 + p1*16 + p2  (:268)
  const size_t foff = static_cast<size_t>(row) * 1536 + static_cast<size_t>(pbase) * 3;   // first of this thread's 12 features
  uint2 pv[3] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
  if (pred) {                                           // fetched together with the pixels: one global-latency phase
#pragma unroll
    for (int j = 0; j < 3; ++j) pv[j] = __ldg(reinterpret_cast<const uint2*>(pred + foff) + j);
  }
  const float mean_c[3] = {0.485f, 0.456f, 0.406f};     // IMAGENET_DEFAULT_MEAN  (:260)
  const float std_c[3] = {0.229f, 0.224f, 0.225f};      // IMAGENET_DEFAULT_STD   (:261)
  float xv[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* src = video + ((static_cast<size_t>(b) * 3 + c) * frames + 2 * t + p0) * size * size +
                       static_cast<size_t>(16 * h + p1) * size + 16 * w + 4 * q;
    float4 v = __ldg(reinterpret_cast<const float4*>(src));
    // videos * std + mean: two separately rounded fp32 ops as in torch (:265)
    xv[c][0] = __fadd_rn(__fmul_rn(v.x, std_c[c]), mean_c[c]);
    xv[c][1] = __fadd_rn(__fmul_rn(v.y, std_c[c]), mean_c[c]);
    xv[c][2] = __fadd_rn(__fmul_rn(v.z, std_c[c]), mean_c[c]);
    xv[c][3] = __fadd_rn(__fmul_rn(v.w, std_c[c]), mean_c[c]);
  }
  // the three channels' statistics go through the block reductions together; per channel the summation tree is
  // warp_sum, then (w0 + w1) + (w2 + w3)
  float mu[3] = {0.f, 0.f, 0.f}, sd[3] = {1.f, 1.f, 1.f};
  if (normalize_target) {
    float s3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) s3[c] = warp_sum((xv[c][0] + xv[c][1]) + (xv[c][2] + xv[c][3]));
    if ((tid & 31) == 0) { sh3[0][tid >> 5][0] = s3[0]; sh3[0][tid >> 5][1] = s3[1]; sh3[0][tid >> 5][2] = s3[2]; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c)
      mu[c] = ((sh3[0][0][c] + sh3[0][1][c]) + (sh3[0][2][c] + sh3[0][3][c])) * (1.0f / 512.0f);   // mean over the 512 pixels (:269)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d0 = xv[c][0] - mu[c], d1 = xv[c][1] - mu[c], d2 = xv[c][2] - mu[c], d3 = xv[c][3] - mu[c];
      s3[c] = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    }
    if ((tid & 31) == 0) { sh3[1][tid >> 5][0] = s3[0]; sh3[1][tid >> 5][1] = s3[1]; sh3[1][tid >> 5][2] = s3[2]; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float var = ((sh3[1][0][c] + sh3[1][1][c]) + (sh3[1][2][c] + sh3[1][3][c])) * (1.0f / 511.0f);   // unbiased (:270)
      // one correctly rounded reciprocal per (tube, channel) instead of 12 IEEE divisions per thread (the kernel was
      // instruction-bound: issue slots 81 % busy, the divisions ~40 % of them); (x - mu) * (1 / sd) is within 1 ulp of
      // the reference's (x - mu) / sd
      sd[c] = __frcp_rn(sqrtf(var) + 1e-6f);
    }
  }
  float l[12];                                           // feature order: l[e*3 + c] = label of pixel pbase+e, channel c  (:276)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int c = 0; c < 3; ++c) l[e * 3 + c] = normalize_target ? (xv[c][e] - mu[c]) * sd[c] : xv[c][e];
  }
  if (labels_out) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      reinterpret_cast<float4*>(labels_out + foff)[j] = make_float4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
  }
  float acc = 0.f;
  if (pred) {
    float d[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      d[4 * j + 0] = bf16_lo(pv[j].x) - l[4 * j + 0]; d[4 * j + 1] = bf16_hi(pv[j].x) - l[4 * j + 1];
      d[4 * j + 2] = bf16_lo(pv[j].y) - l[4 * j + 2]; d[4 * j + 3] = bf16_hi(pv[j].y) - l[4 * j + 3];
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) acc += d[e] * d[e];
    if (dpred) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        uint2 o;
        o.x = pack_bf16(d[4 * j + 0] * gscale, d[4 * j + 1] * gscale);
        o.y = pack_bf16(d[4 * j + 2] * gscale, d[4 * j + 3] * gscale);
        reinterpret_cast<uint2*>(dpred + foff)[j] = o;
      }
    }
  }
  float tot = block_sum_128(acc, sh);
  if (tid == 0 && loss_partials) loss_partials[row] = tot;
}

__global__ void __launch_bounds__(1024) loss_finish_kernel(const float* __restrict__ partials, int n, double inv_count,
                                                           float* __restrict__ loss) {
  pdl_wait();
  pdl_trigger();
  __shared__ double sh[32];
  // per-thread fp32 running sums over 4 independent strided chains (n / 4096 terms each: non-negative, similar
  // magnitude, rel. error ~1e-7), combined in fp64 - fp64 adds run at 1/64 rate, 45 K of them were 20 us on one SM
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = threadIdx.x;
  for (; i + 3 * 1024 < n; i += 4 * 1024) {
    a0 += partials[i]; a1 += partials[i + 1024]; a2 += partials[i + 2048]; a3 += partials[i + 3072];
  }
  for (; i < n; i += 1024) a0 += partials[i];
  double s = (static_cast<double>(a0) + static_cast<double>(a1)) + (static_cast<double>(a2) + static_cast<double>(a3));
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *loss = static_cast<float>(s * inv_count);
  }
}

// =================================================================================================
// (8) helpers
// =================================================================================================
// 32x32 tile transpose through shared memory; block (32,8)
__global__ void cast_weight_kernel(const float* __restrict__ W, int R, int C, __nv_bfloat16* __restrict__ Wb,
                                   __nv_bfloat16* __restrict__ Wt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    float v = (r < R && c < C) ? W[static_cast<size_t>(r) * C + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (Wb && r < R && c < C) Wb[static_cast<size_t>(r) * C + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (Wt) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) Wt[static_cast<size_t>(c) * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
  }
}

__global__ void pack_qkv_bias_kernel(const float* __restrict__ qb, const float* __restrict__ vb, int D,
                                     float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * D) out[i] = i < D ? qb[i] : (i < 2 * D ? 0.f : vb[i - 2 * D]);
}

// grid (ceil(N/256), row chunks); 256 threads = 8 warps; lane owns 8 consecutive columns; 4 row loads in flight.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int M, int N,
                                                          int rows_per_cta, float* __restrict__ out) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (int r = r0 + wid; r < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + 8 * u;
        v[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(X + static_cast<size_t>(rr) * ldx + col)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x); acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
        acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z); acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[wid][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s2 = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) s2 += red[w2][threadIdx.x];
    atomicAdd(out + c, s2);
  }
}

// uint8 NCTHW clip -> ImageNet-normalised f32 (ToTorchFormatTensor(div=True) + GroupNormalize, datasets.py:44-50):
// v = float(x) / 255; out = (v - mean_c) / std_c, with the same two roundings per step as the torch ops.
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                           int64_t n16, int64_t plane16) {
  const float mean_c[3] = {0.485f, 0.456f, 0.406f}, std_c[3] = {0.229f, 0.224f, 0.225f};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / plane16) % 3);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(out) + 4 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 r;
      r.x = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(w[k] & 0xffu), 255.0f), mean_c[c]), std_c[c]);
      r.y = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>((w[k] >> 8) & 0xffu), 255.0f), mean_c[c]), std_c[c]);
      r.z = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>((w[k] >> 16) & 0xffu), 255.0f), mean_c[c]), std_c[c]);
      r.w = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(w[k] >> 24), 255.0f), mean_c[c]), std_c[c]);
      o[k] = r;
    }
  }
}

// =================================================================================================
// (8c) clip preprocessing: crop + bilinear resize + /255 + normalise + NCTHW, and the box transform
// =================================================================================================
// One CTA per frame (b, t).  OpenCV's 8-bit INTER_LINEAR is reproduced bit for bit: taps in float exactly as resize.cpp
// computes them, 11-bit coefficients, int32 horizontal pass, the (>> 4, >> 16, + 2, >> 2) vertical pass.
struct Tap { int idx; int a0; int a1; };
__device__ __forceinline__ Tap linear_tap(int d, int dn, int sn, bool is_y) {
  // separately rounded double operations (no FMA contraction), as the host code of resize.cpp evaluates them
  const double scale = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dn), static_cast<double>(sn)));
  float f = __double2float_rn(__dsub_rn(__dmul_rn(d + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (!is_y) {                                   // columns: taps outside the row are reset; rows are clipped instead
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
  }
  Tap t;
  t.idx = s;
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  return t;
}

__global__ void __launch_bounds__(256) clip_preprocess_kernel(const uint8_t* __restrict__ frames, int T, int H, int W,
                                                              const int32_t* __restrict__ crops,
                                                              const double* __restrict__ boxes_in, int S,
                                                              float* __restrict__ out, double* __restrict__ boxes_out) {
  extern __shared__ int tap_s[];                 // [2][S][3]: x taps then y taps
  const int t = blockIdx.x, b = blockIdx.y;
  const int x_off = crops[b * 4 + 0], y_off = crops[b * 4 + 1], cw = crops[b * 4 + 2], ch = crops[b * 4 + 3];
  for (int i = threadIdx.x; i < 2 * S; i += blockDim.x) {
    const bool is_y = i >= S;
    const Tap tp = linear_tap(is_y ? i - S : i, S, is_y ? ch : cw, is_y);
    tap_s[i * 3 + 0] = tp.idx; tap_s[i * 3 + 1] = tp.a0; tap_s[i * 3 + 2] = tp.a1;
  }
  if (threadIdx.x == 0 && boxes_in != nullptr) {                 // pascal_voc box through Crop + Resize (albumentations semantics)
    const double* bi = boxes_in + (static_cast<size_t>(b) * T + t) * 4;
    double* bo = boxes_out + (static_cast<size_t>(b) * T + t) * 4;
    const double dims[4] = {static_cast<double>(W), static_cast<double>(H), static_cast<double>(W), static_cast<double>(H)};
    const double offs[4] = {static_cast<double>(x_off), static_cast<double>(y_off), static_cast<double>(x_off), static_cast<double>(y_off)};
    const double cdim[4] = {static_cast<double>(cw), static_cast<double>(ch), static_cast<double>(cw), static_cast<double>(ch)};
    double c[4];
    for (int k = 0; k < 4; ++k) {                 // normalise, shift into the crop, re-normalise: separately rounded steps
      const double n = __ddiv_rn(bi[k], dims[k]);
      c[k] = __ddiv_rn(__dsub_rn(__dmul_rn(n, dims[k]), offs[k]), cdim[k]);
      c[k] = fmin(fmax(c[k], 0.0), 1.0);
    }
    if (__dmul_rn(__dsub_rn(c[2], c[0]), __dsub_rn(c[3], c[1])) <= 0.0) { bo[0] = 0.0; bo[1] = 0.0; bo[2] = 1.0; bo[3] = 1.0; }   // transforms.py:120-123
    else { for (int k = 0; k < 4; ++k) bo[k] = __dmul_rn(c[k], static_cast<double>(S)); }
  }
  __syncthreads();
  const uint8_t* src = frames + ((static_cast<size_t>(b) * T + t) * H + y_off) * W * 3 + static_cast<size_t>(x_off) * 3;
  const size_t row_bytes = static_cast<size_t>(W) * 3;
  const float mean_c[3] = {0.485f, 0.456f, 0.406f}, std_c[3] = {0.229f, 0.224f, 0.225f};
  const int groups = S >> 2;                      // 4 output pixels per thread and iteration
  const size_t plane = static_cast<size_t>(S) * S;
  float* dst = out + (static_cast<size_t>(b) * 3 * T + t) * plane;       // channel c plane at + c * T * plane
  for (int g = threadIdx.x; g < S * groups; g += blockDim.x) {
    const int dy = g / groups, dx0 = (g % groups) * 4;
    const int sy = tap_s[(S + dy) * 3], b0 = tap_s[(S + dy) * 3 + 1], b1 = tap_s[(S + dy) * 3 + 2];
    const uint8_t* r0 = src + static_cast<size_t>(min(max(sy, 0), ch - 1)) * row_bytes;
    const uint8_t* r1 = src + static_cast<size_t>(min(max(sy + 1, 0), ch - 1)) * row_bytes;
    float v[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int sx = tap_s[(dx0 + e) * 3], a0 = tap_s[(dx0 + e) * 3 + 1], a1 = tap_s[(dx0 + e) * 3 + 2];
      const int sx1 = min(sx + 1, cw - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int S0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
        const int S1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        int px = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        px = min(max(px, 0), 255);
        v[c][e] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(px), 255.0f), mean_c[c]), std_c[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(c) * T * plane + static_cast<size_t>(dy) * S + dx0) =
          make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

__global__ void __launch_bounds__(256) sq_norm_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float sh[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { float v = x[n4 * 4 + threadIdx.x]; s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

}  // namespace mofo

// =================================================================================================
// C ABI
// =================================================================================================
using namespace mofo;

static int mask_common(bool bb, const double* bb_first, const uint32_t* rng_words, int B, int W, int T, int H, int Wd,
                       int nmask, double ratio_bb, uint8_t* mask, int32_t* vis_idx, int32_t* msk_idx,
                       int32_t* words_used, void* stream) {
  MOFO_CHECK_ARG(B > 0 && W > 0 && T > 0 && H > 0 && Wd > 0, "tube_mask: non-positive size");
  MOFO_CHECK_ARG(H * Wd <= 4096 && nmask >= 0 && nmask <= H * Wd, "tube_mask: grid %dx%d / n_mask %d unsupported", H, Wd, nmask);
  MOFO_CHECK_ARG(rng_words && mask && vis_idx && msk_idx && words_used && (!bb || bb_first), "tube_mask: null pointer");
  const int warps = 1;      // one clip per CTA: the per-clip chain is sequential, so clips spread over SMs
  const int Wsm = W < 2048 ? W : 2048;                                    // staged words per clip
  size_t smem = (static_cast<size_t>(warps) * 3 * H * Wd + 1) * sizeof(int16_t) + static_cast<size_t>(warps) * Wsm * sizeof(uint32_t) + 4;
  dim3 grid((B + warps - 1) / warps), block(warps * 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bb) {
    if (smem > 48 * 1024) MOFO_CUDA(cudaFuncSetAttribute(tube_mask_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tube_mask_kernel<true><<<grid, block, smem, s>>>(bb_first, rng_words, B, W, Wsm, T, H, Wd, nmask, ratio_bb, mask, vis_idx, msk_idx, words_used);
  } else {
    if (smem > 48 * 1024) MOFO_CUDA(cudaFuncSetAttribute(tube_mask_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tube_mask_kernel<false><<<grid, block, smem, s>>>(nullptr, rng_words, B, W, Wsm, T, H, Wd, nmask, 0.0, mask, vis_idx, msk_idx, words_used);
  }
  MOFO_LAUNCH_CHECK("tube_mask_kernel");
  return MOFO_OK;
}

extern "C" {

int mofo_tube_mask_bb(const double* bb_first, const uint32_t* rng_words, int B, int W, int T, int H, int Wd,
                      int n_mask_per_frame, double ratio_bb, uint8_t* mask, int32_t* vis_idx, int32_t* msk_idx,
                      int32_t* words_used, void* stream) {
  return mask_common(true, bb_first, rng_words, B, W, T, H, Wd, n_mask_per_frame, ratio_bb, mask, vis_idx, msk_idx, words_used, stream);
}

int mofo_tube_mask_plain(const uint32_t* rng_words, int B, int W, int T, int H, int Wd, int n_mask_per_frame,
                         uint8_t* mask, int32_t* vis_idx, int32_t* msk_idx, int32_t* words_used, void* stream) {
  return mask_common(false, nullptr, rng_words, B, W, T, H, Wd, n_mask_per_frame, 0.0, mask, vis_idx, msk_idx, words_used, stream);
}

int mofo_mask_indices(const uint8_t* mask, int B, int N, int n_msk, int32_t* vis_idx, int32_t* msk_idx,
                      int32_t* bad_rows, void* stream) {
  MOFO_CHECK_ARG(mask && vis_idx && msk_idx && bad_rows, "mask_indices: null pointer");
  MOFO_CHECK_ARG(B > 0 && N > 0 && n_msk >= 0 && n_msk <= N, "mask_indices: bad shape B=%d N=%d n_msk=%d", B, N, n_msk);
  mask_indices_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(mask, B, N, n_msk, vis_idx, msk_idx, bad_rows);
  MOFO_LAUNCH_CHECK("mask_indices_kernel");
  return MOFO_OK;
}

int mofo_gather_tubes(const float* video, const int32_t* idx, int B, int n_idx, int frames, int size, mofo_bf16* A,
                      void* stream) {
  MOFO_CHECK_ARG(video && idx && A, "gather_tubes: null pointer");
  MOFO_CHECK_ARG(B > 0 && n_idx > 0 && frames % 2 == 0 && size % 16 == 0, "gather_tubes: bad shape B=%d n=%d frames=%d size=%d", B, n_idx, frames, size);
  MOFO_CUDA(launch_pdl(gather_tubes_kernel, dim3(B * n_idx), dim3(192), 0, static_cast<cudaStream_t>(stream), video, idx, n_idx, frames,
                       size, reinterpret_cast<__nv_bfloat16*>(A)));
  return MOFO_OK;
}

int mofo_layernorm_fwd(const float* x, const float* gamma, const float* beta, int M, int D, float eps, int group_rows,
                       int in_group_rows, int in_row_offset, mofo_bf16* y, float* mean, float* rstd, void* stream) {
  MOFO_CHECK_ARG(x && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
  MOFO_CHECK_ARG(M > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAX_CHUNKS && group_rows > 0, "layernorm_fwd: unsupported M=%d D=%d", M, D);
  const int nch = (D + 127) / 128;
#define MOFO_LN_FWD(NCH)                                                                                              \
  MOFO_CUDA(launch_pdl(layernorm_fwd_kernel<NCH>, dim3((M + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), x, gamma, \
                       beta, M, D, eps, group_rows, in_group_rows, in_row_offset, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd))
  if (nch <= 1) MOFO_LN_FWD(1);
  else if (nch == 2) MOFO_LN_FWD(2);
  else if (nch == 3) MOFO_LN_FWD(3);
  else if (nch == 4) MOFO_LN_FWD(4);
  else if (nch <= 6) MOFO_LN_FWD(6);
  else MOFO_LN_FWD(8);
#undef MOFO_LN_FWD
  MOFO_LAUNCH_CHECK("layernorm_fwd_kernel");
  return MOFO_OK;
}

int mofo_layernorm_bwd(const mofo_bf16* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* dres, int M, int D, int group_rows, int in_group_rows, int in_row_offset,
                       float* dx_f32, mofo_bf16* dx_bf16, float* dgamma, float* dbeta, const float* bf16_row_scale,
                       void* stream) {
  MOFO_CHECK_ARG(dy && x && gamma && mean && rstd && dgamma && dbeta, "layernorm_bwd: null pointer");
  MOFO_CHECK_ARG(((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0,
                 "layernorm_bwd: dgamma / dbeta must be 16-byte aligned");
  MOFO_CHECK_ARG(M > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAX_CHUNKS && group_rows > 0, "layernorm_bwd: unsupported M=%d D=%d", M, D);
  const int split = D > 384 ? 2 : 1;                          // warps per row
  MOFO_CHECK_ARG((D >> 2) % split == 0, "layernorm_bwd: D=%d", D);
  const int slots = 8 / split;                                 // rows in flight per CTA (8 warps)
  int grid = (M + slots - 1) / slots;
  static const int cap_mul = [] { const char* e = getenv("MOFO_LN_CAP"); return e ? atoi(e) : 0; }();   // tuning aid
  int cap = sm_count() * (cap_mul > 0 ? cap_mul : 2);          // 2 CTAs / SM are resident (128 registers): one wave
  if (grid > cap) grid = cap;
  size_t smem = static_cast<size_t>(slots) * 2 * D * sizeof(float);
  const int nch = (D / split + 127) / 128;                     // float4 chunks per lane
#define MOFO_LN_BWD(NCH, SPLIT)                                                                                        \
  do {                                                                                                                 \
    if (smem + 1024 > 48 * 1024)                                                                                       \
      MOFO_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<NCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    MOFO_CUDA(launch_pdl(layernorm_bwd_kernel<NCH, SPLIT>, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), \
                         reinterpret_cast<const __nv_bfloat16*>(dy), x, gamma, mean, rstd, dres, M, D, group_rows,     \
                         in_group_rows, in_row_offset, dx_f32, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma,      \
                         dbeta, bf16_row_scale));                                                                      \
  } while (0)
  if (split == 1) {
    if (nch <= 1) MOFO_LN_BWD(1, 1);
    else if (nch == 2) MOFO_LN_BWD(2, 1);
    else MOFO_LN_BWD(3, 1);
  } else {
    if (nch <= 3) MOFO_LN_BWD(3, 2);
    else if (nch == 4) MOFO_LN_BWD(4, 2);
    else if (nch <= 6) MOFO_LN_BWD(6, 2);
    else MOFO_LN_BWD(8, 2);
  }
#undef MOFO_LN_BWD
  MOFO_LAUNCH_CHECK("layernorm_bwd_kernel");
  return MOFO_OK;
}

int mofo_decoder_assemble_fwd(const float* mask_token, const float* pos, const int32_t* msk_idx, int B, int n_vis,
                              int n_msk, int Dd, float* x_full, void* stream) {
  MOFO_CHECK_ARG(mask_token && pos && msk_idx && x_full, "decoder_assemble_fwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && n_vis >= 0 && n_msk > 0 && Dd % 4 == 0, "decoder_assemble_fwd: bad shape");
  const int rows = B * n_msk;
  int grid = (rows + 7) / 8;
  if (grid > 8 * sm_count()) grid = 8 * sm_count();
  MOFO_CUDA(launch_pdl(assemble_fwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), mask_token, pos, msk_idx,
                       rows, n_vis, n_msk, Dd, x_full));
  return MOFO_OK;
}

int mofo_decoder_assemble_bwd(const float* dx_full, int B, int n_vis, int n_msk, int Dd, float* dmask_token,
                              mofo_bf16* dvis, void* stream) {
  MOFO_CHECK_ARG(dx_full && dmask_token && dvis, "decoder_assemble_bwd: null pointer");
  MOFO_CHECK_ARG((reinterpret_cast<uintptr_t>(dmask_token) & 15) == 0, "decoder_assemble_bwd: dmask_token must be 16-byte aligned");
  MOFO_CHECK_ARG(B > 0 && n_vis >= 0 && n_msk > 0 && Dd % 4 == 0, "decoder_assemble_bwd: bad shape");
  const int N = n_vis + n_msk;
  int chunks = (2 * sm_count() + B - 1) / B;                   // ~2 CTAs per SM over the whole batch
  if (chunks < 1) chunks = 1;
  int rows_per_cta = (N + chunks - 1) / chunks;
  if (rows_per_cta < 4 * ASM_RG) rows_per_cta = 4 * ASM_RG;
  dim3 grid((N + rows_per_cta - 1) / rows_per_cta, B);
  const int cx = ((Dd >> 2) + 31) / 32 * 32;
  MOFO_CHECK_ARG(cx * ASM_RG <= 1024, "decoder_assemble_bwd: Dd=%d too wide", Dd);
  const size_t smem = static_cast<size_t>(ASM_RG) * (Dd >> 2) * sizeof(float4);
  MOFO_CUDA(launch_pdl(assemble_bwd_kernel, grid, dim3(cx, ASM_RG), smem, static_cast<cudaStream_t>(stream), dx_full, n_vis, n_msk, Dd,
                       rows_per_cta, dmask_token, reinterpret_cast<__nv_bfloat16*>(dvis)));
  return MOFO_OK;
}

int mofo_box_tokens(const int64_t* boxes, int B, int frames, int size, int mode, uint8_t* inbox, float* weights, void* stream) {
  MOFO_CHECK_ARG(boxes && inbox, "box_tokens: null pointer");
  MOFO_CHECK_ARG(B > 0 && frames > 0 && frames % 2 == 0 && size > 0 && size % 16 == 0 && mode >= 0 && mode <= 2, "box_tokens: bad argument");
  box_tokens_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(boxes), frames, size, mode, inbox, weights);
  MOFO_LAUNCH_CHECK("box_tokens_kernel");
  return MOFO_OK;
}

int mofo_zero_rows(float* x_f32, mofo_bf16* x_bf16, int groups, int group_rows, int n_zero, int D, void* stream) {
  MOFO_CHECK_ARG(x_f32 || x_bf16, "zero_rows: null pointer");
  MOFO_CHECK_ARG(groups > 0 && group_rows > 0 && n_zero >= 0 && n_zero <= group_rows && D > 0 && D % 4 == 0, "zero_rows: bad shape");
  if (n_zero == 0) return MOFO_OK;
  const int64_t total4 = static_cast<int64_t>(groups) * n_zero * (D >> 2);
  int64_t blocks = (total4 + 255) / 256;
  if (blocks > 8L * sm_count()) blocks = 8L * sm_count();
  MOFO_CUDA(launch_pdl(zero_rows_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), x_f32,
                       reinterpret_cast<__nv_bfloat16*>(x_bf16), group_rows, n_zero, D, total4));
  return MOFO_OK;
}

int mofo_token_mean_fwd(const float* x, const float* weights, int B, int N, int D, float* pooled, void* stream) {
  MOFO_CHECK_ARG(x && pooled, "token_mean_fwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && D > 0 && D % 4 == 0 && (reinterpret_cast<uintptr_t>(pooled) & 15) == 0,
                 "token_mean_fwd: bad shape B=%d N=%d D=%d", B, N, D);
  dim3 grid(((D >> 2) + 31) / 32, B);
  MOFO_CUDA(launch_pdl(token_mean_fwd_kernel, grid, dim3(32, TM_RG), 0, static_cast<cudaStream_t>(stream), x, weights, N, D, 1.0f / N, pooled));
  return MOFO_OK;
}

int mofo_token_mean_bwd(const float* dpooled, const float* weights, int B, int N, int D, float* dx_f32, mofo_bf16* dx_bf16,
                        const float* bf16_row_scale, void* stream) {
  MOFO_CHECK_ARG(dpooled && (dx_f32 || dx_bf16), "token_mean_bwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && N > 0 && D > 0 && D % 4 == 0, "token_mean_bwd: bad shape B=%d N=%d D=%d", B, N, D);
  const int64_t total4 = static_cast<int64_t>(B) * N * (D >> 2);
  int64_t blocks = (total4 + 255) / 256;
  if (blocks > 16L * sm_count()) blocks = 16L * sm_count();
  MOFO_CUDA(launch_pdl(token_mean_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                       dpooled, weights, N, D, 1.0f / N, total4, dx_f32, reinterpret_cast<__nv_bfloat16*>(dx_bf16), bf16_row_scale));
  return MOFO_OK;
}

int mofo_masked_softmax_fwd(const float* S, const uint8_t* key_allowed, int B, int rows_per_b, int Nk, float scale, mofo_bf16* P,
                            void* stream) {
  MOFO_CHECK_ARG(S && key_allowed && P, "masked_softmax_fwd: null pointer");
  MOFO_CHECK_ARG(B > 0 && rows_per_b > 0 && Nk > 0 && Nk % 4 == 0 && Nk <= 4 * MS_THREADS * MS_CHUNKS &&
                 static_cast<int64_t>(B) * rows_per_b < (1LL << 31) && (reinterpret_cast<uintptr_t>(S) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(P) & 7) == 0 && (reinterpret_cast<uintptr_t>(key_allowed) & 3) == 0,
                 "masked_softmax_fwd: bad shape B=%d rows=%d Nk=%d (Nk %% 4 == 0, <= %d)", B, rows_per_b, Nk, 4 * MS_THREADS * MS_CHUNKS);
  MOFO_CUDA(launch_pdl(masked_softmax_fwd_kernel, dim3(static_cast<unsigned>(B) * rows_per_b), dim3(MS_THREADS), 0,
                       static_cast<cudaStream_t>(stream), S, key_allowed, rows_per_b, Nk, scale, reinterpret_cast<__nv_bfloat16*>(P)));
  return MOFO_OK;
}

int mofo_masked_softmax_bwd(const mofo_bf16* P, const float* dP, int64_t rows, int Nk, float scale, mofo_bf16* dS, void* stream) {
  MOFO_CHECK_ARG(P && dP && dS, "masked_softmax_bwd: null pointer");
  MOFO_CHECK_ARG(rows > 0 && rows < (1LL << 31) && Nk > 0 && Nk % 4 == 0 && Nk <= 4 * MS_THREADS * MS_CHUNKS &&
                 (reinterpret_cast<uintptr_t>(dP) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(dS) & 7) == 0, "masked_softmax_bwd: bad shape rows=%lld Nk=%d", static_cast<long long>(rows), Nk);
  MOFO_CUDA(launch_pdl(masked_softmax_bwd_kernel, dim3(static_cast<unsigned>(rows)), dim3(MS_THREADS), 0, static_cast<cudaStream_t>(stream),
                       reinterpret_cast<const __nv_bfloat16*>(P), dP, Nk, scale, reinterpret_cast<__nv_bfloat16*>(dS)));
  return MOFO_OK;
}

int mofo_cast_f32_bf16(const float* src, int lds, int M, int N, mofo_bf16* dst, int ldd, void* stream) {
  MOFO_CHECK_ARG(src && dst && M > 0 && N > 0 && N % 4 == 0 && lds >= N && ldd >= N && lds % 4 == 0 && ldd % 4 == 0 &&
                 (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "cast_f32_bf16: bad argument");
  const int64_t total4 = static_cast<int64_t>(M) * (N >> 2);
  int64_t blocks = (total4 + 255) / 256;
  if (blocks > 16L * sm_count()) blocks = 16L * sm_count();
  MOFO_CUDA(launch_pdl(cast_f32_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, lds,
                       total4, N >> 2, reinterpret_cast<__nv_bfloat16*>(dst), ldd));
  return MOFO_OK;
}

int mofo_target_mse(const float* video, const int32_t* msk_idx, const mofo_bf16* pred, int B, int n_msk, int frames,
                    int size, int normalize_target, float grad_scale, float* loss_partials, float* loss,
                    mofo_bf16* dpred, float* labels_out, void* stream) {
  MOFO_CHECK_ARG(video && msk_idx, "target_mse: null pointer");
  MOFO_CHECK_ARG(B > 0 && n_msk > 0 && frames % 2 == 0 && size % 16 == 0, "target_mse: bad shape");
  MOFO_CHECK_ARG(!loss || (loss_partials && pred), "target_mse: loss needs pred and loss_partials");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const double count = static_cast<double>(B) * n_msk * 1536.0;
  const float gscale = static_cast<float>(2.0 / count) * grad_scale;
  MOFO_CUDA(launch_pdl(target_mse_kernel, dim3(B * n_msk), dim3(128), 0, s, video, msk_idx, reinterpret_cast<const __nv_bfloat16*>(pred),
                       n_msk, frames, size, normalize_target, gscale, loss_partials, reinterpret_cast<__nv_bfloat16*>(dpred), labels_out));
  if (loss) MOFO_CUDA(launch_pdl(loss_finish_kernel, dim3(1), dim3(1024), 0, s, static_cast<const float*>(loss_partials), B * n_msk, 1.0 / count, loss));
  return MOFO_OK;
}

int mofo_cast_weight(const float* W, int R, int C, mofo_bf16* W_bf16, mofo_bf16* Wt_bf16, void* stream) {
  MOFO_CHECK_ARG(W && (W_bf16 || Wt_bf16) && R > 0 && C > 0, "cast_weight: bad argument");
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  cast_weight_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(W, R, C, reinterpret_cast<__nv_bfloat16*>(W_bf16),
                                                                          reinterpret_cast<__nv_bfloat16*>(Wt_bf16));
  MOFO_LAUNCH_CHECK("cast_weight_kernel");
  return MOFO_OK;
}

int mofo_pack_qkv_bias(const float* q_bias, const float* v_bias, int D, float* out, void* stream) {
  MOFO_CHECK_ARG(q_bias && v_bias && out && D > 0, "pack_qkv_bias: bad argument");
  pack_qkv_bias_kernel<<<(3 * D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(q_bias, v_bias, D, out);
  MOFO_LAUNCH_CHECK("pack_qkv_bias_kernel");
  return MOFO_OK;
}

int mofo_colsum_bf16(const mofo_bf16* X, int ldx, int M, int N, float* out, void* stream) {
  MOFO_CHECK_ARG(X && out && M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0, "colsum_bf16: bad argument (N, ldx must be multiples of 8)");
  int rows_per_cta = 128;
  dim3 grid((N + 255) / 256, (M + rows_per_cta - 1) / rows_per_cta);
  colsum_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(X), ldx, M, N, rows_per_cta, out);
  MOFO_LAUNCH_CHECK("colsum_bf16_kernel");
  return MOFO_OK;
}

int mofo_normalize_u8(const uint8_t* clip_u8, int B, int frames, int size, float* out, void* stream) {
  MOFO_CHECK_ARG(clip_u8 && out && B > 0 && frames > 0 && size > 0, "normalize_u8: bad argument");
  const int64_t plane = static_cast<int64_t>(frames) * size * size;
  MOFO_CHECK_ARG(plane % 16 == 0 && (reinterpret_cast<uintptr_t>(clip_u8) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "normalize_u8: frames*size*size must be a multiple of 16 and buffers 16-byte aligned");
  const int64_t n16 = static_cast<int64_t>(B) * 3 * plane / 16;
  normalize_u8_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(clip_u8, out, n16, plane / 16);
  MOFO_LAUNCH_CHECK("normalize_u8_kernel");
  return MOFO_OK;
}

int mofo_clip_preprocess(const uint8_t* frames, int B, int T, int H, int W, const int32_t* crops, const double* boxes_in,
                         int out_size, float* clip_out, double* boxes_out, void* stream) {
  MOFO_CHECK_ARG(frames && crops && clip_out && (!boxes_in || boxes_out), "clip_preprocess: null pointer");
  MOFO_CHECK_ARG(B > 0 && T > 0 && H > 1 && W > 1 && out_size > 0 && out_size % 4 == 0 && out_size <= 1024 && B <= 65535,
                 "clip_preprocess: bad shape B=%d T=%d H=%d W=%d out=%d", B, T, H, W, out_size);
  MOFO_CHECK_ARG((reinterpret_cast<uintptr_t>(clip_out) & 15) == 0, "clip_preprocess: clip_out must be 16-byte aligned");
  const size_t smem = static_cast<size_t>(2) * out_size * 3 * sizeof(int);
  clip_preprocess_kernel<<<dim3(T, B), 256, smem, static_cast<cudaStream_t>(stream)>>>(frames, T, H, W, crops, boxes_in, out_size,
                                                                                      clip_out, boxes_out);
  MOFO_LAUNCH_CHECK("clip_preprocess_kernel");
  return MOFO_OK;
}

int mofo_sq_norm_f32(const float* x, int64_t n, float* out, void* stream) {
  MOFO_CHECK_ARG(x && out && n > 0, "sq_norm_f32: bad argument");
  MOFO_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sq_norm_f32: x must be 16-byte aligned");
  int grid = sm_count() * 4;
  sq_norm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  MOFO_LAUNCH_CHECK("sq_norm_kernel");
  return MOFO_OK;
}

}  // extern "C"
