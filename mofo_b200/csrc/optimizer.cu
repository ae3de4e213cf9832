// Fused AdamW over the flat parameter / gradient / moment arenas, emitting the bf16 operand copies (W and W^T) that
// the next step's GEMMs read.  One launch replaces torch's multi-tensor AdamW (optim_factory.py:126-127 creates
// optim.AdamW; utils.py:355-364 drives unscale / clip / step), the ~67 per-step weight casts and the qkv-bias packing.
//
// Work decomposition: a table of tiles.  A tile is either a 32 x 128 block of a 2-D weight (so that the transposed bf16
// copy is written coalesced through a shared-memory transpose) or a run of 4096 consecutive elements of any other
// tensor.  Every thread moves 16-byte vectors and keeps 16 of them in flight (4 rows x {p, g, m, v}).  HBM-bound: 16 B
// read + 12 B written per parameter in fp32, + 4 B for the two bf16 copies.  Optionally the squared gradient norm
// (utils.py:376-388) is accumulated in the same pass, which saves the separate read of the gradient arena.
#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

constexpr int AD_TR = 32, AD_TC = 128, AD_RUN = 4096;

// segs: int64 [n_seg][6] = {arena offset, rows, cols, group, w16 offset (-1: none), wt16 offset (-1: none)}
// tiles: int32 [n_tiles][2] = {segment, tile index within the segment}
// hyper: float [8 + 2*groups] = {beta1, beta2, eps, bias_correction1, sqrt(bias_correction2), -, -, -, lr_0, wd_0, lr_1, ...}
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                    float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                    __nv_bfloat16* __restrict__ w16, const int64_t* __restrict__ segs,
                                                    const int32_t* __restrict__ tiles, const float* __restrict__ hyper,
                                                    const float* __restrict__ clip_coef,
                                                    const float* __restrict__ loss_guard, float* __restrict__ sq_norm_out) {
  __shared__ float tile_s[AD_TR][AD_TC + 1];
  __shared__ float red_s[8];
  if (loss_guard != nullptr) {
    const float l = *loss_guard;
    if (!(fabsf(l) <= 3.0e38f)) return;          // NaN / Inf loss: leave parameters and moments untouched
  }
  const int seg = tiles[2 * blockIdx.x], t = tiles[2 * blockIdx.x + 1];
  const int64_t off = segs[seg * 6 + 0];
  const int rows = static_cast<int>(segs[seg * 6 + 1]), cols = static_cast<int>(segs[seg * 6 + 2]);
  const int group = static_cast<int>(segs[seg * 6 + 3]);
  const int64_t w16_off = segs[seg * 6 + 4], wt16_off = segs[seg * 6 + 5];
  const float beta1 = hyper[0], beta2 = hyper[1], eps = hyper[2], bc1 = hyper[3], bc2_sqrt = hyper[4];
  const float lr = hyper[8 + 2 * group], wd = hyper[9 + 2 * group];
  const float gscale = clip_coef ? *clip_coef : 1.0f;
  const float decay = 1.0f - lr * wd, step_size = lr / bc1;
  float gsq = 0.f;

  auto update1 = [&](float& p, float g, float& m, float& v) {
    gsq = fmaf(g, g, gsq);
    g *= gscale;
    p = p * decay;                                            // param.mul_(1 - lr * weight_decay)
    m = m + (g - m) * (1.0f - beta1);                         // exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (1.0f - beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);                          // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  auto update4 = [&](float4& p, const float4& g, float4& m, float4& v) {
    update1(p.x, g.x, m.x, v.x); update1(p.y, g.y, m.y, v.y); update1(p.z, g.z, m.z, v.z); update1(p.w, g.w, m.w, v.w);
  };

  if (wt16_off >= 0) {                                         // 32 x 128 tile of a 2-D weight: bf16 W and W^T
    const int tiles_c = (cols + AD_TC - 1) / AD_TC;
    const int r0 = (t / tiles_c) * AD_TR, c0 = (t % tiles_c) * AD_TC;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = c0 + 4 * tx;
    const bool vec = (cols & 3) == 0;                          // rows start 16-byte aligned (arena offsets are)
    float4 p4[4], g4[4], m4[4], v4[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k;
      ok[k] = r < rows && c < cols;
      if (ok[k] && vec) {
        const int64_t i = off + static_cast<int64_t>(r) * cols + c;
        p4[k] = *reinterpret_cast<const float4*>(params + i);
        g4[k] = *reinterpret_cast<const float4*>(grads + i);
        m4[k] = *reinterpret_cast<const float4*>(exp_avg + i);
        v4[k] = *reinterpret_cast<const float4*>(exp_avg_sq + i);
      } else if (ok[k]) {                                      // ragged width: scalar accesses, zero padding
        const int64_t i = off + static_cast<int64_t>(r) * cols + c;
        float* pp = &p4[k].x; float* gg = &g4[k].x; float* mm = &m4[k].x; float* vv = &v4[k].x;
        for (int e = 0; e < 4; ++e) {
          const bool in = c + e < cols;
          pp[e] = in ? params[i + e] : 0.f; gg[e] = in ? grads[i + e] : 0.f;
          mm[e] = in ? exp_avg[i + e] : 0.f; vv[e] = in ? exp_avg_sq[i + e] : 1.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rl = ty + 8 * k;
      if (ok[k]) {
        update4(p4[k], g4[k], m4[k], v4[k]);
        const int64_t i = off + static_cast<int64_t>(r0 + rl) * cols + c;
        if (vec) {
          *reinterpret_cast<float4*>(params + i) = p4[k];
          *reinterpret_cast<float4*>(exp_avg + i) = m4[k];
          *reinterpret_cast<float4*>(exp_avg_sq + i) = v4[k];
          uint2 w; w.x = pack_bf16(p4[k].x, p4[k].y); w.y = pack_bf16(p4[k].z, p4[k].w);
          *reinterpret_cast<uint2*>(w16 + w16_off + static_cast<int64_t>(r0 + rl) * cols + c) = w;
        } else {
          const float* pp = &p4[k].x; const float* mm = &m4[k].x; const float* vv = &v4[k].x;
          for (int e = 0; e < 4 && c + e < cols; ++e) {
            params[i + e] = pp[e]; exp_avg[i + e] = mm[e]; exp_avg_sq[i + e] = vv[e];
            w16[w16_off + static_cast<int64_t>(r0 + rl) * cols + c + e] = __float2bfloat16_rn(pp[e]);
          }
        }
        tile_s[rl][4 * tx + 0] = p4[k].x; tile_s[rl][4 * tx + 1] = p4[k].y;
        tile_s[rl][4 * tx + 2] = p4[k].z; tile_s[rl][4 * tx + 3] = p4[k].w;
      }
    }
    __syncthreads();
    // W^T: output row = column c0 + cl of the tile, 32 consecutive r values (64 B); a thread writes one bf16 pair
    const bool pair_ok = (rows & 1) == 0;
#pragma unroll
    for (int it = 0; it < (AD_TR / 2) * AD_TC / 256; ++it) {
      const int o = threadIdx.x + 256 * it;
      const int cl = o >> 4, rp = (o & 15) * 2;
      const int cc = c0 + cl, r = r0 + rp;
      if (cc < cols && r < rows) {
        __nv_bfloat16* dst = w16 + wt16_off + static_cast<int64_t>(cc) * rows + r;
        if (pair_ok && r + 1 < rows) {
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16(tile_s[rp][cl], tile_s[rp + 1][cl]);
        } else {
          dst[0] = __float2bfloat16_rn(tile_s[rp][cl]);
          if (r + 1 < rows) dst[1] = __float2bfloat16_rn(tile_s[rp + 1][cl]);
        }
      }
    }
  } else {                                                     // 4096 consecutive elements
    const int64_t n = static_cast<int64_t>(rows) * cols;
    const int64_t base = static_cast<int64_t>(t) * AD_RUN;
    float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = base + (k * 256 + threadIdx.x) * 4;
      if (j + 3 < n) {
        p4[k] = *reinterpret_cast<const float4*>(params + off + j);
        g4[k] = *reinterpret_cast<const float4*>(grads + off + j);
        m4[k] = *reinterpret_cast<const float4*>(exp_avg + off + j);
        v4[k] = *reinterpret_cast<const float4*>(exp_avg_sq + off + j);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = base + (k * 256 + threadIdx.x) * 4;
      if (j + 3 < n) {
        update4(p4[k], g4[k], m4[k], v4[k]);
        *reinterpret_cast<float4*>(params + off + j) = p4[k];
        *reinterpret_cast<float4*>(exp_avg + off + j) = m4[k];
        *reinterpret_cast<float4*>(exp_avg_sq + off + j) = v4[k];
        if (w16_off >= 0) {
          uint2 w; w.x = pack_bf16(p4[k].x, p4[k].y); w.y = pack_bf16(p4[k].z, p4[k].w);
          *reinterpret_cast<uint2*>(w16 + w16_off + j) = w;
        }
      } else if (j < n) {                                      // ragged tail of the tensor
        for (int64_t i = j; i < n; ++i) {
          float p = params[off + i], m = exp_avg[off + i], v = exp_avg_sq[off + i];
          update1(p, grads[off + i], m, v);
          params[off + i] = p; exp_avg[off + i] = m; exp_avg_sq[off + i] = v;
          if (w16_off >= 0) w16[w16_off + i] = __float2bfloat16_rn(p);
        }
      }
    }
  }
  if (sq_norm_out != nullptr) {                                // block sum of g^2 -> one atomic per CTA
    gsq = warp_sum(gsq);
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = gsq;
    __syncthreads();
    if (threadIdx.x < 32) {
      float s2 = threadIdx.x < 8 ? red_s[threadIdx.x] : 0.f;
      s2 = warp_sum(s2);
      if (threadIdx.x == 0) atomicAdd(sq_norm_out, s2);
    }
  }
}

}  // namespace mofo

using namespace mofo;

extern "C" {

int mofo_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, mofo_bf16* w16,
                    const int64_t* segs, const int32_t* tiles, int n_tiles, const float* hyper, const float* clip_coef,
                    const float* loss_guard, float* sq_norm_out, void* stream) {
  MOFO_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && segs && tiles && hyper, "adamw_step: null pointer");
  MOFO_CHECK_ARG(n_tiles > 0, "adamw_step: empty tile table");
  adamw_kernel<<<n_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq,
                                                                       reinterpret_cast<__nv_bfloat16*>(w16), segs, tiles,
                                                                       hyper, clip_coef, loss_guard, sq_norm_out);
  MOFO_LAUNCH_CHECK("adamw_kernel");
  return MOFO_OK;
}

}  // extern "C"
