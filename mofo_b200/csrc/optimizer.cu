// Fused AdamW over the flat parameter / gradient / moment arenas, emitting the bf16 operand copies (W and W^T) that
// the next step's GEMMs read.  One launch replaces torch's multi-tensor AdamW (optim_factory.py:126-127 creates
// optim.AdamW; utils.py:355-364 drives unscale / clip / step), the ~67 per-step weight casts and the qkv-bias packing.
//
// Work decomposition: a table of tiles.  A tile is either a 32x32 block of a 2-D weight (so that the transposed bf16
// copy is written coalesced through a shared-memory transpose) or a run of 1024 consecutive elements of any other
// tensor.  HBM-bound: 16 B read + 12 B written per parameter in fp32, + 4 B for the two bf16 copies.
#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

// segs: int64 [n_seg][6] = {arena offset, rows, cols, group, w16 offset (-1: none), wt16 offset (-1: none)}
// tiles: int32 [n_tiles][2] = {segment, tile index within the segment}
// hyper: float [8 + 2*groups] = {beta1, beta2, eps, bias_correction1, sqrt(bias_correction2), -, -, -, lr_0, wd_0, lr_1, ...}
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                    float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                    __nv_bfloat16* __restrict__ w16, const int64_t* __restrict__ segs,
                                                    const int32_t* __restrict__ tiles, const float* __restrict__ hyper,
                                                    const float* __restrict__ clip_coef,
                                                    const float* __restrict__ loss_guard) {
  __shared__ float tile_s[32][33];
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

  auto update = [&](int64_t i) -> float {
    const float g = grads[i] * gscale;
    float p = params[i] * decay;                              // param.mul_(1 - lr * weight_decay)
    float m = exp_avg[i];
    m = m + (g - m) * (1.0f - beta1);                         // exp_avg.lerp_(grad, 1 - beta1)
    const float v = exp_avg_sq[i] * beta2 + (1.0f - beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);                          // param.addcdiv_(exp_avg, denom, value=-step_size)
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    params[i] = p;
    return p;
  };

  if (wt16_off >= 0) {                                         // 32x32 tile of a 2-D weight: bf16 W and W^T
    const int tiles_c = (cols + 31) / 32;
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      float p = 0.f;
      if (r < rows && c < cols) {
        p = update(off + static_cast<int64_t>(r) * cols + c);
        w16[w16_off + static_cast<int64_t>(r) * cols + c] = __float2bfloat16_rn(p);
      }
      tile_s[i][tx] = p;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (r < rows && c < cols) w16[wt16_off + static_cast<int64_t>(c) * rows + r] = __float2bfloat16_rn(tile_s[tx][i]);
    }
  } else {                                                     // 1024 consecutive elements
    const int64_t n = static_cast<int64_t>(rows) * cols;
    const int64_t base = static_cast<int64_t>(t) * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = base + k * 256 + threadIdx.x;
      if (i < n) {
        const float p = update(off + i);
        if (w16_off >= 0) w16[w16_off + i] = __float2bfloat16_rn(p);
      }
    }
  }
}

}  // namespace mofo

using namespace mofo;

extern "C" {

int mofo_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, mofo_bf16* w16,
                    const int64_t* segs, const int32_t* tiles, int n_tiles, const float* hyper, const float* clip_coef,
                    const float* loss_guard, void* stream) {
  MOFO_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && segs && tiles && hyper, "adamw_step: null pointer");
  MOFO_CHECK_ARG(n_tiles > 0, "adamw_step: empty tile table");
  adamw_kernel<<<n_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq,
                                                                       reinterpret_cast<__nv_bfloat16*>(w16), segs, tiles,
                                                                       hyper, clip_coef, loss_guard);
  MOFO_LAUNCH_CHECK("adamw_kernel");
  return MOFO_OK;
}

}  // extern "C"
