// Pixel stages of the reference's offline motion-box pipeline, byte / integer work bound by HBM and L2:
//   motion_map_kernel    optical-flow video -> motion-boundary magnitude map
//                        (scripts/data/motion_map_creator.py:160-205, scripts/motion_sts.py:5-37)
//   gauss_pass_kernel    one 1-D pass of scipy.ndimage.gaussian_filter on a uint8 [T,H,W,3] stack, float64 arithmetic in
//                        NI_Correlate1D's order, stored back as uint8 (bounding_box_creator_SSV.py:127,152)
//   box_stats_kernel     per-frame sum / sum of squares after the 0.4*max cut (:139-142)
//   gray_kernel          cv2 BGR2GRAY (:165)
// Everything is bit-exact with the reference's numpy / scipy / cv2 arithmetic (oracle/motion_oracle.py): the float32 and
// float64 operations that decide a byte are written with explicit round-to-nearest intrinsics so that neither
// --use_fast_math nor FMA contraction can change them.
#include <stdlib.h>

#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

// ------------------------------------------------------------------------------------------------------------------------
// Stage A.  One CTA owns a 32 x 8 pixel tile of ONE video and walks its frames in order, keeping the window sums of the u
// and v flow channels (with a one-pixel halo, 'reflect' = edge pixel repeated) in shared memory.  The reference's window
// [lo, hi) never moves backwards, so the tile is updated by subtracting the frames that left and adding the frames that
// entered: every flow byte is read about twice, whatever ws is, instead of ws times.  All sums are exact integers
// (|dx|, |dy| <= 6 * 255 * ws), identical to the reference's float32 accumulation of per-frame stencils.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int MM_TX = 32, MM_TY = 8, MM_HX = MM_TX + 2, MM_HY = MM_TY + 2;

__device__ __forceinline__ void flow_window(int idx /*1-based*/, int T, int ws, int& lo, int& hi) {
  if (ws == 1) { lo = idx - 1; hi = idx; return; }
  const int h = ws / 2;
  if (idx - h >= 0 && idx + h <= T) { lo = idx - h; hi = idx + h; }
  else if (idx - h >= 0) { lo = max(T - ws, 0); hi = T; }          // idx + h > T
  else if (idx + h <= T) { lo = 0; hi = min(ws, T); }              // idx - h < 0
  else { lo = 0; hi = T; }
}

__global__ void __launch_bounds__(MM_TX * MM_TY) motion_map_kernel(const uint8_t* __restrict__ flows, int T, int H, int W, int C,
                                                                   int ws, int border, uint8_t* __restrict__ out, int OC) {
  __shared__ int su[MM_HY][MM_HX], sv[MM_HY][MM_HX];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * MM_TX + tx;
  const int x0 = blockIdx.x * MM_TX, y0 = blockIdx.y * MM_TY;
  const int x = x0 + tx, y = y0 + ty;
  const size_t frame_stride = static_cast<size_t>(H) * W * C;
  // this thread's (up to two) halo-tile positions and their clamped source offsets
  int pos[2], off[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = tid + k * MM_TX * MM_TY;
    pos[k] = p < MM_HX * MM_HY ? p : -1;
    const int hy = p / MM_HX, hx = p % MM_HX;
    const int sy = min(max(y0 + hy - 1, 0), H - 1), sx = min(max(x0 + hx - 1, 0), W - 1);
    off[k] = (sy * W + sx) * C;
  }
  int au[2] = {0, 0}, av[2] = {0, 0};          // running window sums of the thread's halo positions
  int cur_lo = 0, cur_hi = 0;                  // frames currently inside the sums
  for (int t = 0; t < T; ++t) {
    int lo, hi;
    flow_window(t + 1, T, ws, lo, hi);
    if (lo < cur_lo || hi < cur_hi || lo > cur_hi) {      // never taken for the reference's windows; kept for safety
      au[0] = au[1] = av[0] = av[1] = 0;
      cur_lo = cur_hi = lo;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (pos[k] < 0) continue;
      const uint8_t* src = flows + off[k];
      for (int f = cur_lo; f < lo; ++f) { au[k] -= src[f * frame_stride]; av[k] -= src[f * frame_stride + 1]; }
      for (int f = cur_hi; f < hi; ++f) { au[k] += src[f * frame_stride]; av[k] += src[f * frame_stride + 1]; }
      su[pos[k] / MM_HX][pos[k] % MM_HX] = au[k];
      sv[pos[k] / MM_HX][pos[k] % MM_HX] = av[k];
    }
    cur_lo = lo; cur_hi = hi;
    __syncthreads();
    if (x < W && y < H) {
      uint32_t v8 = 0;
      if (!(y < border || x < border || y >= H - border || x >= W - border)) {
        const int cx = tx + 1, cy = ty + 1;
        // ndimage.convolve flips the kernel: dx = sum_rows (left - right), dy = sum_cols (up - down)
        const int dxu = (su[cy - 1][cx - 1] + su[cy][cx - 1] + su[cy + 1][cx - 1]) - (su[cy - 1][cx + 1] + su[cy][cx + 1] + su[cy + 1][cx + 1]);
        const int dyu = (su[cy - 1][cx - 1] + su[cy - 1][cx] + su[cy - 1][cx + 1]) - (su[cy + 1][cx - 1] + su[cy + 1][cx] + su[cy + 1][cx + 1]);
        const int dxv = (sv[cy - 1][cx - 1] + sv[cy][cx - 1] + sv[cy + 1][cx - 1]) - (sv[cy - 1][cx + 1] + sv[cy][cx + 1] + sv[cy + 1][cx + 1]);
        const int dyv = (sv[cy - 1][cx - 1] + sv[cy - 1][cx] + sv[cy - 1][cx + 1]) - (sv[cy + 1][cx - 1] + sv[cy + 1][cx] + sv[cy + 1][cx + 1]);
        // cv2.cartToPolar (float32): sqrt(fma(x, x, y*y)); then (mag_u + mag_v) / 2; astype(uint8) wraps: trunc mod 256
        const float fxu = static_cast<float>(dxu), fyu = static_cast<float>(dyu), fxv = static_cast<float>(dxv), fyv = static_cast<float>(dyv);
        const float mu = __fsqrt_rn(__fmaf_rn(fxu, fxu, __fmul_rn(fyu, fyu)));
        const float mv = __fsqrt_rn(__fmaf_rn(fxv, fxv, __fmul_rn(fyv, fyv)));
        const float m = __fmul_rn(__fadd_rn(mu, mv), 0.5f);
        v8 = static_cast<uint32_t>(m) & 255u;
      }
      uint8_t* dst = out + (static_cast<size_t>(t) * H * W + static_cast<size_t>(y) * W + x) * OC;
      for (int c = 0; c < OC; ++c) dst[c] = static_cast<uint8_t>(v8);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Stage B.
// ------------------------------------------------------------------------------------------------------------------------
// per-frame statistics block: [0] max after the first gaussian, [1] sum, [2] sum of squares after the 0.4*max cut
constexpr int BOX_STATS = 4;

// x < thr for a non-negative integer x  <=>  x < ceil(thr)
__device__ __forceinline__ int int_cut(double thr) {
  if (!(thr > 0.0)) return 0;
  return thr > 1.0e9 ? 1000000000 : static_cast<int>(ceil(thr));
}
__device__ __forceinline__ int cut_from_max(unsigned long long mx, double remove_thrd) {
  return int_cut(__dmul_rn(remove_thrd, static_cast<double>(mx)));                          // frame < 0.4 * max
}
// frame < std_k * (np.std(frame) + eps); the variance is formed from exact integer sums: (n*S2 - S1^2) / n^2, one rounding
__device__ __forceinline__ int cut_from_std(unsigned long long s1, unsigned long long s2, long long n, double std_k, double eps) {
  const unsigned long long num = static_cast<unsigned long long>(n) * s2 - s1 * s1;        // < 2^52 for n <= 2^18 * 16
  const double var = __ddiv_rn(static_cast<double>(num), __dmul_rn(static_cast<double>(n), static_cast<double>(n)));
  return int_cut(__dmul_rn(std_k, __dadd_rn(__dsqrt_rn(var), eps)));
}

// One 1-D pass along AXIS (0 = H, 1 = W, 2 = channel) of a [T,H,W,3] uint8 stack.  Thread = one output byte; blockIdx.y = frame.
// wd[d] = gaussian weight at distance d (0..r).  NI_Correlate1D, symmetric branch:
//   tmp = x[0]*w[0];  for j = -r .. -1:  tmp += (x[j] + x[-j]) * w[j]      (float64, no contraction), then a C cast to uint8.
// The 'reflect' extension (edge sample repeated, any number of reflections) comes from a per-CTA lookup table.
// cut_mode 1: bytes below max(cut(0.4*max), cut(1.5*(std+eps))) of their frame are read as 0 (the two in-place
// thresholdings of the reference folded into the load).  want_max: the frame maximum of the OUTPUT is accumulated.
template <int AXIS>
__global__ void __launch_bounds__(256) gauss_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                                         const double* __restrict__ wd, int r, int cut_mode, double remove_thrd,
                                                         double std_k, double std_eps, unsigned long long* __restrict__ stats,
                                                         int want_max) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int n = AXIS == 0 ? H : AXIS == 1 ? W : 3;
  int* lut = reinterpret_cast<int*>(smem_raw);                       // [n + 2r]: source offset of position k - r
  double* sw = reinterpret_cast<double*>(smem_raw + ((static_cast<size_t>(n + 2 * r) * 4 + 7) & ~size_t(7)));
  __shared__ int s_cut;
  const int stride = AXIS == 0 ? W * 3 : AXIS == 1 ? 3 : 1;
  for (int k = threadIdx.x; k < n + 2 * r; k += blockDim.x) {
    int p = (k - r) % (2 * n);
    if (p < 0) p += 2 * n;
    lut[k] = (p >= n ? 2 * n - 1 - p : p) * stride;
  }
  for (int k = threadIdx.x; k <= r; k += blockDim.x) sw[k] = wd[k];
  const int t = blockIdx.y;
  const long long frame_n = static_cast<long long>(H) * W * 3;
  if (threadIdx.x == 0) {
    int cut = 0;
    if (cut_mode) {
      const unsigned long long* st = stats + static_cast<size_t>(t) * BOX_STATS;
      cut = max(cut_from_max(st[0], remove_thrd), cut_from_std(st[1], st[2], frame_n, std_k, std_eps));
    }
    s_cut = cut;
  }
  __syncthreads();
  const int cut = s_cut;
  const uint8_t* fin = in + static_cast<size_t>(t) * frame_n;
  int vmax = 0;
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < frame_n) {
    const int c = static_cast<int>(e % 3), xw = static_cast<int>((e / 3) % W), yh = static_cast<int>(e / (3LL * W));
    const int pos = AXIS == 0 ? yh : AXIS == 1 ? xw : c;
    const uint8_t* line = fin + (e - static_cast<long long>(pos) * stride);      // element 0 of this output's line
    auto ld = [&](int k) -> int {
      const int v = line[lut[k]];
      return v < cut ? 0 : v;
    };
    double tmp = __dmul_rn(static_cast<double>(ld(pos + r)), sw[0]);
    for (int d = r; d >= 1; --d)
      tmp = __dadd_rn(tmp, __dmul_rn(static_cast<double>(ld(pos + r - d) + ld(pos + r + d)), sw[d]));
    const int v = static_cast<int>(tmp);              // C cast double -> unsigned char for values in [0, 256)
    out[static_cast<size_t>(t) * frame_n + e] = static_cast<uint8_t>(v);
    vmax = v & 255;
  }
  if (want_max) {
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if ((threadIdx.x & 31) == 0 && vmax > 0) atomicMax(stats + static_cast<size_t>(t) * BOX_STATS, static_cast<unsigned long long>(vmax));
  }
}

__global__ void __launch_bounds__(256) box_stats_kernel(const uint8_t* __restrict__ in, long long frame_n, double remove_thrd,
                                                        unsigned long long* __restrict__ stats) {
  const int t = blockIdx.y;
  unsigned long long* st = stats + static_cast<size_t>(t) * BOX_STATS;
  const int cut = cut_from_max(st[0], remove_thrd);
  const uint8_t* fin = in + static_cast<size_t>(t) * frame_n;
  unsigned int s1 = 0, s2 = 0;                      // 16 elements per thread: no overflow
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const long long e = (static_cast<long long>(blockIdx.x) * 16 + k) * 256 + threadIdx.x;
    if (e < frame_n) {
      int v = fin[e];
      v = v < cut ? 0 : v;
      s1 += v; s2 += v * v;
    }
  }
  s1 = __reduce_add_sync(0xffffffffu, s1);
  s2 = __reduce_add_sync(0xffffffffu, s2);
  if ((threadIdx.x & 31) == 0 && s1) {
    atomicAdd(st + 1, static_cast<unsigned long long>(s1));
    atomicAdd(st + 2, static_cast<unsigned long long>(s2));
  }
}

__global__ void __launch_bounds__(256) gray_kernel(const uint8_t* __restrict__ bgr, long long pixels, uint8_t* __restrict__ gray) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  const uint8_t* s = bgr + p * 3;
  gray[p] = static_cast<uint8_t>((s[0] * 3735 + s[1] * 19235 + s[2] * 9798 + 16384) >> 15);
}

}  // namespace mofo

using namespace mofo;

extern "C" {

int mofo_motion_map(const uint8_t* flows, int T, int H, int W, int C, int ws, int border, uint8_t* out, int out_channels,
                    void* stream) {
  MOFO_CHECK_ARG(flows && out, "motion_map: null pointer");
  MOFO_CHECK_ARG(T > 0 && H > 0 && W > 0 && C >= 2 && ws >= 1 && ws <= 4096 && border >= 0 && out_channels >= 1 &&
                     static_cast<int64_t>(H) * W * C < (int64_t(1) << 31),
                 "motion_map: bad shape T=%d H=%d W=%d C=%d ws=%d border=%d out_channels=%d", T, H, W, C, ws, border, out_channels);
  const dim3 grid((W + MM_TX - 1) / MM_TX, (H + MM_TY - 1) / MM_TY);
  MOFO_CHECK_ARG(grid.y <= 65535, "motion_map: H too large");
  motion_map_kernel<<<grid, dim3(MM_TX, MM_TY), 0, static_cast<cudaStream_t>(stream)>>>(flows, T, H, W, C, ws, border, out, out_channels);
  MOFO_LAUNCH_CHECK("motion_map_kernel");
  return MOFO_OK;
}

int mofo_motion_box_filter(const uint8_t* frames, int T, int H, int W, const double* w_before, int r_before, const double* w_after,
                           int r_after, double remove_thrd, double std_k, double std_eps, uint8_t* work, uint64_t* stats,
                           uint8_t* filtered, uint8_t* gray, void* stream) {
  MOFO_CHECK_ARG(frames && w_before && w_after && work && stats && filtered && gray, "motion_box_filter: null pointer");
  MOFO_CHECK_ARG(T > 0 && T <= 65535 && H > 0 && W > 0 && H <= 8192 && W <= 8192 && r_before >= 0 && r_after >= 0 && r_before <= 2048 &&
                     r_after <= 2048 && static_cast<int64_t>(H) * W * 3 <= (int64_t(1) << 22),
                 "motion_box_filter: bad shape T=%d H=%d W=%d r=%d/%d (H*W*3 must not exceed 2^22)", T, H, W, r_before, r_after);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long frame_n = static_cast<long long>(H) * W * 3;
  uint8_t* bufA = work;
  uint8_t* bufB = work + static_cast<size_t>(T) * frame_n;
  unsigned long long* stats_ull = reinterpret_cast<unsigned long long*>(stats);
  MOFO_CUDA(cudaMemsetAsync(stats, 0, static_cast<size_t>(T) * BOX_STATS * sizeof(uint64_t), st));
  const dim3 grid(static_cast<unsigned>((frame_n + 255) / 256), T);
  auto smem = [&](int n, int r) { return ((static_cast<size_t>(n + 2 * r) * 4 + 7) & ~size_t(7)) + static_cast<size_t>(r + 1) * 8; };
  MOFO_CHECK_ARG(smem(max(H, W), max(r_before, r_after)) <= 48 * 1024, "motion_box_filter: lookup table does not fit shared memory");
  // gaussian_filter(sigma = before): H, W, channel axis in turn; the last pass also reduces the frame maximum
  gauss_pass_kernel<0><<<grid, 256, smem(H, r_before), st>>>(frames, bufA, H, W, w_before, r_before, 0, 0.0, 0.0, 0.0, stats_ull, 0);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<0>");
  gauss_pass_kernel<1><<<grid, 256, smem(W, r_before), st>>>(bufA, bufB, H, W, w_before, r_before, 0, 0.0, 0.0, 0.0, stats_ull, 0);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<1>");
  gauss_pass_kernel<2><<<grid, 256, smem(3, r_before), st>>>(bufB, bufA, H, W, w_before, r_before, 0, 0.0, 0.0, 0.0, stats_ull, 1);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<2>");
  box_stats_kernel<<<dim3(static_cast<unsigned>((frame_n + 4095) / 4096), T), 256, 0, st>>>(bufA, frame_n, remove_thrd, stats_ull);
  MOFO_LAUNCH_CHECK("box_stats_kernel");
  // gaussian_filter(sigma = after) of the doubly thresholded frame; the thresholds are applied as the bytes are read
  gauss_pass_kernel<0><<<grid, 256, smem(H, r_after), st>>>(bufA, bufB, H, W, w_after, r_after, 1, remove_thrd, std_k, std_eps, stats_ull, 0);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<0>");
  gauss_pass_kernel<1><<<grid, 256, smem(W, r_after), st>>>(bufB, bufA, H, W, w_after, r_after, 0, 0.0, 0.0, 0.0, stats_ull, 0);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<1>");
  gauss_pass_kernel<2><<<grid, 256, smem(3, r_after), st>>>(bufA, filtered, H, W, w_after, r_after, 0, 0.0, 0.0, 0.0, stats_ull, 0);
  MOFO_LAUNCH_CHECK("gauss_pass_kernel<2>");
  const long long pixels = static_cast<long long>(T) * H * W;
  gray_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, st>>>(filtered, pixels, gray);
  MOFO_LAUNCH_CHECK("gray_kernel");
  return MOFO_OK;
}

}  // extern "C"
