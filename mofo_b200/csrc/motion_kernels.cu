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
                                                                   int ws, int border, uint8_t* __restrict__ out, int OC, int chunk) {
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
  // blockIdx.z = chunk of `chunk` consecutive frames: a chunk rebuilds its first window from scratch (ws extra frame reads
  // from L2) and then slides; 300 CTAs walking 48 frames one after the other were latency-bound (78 us per video)
  const int t_begin = blockIdx.z * chunk, t_end = min(T, t_begin + chunk);
  int cur_lo, cur_hi;                          // frames currently inside the sums
  {
    int lo0, hi0;
    flow_window(t_begin + 1, T, ws, lo0, hi0);
    cur_lo = cur_hi = lo0;
  }
  for (int t = t_begin; t < t_end; ++t) {
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

// ---- specialised passes: the generic kernel above spends ~12 instructions per tap pair (two table lookups, two global byte
// loads, two threshold selects ...) for two float64 operations and ran at 13 % of the float64 issue rate (bench.py --workload
// motion, round 2).  These stage the reflected, thresholded line segments in shared memory ONCE, so a tap pair costs two
// shared byte loads, one integer add, one conversion, DMUL, DADD (+ the weight, a broadcast shared load); the channel pass
// needs no loads at all.  Arithmetic and its order are unchanged.  The host falls back to the generic kernel when a frame is
// too large for the staging buffers.
__device__ __forceinline__ int frame_cut(int cut_mode, const unsigned long long* stats, int t, long long frame_n, double remove_thrd,
                                         double std_k, double std_eps) {
  if (!cut_mode) return 0;
  const unsigned long long* st = stats + static_cast<size_t>(t) * BOX_STATS;
  return max(cut_from_max(st[0], remove_thrd), cut_from_std(st[1], st[2], frame_n, std_k, std_eps));
}
// exact int -> double for 0 <= s < 2^32 without the conversion unit: the bits (0x43300000, s) are the double 2^52 + s, and the
// subtraction is exact.  I2F.F64 issues at a quarter of the DADD rate and was the limiter of the staged passes.
__device__ __forceinline__ double u32_to_double(unsigned int s) { return __dadd_rn(__hiloint2double(0x43300000, static_cast<int>(s)), -4503599627370496.0); }
__device__ __forceinline__ int reflect_idx(int i, int n) {
  int p = i % (2 * n);
  if (p < 0) p += 2 * n;
  return p >= n ? 2 * n - 1 - p : p;
}

// Four outputs per thread, taken from ALIGNED 32-bit shared-memory words (four neighbouring lines side by side): the first
// staged version issued two byte loads and one weight load per tap pair and was bound by the shared-memory pipe (one
// wavefront per instruction: 3 clk per warp and pair against 1.5 clk of float64 work).  Here a step costs two word loads and
// one weight load for four pairs; the four byte sums are formed two at a time in 16-bit lanes (PRMT + IADD).
__device__ __forceinline__ void tap_loop4(const uint8_t* base, int stride, const double* sw, int r, int (&res)[4]) {
  auto w32 = [&](int off) { return *reinterpret_cast<const unsigned int*>(base + off); };
  const unsigned int c0 = w32(0);
  double tmp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) tmp[k] = __dmul_rn(u32_to_double((c0 >> (8 * k)) & 255u), sw[0]);
#pragma unroll 4
  for (int d = r; d >= 1; --d) {
    const unsigned int a = w32(-d * stride), b = w32(d * stride);
    const double w = sw[d];
    const unsigned int se = __byte_perm(a, 0, 0x4240) + __byte_perm(b, 0, 0x4240);      // bytes 0 and 2 in 16-bit lanes
    const unsigned int so = __byte_perm(a, 0, 0x4341) + __byte_perm(b, 0, 0x4341);      // bytes 1 and 3
    tmp[0] = __dadd_rn(tmp[0], __dmul_rn(u32_to_double(se & 0xFFFFu), w));
    tmp[1] = __dadd_rn(tmp[1], __dmul_rn(u32_to_double(so & 0xFFFFu), w));
    tmp[2] = __dadd_rn(tmp[2], __dmul_rn(u32_to_double(se >> 16), w));
    tmp[3] = __dadd_rn(tmp[3], __dmul_rn(u32_to_double(so >> 16), w));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) res[k] = static_cast<int>(tmp[k]) & 255;      // C cast double -> unsigned char
}

// AXIS 0 (H): one CTA = a strip of CW consecutive bytes of the (W*3)-byte rows over the full height of one frame; a thread
// owns four neighbouring bytes of a row.
constexpr int GS_CW = 64;
__global__ void __launch_bounds__(256) gauss_strip_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W3,
                                                            const double* __restrict__ wd, int r, int cut_mode, double remove_thrd,
                                                            double std_k, double std_eps, const unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double* sw = reinterpret_cast<double*>(smem_raw);                               // [r + 1]
  uint8_t* s = smem_raw + static_cast<size_t>(r + 1) * 8;                         // [(H + 2r)][CW]
  const int t = blockIdx.y, x0 = blockIdx.x * GS_CW;
  const long long frame_n = static_cast<long long>(H) * W3;
  const uint8_t* fin = in + static_cast<size_t>(t) * frame_n;
  const int cut = frame_cut(cut_mode, stats, t, frame_n, remove_thrd, std_k, std_eps);
  for (int k = threadIdx.x; k <= r; k += 256) sw[k] = wd[k];
  for (int i = threadIdx.x; i < (H + 2 * r) * GS_CW; i += 256) {
    const int row = i / GS_CW, x = i % GS_CW;
    int v = 0;
    if (x0 + x < W3) v = fin[static_cast<size_t>(reflect_idx(row - r, H)) * W3 + x0 + x];
    s[i] = static_cast<uint8_t>(v < cut ? 0 : v);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < H * (GS_CW / 4); o += 256) {
    const int y = o / (GS_CW / 4), x = (o % (GS_CW / 4)) * 4;
    if (x0 + x >= W3) continue;
    int res[4];
    tap_loop4(s + (y + r) * GS_CW + x, GS_CW, sw, r, res);
    uint8_t* dst = out + static_cast<size_t>(t) * frame_n + static_cast<size_t>(y) * W3 + x0 + x;
    if (x0 + x + 3 < W3 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
      *reinterpret_cast<unsigned int*>(dst) = static_cast<unsigned int>(res[0]) | (res[1] << 8) | (res[2] << 16) | (static_cast<unsigned int>(res[3]) << 24);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x0 + x + k < W3) dst[k] = static_cast<uint8_t>(res[k]);
    }
  }
}

// AXIS 1 (W): one CTA = RB (4, 8 or 16) full rows of one frame, de-interleaved per channel with the reflected halo and
// TRANSPOSED so that the RB rows of one (channel, x) sit side by side: s[c][W + 2r][RB]; a thread owns four rows of one (c, x).
__global__ void __launch_bounds__(256) gauss_rows_w_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                                           int RB, const double* __restrict__ wd, int r, int cut_mode,
                                                           double remove_thrd, double std_k, double std_eps,
                                                           const unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double* sw = reinterpret_cast<double*>(smem_raw);
  uint8_t* s = smem_raw + static_cast<size_t>(r + 1) * 8;
  const int t = blockIdx.y, y0 = blockIdx.x * RB, WE = W + 2 * r, W3 = W * 3;
  const int rows = min(RB, H - y0);
  const long long frame_n = static_cast<long long>(H) * W3;
  const uint8_t* fin = in + static_cast<size_t>(t) * frame_n + static_cast<size_t>(y0) * W3;
  const int cut = frame_cut(cut_mode, stats, t, frame_n, remove_thrd, std_k, std_eps);
  for (int k = threadIdx.x; k <= r; k += 256) sw[k] = wd[k];
  for (int i = threadIdx.x; i < 3 * WE * RB; i += 256) {
    const int row = i % RB, rem = i / RB, xe = rem % WE, c = rem / WE;
    int v = 0;
    if (row < rows) v = fin[static_cast<size_t>(row) * W3 + reflect_idx(xe - r, W) * 3 + c];
    s[i] = static_cast<uint8_t>(v < cut ? 0 : v);
  }
  __syncthreads();
  const int quads = RB / 4;
  for (int o = threadIdx.x; o < 3 * W * quads; o += 256) {
    const int q = o % quads, rem = o / quads, x = rem % W, c = rem / W;
    if (4 * q >= rows) continue;
    int res[4];
    tap_loop4(s + (c * WE + x + r) * RB + 4 * q, RB, sw, r, res);
    uint8_t* dst = out + static_cast<size_t>(t) * frame_n + (static_cast<size_t>(y0 + 4 * q) * W + x) * 3 + c;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * q + k < rows) dst[static_cast<size_t>(k) * W3] = static_cast<uint8_t>(res[k]);
  }
}

// AXIS 2 (channel): thread = one pixel.  The reflected extension of a 3-sample line has period 6, so x[c-d] + x[c+d] takes
// only six values per output channel: they are formed once (18 doubles in registers) and the tap loop is pure DMUL + DADD
// with a broadcast weight load, in scipy's order d = r .. 1.
__global__ void __launch_bounds__(256) gauss_chan_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long pixels,
                                                         long long frame_n, const double* __restrict__ wd, int r, int cut_mode,
                                                         double remove_thrd, double std_k, double std_eps,
                                                         unsigned long long* __restrict__ stats, int want_max) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double* sw = reinterpret_cast<double*>(smem_raw);
  const int t = blockIdx.y;
  const int cut = frame_cut(cut_mode, stats, t, frame_n, remove_thrd, std_k, std_eps);
  for (int k = threadIdx.x; k <= r; k += 256) sw[k] = wd[k];
  __syncthreads();
  const long long px = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  int vmax = 0;
  if (px < pixels) {
    const uint8_t* src = in + static_cast<size_t>(t) * frame_n + px * 3;
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { v[c] = src[c]; v[c] = v[c] < cut ? 0 : v[c]; }
    // ext(k) for k = -6 .. 8 via the period-6 pattern v0 v1 v2 v2 v1 v0
    auto ext = [&](int k) { const int p = ((k % 6) + 6) % 6; return p == 0 || p == 5 ? v[0] : (p == 1 || p == 4 ? v[1] : v[2]); };
    double ps[3][6], tmp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int m = 0; m < 6; ++m) ps[c][m] = static_cast<double>(ext(c - m) + ext(c + m));      // d = m (mod 6)
      tmp[c] = __dmul_rn(static_cast<double>(v[c]), sw[0]);
    }
    int d = r;
    for (; d % 6 != 0; --d) {                     // head: distances down to the next multiple of 6 (at most five steps)
      const int m = d % 6;
      const double w = sw[d];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double q = m == 1 ? ps[c][1] : m == 2 ? ps[c][2] : m == 3 ? ps[c][3] : m == 4 ? ps[c][4] : ps[c][5];
        tmp[c] = __dadd_rn(tmp[c], __dmul_rn(q, w));
      }
    }
    for (; d >= 6; d -= 6) {                      // d, d-1, .. d-5  <->  m = 0, 5, 4, 3, 2, 1
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double w = sw[d - j];
        const int m = (6 - j) % 6;
#pragma unroll
        for (int c = 0; c < 3; ++c) tmp[c] = __dadd_rn(tmp[c], __dmul_rn(ps[c][m], w));
      }
    }
    uint8_t* dst = out + static_cast<size_t>(t) * frame_n + px * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int o = static_cast<int>(tmp[c]);
      dst[c] = static_cast<uint8_t>(o);
      vmax = max(vmax, o & 255);
    }
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
  // frames per CTA: long enough that sliding beats rebuilding (>= 2 ws), short enough to fill the machine (~8 CTAs per SM)
  const int tiles = ((W + MM_TX - 1) / MM_TX) * ((H + MM_TY - 1) / MM_TY);
  int chunk = T;
  if (getenv("MOFO_MOTION_CHUNK")) chunk = atoi(getenv("MOFO_MOTION_CHUNK"));
  else while (chunk > 2 * ws && tiles * ((T + chunk - 1) / chunk) < 8 * sm_count()) chunk = (chunk + 1) / 2;
  if (chunk < 1) chunk = 1;
  const dim3 grid((W + MM_TX - 1) / MM_TX, (H + MM_TY - 1) / MM_TY, (T + chunk - 1) / chunk);
  MOFO_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "motion_map: H or T too large");
  motion_map_kernel<<<grid, dim3(MM_TX, MM_TY), 0, static_cast<cudaStream_t>(stream)>>>(flows, T, H, W, C, ws, border, out, out_channels, chunk);
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
  const bool generic_only = getenv("MOFO_MOTION_GENERIC") != nullptr;      // tests: force the fallback kernels
  // one 1-D pass: the staged kernel when its buffers fit the default 48 KB of shared memory, else the generic one
  auto pass = [&](int axis, const uint8_t* src, uint8_t* dst, const double* w, int r, int cut_mode, int want_max) -> int {
    const size_t wbytes = static_cast<size_t>(r + 1) * 8;
    if (axis == 0) {
      const size_t need = wbytes + static_cast<size_t>(H + 2 * r) * GS_CW;
      if (!generic_only && need <= 48 * 1024) {
        gauss_strip_h_kernel<<<dim3((W * 3 + GS_CW - 1) / GS_CW, T), 256, need, st>>>(src, dst, H, W * 3, w, r, cut_mode, remove_thrd, std_k,
                                                                                      std_eps, stats_ull);
      } else {
        gauss_pass_kernel<0><<<grid, 256, smem(H, r), st>>>(src, dst, H, W, w, r, cut_mode, remove_thrd, std_k, std_eps, stats_ull, want_max);
      }
    } else if (axis == 1) {
      const size_t per_row = static_cast<size_t>(3) * (W + 2 * r);
      const size_t rb_fit = (48 * 1024 - wbytes) / per_row;
      const int RB = rb_fit >= 16 ? 16 : rb_fit >= 8 ? 8 : rb_fit >= 4 ? 4 : 0;
      if (!generic_only && RB >= 4) {
        gauss_rows_w_kernel<<<dim3((H + RB - 1) / RB, T), 256, wbytes + RB * per_row, st>>>(src, dst, H, W, RB, w, r, cut_mode, remove_thrd,
                                                                                            std_k, std_eps, stats_ull);
      } else {
        gauss_pass_kernel<1><<<grid, 256, smem(W, r), st>>>(src, dst, H, W, w, r, cut_mode, remove_thrd, std_k, std_eps, stats_ull, want_max);
      }
    } else {
      if (!generic_only) {
        const long long px = static_cast<long long>(H) * W;
        gauss_chan_kernel<<<dim3(static_cast<unsigned>((px + 255) / 256), T), 256, wbytes, st>>>(src, dst, px, frame_n, w, r, cut_mode, remove_thrd,
                                                                                               std_k, std_eps, stats_ull, want_max);
      } else {
        gauss_pass_kernel<2><<<grid, 256, smem(3, r), st>>>(src, dst, H, W, w, r, cut_mode, remove_thrd, std_k, std_eps, stats_ull, want_max);
      }
    }
    MOFO_LAUNCH_CHECK("gauss pass");
    return MOFO_OK;
  };
  int rc;
  // gaussian_filter(sigma = before): H, W, channel axis in turn; the last pass also reduces the frame maximum
  if ((rc = pass(0, frames, bufA, w_before, r_before, 0, 0)) != MOFO_OK) return rc;
  if ((rc = pass(1, bufA, bufB, w_before, r_before, 0, 0)) != MOFO_OK) return rc;
  if ((rc = pass(2, bufB, bufA, w_before, r_before, 0, 1)) != MOFO_OK) return rc;
  box_stats_kernel<<<dim3(static_cast<unsigned>((frame_n + 4095) / 4096), T), 256, 0, st>>>(bufA, frame_n, remove_thrd, stats_ull);
  MOFO_LAUNCH_CHECK("box_stats_kernel");
  // gaussian_filter(sigma = after) of the doubly thresholded frame; the thresholds are applied as the bytes are read
  if ((rc = pass(0, bufA, bufB, w_after, r_after, 1, 0)) != MOFO_OK) return rc;
  if ((rc = pass(1, bufB, bufA, w_after, r_after, 0, 0)) != MOFO_OK) return rc;
  if ((rc = pass(2, bufA, filtered, w_after, r_after, 0, 0)) != MOFO_OK) return rc;
  const long long pixels = static_cast<long long>(T) * H * W;
  gray_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, st>>>(filtered, pixels, gray);
  MOFO_LAUNCH_CHECK("gray_kernel");
  return MOFO_OK;
}

}  // extern "C"
