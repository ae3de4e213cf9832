"""TEST INFRASTRUCTURE ONLY - CPU restatement of the pixel stages of the reference's offline motion-box pipeline.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the product path never does.

Stage A - motion map (scripts/data/motion_map_creator.py:160-205 calling scripts/motion_sts.py:5-37):
  for frame idx = 1..T of an optical-flow video [T,H,W,3] uint8 (channel 0 = u, 1 = v) take a window of ws flow frames
  (:163-172), per channel sum the 3x3 Prewitt responses over the window (compute_motion_boudary, motion_sts.py:5-27:
  scipy.ndimage.convolve, mode 'reflect', float32), magnitude = cv2.cartToPolar (:183-184), average the two channels (:185),
  zero an 8-pixel border (zero_boundary, motion_sts.py:29-37), replicate to 3 channels (:191) and cast to uint8 (:228).
  Pinned in tests/test_oracle_motion.py against the reference's own functions (imported from /root/reference when present),
  scipy.ndimage and cv2 (both in this image), bit for bit.  Facts the pin established:
    * every intermediate up to the magnitude is an exact integer in float32, so "sum the window, then one stencil" is identical
      to the reference's "stencil every frame, then sum";
    * cv2.cartToPolar's float32 magnitude is sqrt(fma(x, x, y*y)) on this build (not sqrt(x*x + y*y));
    * ndarray.astype(uint8) of the out-of-range float32 magnitudes (up to ~8.6e3) wraps: trunc(x) mod 256.  That is what
      the reference writes to the motion-map video, so it is what is reproduced.

Stage B - per-frame filtering in front of the contour search (scripts/data/SSV2/bounding_box_creator_SSV.py:125-166):
  gaussian_filter(frame, sigma=1) on the uint8 [H,W,3] array (scipy filters ALL three axes, the channel axis included, and
  stores every 1-D pass back into uint8), zero everything below 0.4 * max, zero everything below 1.5 * (std + 1e-5),
  gaussian_filter(sigma=30), cv2.cvtColor(BGR2GRAY).  scipy.ndimage is a dependency that is not vendored in
  /root/reference (the reference pins no version; this image has scipy 1.18.1): NI_Correlate1D's symmetric branch is
  restated here - tmp = x[0]*w[0]; for j = -r..-1: tmp += (x[j] + x[-j]) * w[j] in float64, 'reflect' extension
  (d c b a | a b c d | d c b a), C cast to uint8 - and pinned against scipy itself, on float64 data (which exposes the
  summation order) and on uint8 data.  cv2's BGR2GRAY for uint8 is (B*3735 + G*19235 + R*9798 + 16384) >> 15, pinned against cv2.

The contour search / ranking / temporal smoothing that follows (:168-475) is sequential host code in the reference and stays
host code.
"""
import numpy as np


# ---------------------------------------------------------------- stage A ----------------------------------------------------
def flow_window(idx, T, ws):
    """[lo, hi) of the flow frames used for the 1-based frame idx (motion_map_creator.py:160-172)."""
    if ws == 1:
        return idx - 1, idx
    h = ws // 2
    if idx - h >= 0 and idx + h <= T:
        return idx - h, idx + h
    if idx - h >= 0 and idx + h > T:
        return max(T - ws, 0), T
    if idx + h <= T and idx - h < 0:
        return 0, min(ws, T)
    return 0, T


def prewitt_reflect(s):
    """(dx, dy) = (ndimage.convolve(s, mx), ndimage.convolve(s, my)) for the reference's mx / my, mode 'reflect', on int64."""
    p = np.pad(s, 1, mode="symmetric")            # numpy 'symmetric' = scipy 'reflect' (edge sample repeated)
    dx = (p[:-2, :-2] + p[1:-1, :-2] + p[2:, :-2]) - (p[:-2, 2:] + p[1:-1, 2:] + p[2:, 2:])
    dy = (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:]) - (p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:])
    return dx, dy


def magnitude_f32(x, y):
    """cv2.cartToPolar's magnitude for float32 inputs: sqrt(fma(x, x, y*y)), every step rounded to float32."""
    x = x.astype(np.float32); y = y.astype(np.float32)
    yy = (y * y).astype(np.float32)
    s = (x.astype(np.float64) * x.astype(np.float64) + yy.astype(np.float64)).astype(np.float32)   # one rounding: the fma
    return np.sqrt(s, dtype=np.float32)


def wrap_u8(x):
    """ndarray.astype(np.uint8) for non-negative finite floats as this platform does it: trunc(x) mod 256."""
    return (np.trunc(x).astype(np.int64) & 255).astype(np.uint8)


def motion_map(flows, ws=8, border=8):
    """[T,H,W,>=2] uint8 flow frames -> [T,H,W] uint8 motion map (one channel; the reference replicates it three times)."""
    T, H, W = flows.shape[:3]
    out = np.zeros((T, H, W), np.uint8)
    f = flows.astype(np.int64)
    for idx in range(1, T + 1):
        lo, hi = flow_window(idx, T, ws)
        mags = []
        for ch in (0, 1):
            dx, dy = prewitt_reflect(f[lo:hi, :, :, ch].sum(0))
            mags.append(magnitude_f32(dx, dy))
        m = ((mags[0] + mags[1]).astype(np.float32) / np.float32(2)).astype(np.float32)
        m[:border, :] = 0; m[:, :border] = 0; m[-border:, :] = 0; m[:, -border:] = 0
        out[idx - 1] = wrap_u8(m)
    return out


# ---------------------------------------------------------------- stage B ----------------------------------------------------
def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) (float64), radius = int(truncate * sigma + 0.5)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum(), radius


def reflect_index(i, n):
    """scipy 'reflect' extension (d c b a | a b c d | d c b a), any distance outside [0, n)."""
    p = np.mod(i, 2 * n)
    return np.where(p >= n, 2 * n - 1 - p, p)


def correlate1d_symmetric(a, w, axis):
    """NI_Correlate1D, symmetric branch, on float64 data: returns float64 (the caller casts)."""
    r = (len(w) - 1) // 2
    a = np.moveaxis(np.asarray(a, np.float64), axis, 0)
    n = a.shape[0]
    pos = np.arange(n)
    tmp = a * w[r]
    for j in range(-r, 0):
        lo = a[reflect_index(pos + j, n)]
        hi = a[reflect_index(pos - j, n)]
        tmp = tmp + (lo + hi) * w[r + j]
    return np.moveaxis(tmp, 0, axis)


def gaussian_filter_u8(img, sigma):
    """scipy.ndimage.gaussian_filter(img, sigma) for a uint8 N-d array: every axis in turn, each pass stored as uint8."""
    w, _ = gaussian_weights(sigma)
    out = img
    for axis in range(img.ndim):
        out = correlate1d_symmetric(out, w, axis).astype(np.uint8)      # values are in [0, 255]: C cast = truncation
    return out


def bgr2gray_u8(img):
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8."""
    i = img.astype(np.int64)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def filter_frame(frame, before_sigma=1, remove_thrd=0.4, std_k=1.5, after_sigma=30):
    """bounding_box_creator_SSV.py:125-166 for one [H,W,3] uint8 motion-map frame -> (filtered [H,W,3] uint8, gray [H,W] uint8)."""
    f = gaussian_filter_u8(frame, before_sigma)
    mx = f.max()
    f = f.copy()
    f[f < remove_thrd * mx] = 0
    sigma = np.std(f) + 1e-5
    f[f < std_k * sigma] = 0
    f = gaussian_filter_u8(f, after_sigma)
    return f, bgr2gray_u8(f)
