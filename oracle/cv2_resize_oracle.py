"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy restatement of the two `cv2.resize` modes the reference's frame ingest uses (run.py:179-183, :263, :415-421;
utils/utils.py:165-173): INTER_LINEAR on uint8 frames and INTER_NEAREST on uint8 label maps.  The arithmetic lives in
OpenCV (imgproc/src/resize.cpp; the reference pins opencv 3.4.2 in environment.yml, this container has 4.13 -- the
uint8 linear path has used the same 11-bit fixed-point scheme throughout), so this oracle is PINNED against the
library itself: oracle/make_resize_golden.py runs cv2 here and commits tests/golden/resize_*.npz, and
tests/test_oracle.py also compares against cv2 directly whenever it is importable.

INTER_LINEAR, uint8 (resize.cpp: resize() -> ResizeFunc<HResizeLinear<uchar,int,short>, VResizeLinear<...>>):
  scale = 1 / (dst / src)  (double);  f = (float)((d + 0.5) * scale - 0.5);  i = floor(f);  f -= i
  x direction: i < 0 -> (i, f) = (0, 0);  i >= src-1 -> (i, f) = (src-1, 0)     (tap AND weight are clamped)
  y direction: only the row indices i, i+1 are clipped to [0, src-1]             (the weights are not)
  weights: saturate_cast<short>(w * 2048) with round-half-even
  rows:  S = src[y][x0] * a0 + src[y][x1] * a1                                  (int, scaled by 2^11)
  out  = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
  exact 2x decimation in both directions is rerouted to INTER_AREA: out = (s00 + s01 + s10 + s11 + 2) >> 2
INTER_NEAREST: src index = min(floor(d * (src / dst)), src - 1) per axis (resizeNN).
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_coeffs(dn, sn, clamp_weight):
    scale = 1.0 / (np.float64(dn) / np.float64(sn))
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    i = np.floor(f).astype(np.int32)
    f = (f - i.astype(np.float32)).astype(np.float32)
    if clamp_weight:
        lo = i < 0
        i[lo] = 0
        f[lo] = 0
        hi = i >= sn - 1
        f[hi] = 0
        i[hi] = sn - 1
    c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    return np.clip(i, 0, sn - 1), np.clip(i + 1, 0, sn - 1), c0, c1


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh)) for uint8 [H,W] or [H,W,C]."""
    src = np.asarray(src, dtype=np.uint8)
    sh, sw = src.shape[:2]
    s = src.reshape(sh, sw, -1).astype(np.int32)
    if sw == 2 * dw and sh == 2 * dh:                              # INTER_LINEAR -> INTER_AREA fast path
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape((dh, dw) + src.shape[2:])
    x0, x1, a0, a1 = linear_coeffs(dw, sw, True)
    y0, y1, b0, b1 = linear_coeffs(dh, sh, False)
    r0 = s[y0][:, x0] * a0[None, :, None] + s[y0][:, x1] * a1[None, :, None]
    r1 = s[y1][:, x0] * a0[None, :, None] + s[y1][:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def resize_nearest_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST)."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    xi = np.minimum(np.floor(np.arange(dw) * (1.0 / (np.float64(dw) / sw))).astype(np.int64), sw - 1)
    yi = np.minimum(np.floor(np.arange(dh) * (1.0 / (np.float64(dh) / sh))).astype(np.int64), sh - 1)
    return src[yi][:, xi]


def ingest_frame(frame_bgr, height, width):
    """run.py:415-416: resize to the student size, then BGR -> RGB."""
    return resize_linear_u8(frame_bgr, width, height)[..., ::-1].copy()
