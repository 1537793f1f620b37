"""numpy restatement of the pixel arithmetic behind the reference's KITTI training augmentation list
(configs/kitti_wpose_example:123-158 -> vision_base/data/augmentations/augmentations.py:91-109,200-226,436-498,527-592), as ONE
per-pixel function of the raw uint8 frames and a small parameter block ("plan").

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The heavy arithmetic of those classes lives in OpenCV (``cv2.warpAffine``,
``cv2.cvtColor``; opencv-python, unpinned in requirement.txt, 4.13.0 in this image), so it is restated here from OpenCV's published
algorithm and pinned twice: against cv2 itself on random inputs, and against golden vectors produced by the reference's own
pipeline (tests/golden/aug_train.npz) -- tests/test_device_augment_cpu.py.

cv2.warpAffine (imgwarp.cpp): the 2x3 matrix is inverted in double; source coordinates are 10-bit fixed point
(``rint(coef * 1024)``), rounded to 1/32 pixel for INTER_LINEAR (to the pixel for INTER_NEAREST); bilinear weights are float
products of (1 - f, f) with f = k/32; taps outside the image read the constant border 0.
cv2.resize (resize.cpp): source coordinate (d + 0.5) * scale - 0.5 with scale = 1 / (n_out / n_in) in double, split into
floor + float fraction, clamped at the borders; horizontal pass then vertical pass in float; INTER_NEAREST: floor(d * scale).
The Resize transform (augmentations.py:112-198) then zero-pads (or crops) to the requested size.
cv2.cvtColor RGB<->HSV on float32 (color_hsv.simd.hpp): H in [0, 360), S = (V - min) / (|V| + eps), no clipping anywhere.
"""
import numpy as np

F32 = np.float32
EPS = np.finfo(np.float32).eps
OP_NONE, OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION = 0, 1, 2, 3
PLAN_SIZE = 16           # [0:6] geometry, 6 mirror, 7:10 op codes, 10:13 op values, 13 h0, 14 w0, 15 geometry mode
GEOM_AFFINE, GEOM_RESIZE = 0, 1   # [0:6] = inverse affine (row major)  |  scale_x, scale_y, w_eff, h_eff, -, -


def invert_affine(M) -> np.ndarray:
    """cv2.warpAffine's inversion of a forward 2x3 matrix (imgwarp.cpp, ``!(flags & WARP_INVERSE_MAP)`` branch), in double."""
    M = np.asarray(M, dtype=np.float64).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def _fixed_point_grid(Minv, out_h, out_w, mirror, round_delta):
    x = np.arange(out_w, dtype=np.float64)
    if mirror:
        x = out_w - 1 - x                 # RandomMirror after the warp: output column x shows warped column W-1-x
    y = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(Minv[0, 0] * x * 1024).astype(np.int64)
    bdelta = np.rint(Minv[1, 0] * x * 1024).astype(np.int64)
    X0 = np.rint((Minv[0, 1] * y + Minv[0, 2]) * 1024).astype(np.int64) + round_delta
    Y0 = np.rint((Minv[1, 1] * y + Minv[1, 2]) * 1024).astype(np.int64) + round_delta
    return X0[:, None] + adelta[None, :], Y0[:, None] + bdelta[None, :]


def _tap(img, yy, xx, h0, w0):
    ok = (yy >= 0) & (yy < h0) & (xx >= 0) & (xx < w0)
    v = img[np.clip(yy, 0, h0 - 1), np.clip(xx, 0, w0 - 1)]
    return np.where(ok[..., None] if v.ndim == 3 else ok, v, v.dtype.type(0))


def warp_linear(img, Minv, out_h, out_w, mirror=False, h0=None, w0=None):
    """img [H,W,C] (any real dtype, read as float32) -> float32 [out_h,out_w,C]; (h0, w0) = valid region of a padded source."""
    h0, w0 = (img.shape[0] if h0 is None else h0), (img.shape[1] if w0 is None else w0)
    img = img.astype(F32)
    X, Y = _fixed_point_grid(Minv, out_h, out_w, mirror, 16)
    X, Y = X >> 5, Y >> 5
    sx, sy = X >> 5, Y >> 5
    fx, fy = (X & 31).astype(F32) / F32(32), (Y & 31).astype(F32) / F32(32)
    one = F32(1)
    w00, w01 = ((one - fy) * (one - fx))[..., None], ((one - fy) * fx)[..., None]
    w10, w11 = (fy * (one - fx))[..., None], (fy * fx)[..., None]
    return (_tap(img, sy, sx, h0, w0) * w00 + _tap(img, sy, sx + 1, h0, w0) * w01
            + _tap(img, sy + 1, sx, h0, w0) * w10 + _tap(img, sy + 1, sx + 1, h0, w0) * w11)


def warp_nearest(img, Minv, out_h, out_w, mirror=False, h0=None, w0=None):
    h0, w0 = (img.shape[0] if h0 is None else h0), (img.shape[1] if w0 is None else w0)
    X, Y = _fixed_point_grid(Minv, out_h, out_w, mirror, 512)
    return _tap(img, Y >> 10, X >> 10, h0, w0)


def _resize_coords(n_out, n_in, scale):
    f = (np.arange(n_out) + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(F32)
    lo = s < 0
    f, s = np.where(lo, F32(0), f), np.where(lo, 0, s)
    hi = s >= n_in - 1
    return np.where(hi, n_in - 1, s), np.where(hi, F32(0), f).astype(F32)


def resize_pad_linear(img, scale_x, scale_y, w_eff, h_eff, out_h, out_w, mirror=False, h0=None, w0=None):
    """cv2.resize(img, (w_eff, h_eff)) placed at the top-left of a zero [out_h, out_w] canvas (cropped if larger), then mirrored."""
    h0, w0 = (img.shape[0] if h0 is None else h0), (img.shape[1] if w0 is None else w0)
    img = img[:h0, :w0].astype(F32)
    sx, fx = _resize_coords(w_eff, w0, scale_x)
    sy, fy = _resize_coords(h_eff, h0, scale_y)
    sx1, sy1 = np.minimum(sx + 1, w0 - 1), np.minimum(sy + 1, h0 - 1)
    a0, a1 = (F32(1) - fx)[None, :, None], fx[None, :, None]
    b0, b1 = (F32(1) - fy)[:, None, None], fy[:, None, None]
    r0 = img[sy][:, sx] * a0 + img[sy][:, sx1] * a1
    r1 = img[sy1][:, sx] * a0 + img[sy1][:, sx1] * a1
    canvas = np.zeros((out_h, out_w, img.shape[2]), dtype=F32)
    h, w = min(h_eff, out_h), min(w_eff, out_w)
    canvas[:h, :w] = (r0 * b0 + r1 * b1)[:h, :w]
    return canvas[:, ::-1] if mirror else canvas


def resize_pad_nearest(img, scale_x, scale_y, w_eff, h_eff, out_h, out_w, mirror=False, h0=None, w0=None):
    h0, w0 = (img.shape[0] if h0 is None else h0), (img.shape[1] if w0 is None else w0)
    sx = np.minimum(np.floor(np.arange(w_eff) * scale_x).astype(np.int64), w0 - 1)
    sy = np.minimum(np.floor(np.arange(h_eff) * scale_y).astype(np.int64), h0 - 1)
    canvas = np.zeros((out_h, out_w), dtype=img.dtype)
    h, w = min(h_eff, out_h), min(w_eff, out_w)
    canvas[:h, :w] = img[sy][:, sx][:h, :w]
    return canvas[:, ::-1] if mirror else canvas


def rgb2hsv(img):
    r, g, b = img[..., 0], img[..., 1], img[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = diff / (np.abs(v) + EPS)
    d = F32(60) / (diff + EPS)
    h = np.where(v == r, (g - b) * d, np.where(v == g, (b - r) * d + F32(120), (r - g) * d + F32(240)))
    h = np.where(h < 0, h + F32(360), h)
    return np.stack([h, s, v], -1).astype(F32)


def hsv2rgb(img):
    h, s, v = img[..., 0] * F32(6.0 / 360.0), img[..., 1], img[..., 2]
    h = np.where(h < 0, h + F32(6) * np.ceil(-h / F32(6)), h)
    h = np.where(h >= 6, h - F32(6) * np.floor(h / F32(6)), h)
    sector = np.floor(h)
    f = h - sector
    sector = sector.astype(np.int64)
    f = np.where(sector >= 6, F32(0), f)
    sector = np.where(sector >= 6, 0, sector)
    one = F32(1)
    tab = np.stack([v, v * (one - s), v * (one - s * f), v * (one - s * (one - f))], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])[sector]        # (b, g, r) table columns
    pick = lambda k: np.take_along_axis(tab, sd[..., k:k + 1], -1)[..., 0]                          # noqa: E731
    out = np.stack([pick(2), pick(1), pick(0)], -1)
    return np.where((s == 0)[..., None], np.repeat(v[..., None], 3, -1), out).astype(F32)


def colour_chain(img, codes, values):
    """RandomBrightness / RandomContrast / (HSV, RandomSaturation, RGB) in the drawn order, on float32 [H,W,3] in 0..255."""
    for code, value in zip(codes, values):
        code = int(code)
        if code == OP_BRIGHTNESS:
            img = img + F32(value)
        elif code == OP_CONTRAST:
            img = img * F32(value)
        elif code == OP_SATURATION:
            hsv = rgb2hsv(img)
            if not np.isnan(value):                     # NaN = the saturation factor was not drawn: the round trip alone
                hsv[..., 1] *= F32(value)
            img = hsv2rgb(hsv)
    return img


def apply_plan(frames_u8, mask_u8, plan, out_h, out_w, mean, std):
    """frames_u8 [F,H,W,3], mask_u8 [H,W] or None, plan [PLAN_SIZE] float64 ->
    image [F,3,out_h,out_w] float32 (augmented, normalised), original [F,3,out_h,out_w] float32 (warped / 255), mask float64."""
    mirror = bool(plan[6])
    h0, w0 = int(plan[13]), int(plan[14])
    if int(plan[15]) == GEOM_RESIZE:
        geom = (float(plan[0]), float(plan[1]), int(plan[2]), int(plan[3]), out_h, out_w, mirror, h0, w0)
        linear, nearest = (lambda im: resize_pad_linear(im, *geom)), (lambda im: resize_pad_nearest(im, *geom))
    else:
        Minv = np.asarray(plan[0:6], dtype=np.float64).reshape(2, 3)
        linear = lambda im: warp_linear(im, Minv, out_h, out_w, mirror, h0, w0)          # noqa: E731
        nearest = lambda im: warp_nearest(im, Minv, out_h, out_w, mirror, h0, w0)        # noqa: E731
    mean, std = np.asarray(mean, dtype=F32), np.asarray(std, dtype=F32)
    images, originals = [], []
    for frame in frames_u8:
        warped = linear(frame)
        originals.append((warped / F32(255)).transpose(2, 0, 1))
        img = colour_chain(warped, plan[7:10], plan[10:13])
        images.append((((img / F32(255)) - mean) / std).transpose(2, 0, 1))
    mask = None if mask_u8 is None else nearest(mask_u8).astype(np.float64)
    return np.stack(images).astype(F32), np.stack(originals).astype(F32), mask
