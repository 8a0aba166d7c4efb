"""ORACLE -- test infrastructure only (never imported by the product path).

numpy restatement of the reference's image pre-processing, the step right before HydraNet.forward
(SURVEY.md section 8 row f-1):

  model/demo.py:191-196
      img = cv2.cvtColor(input_img, cv2.COLOR_BGR2RGB)
      img = cv2.resize(img, net_input_size)                 # uint8, INTER_LINEAR
      img = img.astype(np.float32)
      img = imagenet_normalize(img)                         # demo.py:26-40, float64 arithmetic
      img = np.expand_dims(np.transpose(img, (2, 0, 1)), 0)
      img = torch.tensor(img).cuda().float()

cv2.resize is third-party (OpenCV; un-vendored, un-pinned by the reference; this image has 4.13.0).
Its uint8 INTER_LINEAR path is restated here from its published algorithm (imgproc/resize.cpp):
  * scale = 1 / (dst / src) in double; source coordinate fx = (float)((dx + 0.5) * scale - 0.5);
    sx = floor(fx); fx -= sx
  * horizontally, taps outside the row collapse onto the border pixel with weight 1 (fx = 0);
    vertically the fraction is KEPT and the two row indices are clipped (both rows may be the same)
  * coefficients are rounded to 11-bit fixed point: saturate_cast<short>(w * 2048) (round half even)
  * horizontal pass in int32: S0*a0 + S1*a1; vertical pass
        dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
  * an exact 2x2 down-scale in both directions switches to INTER_AREA: (a + b + c + d + 2) >> 2
Pinned bit-for-bit against cv2 4.13.0 in tests/test_cpu_oracle.py (20 size pairs: up, down, 2x area,
identity, odd sizes) and against tests/golden/preprocess.npz, which oracle/make_golden.py generates by
running cv2 and the reference's own imagenet_normalize (loaded from /root/reference at generation time).
"""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406])
STD = np.array([0.229, 0.224, 0.225])


def linear_coeffs(dst, src, clamp_fraction):
    """(index int32[dst], w0 int32[dst], w1 int32[dst]) of cv2's fixed-point INTER_LINEAR along one axis."""
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:  # x direction (resize.cpp: "if( sx < 0 ) fx = 0, sx = 0; if( sx >= ssize.width-1 ) fx = 0, sx = width-1")
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return s, w0, w1


def resize_u8(img, out_w, out_h):
    """cv2.resize(img, (out_w, out_h)) for uint8 HxWxC, default interpolation."""
    h, w = img.shape[:2]
    i = img.astype(np.int32)
    if w == 2 * out_w and h == 2 * out_h:
        return ((i[0::2, 0::2] + i[0::2, 1::2] + i[1::2, 0::2] + i[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, ax0, ax1 = linear_coeffs(out_w, w, True)
    sy, by0, by1 = linear_coeffs(out_h, h, False)
    sx1 = np.minimum(sx + 1, w - 1)
    rows = i[:, sx] * ax0[None, :, None] + i[:, sx1] * ax1[None, :, None]  # [h, out_w, C]
    r0 = rows[np.clip(sy, 0, h - 1)]
    r1 = rows[np.clip(sy + 1, 0, h - 1)]
    out = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def imagenet_normalize(img):
    """demo.py:26-40: float32 image / int64 array -> float64; (x / 255 - mean) / std in float64."""
    img = img / np.array([255, 255, 255])
    img = img - MEAN
    img = img / STD
    return img


def preprocess(img_bgr, out_w, out_h):
    """demo.py:191-196 for one HxWx3 uint8 BGR image -> float32 [3, out_h, out_w]."""
    rgb = img_bgr[:, :, ::-1]
    r = resize_u8(np.ascontiguousarray(rgb), out_w, out_h).astype(np.float32)
    return np.transpose(imagenet_normalize(r), (2, 0, 1)).astype(np.float32)


def normalize_lut():
    """float32 [3 (RGB), 256]: the value the pipeline above yields for every uint8 level."""
    v = np.arange(256, dtype=np.float32)[None, :].astype(np.float64)  # float32 / int64 array promotes to float64
    return (((v / 255.0) - MEAN[:, None]) / STD[:, None]).astype(np.float32)
