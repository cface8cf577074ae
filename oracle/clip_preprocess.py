"""CPU restatement of the crop -> CLIP-input path.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path is csrc/crop_resize.cu.

Reference lines restated:
  * `VideoCapture.save_object` crop rectangle, clearcam.py:373-397 — the track box grown to twice its size about its
    centre with integer arithmetic, clamped to the frame, dropped when a side is under 100 px  -> crop_rect()
  * `ObjectFinder.preprocess`, models/objects.py:237-242 — cv2.resize(img, (224,224), INTER_CUBIC), /255, (x-.5)/.5, CHW
    -> resize_cubic_u8() + normalize()

cv2.resize is a third-party dependency (opencv-python, unpinned in the reference's requirements); the arithmetic
restated is OpenCV's own 8-bit bicubic (imgproc resize.cpp: a = -0.75 kernel evaluated in float32, taps rounded to
11-bit fixed point, exact int32 horizontal pass, vertical pass in float32 as S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with
round-half-even and saturation; columns past the last full group of 8 interleaved elements take the integer
`(sum + 2^21) >> 22` form).  tests/test_clip_preprocess_cpu.py pins it bit-for-bit against cv2 itself with
`cv2.ipp.setUseIPP(False)`.  Wheels that bundle Intel IPP (x86 PyPI builds) route this call to IPP's closed-source
bicubic by default, which differs from OpenCV's own code by one grey level on about 4 % of the pixels (measured, same
test); builds without IPP (ARM, macOS, distro packages) run the arithmetic restated here."""
import numpy as np

_A = np.float32(-0.75)
_1, _2, _3, _4, _5, _8 = (np.float32(v) for v in (1, 2, 3, 4, 5, 8))


def cubic_taps(src: int, dst: int):
    """Per destination index: first source index - 1 ... and the four 11-bit fixed-point taps."""
    scale = 1.0 / (float(dst) / float(src))                       # hal::resize: scale = 1/inv_scale, float64
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    x = (f - s.astype(np.float32)).astype(np.float32)
    x1 = x + _1
    c0 = ((_A * x1 - _5 * _A) * x1 + _8 * _A) * x1 - _4 * _A
    c1 = ((_A + _2) * x - (_A + _3)) * x * x + _1
    xm = _1 - x
    c2 = ((_A + _2) * xm - (_A + _3)) * xm * xm + _1
    c3 = _1 - c0 - c1 - c2
    taps = np.rint(np.stack([c0, c1, c2, c3], -1).astype(np.float32) * np.float32(2048)).astype(np.int64)
    return s, taps


def resize_cubic_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """img: (H,W,C) uint8 -> (dh,dw,C) uint8, OpenCV INTER_CUBIC arithmetic."""
    H, W, C = img.shape
    if (H, W) == (dh, dw):
        return img.copy()
    sx, ax = cubic_taps(W, dw)
    sy, ay = cubic_taps(H, dh)
    xi = np.clip(sx[:, None] + np.arange(-1, 3)[None], 0, W - 1)
    yi = np.clip(sy[:, None] + np.arange(-1, 3)[None], 0, H - 1)
    hor = (img.astype(np.int64)[:, xi, :] * ax[None, :, :, None]).sum(2)          # (H,dw,C) exact integers
    rows = hor[yi]                                                                   # (dh,4,dw,C)
    b = (ay.astype(np.float32) * np.float32(1.0 / (2048 * 2048))).astype(np.float32)
    r = rows.astype(np.float32)
    t = r[:, 3] * b[:, 3, None, None]
    for k in (2, 1, 0):
        t = r[:, k] * b[:, k, None, None] + t                                        # float32, one rounding per op
    out = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    tail = (dw * C) // 8 * 8                                                         # elements past the last SIMD group
    if tail < dw * C:
        v = (rows * ay[:, :, None, None]).sum(1)
        vi = np.clip((v + (1 << 21)) >> 22, 0, 255).astype(np.uint8).reshape(dh, dw * C)
        flat = out.reshape(dh, dw * C)
        flat[:, tail:] = vi[:, tail:]
        out = flat.reshape(dh, dw, C)
    return out


def normalize(img_u8: np.ndarray) -> np.ndarray:
    """(S,S,3) uint8 -> (3,S,S) float32, models/objects.py:239-241."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.transpose(x, (2, 0, 1))


def crop_rect(box_xyxy, W: int, H: int, min_side: int = 100):
    """clearcam.py:381-395.  box: x1,y1,x2,y2 (floats, truncated like int()).  -> (x1,y1,x2,y2) ints or None."""
    x1, y1, x2, y2 = (int(v) for v in box_xyxy)
    cx, cy = (x1 + x2) // 2, (y1 + y2) // 2
    hw, hh = (x2 - x1) // 2 * 2, (y2 - y1) // 2 * 2
    nx1, nx2 = max(0, min(cx - hw, W)), max(0, min(cx + hw, W))
    ny1, ny2 = max(0, min(cy - hh, H)), max(0, min(cy + hh, H))
    if (ny2 - ny1) < min_side or (nx2 - nx1) < min_side:
        return None
    return nx1, ny1, nx2, ny2


def preprocess_crops(frame_bgr: np.ndarray, rects, size: int = 224) -> np.ndarray:
    """frame (H,W,3) BGR uint8, rects [(x1,y1,x2,y2)] -> (K,3,size,size) float32: crop (clearcam.py:396), BGR->RGB
    (models/objects.py:249), preprocess (:237-242) — the reference's path minus its JPEG write/read in between."""
    out = np.empty((len(rects), 3, size, size), np.float32)
    for i, (x1, y1, x2, y2) in enumerate(rects):
        crop = frame_bgr[y1:y2, x1:x2, ::-1]
        out[i] = normalize(resize_cubic_u8(np.ascontiguousarray(crop), size, size))
    return out
