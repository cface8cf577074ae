"""Golden vectors for the crop -> CLIP-input path from cv2 itself (the reference's dependency for models/objects.py:238).

TEST INFRASTRUCTURE.  cv2 is asked for its own bicubic (`cv2.ipp.setUseIPP(False)`; see oracle/clip_preprocess.py for why)
on crops of a synthetic street-like frame, through the reference's exact calls (clearcam.py:396 crop, models/objects.py:249
cvtColor, :238 resize).  Writes tests/golden/clip_preprocess.npz: frame (H,W,3) uint8 BGR, rects (K,4), resized
(K,224,224,3) uint8 RGB.

    python oracle/make_golden_preprocess.py
"""
from pathlib import Path

import cv2
import numpy as np

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def frame(H=540, W=720, seed=2):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    f = np.stack([96 + 80 * np.sin(xx / 37 + c) * np.cos(yy / 53 - c) for c in range(3)], -1)
    for _ in range(40):                                       # hard-edged rectangles so that the cubic overshoots and saturates
        x, y, w, h = g.integers(0, W - 40), g.integers(0, H - 40), g.integers(8, 160), g.integers(8, 160)
        f[y:y + h, x:x + w] = g.integers(0, 256, 3)
    f += g.normal(0, 6, f.shape)
    return np.clip(f, 0, 255).astype(np.uint8)


def main():
    cv2.ipp.setUseIPP(False)
    fr = frame()
    rects = np.array([[40, 30, 300, 420], [350, 100, 574, 324], [500, 0, 720, 131], [0, 400, 101, 540]], np.int32)
    res = []
    for x1, y1, x2, y2 in rects:
        crop = cv2.cvtColor(fr[y1:y2, x1:x2], cv2.COLOR_BGR2RGB)
        res.append(cv2.resize(crop, (224, 224), interpolation=cv2.INTER_CUBIC))
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "clip_preprocess.npz", frame=fr, rects=rects, resized=np.stack(res))
    print("wrote", OUT / "clip_preprocess.npz", cv2.__version__)


if __name__ == "__main__":
    main()
