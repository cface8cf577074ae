"""The crop -> CLIP-input oracle (oracle/clip_preprocess.py) against cv2's own arithmetic: committed golden vectors
(oracle/make_golden_preprocess.py) and, where cv2 is importable, cv2 itself over random shapes."""
from pathlib import Path

import numpy as np
import pytest

from oracle import clip_preprocess as cp

GOLD = Path(__file__).parent / "golden" / "clip_preprocess.npz"


def test_golden_vectors_from_cv2():
    g = np.load(GOLD)
    for (x1, y1, x2, y2), want in zip(g["rects"], g["resized"]):
        crop = np.ascontiguousarray(g["frame"][y1:y2, x1:x2, ::-1])
        np.testing.assert_array_equal(cp.resize_cubic_u8(crop, 224, 224), want)
    x = cp.preprocess_crops(g["frame"], g["rects"])
    assert x.shape == (4, 3, 224, 224) and x.dtype == np.float32
    np.testing.assert_array_equal(x[1], np.transpose((g["resized"][1].astype(np.float32) / 255.0 - 0.5) / 0.5, (2, 0, 1)))
    assert x.min() == -1.0 and x.max() == 1.0               # the rectangles make the cubic saturate both ways


def test_against_cv2_random_shapes():
    cv2 = pytest.importorskip("cv2")
    had = cv2.ipp.useIPP()
    g = np.random.default_rng(1)
    try:
        cv2.ipp.setUseIPP(False)                            # OpenCV's own code path (what non-IPP builds run)
        for t in range(60):
            H, W = int(g.integers(5, 500)), int(g.integers(5, 700))
            dw, dh = (224, 224) if t < 30 else (int(g.integers(3, 300)), int(g.integers(3, 300)))
            C = 3 if t < 30 else int(g.choice([1, 3, 4]))
            img = g.integers(0, 256, (H, W, C), dtype=np.uint8)
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC).reshape(dh, dw, C)
            np.testing.assert_array_equal(cp.resize_cubic_u8(img, dw, dh), ref, err_msg=f"{(H, W, C)} -> {(dh, dw)}")
        if had:                                             # IPP's bicubic: never more than one grey level away
            cv2.ipp.setUseIPP(True)
            img = g.integers(0, 256, (300, 400, 3), dtype=np.uint8)
            ipp = cv2.resize(img, (224, 224), interpolation=cv2.INTER_CUBIC).astype(int)
            d = np.abs(ipp - cp.resize_cubic_u8(img, 224, 224))
            assert d.max() <= 1 and (d > 0).mean() < 0.10
    finally:
        cv2.ipp.setUseIPP(had)


def test_crop_rect_follows_save_object():
    # clearcam.py:381-395 worked by hand: box 100.9,50.2,200.7,300.9 -> ints 100,50,200,300 -> centre 150,175, half 50,125
    # doubled to 100,250 (even) -> 50..250 x -75..425 -> clamped to the frame
    assert cp.crop_rect([100.9, 50.2, 200.7, 300.9], 640, 360) == (50, 0, 250, 360)
    assert cp.crop_rect([10, 10, 40, 200], 640, 360) is None                   # 2*(30//2*2)=60 wide < 100
    assert cp.crop_rect([600, 300, 700, 420], 640, 360) is None                # clamped to 550..640: 90 px wide
    assert cp.crop_rect([-20.5, 5, 101, 150], 640, 360) == (0, 0, 160, 221)    # int() truncates toward zero: -20
    from clearcam_b200.models.objects import ObjectFinder
    g = np.random.default_rng(0)
    for _ in range(200):
        b = np.sort(g.uniform(-50, 700, 4).reshape(2, 2), 0).T.reshape(-1)[[0, 2, 1, 3]]
        assert ObjectFinder.crop_rect(b, 640, 360) == cp.crop_rect(b, 640, 360)
