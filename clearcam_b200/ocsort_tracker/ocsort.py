"""`OCSort` with the reference's constructor and `update` contract (ocsort_tracker/ocsort.py:163-177), running in the
C++ tracker of libclearcam_b200 (csrc/ocsort.cu) instead of per-frame numpy.

    tracker = ocsort.OCSort(max_age=100)                 # clearcam.py:239
    online_targets = tracker.update(preds, thresh)       # clearcam.py:585, preds = the detector's (300,6) array

`update_many` steps the trackers of several cameras on one batched detector result in a single library call."""
import ctypes
from typing import List, Optional, Sequence

import numpy as np

from .._lib import check, lib
from .STrack import STrack

_CAP = 1024          # rows of the result buffer (the detector emits at most 300 boxes per frame)


def _rows(output_results) -> np.ndarray:
    a = output_results.numpy() if hasattr(output_results, "numpy") and not isinstance(output_results, np.ndarray) else output_results
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 6:
        raise ValueError("detections must be (n, 6): x1, y1, x2, y2, score, class")
    return a if a.shape[1] == 6 else np.ascontiguousarray(a[:, :6])


def _tracks(buf: np.ndarray, n: int) -> List[STrack]:
    return [STrack(tlwh=r[:4], score=r[4], class_id=r[5], track_id=r[6], age=r[7], speed=r[8]) for r in buf[:n]]


class OCSort:
    def __init__(self, det_thresh=0.25, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3, asso_func="iou", inertia=0.2,
                 use_byte=False):
        if asso_func != "iou":
            raise ValueError("only the IoU association the reference uses is implemented")    # ocsort.py:175
        self.max_age, self.min_hits, self.iou_threshold, self.delta_t = max_age, min_hits, iou_threshold, delta_t
        self.inertia, self.use_byte = inertia, use_byte
        self.frame_count = 0
        self._h = ctypes.c_void_p()
        check(lib().cc_ocsort_create(int(max_age), int(min_hits), float(iou_threshold), int(delta_t), float(inertia),
                                     int(bool(use_byte)), ctypes.byref(self._h)), "cc_ocsort_create")
        self._out = np.empty((_CAP, 9), np.float64)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().cc_ocsort_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __len__(self):
        return lib().cc_ocsort_num_tracks(self._h)

    def update(self, output_results, det_thresh=0.25) -> List[STrack]:
        if output_results is None:                          # ocsort.py:185-186
            return np.empty((0, 5))
        a = _rows(output_results)
        n = ctypes.c_int()
        check(lib().cc_ocsort_update(self._h, a.ctypes.data, a.shape[0], float(det_thresh), self._out.ctypes.data, _CAP,
                                     ctypes.byref(n)), "cc_ocsort_update")
        self.frame_count += 1
        return _tracks(self._out, n.value)


def update_many(trackers: Sequence[Optional[OCSort]], results, det_thresh) -> List[List[STrack]]:
    """Step tracker b on results[b] (a (B,n,6) float32 array — one batched detector output) for every b; entries of
    `trackers` may be None (camera without a tracker).  det_thresh: scalar or one value per camera."""
    r = results.numpy() if hasattr(results, "numpy") and not isinstance(results, np.ndarray) else results
    r = np.ascontiguousarray(r, dtype=np.float32)
    B = len(trackers)
    if r.ndim != 3 or r.shape[0] != B or r.shape[2] != 6:
        raise ValueError("results must be (len(trackers), n, 6)")
    thr = np.ascontiguousarray(np.broadcast_to(np.asarray(det_thresh, np.float32), (B,)))
    hs = (ctypes.c_void_p * B)(*[t._h if t is not None else None for t in trackers])
    out = np.empty((B, _CAP, 9), np.float64)
    n = np.zeros(B, np.int32)
    check(lib().cc_ocsort_update_batch(hs, B, r.ctypes.data, r.shape[1], thr.ctypes.data, out.ctypes.data, _CAP, n.ctypes.data),
          "cc_ocsort_update_batch")
    for t in trackers:
        if t is not None:
            t.frame_count += 1
    return [_tracks(out[b], int(n[b])) if trackers[b] is not None else [] for b in range(B)]
