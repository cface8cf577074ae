"""Result record of the tracker — the fields the reference's callers read (ocsort_tracker/STrack.py:4-18;
clearcam.py:583-621 uses tlwh, score, class_id, track_id, tracklet_len, speed)."""
import numpy as np


class STrack:
    __slots__ = ("_tlwh", "score", "class_id", "track_id", "tracklet_len", "speed")

    def __init__(self, tlwh, score, class_id, track_id=None, age=0, speed=0):
        self._tlwh = np.asarray(tlwh, dtype=np.float64)
        self.score, self.class_id, self.track_id, self.tracklet_len, self.speed = score, class_id, track_id, age, speed

    @property
    def tlwh(self):
        """(top-left x, top-left y, width, height)."""
        return self._tlwh.copy()

    @property
    def tlbr(self):
        t = self._tlwh.copy()
        t[2:] += t[:2]
        return t

    def __repr__(self):
        return f"STrack(id={int(self.track_id)}, cls={int(self.class_id)}, tlwh={self._tlwh.round(1).tolist()}, score={self.score:.3f})"
