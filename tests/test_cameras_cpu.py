"""Host logic of the multi-camera batcher with a stand-in detector; the GPU run is in test_yolo_gpu.py."""
import numpy as np
import pytest
import torch

from clearcam_b200.cameras import CameraBatch
from clearcam_b200.ocsort_tracker import ocsort


class FakeDetector:
    """Returns one box whose position encodes the frame's mean value, so that results can be traced to inputs."""
    def __init__(self):
        self.calls = []

    def detect_batch(self, frames):
        self.calls.append(tuple(frames.shape))
        B = frames.shape[0]
        out = torch.zeros(B, 300, 6)
        m = frames.reshape(B, -1).float().mean(1)
        out[:, 0, 0], out[:, 0, 1], out[:, 0, 2], out[:, 0, 3] = m, m, m + 50, m + 80
        out[:, 0, 4], out[:, 0, 5] = 0.9, 2
        return out


def _frames(t):
    f = {f"cam{i}": np.full((48, 64, 3), 10 * i + t, np.uint8) for i in range(3)}
    f["wide"] = np.full((36, 96, 3), 100 + t, np.uint8)
    f["float"] = np.full((48, 64, 3), 70.0 + t, np.float32)
    f["idle"] = None
    return f


def test_groups_by_shape_and_dtype():
    g = CameraBatch.group_by_shape(_frames(0))
    assert sorted(map(len, g.values())) == [1, 1, 3] and "idle" not in sum(g.values(), [])
    with pytest.raises(ValueError):
        CameraBatch.group_by_shape({"bad": np.zeros((4, 4), np.uint8)})


def test_step_routes_results_to_cameras_and_tracks():
    det = FakeDetector()
    cb = CameraBatch(det)
    cb.add_camera("cam1", thresh=0.95)                      # its 0.9 box never passes this camera's threshold
    cb.add_camera("cam2", classes={0})                      # class filter drops class 2
    for t in range(4):
        res = cb.step(_frames(t))
    # one call per shape group; the 3-camera group is padded to the batch bucket 4 (zero frame, result dropped)
    assert sorted(det.calls[-3:]) == [(1, 36, 96, 3), (1, 48, 64, 3), (4, 48, 64, 3)]
    assert set(res) == {"cam0", "cam1", "cam2", "wide", "float"}
    for name, base in [("cam0", 0), ("wide", 100), ("float", 70)]:
        r = res[name]
        assert r.rows.shape == (300, 6) and abs(r.rows[0, 0] - (base + 3)) < 1e-4
        assert len(r.targets) == 1 and r.preds.shape == (1, 7)
        assert abs(r.preds[0, 0] - (base + 3)) < 1e-4 and r.preds[0, 6] == 1
    assert res["cam1"].targets == [] and res["cam2"].targets == [] and res["cam2"].preds.shape == (0, 7)
    # same answers as one tracker stepped alone on that camera's rows
    solo = ocsort.OCSort(max_age=100)
    for t in range(4):
        rows = det.detect_batch(torch.from_numpy(_frames(t)["cam0"])[None])[0].numpy()
        exp = solo.update(rows, 0.5)
    np.testing.assert_array_equal(exp[0].tlwh, res["cam0"].targets[0].tlwh)


def test_varying_camera_count_uses_a_bounded_set_of_batch_sizes():
    """The number of cameras with a new frame changes every step; the detector must only ever see bucket sizes (each
    distinct batch size is a cached plan in the library)."""
    from clearcam_b200.utils.helpers import batch_bucket
    assert [batch_bucket(n) for n in (1, 2, 3, 5, 9, 13, 17, 25, 33, 100)] == [1, 2, 4, 8, 12, 16, 24, 32, 48, 112]
    assert all(batch_bucket(n) >= n and batch_bucket(n) <= max(4 * n // 3 + 1, n + 15) for n in range(1, 300))
    det = FakeDetector()
    cb = CameraBatch(det)
    rng = np.random.default_rng(0)
    for t in range(40):
        k = int(rng.integers(1, 40))
        res = cb.step({f"c{i}": np.full((16, 16, 3), (i + t) % 200, np.uint8) for i in range(k)})
        assert len(res) == k and all(abs(res[f"c{i}"].rows[0, 0] - (i + t) % 200) < 1e-4 for i in range(k))
    assert {c[0] for c in det.calls} <= {1, 2, 4, 8, 12, 16, 24, 32, 48}
    assert len(cb._stage) <= 9
    # a removed and re-added camera (fresh mailbox, frame_num restarts) is not mistaken for "already seen"
    cb._seen["c0"] = 0
    cb.remove_camera("c0")
    assert "c0" not in cb._seen


def test_jit_infer_is_a_plain_call_keyed_by_shape():
    from clearcam_b200.utils.helpers import jit_infer
    cache, calls = {}, []
    f = lambda x: calls.append(x.shape) or x.sum()          # noqa: E731
    assert jit_infer(f, np.ones((2, 3)), cache) == 6 and jit_infer(f, np.ones((2, 3)), cache) == 6
    assert jit_infer(f, np.ones((4, 3)), cache) == 12
    assert set(cache) == {(2, 3), (4, 3)} and len(calls) == 3
