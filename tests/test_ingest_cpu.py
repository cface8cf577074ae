"""Frame ingest: rawvideo bgr24 bytes from a pipe -> mailbox -> batched detector, host logic on the CPU."""
import io
import os
import threading
import time

import numpy as np
import torch

from clearcam_b200.cameras import CameraBatch
from clearcam_b200.ingest import FrameMailbox, PipeReader

H, W = 24, 40


def _frame(k):
    return np.full((H, W, 3), k % 251, np.uint8) + np.arange(W, dtype=np.uint8)[None, :, None]


def test_mailbox_reads_whole_frames_from_a_chunked_stream():
    class Dribble(io.RawIOBase):                      # a pipe delivers whatever is available: here 1000 bytes at a time
        def __init__(self, data): self.data, self.pos = data, 0
        def readable(self): return True
        def readinto(self, b):
            n = min(len(b), 1000, len(self.data) - self.pos)
            b[:n] = self.data[self.pos:self.pos + n]
            self.pos += n
            return n
    data = b"".join(_frame(k).tobytes() for k in range(3)) + _frame(3).tobytes()[:100]     # last frame cut short
    mb, s = FrameMailbox(H, W, pin=False), Dribble(data)
    assert mb.latest() is None
    assert mb.fill(s) and mb.frame_num == 0
    n, f = mb.latest()
    assert n == 0 and f.shape == (H, W, 3) and np.array_equal(f.numpy(), _frame(0))
    assert mb.latest(last_seen=0) is None             # nothing new (clearcam.py:446)
    assert mb.fill(s) and mb.fill(s)                  # two more frames arrive while the consumer still holds frame 0
    assert np.array_equal(f.numpy(), _frame(0))       # ... which stays intact
    n, g = mb.latest(last_seen=0)
    assert n == 2 and np.array_equal(g.numpy(), _frame(2))
    assert not mb.fill(s) and mb.frame_num == 2       # short read: not published


def test_pipe_reader_thread_and_restart_after_failed_reads():
    opened = []

    def open_stream():
        r, w = os.pipe()
        k0 = 10 * len(opened)
        opened.append(w)

        def feed():
            try:
                with os.fdopen(w, "wb") as out:
                    for k in range(k0, k0 + 4):
                        out.write(_frame(k).tobytes())
                        out.flush()
                        time.sleep(0.01)
                    out.write(b"\0" * 7)              # stream dies mid-frame; EOF follows
            except (BrokenPipeError, OSError):        # the reader went away first (end of the test)
                pass
        threading.Thread(target=feed, daemon=True).start()
        return os.fdopen(r, "rb", buffering=0)

    mb = FrameMailbox(H, W, pin=False)
    rd = PipeReader(mb, open_stream, max_fail=2, retry_sleep=0.01)
    rd.start()
    seen, last, t0 = [], -1, time.time()
    while time.time() - t0 < 10 and not (rd.restarts >= 1 and len(seen) >= 6):
        got = mb.latest(last)
        if got is None:
            time.sleep(0.002)
            continue
        last, f = got
        seen.append(int(f[0, 0, 0]))
    rd.stop()
    assert rd.restarts >= 1 and len(opened) >= 2      # more than max_fail short reads in a row -> stream reopened (:408-411)
    assert seen == sorted(seen) and any(v >= 10 for v in seen) and any(v < 10 for v in seen)


def test_camera_batch_takes_unseen_frames_from_mailboxes():
    class Det:
        def detect_batch(self, frames):
            B = frames.shape[0]
            out = torch.zeros(B, 300, 6)
            out[:, 0] = torch.tensor([1.0, 2.0, 60.0, 90.0, 0.9, 0.0])
            out[:, 0, 0] = frames.reshape(B, -1)[:, 0].float()
            return out
    boxes = {n: FrameMailbox(H, W, pin=False) for n in ("a", "b")}
    boxes["a"].fill(io.BytesIO(_frame(5).tobytes()))
    cb = CameraBatch(Det())
    res = cb.step_mailboxes(boxes)
    assert set(res) == {"a"} and res["a"].rows[0, 0] == 5          # b has no frame yet
    assert cb.step_mailboxes(boxes) == {}                            # nothing new
    boxes["a"].fill(io.BytesIO(_frame(6).tobytes()))
    boxes["b"].fill(io.BytesIO(_frame(9).tobytes()))
    res = cb.step_mailboxes(boxes)
    assert set(res) == {"a", "b"} and res["a"].rows[0, 0] == 6 and res["b"].rows[0, 0] == 9
    assert len(res["a"].targets) == 1 and res["a"].targets[0].tracklet_len >= 1


def test_reader_survives_a_camera_that_is_down_and_stop_unblocks_a_silent_one():
    """open_stream() raising (ffmpeg / camera down) must not kill the reader thread: it retries with back-off and counts
    the failures; stop() closes the stream so a read blocked on a silent camera returns (the reference's frame_loop keeps
    retrying, clearcam.py:401-421)."""
    attempts = []
    r_fd, w_fd = os.pipe()                                  # a camera that connects but never sends a byte

    def open_stream():
        attempts.append(time.time())
        if len(attempts) <= 3:
            raise OSError("camera unreachable")
        return os.fdopen(r_fd, "rb", buffering=0)

    mb = FrameMailbox(H, W, pin=False)
    rd = PipeReader(mb, open_stream, max_fail=2, retry_sleep=0.01)
    rd.start()
    t0 = time.time()
    while time.time() - t0 < 5 and len(attempts) < 4:
        time.sleep(0.005)
    assert rd.is_alive() and rd.open_errors == 3 and isinstance(rd.last_error, OSError)
    time.sleep(0.05)                                        # now blocked in readinto() on the silent pipe
    os.write(w_fd, _frame(7).tobytes())
    t0 = time.time()
    while time.time() - t0 < 5 and mb.latest(-1) is None:
        time.sleep(0.005)
    assert mb.latest(-1) is not None                        # the stream opened on the 4th attempt delivers
    rd.stop()
    os.close(w_fd)
    rd.join(timeout=5)
    assert not rd.is_alive()
