"""Real YOLOv9-t weights + real frame: CUDA path vs fp32 oracle (and vs the reference's recorded detections)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "yolov9t_mot16.npz"))
P = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
frame = torch.from_numpy(z["frame"])
for res in (960, 640):
    want = o.detect("t", P, frame, res)[0]
    wq = o.detect("t", P, frame, res, quant="bf16")[0]
    m = YOLOv9("t", res, weights=P)
    got = torch.from_numpy(m(frame.numpy()).numpy())
    for tag, ref in (("fp32 oracle", want), ("bf16 mirror", wq)):
        A, B = ref[ref[:, 4] > 0], got[got[:, 4] > 0]
        d = (A[:, None, :4] - B[None, :, :4]).abs().max(-1)[0] + (A[:, None, 5] != B[None, :, 5]) * 1e6
        md, mi = d.min(1)
        ok = md < 1e5
        dc = (A[ok, 4] - B[mi[ok], 4]).abs()
        print(f"res {res} vs {tag}: oracle {len(A)} cuda {len(B)} class-matched {int(ok.sum())}  box |d| px: median {md[ok].median():.4f} "
              f"p90 {md[ok].quantile(0.9):.4f} max {md[ok].max():.4f}   conf |d|: median {dc.median():.5f} max {dc.max():.5f}  "
              f"rank-identical rows {(int(((ref - got).abs().max(1)[0] < 0.5).sum()))}/300")
