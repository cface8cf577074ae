"""Per-op device time table of one YOLOv9 forward (cc_yolo_profile). usage: prof_layers.py [size] [B] [res]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9
size = sys.argv[1] if len(sys.argv) > 1 else "c"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
res = int(sys.argv[3]) if len(sys.argv) > 3 else 640
fr = o.synthetic_frames(4, res, res, seed=0)
P = o.synthetic_weights(size, seed=0, calib=fr[:2].flip(-1).permute(0, 3, 1, 2).float() / 255)
m = YOLOv9(size, res, weights=P)
frames = fr[torch.arange(B) % 4].cuda()
for _ in range(3):
    prof = m.profile(frames)
tot = sum(r["ms"] for r in prof)
print(f"total {tot:.3f} ms for B={B}  ({B/tot*1000:.0f} fps)")
print(f"{'#':>3} {'kind':12s} {'name':34s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s':>7s} {'ideal_ms':>8s} {'gap_ms':>7s}")
ideal_tot = 0.0
for i, r in enumerate(prof):
    tf = r["flops"] / r["ms"] / 1e9 if r["ms"] > 0 else 0
    gb = r["bytes"] / r["ms"] / 1e6 if r["ms"] > 0 else 0
    ideal = max(r["flops"] / 1443e9, r["bytes"] / 6569e6)
    ideal_tot += ideal
    print(f"{i:3d} {r['kind']:12s} {r['name']:34s} {r['ms']:8.4f} {tf:8.1f} {gb:7.0f} {ideal:8.4f} {r['ms']-ideal:7.4f}")
print(f"sum of per-layer ideals (conv_gemm only): {ideal_tot:.3f} ms")
