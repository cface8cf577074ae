"""GPU diagnostic: per-layer deviation of the CUDA path from the bf16-mirror oracle (finds where errors enter)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9

size = sys.argv[1] if len(sys.argv) > 1 else "c"
res = int(sys.argv[2]) if len(sys.argv) > 2 else 320
B = 2
fr = o.synthetic_frames(B, res, res, seed=0)
x = fr.flip(-1).permute(0, 3, 1, 2).float() / 255
P = o.synthetic_weights(size, seed=0, calib=x)
tq, tf = [], []
with torch.no_grad():
    o.forward_raw(size, P, x, quant="bf16", taps=tq)
    o.forward_raw(size, P, x, taps=tf)
m = YOLOv9(size, res, weights=P)
m.detect_batch(fr)
torch.cuda.synchronize()
print("layer  shape                rms      cuda-vs-mirror(max,rms)    mirror-vs-fp32(max,rms)   frac>1ulp")
for i in range(len(tq)):
    if not isinstance(tq[i], torch.Tensor) or tq[i].dim() != 4 or tq[i].shape[1] == 3:
        continue
    g = m.layer_output(i, B, res, res)
    if g is None:
        continue
    g = g.cpu()
    d = (g - tq[i]).abs(); d2 = (tq[i] - tf[i]).abs()
    rms = tq[i].pow(2).mean().sqrt()
    ulp = tq[i].abs().clamp(min=1e-3) * 2 ** -8
    print(f"{i:3d}  {str(tuple(g.shape)):20s} {rms:7.3f}   {d.max():9.4g} {d.pow(2).mean().sqrt():9.4g}      {d2.max():9.4g} {d2.pow(2).mean().sqrt():9.4g}    {(d > ulp).float().mean():.4f}")
