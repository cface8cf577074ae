"""GPU timing of the CLIP towers. usage: diag_clip.py [arch] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import clip as oc
from clearcam_b200.models.objects import OpenCLIP
arch = sys.argv[1] if len(sys.argv) > 1 else "ViT-B/32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = oc.CONFIGS[arch]
P = oc.synthetic_weights(cfg, seed=0)
m = OpenCLIP(weights=P, arch=arch)
x = oc.synthetic_images(8, cfg.image_size, seed=1)[torch.arange(B) % 8].cuda()
for _ in range(2):
    prof, tot = m.profile(x=x)
ms = sum(r["ms"] for r in prof)
print(f"{arch} image B={B}: {ms:.3f} ms -> {B/ms*1000:.0f} img/s, {tot/ms/1e9:.1f} TFLOP/s")
agg = {}
for r in prof:
    a = agg.setdefault(r["name"], [0.0, 0.0, 0]); a[0] += r["ms"]; a[1] += r["flops"]; a[2] += 1
for k, (t, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"   {k:22s} x{n:3d} {t:9.3f} ms  {f/t/1e9 if t>0 else 0:8.1f} TFLOP/s  {100*t/ms:5.1f}%")
ids = torch.tensor([m.tokenize("a photo of a cat")] * 32)
for _ in range(2):
    prof, tot = m.profile(ids=ids)
ms = sum(r["ms"] for r in prof)
print(f"{arch} text B=32: {ms:.3f} ms -> {32/ms*1000:.0f} q/s, {tot/ms/1e9:.1f} TFLOP/s")
