"""One CLIP image-tower forward (for ncu). usage: run_clip_once.py [arch] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import clip as oc
from clearcam_b200.models.objects import OpenCLIP

arch = sys.argv[1] if len(sys.argv) > 1 else "ViT-L/14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = oc.CONFIGS[arch]
m = OpenCLIP(weights=oc.synthetic_weights(cfg, seed=0), arch=arch)
x = oc.synthetic_images(8, cfg.image_size, seed=1)[torch.arange(B) % 8].cuda()
for _ in range(2):
    m.precompute_embedding(x)
torch.cuda.synchronize()
print("done")
