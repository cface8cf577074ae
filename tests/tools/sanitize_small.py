"""Small forwards of both models for compute-sanitizer (memcheck / racecheck): YOLOv9-c 2 frames 192x256 -> res 256 (letterbox,
fused upsample/concat, SPP3, pools, head, decode, postprocess, standalone postprocess), the fp32-accurate mode of the same, CLIP ViT-tiny image + text.
usage: compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import clip as oc
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9, postprocess
from clearcam_b200.models.objects import OpenCLIP

fr = o.synthetic_frames(2, 192, 256, seed=11)
pre = torch.stack([o.preprocess(f, 256) for f in fr])
x = pre.flip(-1).permute(0, 3, 1, 2).float() / 255
P = o.synthetic_weights("c", seed=11, calib=x)
for precise in (False, True):
    m = YOLOv9("c", 256, weights=P, precise=precise)
    out, raw = m.detect_batch(fr, raw=True)
    again = postprocess(raw).tensor          # the standalone postprocess(output) of the reference's API on the head tap
    torch.cuda.synchronize()
    print("detector", "fp32-accurate" if precise else "default", int((out[..., 4] > 0).sum()), int((again[..., 4] > 0).sum()))
    del m
cfg = oc.CONFIGS["ViT-tiny"]
cm = OpenCLIP(weights=oc.synthetic_weights(cfg, seed=3), arch="ViT-tiny")
e = cm.precompute_embedding(oc.synthetic_images(5, cfg.image_size, seed=5)).tensor
t = cm.encode_text_batch(["a person", "a red car"]).tensor
torch.cuda.synchronize()
print("clip", tuple(e.shape), tuple(t.shape))
