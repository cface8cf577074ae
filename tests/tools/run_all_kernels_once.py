"""One pass over every kernel kind of the hot path (for `ncu --set full`): detector forward (YOLOv9-c, B=32, 1080p frames ->
letterbox 640), CLIP ViT-B/32 image tower from device crops (crop_resize kernel), text tower, score search and top-k search."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import clip as oc
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9
from clearcam_b200.models.objects import ObjectFinder, search_scores

B = 32
fr = o.synthetic_frames(2, 1080, 1920, seed=0)
calib = torch.stack([o.preprocess(f, 640) for f in fr]).flip(-1).permute(0, 3, 1, 2).float() / 255
m = YOLOv9("c", 640, weights=o.synthetic_weights("c", seed=0, calib=calib))
frames = fr[torch.arange(B) % 2].cuda()
for _ in range(2):
    m.detect_batch(frames)
cfg = oc.CONFIGS["ViT-B/32"]
fin = ObjectFinder()
fin.init_clip(weights=oc.synthetic_weights(cfg, seed=0), arch="ViT-B/32")
rects = [(f % B, 100 + 7 * f, 50 + 3 * f, 400 + 7 * f, 420 + 3 * f) for f in range(64)]
for _ in range(2):
    emb = fin.embed_crops(frames, rects).tensor
q = fin.model.encode_text_batch(["a person walking a dog", "red car"]).tensor
sc = search_scores(emb, q)
torch.cuda.synchronize()
print("ok", tuple(emb.shape), tuple(sc.shape))
