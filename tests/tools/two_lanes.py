"""GPU experiment: does running the detector as L independent half/quarter-batch lanes on L streams beat one B=32 launch
chain?  (Layer-by-layer execution leaves every SM idle while a layer's last tiles drain and the next layer's first operands
arrive; a second, independent chain can fill those bubbles.)  usage: two_lanes.py [B] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
HW = 640
fr = o.synthetic_frames(2, HW, HW, seed=0)
P = o.synthetic_weights("c", seed=0, calib=fr.flip(-1).permute(0, 3, 1, 2).float() / 255)
base = o.synthetic_frames(4, HW, HW, seed=100)
nbuf = 6
g = torch.Generator(device="cuda").manual_seed(0)
batches = []
for i in range(nbuf):
    fb = base[torch.arange(B) % 4].cuda()
    batches.append(((fb // 2) + torch.randint(0, 8, fb.shape, device="cuda", dtype=torch.uint8, generator=g) + i).contiguous())


def timed(fn):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


ref_model = YOLOv9("c", HW, weights=P)
ref = ref_model.detect_batch(batches[0]).clone()
ms1 = timed(lambda i: ref_model.detect_batch(batches[i % nbuf]))
print(f"1 lane  x B={B}: {ms1:.3f} ms/step  {B / ms1 * 1e3:.0f} frames/s")

for L in (2, 4):
    if B % L:
        continue
    models = [ref_model] + [YOLOv9("c", HW, weights=P) for _ in range(L - 1)]
    streams = [torch.cuda.Stream() for _ in range(L)]
    h = B // L
    outs = [None] * L

    def step(i):
        main = torch.cuda.current_stream()
        fb = batches[i % nbuf]
        for l in range(L):
            streams[l].wait_stream(main)
            outs[l] = models[l].detect_batch(fb[l * h:(l + 1) * h], stream=streams[l])
        for l in range(L):
            main.wait_stream(streams[l])

    step(0)
    torch.cuda.synchronize()
    got = torch.cat(outs)
    same = torch.equal(got, ref)
    ms = timed(step)
    print(f"{L} lanes x B={h}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} frames/s  identical to one lane: {same}")
