"""GPU diagnostic: CUDA YOLOv9 path vs the CPU oracle (fp32 and bf16-mirror). Prints deviation statistics."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9


def match_stats(a, b, tag):
    """a,b: (300,6). set-match rows with conf>0 by class + box distance."""
    A = a[a[:, 4] > 0]; Bq = b[b[:, 4] > 0]
    if len(A) == 0 or len(Bq) == 0:
        print(f"   {tag}: n_oracle={len(A)} n_cuda={len(Bq)}"); return
    d = (A[:, None, :4] - Bq[None, :, :4]).abs().max(-1)[0] + (A[:, None, 5] != Bq[None, :, 5]) * 1e6
    md, mi = d.min(1)
    ok = md < 1e5
    print(f"   {tag}: n_oracle={len(A)} n_cuda={len(Bq)} matched(cls)={int(ok.sum())} box max|d| p50={md[ok].median():.4g} "
          f"p99={md[ok].quantile(0.99):.4g} max={md[ok].max():.4g}  conf max|d|={(A[ok,4]-Bq[mi[ok],4]).abs().max():.4g}  "
          f"rank-identical rows={int(((a-b).abs().max(1)[0]<1e-2).sum())}/300")


def run(size, res, B, H, W, seed=0, dtype=torch.uint8):
    print(f"== size {size} res {res} B {B} frame {H}x{W} {dtype}")
    fr = o.synthetic_frames(B, H, W, seed=seed)
    if dtype == torch.float32:
        fr = fr.float()
    pre = torch.stack([o.preprocess(f, res) for f in fr[:2]])
    calib = pre.flip(-1).permute(0, 3, 1, 2).float() / 255
    P = o.synthetic_weights(size, seed=seed, calib=calib)
    t = time.time(); ref = o.detect(size, P, fr, res); t_cpu = time.time() - t
    refq = o.detect(size, P, fr, res, quant="bf16")
    x = torch.stack([o.preprocess(f, res) for f in fr]).flip(-1).permute(0, 3, 1, 2).float() / 255
    with torch.no_grad():
        raw_ref = o.forward_raw(size, P, x); raw_q = o.forward_raw(size, P, x, quant="bf16")
    m = YOLOv9(size, res, weights=P)
    out, raw = m.detect_batch(fr, raw=True)
    torch.cuda.synchronize()
    out = out.cpu(); raw = raw.cpu()
    print("  plan:", m.plan_info(B, H, W, is_f32=dtype == torch.float32), f"cpu oracle {t_cpu/B*1000:.0f} ms/frame")
    for tag, r in (("raw vs fp32 oracle", raw_ref), ("raw vs bf16-mirror", raw_q)):
        db = (raw[:, :4] - r[:, :4]).abs(); dc = (raw[:, 4:] - r[:, 4:]).abs()
        print(f"  {tag}: box |d| mean={db.mean():.4g} p99={db.flatten().quantile(0.99) if db.numel()<1.6e7 else -1:.4g} max={db.max():.4g} ; prob |d| mean={dc.mean():.3g} max={dc.max():.3g}")
    print("  mirror-vs-fp32 oracle itself: box max", (raw_q[:, :4] - raw_ref[:, :4]).abs().max().item(), "prob max", (raw_q[:, 4:] - raw_ref[:, 4:]).abs().max().item())
    for b in range(min(B, 3)):
        match_stats(ref[b], out[b], f"img{b} final vs fp32  ")
        match_stats(refq[b], out[b], f"img{b} final vs mirror")
    # timing
    for _ in range(3): m.detect_batch(fr)
    torch.cuda.synchronize()
    frd = fr.cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.detect_batch(frd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    info = m.plan_info(B, H, W, is_f32=dtype == torch.float32)
    print(f"  GPU: {ms:.3f} ms/batch -> {B/ms*1000:.0f} fps, {info['conv_flops']/ms/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c320", "t320", "e320", "lb", "c640"]
    if "c320" in which: run("c", 320, 2, 320, 320)
    if "t320" in which: run("t", 320, 2, 320, 320)
    if "e320" in which: run("e", 320, 2, 320, 320)
    if "lb" in which: run("c", 320, 2, 270, 480); run("c", 320, 2, 270, 480, dtype=torch.float32)
    if "c640" in which: run("c", 640, 8, 640, 640)
    if "c640b32" in which: run("c", 640, 32, 640, 640)
