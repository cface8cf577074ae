"""GPU diagnostic: the fp32-accurate mode (and the default mode) against the fp32 CPU oracle — unconditioned deviations of
the head output (every anchor) and of the final detections. usage: diag_precise.py [size res B H W]..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import yolov9 as o
from clearcam_b200.detection.yolov9 import YOLOv9


def stats(tag, raw, ref):
    db = (raw[:, :4] - ref[:, :4]).abs().flatten()
    dp = (raw[:, 4:] - ref[:, 4:]).abs().flatten()
    q = lambda t, p: float(torch.quantile(t[:: max(1, t.numel() // 2000000)], p))
    ids = (raw[:, 4:].argmax(1) == ref[:, 4:].argmax(1))
    conf = ref[:, 4:].max(1)[0] >= 0.25
    print(f"  {tag}: box |d| px p50={q(db,.5):.3g} p99={q(db,.99):.3g} p99.9={q(db,.999):.3g} max={db.max():.3g} | prob |d| p50={q(dp,.5):.3g} "
          f"p99={q(dp,.99):.3g} max={dp.max():.3g} | argmax agree all={ids.float().mean():.6f} on conf>=.25 anchors={ids[conf].float().mean() if conf.any() else 1:.6f} (n={int(conf.sum())})")


def final(tag, ref, out):
    for b in range(min(len(ref), 2)):
        A, Bq = ref[b][ref[b][:, 4] > 0], out[b][out[b][:, 4] > 0]
        same = len(A) == len(Bq)
        if same and len(A):
            d = (A - Bq).abs()
            print(f"  {tag} img{b}: n={len(A)} rows in the same order: box max|d|={d[:, :4].max():.3g} conf max|d|={d[:, 4].max():.3g} class equal={bool((A[:, 5] == Bq[:, 5]).all())}")
        else:
            print(f"  {tag} img{b}: n_oracle={len(A)} n_cuda={len(Bq)}")


def run(size, res, B, H, W, seed=7):
    print(f"== size {size} res {res} B {B} frame {H}x{W}")
    fr = o.synthetic_frames(B, H, W, seed=seed)
    pre = torch.stack([o.preprocess(f, res) for f in fr])
    x = pre.flip(-1).permute(0, 3, 1, 2).float() / 255
    P = o.synthetic_weights(size, seed=seed, calib=x[:2])
    with torch.no_grad():
        raw_ref = o.forward_raw(size, P, x)
        torch.set_num_threads(1)
        raw_ref1 = o.forward_raw(size, P, x[:1])
        torch.set_num_threads(8)
    ref = o.detect(size, P, fr, res)
    stats("fp32 oracle 1 thread vs N threads (noise floor)", raw_ref1, raw_ref[:1])
    for precise in (True, False):
        m = YOLOv9(size, res, weights=P, precise=precise)
        out, raw = m.detect_batch(fr, raw=True)
        torch.cuda.synchronize()
        out, raw = out.cpu(), raw.cpu()
        tag = "fp32-accurate" if precise else "default bf16 "
        stats(tag, raw, raw_ref)
        final(tag, ref, out)
        frd = fr.cuda()
        for _ in range(2): m.detect_batch(frd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m.detect_batch(frd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"  {tag}: {ms:.3f} ms/batch -> {B/ms*1000:.0f} fps; workspace {m.plan_info(B, H, W)['act_bytes']/1e6:.0f} MB")
        del m


if __name__ == "__main__":
    a = sys.argv[1:]
    cases = [("c", 320, 2, 320, 320), ("c", 640, 2, 640, 640), ("t", 320, 2, 320, 320), ("e", 256, 2, 256, 256)] if not a else \
        [(a[i], int(a[i + 1]), int(a[i + 2]), int(a[i + 3]), int(a[i + 4])) for i in range(0, len(a), 5)]
    for c in cases:
        run(*c)
