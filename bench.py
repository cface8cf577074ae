#!/usr/bin/env python
"""Headline benchmark (driver contract in the task statement).

Workload at N GPUs: YOLOv9-c, 32 synthetic 640x640x3 uint8 BGR frames per GPU per step (BASELINE.json configs[1],
weak scaling: frames shard by batch, no data-path collective), whole path = stem(/255,BGR flip) -> 144 convs ->
DFL decode -> top-300 + suppression -> scale_boxes.  Weights: seeded synthetic (oracle.synthetic_weights).

  value     : frames/s with the uint8 frames already resident in HBM (rotating through > L2-size worth of inputs)
  e2e       : frames/s through the public API (YOLOv9.detect_batch) from PINNED HOST frames, H2D and the D2H read
              of the (B,300,6) result inside the timed region
  roofline  : conv_gemm_kernel (tcgen05) = algorithmic conv FLOPs per step / summed device time of its launches,
              measured live with CUDA events (cc_yolo_profile), against MEASURED_PEAKS.json bf16 peak
  cpu_baseline / --impl reference : the torch-CPU oracle (the reference's tinygrad path cannot run here) on the
              box's host cores, bounded sample.

oracle/ is imported here for two things only: the CPU legs above, and — before any timed region — the seeded synthetic
weights and frames every arm runs on (input generation).  Every timed GPU region calls clearcam_b200 alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SIZE, RES, BATCH, HW = "c", 640, 32, 640
GFLOP_PER_FRAME = 102.14  # SURVEY.md §8(d), YOLOv9-c 640x640 (2*MAC over convs)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tflops": p["bf16_tflops"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"tflops": 1590.0, "tflops_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_weights():
    from oracle import yolov9 as o
    fr = o.synthetic_frames(2, HW, HW, seed=0)
    x = fr.flip(-1).permute(0, 3, 1, 2).float() / 255
    return o.synthetic_weights(SIZE, seed=0, calib=x)


def pick_cpu_threads(P, fr):
    """The torch-CPU oracle is NOT fastest with every hardware thread of a 128-thread host (measured: 0.5 frames/s with 128
    threads against 4.8 on 8): give the CPU side its best thread count — one warm-up, then one timed frame per candidate."""
    from oracle import yolov9 as o
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, ncpu // 4, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        o.detect(SIZE, P, fr[:1], RES)
        t0 = time.time()
        o.detect(SIZE, P, fr[:1], RES)
        dt = time.time() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_fps(P, seconds_budget=20.0, batch=4):
    """torch-CPU oracle on the host threads that serve it best, bounded sample of the same workload."""
    from oracle import yolov9 as o
    fr = o.synthetic_frames(batch, HW, HW, seed=1)
    pick_cpu_threads(P, fr)
    n, t0 = 0, time.time()
    while True:
        o.detect(SIZE, P, fr, RES)
        n += batch
        if time.time() - t0 > seconds_budget or n >= 64:
            break
    dt = time.time() - t0
    return n / dt, n, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = make_weights()
    from oracle import yolov9 as o
    batch = 4   # measured: 4 frames per call is the fastest per-frame configuration of the torch-CPU oracle
    fr = o.synthetic_frames(batch, HW, HW, seed=1)
    pick_cpu_threads(P, fr)
    for _ in range(max(1, min(args.warmup, 3))):
        o.detect(SIZE, P, fr, RES)
    steps = max(1, min(args.steps, 12))
    t0 = time.time()
    for _ in range(steps):
        o.detect(SIZE, P, fr, RES)
    dt = time.time() - t0
    fps = steps * batch / dt
    line = {"impl": "reference", "metric": "frames/s YOLOv9-c 640px", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1000, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"YOLOv9-c {HW}x{HW} uint8 frames, CPU oracle (torch fp32 restatement of detection/yolov9.py; "
                                   f"tinygrad DEV=CPU cannot run here), {batch} frames/step"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{steps} steps x {batch} frames"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the seeded synthetic weights are generated on the CPU by every rank: share the host cores instead of running
        # world x all-threads on top of each other
        torch.set_num_threads(max(1, min(16, (os.cpu_count() or 8) // world)))
    else:
        torch.set_num_threads(min(16, os.cpu_count() or 8))    # 16 threads serve the torch-CPU weight generation best
    from oracle import yolov9 as o
    from clearcam_b200.detection.yolov9 import YOLOv9

    B = args.batch
    W = max(args.warmup, 3)
    K = args.steps
    P = make_weights()
    model = YOLOv9(SIZE, RES, weights=P)
    # > L2 (126 MB) worth of distinct device-resident input batches, rotated between steps
    nbuf = max(2, int(160e6 // (B * HW * HW * 3)) + 1)
    base = o.synthetic_frames(4, HW, HW, seed=100 + rank)
    g = torch.Generator(device="cuda").manual_seed(rank)
    dev_batches = []
    for i in range(nbuf):
        idx = torch.arange(B) % 4
        fb = base[idx].cuda()
        noise = torch.randint(0, 8, fb.shape, device="cuda", dtype=torch.uint8, generator=g)
        dev_batches.append(((fb // 2) + noise + i).contiguous())
    host_batches = [b.cpu().pin_memory() for b in dev_batches[:2]]
    info = model.plan_info(B, HW, HW)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for i in range(W):
        model.detect_batch(dev_batches[i % nbuf])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        model.detect_batch(dev_batches[i % nbuf])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total / 1000.0)

    # ---- end to end through the public API from pinned host memory
    for _ in model.detect_pipelined(host_batches[i % 2] for i in range(W)):
        pass
    barrier()
    e0.record()
    n_out = 0
    for r in model.detect_pipelined(host_batches[i % 2] for i in range(K)):
        n_out += r.shape[0]
    e1.record()
    barrier()
    assert n_out == B * K
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * B * K / (float(t.item()) / 1000.0)

    # ---- CLIP image tower (the second hot loop): device-resident crops, all ranks, in-place all-gather when N>1
    clip_res = {}
    try:
        from oracle import clip as oc
        from clearcam_b200.models.objects import OpenCLIP
        for arch, cb in (("ViT-B/32", 256), ("ViT-L/14", 64)):
            cfg = oc.CONFIGS[arch]
            cm = OpenCLIP(weights=oc.synthetic_weights(cfg, seed=0), arch=arch)
            xs = [oc.synthetic_images(8, cfg.image_size, seed=10 + i)[torch.arange(cb) % 8].cuda() for i in range(3)]
            for i in range(3):
                cm.precompute_embedding(xs[i % 3], gather=world > 1)
            barrier()
            steps_c = max(3, min(K, 10))
            e0.record()
            for i in range(steps_c):
                cm.precompute_embedding(xs[i % 3], gather=world > 1)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ips = world * cb * steps_c / (float(t.item()) / 1000.0)
            clip_res[arch] = {"batch_per_gpu": cb, "images_per_s": ips, "tflops": ips * oc.flops_image(cfg) / 1e12,
                              "gflop_per_image": oc.flops_image(cfg) / 1e9, "all_gather": world > 1}
            del cm, xs
            torch.cuda.empty_cache()
    except Exception as ex:  # the detector line must still be printed
        clip_res = {"error": repr(ex)}

    # ---- roofline of the dominant kernel (rank 0): live per-op CUDA-event timing
    line = None
    if rank == 0:
        pk = peaks()
        prof = None
        for i in range(3):
            prof = model.profile(dev_batches[i % nbuf])
        by = {}
        for r in prof:
            d = by.setdefault(r["kind"], {"ms": 0.0, "flops": 0.0, "n": 0})
            d["ms"] += r["ms"]; d["flops"] += r["flops"]; d["n"] += 1
        gm = by.get("conv_gemm", {"ms": 1e-9, "flops": 0.0, "n": 0})
        achieved = gm["flops"] / (gm["ms"] / 1000.0) / 1e12
        total_prof_ms = sum(d["ms"] for d in by.values())
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel", "achieved": achieved, "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"], "peak_src": pk["src"] + " (sustained bf16)",
                # dram__bytes_read+write of the conv_gemm launches of one step / launches, from the committed ncu launch
                # list profiles/r02_launches_summary.csv (77.16 MB per launch; algorithmic below for comparison)
                "traffic": 77.16e6, "algorithmic_bytes": sum(r["bytes"] for r in prof if r["kind"] == "conv_gemm") / max(gm["n"], 1),
                "launches": gm["n"], "share_of_step": gm["ms"] / total_prof_ms,
                "whole_step_tflops": B * GFLOP_PER_FRAME / 1000.0 / (ms_total / K / 1000.0),
                "per_kind_ms": {k: round(v["ms"], 4) for k, v in by.items()}}
        cpu = None
        if not args.no_cpu and world == 1:                  # rank 0 at N=1 only: the other ranks must not sit in a barrier
            fps, n, cores = cpu_reference_fps(P)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"{n} frames of the same 640x640 workload through oracle.detect (torch fp32, best of several host thread counts)"}
        line = {"metric": "frames/s YOLOv9-c 640px", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"YOLOv9-c, {B} uint8 {HW}x{HW}x3 BGR frames per GPU per step (BASELINE configs[1]), "
                                       "stem->144 convs->DFL decode->top300+suppression->scale_boxes",
                           "global_batch": world * B, "res": RES, "weights": "seeded synthetic",
                           "l2": f"{nbuf} rotating input batches ({nbuf * B * HW * HW * 3 / 1e6:.0f} MB) + {info['act_bytes'] / 1e9:.1f} GB activations per step (> 126 MB L2)",
                           "parallelism": f"dp{world} (frames sharded by batch, no collective)"},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * HW * HW * 3, "d2h_bytes_per_step": B * 300 * 6 * 4},
                "gpu_launches": info["launches"] * K,
                "roofline": roof, "cpu_baseline": cpu, "clip": clip_res}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
