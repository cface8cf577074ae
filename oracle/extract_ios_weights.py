"""Recover the reference's real YOLOv9-t fp32 weights from its iOS bundle blob.

/root/reference/ios/clearcam/yolov9t is a concatenation of records [sha256(payload):32][len:u64 LE][payload]; the
last record is a text op list (BufferAlloc/CopyIn/ProgramAlloc/ProgramExec/CopyOut) of a serialized tinygrad graph
(format parsed by ios/clearcam/Yolo.m:128-196).  The CopyIn payloads, in op order, are the model parameters in the
python definition order of YOLOv9("t") (weight then bias per conv; DDetect: two anchor/stride placeholders, the DFL
weight, then cv2[*], cv3[*]).  This walks oracle.yolov9.conv_table("t") and consumes buffers by element count.
"""
import re
import struct
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BLOB = "/root/reference/ios/clearcam/yolov9t"


def parse_blob(path=BLOB):
    raw = open(path, "rb").read()
    recs, ptr = {}, 0
    last = None
    while ptr < len(raw):
        h = raw[ptr:ptr + 32].hex()
        (n,) = struct.unpack("<Q", raw[ptr + 32:ptr + 40])
        recs[h] = raw[ptr + 40:ptr + 40 + n]
        last = h
        ptr += 40 + n
    ops = recs[last].decode("utf-8")
    copyins = re.findall(r"CopyIn\(session=SessionKey\([^)]*\), buffer_num=(\d+), datahash='([0-9a-f]{64})'", ops)
    return [(int(b), recs[h]) for b, h in copyins]


def extract(path=BLOB):
    import torch
    from oracle import yolov9 as o
    bufs = parse_blob(path)
    arrays = [np.frombuffer(d, dtype=np.float32) for _, d in bufs[:-1]]     # last CopyIn = the uint8 input frame
    table = o.conv_table("t")
    # python definition order: all non-detect convs in table order, then DDetect: anchors, strides, dfl, cv2.*, cv3.*
    det = max(int(n.split(".")[1]) for n, *_ in table)
    body = [t for t in table if not t[0].startswith(f"model.{det}.")]
    head = [t for t in table if t[0].startswith(f"model.{det}.")]
    P, i = {}, 0

    def take(n):
        nonlocal i
        while arrays[i].size != n:      # skip placeholders (anchors / strides)
            i += 1
        a = arrays[i]
        i += 1
        return a

    for name, cin, cout, k, s, g, act in body:
        P[name + ".weight"] = torch.from_numpy(take(cout * (cin // g) * k * k).reshape(cout, cin // g, k, k).copy())
        P[name + ".bias"] = torch.from_numpy(take(cout).copy())
    P[f"model.{det}.dfl.conv.weight"] = torch.from_numpy(take(16).reshape(1, 16, 1, 1).copy())
    for name, cin, cout, k, s, g, act in head:
        P[name + ".weight"] = torch.from_numpy(take(cout * (cin // g) * k * k).reshape(cout, cin // g, k, k).copy())
        P[name + ".bias"] = torch.from_numpy(take(cout).copy())
    return P, i, len(arrays)


if __name__ == "__main__":
    P, used, total = extract()
    n = sum(v.numel() for v in P.values())
    print("tensors", len(P), "params", n, "buffers used", used, "of", total)
    print("dfl", P[[k for k in P if "dfl" in k][0]].flatten()[:16])
