"""Zero-padded equivalents of the detector sizes whose channel widths are not multiples of 16.

The tensor-core conv takes channel counts in multiples of 16 and 16-byte aligned slices.  YOLOv9-t (hidden width 24) and
YOLOv9-m (60/90/184/360 ...) are not like that (`SIZES`, detection/yolov9.py:461-464).  Instead of a slow generic path,
such a model is run as the same graph with every offending width rounded up and the extra channels' weights and biases
set to zero: a zero-weight, zero-bias channel is exactly 0 after SiLU, through the pools, the upsample and every concat,
and contributes exactly nothing to its consumers, so the padded model computes the same function.  What needs care is
WHERE the real channels sit once tensors that are chunked or concatenated are padded segment by segment; this module
walks the generic t/s/m/c graph (detection/yolov9.py:299-327) with that bookkeeping and scatters the reference's weights
into the padded layout.  The library knows the padded tables as sizes "t@16" and "m@16".

    sd16 = pad_state_dict("m", state_dict)      # -> weights for cc_yolo_create("m@16", ...)
"""
from typing import Dict, List, Sequence

import numpy as np

# index names follow YOLOv9.__init__ (detection/yolov9.py:302): a, b, c, d, e, f, g, h, i, j, k, l, m, n, p, q, r, s, t, u, v, w
SIZES = {
    "t": [16, 64, 96, 24, 128, 256, 224, 160, 48, 144, 192, 80, 32, 16, 3, 96, 32, 64, 128, 64, 64, 128],
    "m": [32, 240, 360, 90, 480, 960, 840, 600, 184, 544, 720, 240, 128, 60, 1, 360, 120, 64, 128, 240, 240, 480],
}
# t: hidden 24 -> 32, head class branch 80 -> 96 (Cin = 80 is not a multiple of 32, which the 3x3 halo mainloop needs).
# m: 360 -> 384 rather than 368: the conv kernel tiles Cout by its largest divisor that is a multiple of 16 and at most
# 256, and 368 = 16 x 23 would leave 16-wide tiles (measured 3x slower end to end); 384 tiles as 2 x 192.
PADDED = {
    "t": [16, 64, 96, 32, 128, 256, 224, 160, 48, 144, 192, 96, 32, 16, 3, 96, 32, 64, 128, 64, 64, 128],
    "m": [32, 240, 384, 96, 480, 960, 864, 624, 192, 576, 720, 240, 128, 64, 1, 384, 128, 64, 128, 240, 240, 480],
}


def padded_size(size: str):
    """Name of the padded table the library runs `size` as, or None when the size needs no padding (s, c, e)."""
    return f"{size}@16" if size in PADDED else None


def _segments(logical: Sequence[int], physical: Sequence[int]) -> np.ndarray:
    """Physical channel index of every logical channel of a tensor made of back-to-back segments."""
    out, off = [], 0
    for lo, ph in zip(logical, physical):
        assert ph >= lo
        out.append(off + np.arange(lo))
        off += ph
    return np.concatenate(out)


class _Walk:
    """Both channel tables side by side: every conv gets (input map, output map, padded cin, padded cout)."""

    def __init__(self, z, zp):
        self.z, self.zp = z, zp
        self.convs: Dict[str, tuple] = {}

    def conv(self, key, cin, cout, cin_p, cout_p, in_map=None, out_map=None):
        self.convs[key] = (np.arange(cin) if in_map is None else in_map, np.arange(cout) if out_map is None else out_map,
                           cin_p, cout_p)

    def repncsp(self, pfx, a, b, ap, bp, n):            # detection/yolov9.py:92-105, RepNCSP(a -> a, hidden b)
        self.conv(pfx + ".cv1.conv", a, b, ap, bp)
        self.conv(pfx + ".cv2.conv", a, b, ap, bp)
        for i in range(n):
            self.conv(f"{pfx}.m.{i}.cv1.conv", b, b, bp, bp)
            self.conv(f"{pfx}.m.{i}.cv2.conv", b, b, bp, bp)
        self.conv(pfx + ".cv3.conv", 2 * b, a, 2 * bp, ap, in_map=_segments([b, b], [bp, bp]))

    def elan4(self, pfx, in_map, cin_p, b, c, bp, cp, n):   # :107-125 with c3 = 4b, c4 = 2b
        cin = len(in_map)
        self.conv(pfx + ".cv1.conv", cin, 4 * b, cin_p, 4 * bp, in_map=in_map, out_map=_segments([2 * b, 2 * b], [2 * bp, 2 * bp]))
        for br in ("cv2", "cv3"):
            self.repncsp(f"{pfx}.{br}.0", 2 * b, b, 2 * bp, bp, n)
            self.conv(f"{pfx}.{br}.1.conv", 2 * b, 2 * b, 2 * bp, 2 * bp)
        self.conv(pfx + ".cv4.conv", 8 * b, c, 8 * bp, cp, in_map=_segments([2 * b] * 4, [2 * bp] * 4))

    def run(self, small: bool):
        (a, b, c, d, e, f, g, h, i, j, k, l, m, n, p, q, r, s, t, u, v, w) = self.z
        (A, B, C, D, E, F, G, H, I, J, K, L, M, N, P, Q, R, S, T, U, V, W) = self.zp
        assert (a, p, s, t) == (A, P, S, T)
        ar = np.arange
        self.conv("model.0.conv", 3, a, 3, A)
        self.conv("model.1.conv", a, 2 * a, A, 2 * A)
        if small:                                        # ELAN1, :65-80: widths (2a -> m), chunks m/2, cv4 input b
            assert (m, b) == (M, B)
            self.conv("model.2.cv1.conv", 2 * a, m, 2 * A, M)
            self.conv("model.2.cv2.conv", a, a, A, A)
            self.conv("model.2.cv3.conv", a, a, A, A)
            self.conv("model.2.cv4.conv", b, m, B, M)
        else:
            self.elan4("model.2", ar(s), S, 32, t, 32, T, p)
        self.conv("model.3.cv1.conv", m, u, M, U)                                   # AConv :54-63
        self.elan4("model.4", ar(b), B, n, v, N, V, p)
        self.conv("model.5.cv1.conv", b, q, B, Q)
        self.elan4("model.6", ar(c), C, d, c, D, C, p)
        self.conv("model.7.cv1.conv", q, e, Q, E)
        self.elan4("model.8", ar(w), W, r, w, R, W, p)
        self.conv("model.9.cv1.conv", w, b, W, B)                                   # SPPELAN :134-149
        self.conv("model.9.cv5.conv", f, w, 4 * B, W, in_map=_segments([b] * 4, [B] * 4))
        assert g == w + c and h == c + v and j == i + c and k == b + w              # the concats of :315-326
        self.elan4("model.12", _segments([w, c], [W, C]), W + C, d, c, D, C, p)     # cat(up(9), 6)
        self.elan4("model.15", _segments([c, v], [C, V]), C + V, n, b, N, B, p)     # cat(up(12), 4)
        self.conv("model.16.cv1.conv", v, i, V, I)
        self.elan4("model.18", _segments([i, c], [I, C]), I + C, d, c, D, C, p)     # cat(16, 12)
        self.conv("model.19.cv1.conv", q, b, Q, B)
        self.elan4("model.21", _segments([b, w], [B, W]), B + W, r, w, R, W, p)     # cat(19, 9)
        for jx, (ch, chp) in enumerate(zip((b, c, w), (B, C, W))):                  # DDetect first convs :171-194
            self.conv(f"model.22.cv2.{jx}.0.conv", ch, 64, chp, 64)
            self.conv(f"model.22.cv3.{jx}.0.conv", ch, l, chp, L)
            self.conv(f"model.22.cv3.{jx}.1.conv", l, l, L, L)
            self.conv(f"model.22.cv3.{jx}.2", l, 80, L, 80)
        return self.convs


def layer_channel_maps(size: str) -> List[np.ndarray]:
    """Per graph layer (model.0 ... model.21): physical channel index of every logical channel of its output in the padded
    model — plain tensors are padded at the end, concat outputs segment by segment."""
    (a, b, c, d, e, f, g, h, i, j, k, l, m, n, p, q, r, s, t, u, v, w) = SIZES[size]
    (A, B, C, D, E, F, G, H, I, J, K, L, M, N, P, Q, R, S, T, U, V, W) = PADDED[size]
    ar = np.arange
    l2 = m if size == "t" else t
    return [ar(a), ar(2 * a), ar(l2), ar(u), ar(v), ar(q), ar(c), ar(e), ar(w), ar(w), ar(w), _segments([w, c], [W, C]), ar(c), ar(c),
            _segments([c, v], [C, V]), ar(b), ar(i), _segments([i, c], [I, C]), ar(c), ar(b), _segments([b, w], [B, W]), ar(w)]


def pad_state_dict(size: str, sd: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Reference-named fp32 weights of YOLOv9-`size` -> weights of the zero-padded equivalent (same key names).
    Keys the walk does not list (the head's inner convs, the DFL weight) have unchanged shapes and pass through."""
    if size not in PADDED:
        return dict(sd)
    convs = _Walk(SIZES[size], PADDED[size]).run(small=(size == "t"))
    out = {}
    for key, val in sd.items():
        base, _, leaf = key.rpartition(".")
        if base not in convs or leaf not in ("weight", "bias"):
            out[key] = val
            continue
        in_map, out_map, cin_p, cout_p = convs[base]
        a = np.asarray(val, np.float32)
        if leaf == "bias":
            assert a.shape == (len(out_map),), f"{key}: bias {a.shape}, expected {len(out_map)}"
            b = np.zeros(cout_p, np.float32)
            b[out_map] = a
        else:
            assert a.shape[:2] == (len(out_map), len(in_map)), f"{key}: weight {a.shape}, expected {(len(out_map), len(in_map))}"
            b = np.zeros((cout_p, cin_p) + a.shape[2:], np.float32)
            b[np.ix_(out_map, in_map)] = a
        out[key] = b
    missing = [k for k in convs if k + ".weight" not in sd]
    if missing:
        raise KeyError(f"pad_state_dict({size!r}): state dict lacks {missing[:3]} ...")
    return out
