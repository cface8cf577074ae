"""`jit_infer` of the reference's `utils/helpers.py`, with the same signature (the helper `clearcam.py:583` calls the
detector through).  The other helper on the path, `resize` (utils/helpers.py:127-131), exists only inside
`YOLOv9.preprocess` / `cc_letterbox`, where the reference uses it."""


def jit_infer(fn, x, jit_cache):
    """`jit_infer(model, frame, cache)` (utils/helpers.py:214-221) captures one TinyJit graph per input shape.  Here the
    per-shape capture lives inside the library (a plan = buffers + TMA descriptors + launch list, built on the first call
    with a new shape and cached), so this is a plain call; `jit_cache` only records which shapes were seen, as the
    reference's dict does."""
    jit_cache.setdefault(tuple(x.shape), True)
    return fn(x)


def batch_bucket(n: int) -> int:
    """Batch sizes the library is called with when the caller's count varies from step to step (cameras that delivered a
    new frame, objects cropped out of a frame): every distinct batch size is a distinct cached plan (the reference's
    TinyJit would likewise capture one graph per shape, :214-221), so counts are padded up to 1, 2, 4, 8, 12, 16, 24, 32,
    then multiples of 16 — at most a third of padding, and a bounded set of plans."""
    for b in (1, 2, 4, 8, 12, 16, 24, 32):
        if n <= b:
            return b
    return (n + 15) // 16 * 16
