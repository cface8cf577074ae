"""ctypes binding of libclearcam_b200.so (C-ABI declared in include/clearcam_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a CCError is raised
(the reference's callers catch exceptions and supervise restarts themselves, clearcam.py:543-546).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CC_LIB: developer override used for same-box A/B runs of two builds (tools/ab.py); production loads the in-tree library
LIB_PATH = os.environ.get("CC_LIB") or os.path.join(_HERE, "libclearcam_b200.so")


class CCError(RuntimeError):
    pass


_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_SIGS = {
    "cc_version": (_i, []),
    "cc_last_error": (ctypes.c_char_p, []),
    "cc_device_check": (_i, []),
    "cc_conv2d": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "cc_yolo_create": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_vp),
                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_vp)]),
    "cc_yolo_create_ex": (_i, [ctypes.c_char_p, _i, _i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_vp),
                               ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_vp)]),
    "cc_yolo_destroy": (_i, [_vp]),
    "cc_yolo_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "cc_yolo_workspace_bytes": (_i, [_vp, _i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_size_t)]),
    "cc_yolo_set_workspace": (_i, [_vp, _vp, ctypes.c_size_t]),
    "cc_clip_workspace_bytes": (_i, [_vp, _i, _i, ctypes.POINTER(ctypes.c_size_t)]),
    "cc_clip_set_workspace": (_i, [_vp, _vp, ctypes.c_size_t]),
    "cc_yolo_plan_info": (_i, [_vp, _i, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                               ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "cc_yolo_profile": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, ctypes.POINTER(_f), ctypes.POINTER(ctypes.c_double),
                             ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_i), _vp]),
    "cc_yolo_trace": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_char_p),
                           ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i), _vp]),
    "cc_yolo_layer_output": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, ctypes.POINTER(_i), ctypes.POINTER(_i),
                                  ctypes.POINTER(_i), _vp]),
    "cc_detect_postprocess": (_i, [_vp, _i, _i, _i, _f, _i, _f, _f, _f, _f, _f, _vp, _vp]),
    "cc_detect_pred_from_raw": (_i, [_vp, _i, _i, _i, _f, _vp, _vp]),
    "cc_detect_decode": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_i), ctypes.POINTER(_i), _i, _f,
                              _vp, _vp, _vp]),
    "cc_clip_create": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_int64),
                            ctypes.POINTER(_vp)]),
    "cc_clip_destroy": (_i, [_vp]),
    "cc_clip_encode_image": (_i, [_vp, _vp, _i, _vp, ctypes.c_longlong, _vp]),
    "cc_clip_encode_text": (_i, [_vp, _vp, _i, _vp, ctypes.c_longlong, _vp]),
    "cc_clip_profile": (_i, [_vp, _i, _vp, _i, _vp, _i, ctypes.POINTER(_f), ctypes.POINTER(ctypes.c_double),
                             ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_double), _vp]),
    "cc_search_scores": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "cc_search_topk": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "cc_clip_preprocess": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "cc_ocsort_create": (_i, [_i, _i, ctypes.c_double, _i, ctypes.c_double, _i, ctypes.POINTER(_vp)]),
    "cc_ocsort_destroy": (_i, [_vp]),
    "cc_ocsort_update": (_i, [_vp, _vp, _i, _f, _vp, _i, ctypes.POINTER(_i)]),
    "cc_ocsort_update_batch": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "cc_ocsort_num_tracks": (_i, [_vp]),
    "cc_letterbox": (_i, [_vp, _i, _i, _i, _i, _i, _vp, ctypes.POINTER(_i), ctypes.POINTER(_i), _vp]),
}


def lib():
    """Load (once) and return the ctypes handle; raises CCError when the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CCError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name, None)
            if fn is None and os.environ.get("CC_LIB"):
                continue              # an older build under A/B comparison may lack newer entry points
            if fn is None:
                raise CCError(f"{LIB_PATH} does not export {name}: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().cc_last_error()
        raise CCError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def declared_symbols():
    return sorted(_SIGS)


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)
