"""ctypes binding of libaidet_b200.so (the C ABI declared in include/aidet_b200.h).

There is no CPU path and no fallback: if the shared library is missing, or a call
returns an error code, this module raises.  PyTorch is used for device memory and
streams only; every pointer crossing the boundary is a raw `data_ptr()`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AIDET_B200_LIB") or os.path.join(_HERE, "libaidet_b200.so")   # override: kernel-tuning builds

MODE_IOU, MODE_IOF = 0, 1
CMP_GT, CMP_GE = 0, 1
PROF_RIOU, PROF_NMS_MASK, PROF_ROI_FWD, PROF_ROI_BWD = 0, 1, 2, 3

_lib = None

_SIGNATURES = {
    "aidet_last_error": (C.c_char_p, []),
    "aidet_version": (C.c_int, []),
    "aidet_prof_enable": (C.c_int, [C.c_int]),
    "aidet_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "aidet_launch_count": (C.c_longlong, []),
    "aidet_ffma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "aidet_riou_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "aidet_riou_matrix_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_longlong, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "aidet_riou_matrix_mcast_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_longlong, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "aidet_riou_matrix_multi_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_int, C.c_longlong, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "aidet_riou_aligned_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "aidet_riou_aligned_grad_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "aidet_assign_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "aidet_max_iou_assign_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_int, C.c_void_p]),
    "aidet_assign_wrt_overlaps_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_float,
                                                C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "aidet_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "aidet_nms_batched_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_int, C.c_void_p]),
    "aidet_scene_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "aidet_scene_tile_nms_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.c_int, C.c_void_p]),
    "aidet_scene_merge_nms_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                            C.c_void_p]),
    "aidet_scene_compact_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                          C.c_void_p]),
    "aidet_soft_nms_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "aidet_rroi_align_fwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aidet_rroi_align_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aidet_rroi_align_bwd_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                          C.c_int, C.c_int]),
    "aidet_rroi_align_bwd_gather_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                                                  C.c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load the CUDA library once.  Raises (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "aidet_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C aidet_b200/csrc`).  There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().aidet_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t)))
    if not t.is_cuda:
        raise NotImplementedError("%s is a CPU tensor: aidet_b200 has no CPU implementation" % name)


def prof_enable(on=True):
    """0 / False: off; 1 / True: CUDA-event timing of the dominant kernel of each op; 2: also the phase stamps of the
    fused NMS kernel (left at the end of its workspace, scripts/r2_nms_phases.py)."""
    check(lib().aidet_prof_enable(int(on)), "aidet_prof_enable")


def prof_read(kind, reset=True):
    ms = C.c_double(0.0)
    cnt = C.c_longlong(0)
    check(lib().aidet_prof_read(kind, C.byref(ms), C.byref(cnt), int(bool(reset))), "aidet_prof_read")
    return ms.value, cnt.value


def launch_count():
    return int(lib().aidet_launch_count())


def ffma_peak_tflops(device=0, iters=4096):
    out = C.c_double(0.0)
    with torch.cuda.device(device):
        check(lib().aidet_ffma_peak(device, iters, C.byref(out), stream_ptr(device)), "aidet_ffma_peak")
    return out.value
