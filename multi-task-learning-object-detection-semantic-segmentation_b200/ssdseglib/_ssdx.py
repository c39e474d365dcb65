"""ctypes binding of libssdx.so (include/ssdx.h) and the host/device plumbing around it.

There is no CPU fallback here by design: if the shared library is missing the import fails, and if no
CUDA device is present every compute call raises.  PyTorch is used only for device memory, pinned
staging buffers, streams and DLPack exchange.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Any, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("SSDX_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libssdx.so"))

if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"libssdx.so not found at {_LIB_PATH}: build it with `python build.py` in "
        f"{os.path.dirname(_HERE)} (or `python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
        "This package has no CPU fallback.")

lib = C.CDLL(_LIB_PATH)

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/ssdx.h (tests check this)
SIGNATURES = {
    "ssdx_version": (C.c_int, []),
    "ssdx_last_error": (C.c_char_p, []),
    "ssdx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ssdx_anchors_create": (C.c_int, [_vp, _vp, C.c_int32, C.c_int, C.POINTER(_vp)]),
    "ssdx_anchors_destroy": (C.c_int, [_vp]),
    "ssdx_anchors_count": (C.c_int, [_vp, c_i32_p]),
    "ssdx_anchors_device": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "ssdx_anchors_grid_info": (C.c_int, [_vp, c_i32_p, c_i32_p]),
    "ssdx_set_option": (C.c_int, [C.c_char_p, C.c_char_p]),
    "ssdx_grid_probe": (C.c_int, [c_float_p, C.c_int32, c_i32_p, c_i32_p, c_i32_p, C.c_int32]),
    "ssdx_grid_window_probe": (C.c_int, [c_float_p, C.c_int32, c_float_p, c_float_p, C.c_int32, c_i32_p]),
    "ssdx_encode_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "ssdx_encode": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, c_float_p, _vp, _vp,
                              _vp, _vp, C.c_size_t, _vp]),
    "ssdx_decode": (C.c_int, [_vp, _vp, C.c_int32, c_float_p, C.c_int, _vp, _vp]),
    "ssdx_nms_workspace_bytes": (C.c_size_t, [C.c_int32] * 5),
    "ssdx_nms": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                           _vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_nms_scaled": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_nms_select": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _vp,
                                  C.c_size_t, C.c_int, _vp]),
    "ssdx_nms_finish": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_decode_nms": (C.c_int, [_vp, _vp, c_float_p, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ssdx_seg_class_presence": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_scale_scores": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, C.c_int, _vp]),
    "ssdx_eval_tp": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_float, _vp, _vp, C.c_int, _vp]),
    "ssdx_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "ssdx_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "ssdx_comm_create": (C.c_int, [C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.c_int, C.POINTER(_vp)]),
    "ssdx_comm_destroy": (C.c_int, [_vp]),
    "ssdx_allgather_counts": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp]),
    "ssdx_sum_counts": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int, _vp]),
    "ssdx_allgather_eval": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "ssdx_average_precision_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "ssdx_average_precision": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, _vp, C.c_size_t,
                                         C.c_int, _vp]),
    "ssdx_iou_pred_vs_true": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int, _vp]),
    "ssdx_metric_iou_boxes": (C.c_int, [_vp, _vp, _vp, C.c_int32, c_float_p, _vp, _vp]),
    "ssdx_softmax_scores": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, C.c_int, _vp]),
    "ssdx_nms_logits_workspace_bytes": (C.c_size_t, [C.c_int32] * 5),
    "ssdx_nms_logits": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_decode_nms_logits": (C.c_int, [_vp, _vp, c_float_p, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_float, C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ssdx_localization_loss": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int, _vp]),
    "ssdx_confidence_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "ssdx_confidence_loss": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _vp, _vp, _vp,
                                       _vp, C.c_size_t, C.c_int, _vp]),
    "ssdx_parse_gt_csv": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_float, c_i32_p, c_float_p, C.c_int32,
                                    c_i32_p]),
    "ssdx_read_gt_csv_batch": (C.c_int, [C.POINTER(C.c_char_p), C.c_int32, C.c_int, C.POINTER(C.c_uint8), C.c_float,
                                         C.c_int32, c_i32_p, c_float_p, c_i32_p, C.c_int32]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

SSDX_ERR_INVALID, SSDX_ERR_CUDA, SSDX_ERR_WORKSPACE, SSDX_ERR_NCCL = -1, -2, -3, -4
DECODE_YXYX, DECODE_CENTROIDS_MASKED, DECODE_CORNERS_MASKED = 0, 1, 2


def check(status: int) -> None:
    """0 = OK.  Invalid arguments raise ValueError (the reference raises ValueError at its validation
    sites, boxes.py:44,55, datacoder.py:56,75,108); everything else raises RuntimeError."""
    if status == 0:
        return
    msg = lib.ssdx_last_error().decode(errors="replace")
    if status == SSDX_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"libssdx error {status}: {msg}")


def version() -> int:
    return lib.ssdx_version()


def grid_probe(corners_xyxy: np.ndarray) -> np.ndarray:
    """Host-only: the (first anchor, rows, columns, shapes) rows of the regular grids ssdx_anchors_create finds in an
    (A,4) anchor table, shape (0,4) when the table is not a union of grids (encode then uses the tile match)."""
    c = np.ascontiguousarray(corners_xyxy, dtype=np.float32).reshape(-1, 4)
    n, ns = C.c_int32(0), C.c_int32(0)
    out = np.zeros((256, 4), np.int32)
    check(lib.ssdx_grid_probe(c.ctypes.data_as(c_float_p), c.shape[0], C.byref(n), C.byref(ns),
                              out.ctypes.data_as(c_i32_p), 256))
    return out[: n.value].copy()


def grid_windows(corners_xyxy: np.ndarray, gt_xyxy: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Host-only: (n_boxes, num_grids, 4) = (first row, rows, first column, columns) of the cell windows the grid
    encode drivers enumerate for the IoU bounds t (the same code the kernels run, compiled for the host)."""
    c = np.ascontiguousarray(corners_xyxy, dtype=np.float32).reshape(-1, 4)
    g = np.ascontiguousarray(gt_xyxy, dtype=np.float32).reshape(-1, 4)
    tt = np.ascontiguousarray(t, dtype=np.float32).reshape(-1)
    nq = len(grid_probe(c))
    out = np.zeros((g.shape[0], nq, 4), np.int32)
    check(lib.ssdx_grid_window_probe(c.ctypes.data_as(c_float_p), c.shape[0], g.ctypes.data_as(c_float_p),
                                     tt.ctypes.data_as(c_float_p), g.shape[0], out.ctypes.data_as(c_i32_p)))
    return out


def set_option(key: str, value: str) -> None:
    """Process-level dispatch override (include/ssdx.h: ssdx_set_option), e.g. ("encode_path", "tile_fused").
    Every path of an entry point gives identical results; tests and tuning runs pin one."""
    check(lib.ssdx_set_option(key.encode(), value.encode()))


# ---------------------------------------------------------------------------------------------
# device plumbing (torch)
# ---------------------------------------------------------------------------------------------
def _torch():
    import torch

    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("ssdseglib (B200 build) needs a CUDA device: there is no CPU fallback for the "
                           "encode/decode/NMS/eval path")
    return torch


def current_device() -> int:
    return require_cuda().cuda.current_device()


def stream_ptr(device: int) -> int:
    return _torch().cuda.current_stream(device).cuda_stream


_tls = threading.local()


def workspace(nbytes: int, device: int):
    """Per-thread, per-(device, stream) scratch buffer, grown geometrically (256-byte aligned by the
    caching allocator).  Stream-ordered reuse is safe because every call on it is enqueued on that stream."""
    torch = _torch()
    cache = getattr(_tls, "ws", None)
    if cache is None:
        cache = _tls.ws = {}
    key = (device, stream_ptr(device))
    buf = cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=f"cuda:{device}")
        cache[key] = buf
    return buf


class _PinnedSlot:
    """Pinned staging buffer + the event of the last H2D copy that read it (the buffer must not be
    overwritten by the host before that copy has executed)."""

    __slots__ = ("buf", "event")

    def __init__(self, buf):
        self.buf = buf
        self.event = None


def _pinned(shape, dtype, tag: str) -> _PinnedSlot:
    """Cached pinned staging buffer per (thread, tag, shape, dtype)."""
    torch = _torch()
    cache = getattr(_tls, "pin", None)
    if cache is None:
        cache = _tls.pin = {}
    key = (tag, tuple(shape), dtype)
    slot = cache.get(key)
    if slot is None:
        if len(cache) > 64:
            cache.clear()
        slot = cache[key] = _PinnedSlot(torch.empty(tuple(shape), dtype=dtype, pin_memory=True))
    elif slot.event is not None:
        slot.event.synchronize()
    return slot


_NP2T = {np.dtype(np.float32): "float32", np.dtype(np.int32): "int32", np.dtype(np.uint8): "uint8",
         np.dtype(np.int64): "int64"}


def to_device(x: Any, np_dtype, device: Optional[int] = None, tag: str = "x") -> Tuple[Any, str]:
    """Return (contiguous CUDA torch tensor of dtype, kind) for NumPy / list / torch / DLPack input.
    kind is 'numpy' (host input: staged through pinned memory, result goes back to NumPy), 'torch'
    (CUDA or CPU torch tensor; result is a torch tensor on the same device) or 'dlpack'."""
    torch = require_cuda()
    tdt = getattr(torch, _NP2T[np.dtype(np_dtype)])
    if isinstance(x, torch.Tensor):
        dev = device if device is not None else (x.device.index if x.is_cuda else torch.cuda.current_device())
        t = x.to(device=f"cuda:{dev}", dtype=tdt, non_blocking=True).contiguous()
        return t, ("torch" if x.is_cuda else "torch_cpu")
    if not isinstance(x, np.ndarray) and hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
        if not t.is_cuda:
            return to_device(t.numpy(), np_dtype, device, tag)
        return t.to(dtype=tdt).contiguous(), "dlpack"
    dev = device if device is not None else torch.cuda.current_device()
    a = np.ascontiguousarray(np.asarray(x), dtype=np_dtype)
    h = torch.from_numpy(a)
    if a.size and not h.is_pinned():
        slot = _pinned(a.shape, tdt, tag)
        slot.buf.copy_(h)
        d = slot.buf.to(f"cuda:{dev}", non_blocking=True)
        slot.event = torch.cuda.Event()
        slot.event.record(torch.cuda.current_stream(dev))
        return d, "numpy"
    return h.to(f"cuda:{dev}", non_blocking=True), "numpy"


def from_device(t, kind: str):
    """Mirror the input kind on the way out."""
    if kind == "numpy":
        if t.numel() * t.element_size() >= (1 << 16):
            # device -> pinned host block (torch's caching host allocator hands freed blocks out again) at PCIe rate;
            # the returned array owns the block through its base tensor, so consecutive calls never alias
            torch = _torch()
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device.index).synchronize()
            return h.numpy()
        return t.cpu().numpy()
    if kind == "torch_cpu":
        return t.cpu()
    return t


def empty(shape, np_dtype, device: int):
    torch = _torch()
    return torch.empty(tuple(shape), dtype=getattr(torch, _NP2T[np.dtype(np_dtype)]), device=f"cuda:{device}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class Anchors:
    """Owner of one ssdx_anchors handle (device-resident, immutable)."""

    def __init__(self, corners_xyxy: np.ndarray, centroids: Optional[np.ndarray], device: Optional[int] = None):
        require_cuda()
        self.device = current_device() if device is None else int(device)
        corners = np.ascontiguousarray(corners_xyxy, dtype=np.float32)
        if corners.ndim != 2 or corners.shape[1] != 4:
            raise ValueError("anchors must have shape (A, 4)")
        self.num_anchors = int(corners.shape[0])
        cent = None if centroids is None else np.ascontiguousarray(centroids, dtype=np.float32)
        if cent is not None and cent.shape != corners.shape:
            raise ValueError("anchor centroids must have the same shape as the corners")
        h = _vp()
        check(lib.ssdx_anchors_create(corners.ctypes.data, None if cent is None else cent.ctypes.data,
                                      self.num_anchors, self.device, C.byref(h)))
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                lib.ssdx_anchors_destroy(h)
            except Exception:
                pass
            self.handle = None


def std_array(std) -> Any:
    arr = (C.c_float * 4)(*[float(np.float32(s)) for s in std])
    return arr


# ---------------------------------------------------------------------------------------------
# thin typed wrappers over the exports (device tensors in, device tensors out)
# ---------------------------------------------------------------------------------------------
def encode(anchors: Anchors, gt_xyxy, gt_label, gt_count, num_classes: int, iou_threshold: float, std,
           want_match: bool = False):
    B, G = int(gt_xyxy.shape[0]), int(gt_xyxy.shape[1])
    A, dev = anchors.num_anchors, anchors.device
    labels = empty((B, A, num_classes), np.float32, dev)
    offsets = empty((B, A, 4), np.float32, dev)
    match = empty((B, A), np.int32, dev) if want_match else None
    nbytes = lib.ssdx_encode_workspace_bytes(A, B, G)
    ws = workspace(nbytes, dev)
    check(lib.ssdx_encode(anchors.handle, ptr(gt_xyxy), ptr(gt_label), ptr(gt_count), B, G, num_classes,
                          float(iou_threshold), std_array(std), ptr(labels), ptr(offsets), ptr(match), ptr(ws),
                          ws.numel(), stream_ptr(dev)))
    return labels, offsets, match


def decode(anchors: Anchors, offsets, std, mode: int):
    out = empty(tuple(offsets.shape), np.float32, anchors.device)
    B = int(offsets.numel() // (anchors.num_anchors * 4))
    check(lib.ssdx_decode(anchors.handle, ptr(offsets), B, std_array(std), mode, ptr(out),
                          stream_ptr(anchors.device)))
    return out


def nms(boxes_yxyx, scores, max_per_class: int, max_total: int, iou_threshold: float, score_threshold: float,
        want_aux: bool = False, class_scale=None):
    B, A, Cn = (int(v) for v in scores.shape)
    dev = boxes_yxyx.device.index
    out = empty((B, max_total, 6), np.float32, dev)
    valid = empty((B,), np.int32, dev) if want_aux else None
    aidx = empty((B, max_total), np.int32, dev) if want_aux else None
    nbytes = lib.ssdx_nms_workspace_bytes(B, A, Cn, max_per_class, max_total)
    ws = workspace(nbytes, dev)
    check(lib.ssdx_nms_scaled(ptr(boxes_yxyx), ptr(scores), ptr(class_scale), B, A, Cn, int(max_per_class),
                              int(max_total), float(iou_threshold), float(score_threshold), ptr(out), ptr(valid),
                              ptr(aidx), ptr(ws), ws.numel(), dev, stream_ptr(dev)))
    return out, valid, aidx


def nms_logits(boxes_yxyx, logits, max_per_class: int, max_total: int, iou_threshold: float, score_threshold: float,
               want_aux: bool = False, class_scale=None):
    """NMS on raw class logits: soft-max fused into the score filter (ssdx_nms_logits)."""
    B, A, Cn = (int(v) for v in logits.shape)
    dev = boxes_yxyx.device.index
    out = empty((B, max_total, 6), np.float32, dev)
    valid = empty((B,), np.int32, dev) if want_aux else None
    aidx = empty((B, max_total), np.int32, dev) if want_aux else None
    ws = workspace(lib.ssdx_nms_logits_workspace_bytes(B, A, Cn, max_per_class, max_total), dev)
    check(lib.ssdx_nms_logits(ptr(boxes_yxyx), ptr(logits), ptr(class_scale), B, A, Cn, int(max_per_class),
                              int(max_total), float(iou_threshold), float(score_threshold), ptr(out), ptr(valid),
                              ptr(aidx), ptr(ws), ws.numel(), dev, stream_ptr(dev)))
    return out, valid, aidx


def decode_nms_logits(anchors: "Anchors", offsets, std, logits, max_per_class: int, max_total: int,
                      iou_threshold: float, score_threshold: float, want_aux: bool = False, class_scale=None):
    B, A, Cn = (int(v) for v in logits.shape)
    dev = anchors.device
    out = empty((B, max_total, 6), np.float32, dev)
    valid = empty((B,), np.int32, dev) if want_aux else None
    aidx = empty((B, max_total), np.int32, dev) if want_aux else None
    ws = workspace(lib.ssdx_nms_logits_workspace_bytes(B, A, Cn, max_per_class, max_total), dev)
    check(lib.ssdx_decode_nms_logits(anchors.handle, ptr(offsets), std_array(std), ptr(logits), ptr(class_scale), B, Cn,
                                     int(max_per_class), int(max_total), float(iou_threshold), float(score_threshold),
                                     ptr(out), ptr(valid), ptr(aidx), ptr(ws), ws.numel(), stream_ptr(dev)))
    return out, valid, aidx


def softmax_scores(logits, class_scale=None):
    """logits: CUDA float32 (..., C) -> probabilities of the same shape [* class_scale]."""
    dev = logits.device.index
    Cn = int(logits.shape[-1])
    out = empty(tuple(logits.shape), np.float32, dev)
    check(lib.ssdx_softmax_scores(ptr(logits), int(logits.numel() // Cn), Cn, ptr(class_scale), ptr(out), dev,
                                  stream_ptr(dev)))
    return out


def decode_nms(anchors: "Anchors", offsets, std, scores, max_per_class: int, max_total: int, iou_threshold: float,
               score_threshold: float, want_aux: bool = False, class_scale=None):
    B, A, Cn = (int(v) for v in scores.shape)
    dev = anchors.device
    out = empty((B, max_total, 6), np.float32, dev)
    valid = empty((B,), np.int32, dev) if want_aux else None
    aidx = empty((B, max_total), np.int32, dev) if want_aux else None
    ws = workspace(lib.ssdx_nms_workspace_bytes(B, A, Cn, max_per_class, max_total), dev)
    check(lib.ssdx_decode_nms(anchors.handle, ptr(offsets), std_array(std), ptr(scores), ptr(class_scale), B, Cn,
                              int(max_per_class), int(max_total), float(iou_threshold), float(score_threshold),
                              ptr(out), ptr(valid), ptr(aidx), ptr(ws), ws.numel(), stream_ptr(dev)))
    return out, valid, aidx


class NmsSelection:
    """Phase 1 of the NMS (candidate lists in a private workspace), to be finished with `nms_finish`."""

    def __init__(self, scores, class_scale, max_per_class, max_total, score_threshold, ws, event):
        self.scores, self.class_scale, self.ws, self.event = scores, class_scale, ws, event
        self.max_per_class, self.max_total, self.score_threshold = max_per_class, max_total, score_threshold


def nms_select(scores, max_per_class: int, max_total: int, score_threshold: float, class_scale=None) -> NmsSelection:
    """Runs on the CURRENT stream (which may differ from the stream `nms_finish` is called on); the returned object
    carries the event phase 2 waits for."""
    torch = _torch()
    B, A, Cn = (int(v) for v in scores.shape)
    dev = scores.device.index
    ws = torch.empty(int(lib.ssdx_nms_workspace_bytes(B, A, Cn, max_per_class, max_total)), dtype=torch.uint8,
                     device=f"cuda:{dev}")
    check(lib.ssdx_nms_select(ptr(scores), ptr(class_scale), B, A, Cn, int(max_per_class), int(max_total),
                              float(score_threshold), ptr(ws), ws.numel(), dev, stream_ptr(dev)))
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return NmsSelection(scores, class_scale, int(max_per_class), int(max_total), float(score_threshold), ws, ev)


def nms_finish(sel: NmsSelection, boxes_yxyx, iou_threshold: float, want_aux: bool = False):
    torch = _torch()
    B, A, Cn = (int(v) for v in sel.scores.shape)
    dev = boxes_yxyx.device.index
    cur = torch.cuda.current_stream(dev)
    cur.wait_event(sel.event)
    for t in (sel.scores, sel.class_scale, sel.ws):
        if t is not None:
            t.record_stream(cur)
    out = empty((B, sel.max_total, 6), np.float32, dev)
    valid = empty((B,), np.int32, dev) if want_aux else None
    aidx = empty((B, sel.max_total), np.int32, dev) if want_aux else None
    check(lib.ssdx_nms_finish(ptr(boxes_yxyx), ptr(sel.scores), ptr(sel.class_scale), B, A, Cn, sel.max_per_class,
                              sel.max_total, float(iou_threshold), sel.score_threshold, ptr(out), ptr(valid),
                              ptr(aidx), ptr(sel.ws), sel.ws.numel(), dev, stream_ptr(dev)))
    return out, valid, aidx


def metric_iou_boxes(anchors: "Anchors", y_true, y_pred, std):
    B = int(y_true.shape[0])
    out = empty((B,), np.float32, anchors.device)
    check(lib.ssdx_metric_iou_boxes(anchors.handle, ptr(y_true), ptr(y_pred), B, std_array(std), ptr(out),
                                    stream_ptr(anchors.device)))
    return out


def localization_loss(y_true, y_pred, with_gradient: bool = False):
    """y_true, y_pred: CUDA float32 (B,A,4).  Returns loss (B,) [, gradient (B,A,4)]."""
    dev = y_true.device.index
    B, A = int(y_true.shape[0]), int(y_true.shape[1])
    out = empty((B,), np.float32, dev)
    grad = empty((B, A, 4), np.float32, dev) if with_gradient else None
    check(lib.ssdx_localization_loss(ptr(y_true), ptr(y_pred), B, A, ptr(out), ptr(grad) if with_gradient else None, dev,
                                     stream_ptr(dev)))
    return (out, grad) if with_gradient else out


def confidence_loss(y_true, y_pred, epsilon: float = 1e-7, with_gradient: bool = False, with_keep: bool = False):
    """y_true, y_pred: CUDA float32 (B,A,C).  Returns loss (B,) [, gradient (B,A,C)] [, kept-background mask (B,A) u8]."""
    dev = y_true.device.index
    B, A, Cn = (int(v) for v in y_true.shape)
    out = empty((B,), np.float32, dev)
    grad = empty((B, A, Cn), np.float32, dev) if with_gradient else None
    keep = empty((B, A), np.uint8, dev) if with_keep else None
    nbytes = int(lib.ssdx_confidence_loss_workspace_bytes(B, A))
    ws = workspace(max(nbytes, 256), dev)
    check(lib.ssdx_confidence_loss(ptr(y_true), ptr(y_pred), B, A, Cn, float(np.float32(epsilon)),
                                   float(np.float32(1 - epsilon)), ptr(out), ptr(grad) if with_gradient else None,
                                   ptr(keep) if with_keep else None, ptr(ws), ws.numel(), dev, stream_ptr(dev)))
    res = [out]
    if with_gradient:
        res.append(grad)
    if with_keep:
        res.append(keep)
    return res[0] if len(res) == 1 else tuple(res)


def seg_class_presence(seg_probs, depth: int = 4):
    """seg_probs: CUDA float32 tensor (..., Cm).  Returns a CUDA float32 tensor (depth,) of 0/1."""
    dev = seg_probs.device.index
    cm = int(seg_probs.shape[-1])
    out = empty((depth,), np.float32, dev)
    ws = workspace(256, dev)
    check(lib.ssdx_seg_class_presence(ptr(seg_probs), int(seg_probs.numel() // cm), cm, int(depth), ptr(out), ptr(ws),
                                      ws.numel(), dev, stream_ptr(dev)))
    return out


def scale_scores(scores, class_scale):
    dev = scores.device.index
    Cn = int(scores.shape[-1])
    out = empty(tuple(scores.shape), np.float32, dev)
    check(lib.ssdx_scale_scores(ptr(scores), int(scores.numel() // Cn), Cn, ptr(class_scale), ptr(out), dev,
                                stream_ptr(dev)))
    return out


def eval_tp(pred_label, pred_xyxy, gt_xyxy, gt_label, gt_count, num_classes: int, background_label: int,
            iou_threshold: float, counts=None):
    torch = _torch()
    B, K = int(pred_label.shape[0]), int(pred_label.shape[1])
    G = int(gt_xyxy.shape[1])
    if not (int(gt_count.shape[0]) == B and int(gt_xyxy.shape[0]) == B and int(gt_label.shape[0]) == B
            and int(pred_xyxy.shape[0]) == B):
        raise ValueError(f"eval_tp: {B} prediction rows but {int(gt_count.shape[0])} ground-truth rows")
    dev = pred_label.device.index
    tp = empty((B, K), np.uint8, dev)
    if counts is None:
        counts = torch.zeros((num_classes, 3), dtype=torch.int64, device=f"cuda:{dev}")
    check(lib.ssdx_eval_tp(ptr(pred_label), ptr(pred_xyxy), ptr(gt_xyxy), ptr(gt_label), ptr(gt_count), B, K, G,
                           num_classes, int(background_label), float(iou_threshold), ptr(tp), ptr(counts), dev,
                           stream_ptr(dev)))
    return tp, counts


def average_precision(tp, confidence, label, counts, num_classes: int, background_label: int):
    """tp (N,) uint8, confidence (N,) float32, label (N,) int32 class indices, counts (C,3) int64: CUDA tensors.
    Returns a CUDA float64 tensor (C,)."""
    torch = _torch()
    dev = tp.device.index
    n = int(tp.numel())
    out = torch.empty((num_classes,), dtype=torch.float64, device=f"cuda:{dev}")
    ws = workspace(int(lib.ssdx_average_precision_workspace_bytes(n, num_classes)), dev)
    check(lib.ssdx_average_precision(ptr(tp), ptr(confidence), ptr(label), n, int(num_classes), ptr(counts),
                                     int(background_label), ptr(out), ptr(ws), ws.numel(), dev, stream_ptr(dev)))
    return out


def iou_pred_vs_true(pred_label, pred_xyxy, gt_label, gt_xyxy):
    P, G = int(pred_label.shape[0]), int(gt_label.shape[0])
    dev = pred_xyxy.device.index
    out = empty((P, G), np.float32, dev)
    check(lib.ssdx_iou_pred_vs_true(ptr(pred_label), ptr(pred_xyxy), ptr(gt_label), ptr(gt_xyxy), P, G, ptr(out),
                                    dev, stream_ptr(dev)))
    return out


def parse_gt_csv(text: bytes, crlf_only: bool, flip_width: float = 0.0, capacity: int = 0):
    """Host-side C parser; returns (labels int32 (G,), boxes float32 (G,4))."""
    if isinstance(text, str):
        text = text.encode()
    cap = capacity if capacity > 0 else max(16, text.count(b"\n") + 2)
    labels = np.empty(cap, np.int32)
    boxes = np.empty((cap, 4), np.float32)
    n = C.c_int32(0)
    st = lib.ssdx_parse_gt_csv(text, len(text), 1 if crlf_only else 0, float(flip_width),
                               labels.ctypes.data_as(c_i32_p), boxes.ctypes.data_as(c_float_p), cap, C.byref(n))
    check(st)
    return labels[: n.value].copy(), boxes[: n.value].copy()


def read_gt_csv_batch(paths, crlf_only: bool, flips=None, flip_width: float = 0.0, gmax: int = 32, threads: int = 0):
    """Multithreaded host-side ingest of one CSV file per image (C++ worker threads, the GIL is released).
    Returns padded (labels (N, G) int32, boxes (N, G, 4) float32, counts (N,) int32) with G = max(count, 1);
    retried once with the exact size if a file holds more than `gmax` records."""
    n = len(paths)
    enc = [p if isinstance(p, bytes) else str(p).encode() for p in paths]
    arr = (C.c_char_p * max(n, 1))(*enc)
    flip_arr = None
    if flips is not None:
        fl = np.ascontiguousarray(np.asarray(flips).astype(np.uint8))
        if fl.shape != (n,):
            raise ValueError("flips must have one entry per path")
        flip_arr = fl.ctypes.data_as(C.POINTER(C.c_uint8))
    counts = np.zeros(max(n, 1), np.int32)
    g = max(int(gmax), 1)
    for _ in range(2):
        labels = np.empty((max(n, 1), g), np.int32)
        boxes = np.empty((max(n, 1), g, 4), np.float32)
        st = lib.ssdx_read_gt_csv_batch(arr, n, 1 if crlf_only else 0, flip_arr, float(flip_width), g,
                                        labels.ctypes.data_as(c_i32_p), boxes.ctypes.data_as(c_float_p),
                                        counts.ctypes.data_as(c_i32_p), int(threads))
        if st != SSDX_ERR_WORKSPACE:
            break
        g = int(counts[:n].max())
    check(st)
    gfit = max(int(counts[:n].max()) if n else 0, 1)
    return labels[:n, :gfit].copy(), boxes[:n, :gfit].copy(), counts[:n].copy()
