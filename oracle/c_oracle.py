"""ctypes front-end of oracle/ssd_oracle.c (multithreaded C port of the reference path).

TEST / BASELINE INFRASTRUCTURE ONLY - same rules as ssd_oracle.py: imported by tests/, smoke() and the
cpu_baseline / `--impl reference` legs of bench.py, never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


CFLAGS = ["-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math"]  # one rounding per fp32 op (no FMA contraction)


def _cpu_stamp() -> str:
    """-march=native code only runs on the CPU it was built for: the library is rebuilt when the host changes
    (the build container and the GPU box need not have the same CPU)."""
    import hashlib

    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1((flags + " ".join(CFLAGS)).encode()).hexdigest()


def build(force: bool = False) -> str:
    """gcc -O3 -march=native -fopenmp -ffp-contract=off ....  Tries the system gcc first (an
    image-provided $CC may lack libgomp) and falls back to a single-threaded build if no compiler has OpenMP."""
    src = os.path.join(HERE, "ssd_oracle.c")
    stamp_file = LIB + ".stamp"
    stamp = _cpu_stamp()
    try:
        with open(stamp_file) as f:
            same_host = f.read().strip() == stamp
    except OSError:
        same_host = False
    if not force and same_host and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    base = CFLAGS + ["-fPIC", "-shared", "-fvisibility=hidden", "-o", LIB, src, "-lm"]
    err = ""
    for omp in (["-fopenmp"], []):
        for cc in ("/usr/bin/gcc", "gcc", "cc", os.environ.get("CC") or "gcc"):
            r = subprocess.run([cc] + omp + base, capture_output=True, text=True)
            if r.returncode == 0:
                with open(stamp_file, "w") as f:
                    f.write(stamp)
                return LIB
            err = r.stderr
    raise RuntimeError("could not build oracle/ssd_oracle.c: " + err)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def encode_batch(anchors_xyxy, gt_labels, gt_xyxy, gt_counts, num_classes, iou_threshold=0.5,
                 std=(0.1, 0.1, 0.2, 0.2), nthreads=1, want_match=True):
    anc, gt, gl, gc = _f32(anchors_xyxy), _f32(gt_xyxy), _i32(gt_labels), _i32(gt_counts)
    B, G = gt.shape[0], gt.shape[1]
    A = anc.shape[0]
    labels = np.empty((B, A, num_classes), np.float32)
    offsets = np.empty((B, A, 4), np.float32)
    match = np.empty((B, A), np.int32) if want_match else None
    sd = _f32(std)
    lib().oracle_encode_batch(_p(anc), A, _p(gt), _p(gl), _p(gc), B, G, int(num_classes),
                              C.c_float(np.float32(iou_threshold)), _p(sd), _p(labels), _p(offsets), _p(match),
                              int(nthreads))
    return labels, offsets, match


def decode_batch(anchors_cxcywh, offsets, std=(0.1, 0.1, 0.2, 0.2), mode=0, nthreads=1):
    cent, off = _f32(anchors_cxcywh), _f32(offsets)
    A = cent.shape[0]
    B = off.size // (A * 4)
    out = np.empty_like(off)
    sd = _f32(std)
    lib().oracle_decode_batch(_p(cent), A, _p(off), B, _p(sd), int(mode), _p(out), int(nthreads))
    return out


def combined_nms(boxes_yxyx, scores, max_per_class, max_total, iou_thr, score_thr, nthreads=1):
    bx, sc = _f32(boxes_yxyx), _f32(scores)
    B, A, Cn = sc.shape
    out = np.empty((B, max_total, 6), np.float32)
    valid = np.empty(B, np.int32)
    aidx = np.empty((B, max_total), np.int32)
    lib().oracle_combined_nms(_p(bx), _p(sc), B, A, Cn, int(max_per_class), int(max_total),
                              C.c_float(np.float32(iou_thr)), C.c_float(np.float32(score_thr)), _p(out), _p(valid),
                              _p(aidx), int(nthreads))
    return out, valid, aidx


def eval_tp(pred_label, pred_xyxy, gt_xyxy, gt_label, gt_count, num_classes, bg, thr, nthreads=0):
    pl, pb, gb, gl, gc = _i32(pred_label), _f32(pred_xyxy), _f32(gt_xyxy), _i32(gt_label), _i32(gt_count)
    B, K = pl.shape
    G = gb.shape[1]
    tp = np.empty((B, K), np.uint8)
    counts = np.zeros((num_classes, 3), np.int64)
    lib().oracle_eval_tp(_p(pl), _p(pb), _p(gb), _p(gl), _p(gc), B, K, G, int(num_classes), int(bg),
                         C.c_float(np.float32(thr)), _p(tp), _p(counts), int(nthreads) or (os.cpu_count() or 1))
    return tp, counts
