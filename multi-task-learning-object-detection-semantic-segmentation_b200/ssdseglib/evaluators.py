"""Detection evaluation - mirror of the detection side of the reference's `ssdseglib/evaluators.py`.

`_iou_boxes_pred_vs_true` and `average_precision_object_detection` keep the reference signatures
(evaluators.py:6-11, :65-73).  IoU, TP flags and per-class counts are computed by libssdx on the GPU for the
whole batch at once; the per-class sort / cumulative sums / trapezoid (evaluators.py:164-185) are float64
host arithmetic on the gathered (tp, confidence) pairs, with the reference's exact formula.
`jaccard_iou_semantic_segmentation` (mask IoU) is not part of this path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ssdx


def _iou_boxes_pred_vs_true(labels_pred, boxes_pred, labels_true, boxes_true) -> np.ndarray:
    """IoU between every predicted box and every ground-truth box, zero where the labels differ
    (evaluators.py:6-63).  Boxes are (xmin, ymin, xmax, ymax) with inclusive pixel indices.
    Returns (P, G) float32; (P, 1) zeros when there is no ground truth."""
    P = int(np.shape(boxes_pred)[0])
    if len(labels_true) == 0:
        return np.zeros((P, 1), dtype=np.float32)
    lp, kind = _ssdx.to_device(np.asarray(labels_pred).reshape(-1), np.int32, None, "iou_lp")
    dev = lp.device.index
    bp, _ = _ssdx.to_device(np.asarray(boxes_pred).reshape(-1, 4), np.float32, dev, "iou_bp")
    lt, _ = _ssdx.to_device(np.asarray(labels_true).reshape(-1), np.int32, dev, "iou_lt")
    bt, _ = _ssdx.to_device(np.asarray(boxes_true).reshape(-1, 4), np.float32, dev, "iou_bt")
    return _ssdx.from_device(_ssdx.iou_pred_vs_true(lp, bp, lt, bt), "numpy")


def read_ground_truth(path_files_labels_boxes: Sequence[str], threads: int = 0):
    """Read the `label,xmin,ymin,xmax,ymax` CSV files of a dataset (evaluators.py:109-125) into padded
    arrays: labels (N, Gmax) int32, boxes (N, Gmax, 4) float32, counts (N,) int32.  The files are read and parsed
    by C++ worker threads (`ssdx_read_gt_csv_batch`); any newline convention is accepted, as csv.reader does."""
    return _ssdx.read_gt_csv_batch(list(path_files_labels_boxes), crlf_only=False, threads=threads)


def true_positives_and_counts(labels_pred_batch, boxes_pred_batch, gt_labels, gt_boxes, gt_counts,
                              iou_threshold: float, num_classes: int, label_code_background: int,
                              process_group=None):
    """Device part of the evaluation for a batch that is already in arrays.

    Returns (tp (N, K) uint8, counts (num_classes, 3) int64 = [TP, FP, ground-truth boxes] per class index).
    With `process_group` (torch.distributed) the counts are all-gathered over the ranks and summed - the
    only collective on this path."""
    lp, kind = _ssdx.to_device(labels_pred_batch, np.int32, None, "ev_lp")
    dev = lp.device.index
    bp, _ = _ssdx.to_device(boxes_pred_batch, np.float32, dev, "ev_bp")
    gl, _ = _ssdx.to_device(gt_labels, np.int32, dev, "ev_gl")
    gb, _ = _ssdx.to_device(gt_boxes, np.float32, dev, "ev_gb")
    gc, _ = _ssdx.to_device(gt_counts, np.int32, dev, "ev_gc")
    tp, counts = _ssdx.eval_tp(lp, bp, gb, gl, gc, num_classes, label_code_background, iou_threshold)
    if process_group is not None:
        from ._dist import all_gather_counts

        counts = all_gather_counts(counts, process_group).sum(dim=0)
    return _ssdx.from_device(tp, kind), _ssdx.from_device(counts, kind)


def average_precision_from_flags(tp, confidences, labels_pred, n_gt_per_label: Dict[int, int],
                                 label_code_background: int) -> Dict[int, float]:
    """evaluators.py:159-187 on flat arrays: per class, sort by confidence descending
    (`argsort(...)[::-1]`), precision = cumsum(tp)/rank, recall = cumsum(tp)/n_gt, AP = trapezoid(precision,
    recall) with no (0, .) sentinel; 0.0 when the class has no ground truth or no prediction."""
    ap: Dict[int, float] = {}
    tp = np.asarray(tp).reshape(-1)
    conf = np.asarray(confidences, dtype=np.float32).reshape(-1)
    lab = np.asarray(labels_pred).reshape(-1)
    for label, n_gt in n_gt_per_label.items():
        sel = lab == label
        if n_gt == 0 or not sel.any():
            ap[label] = 0.0
            continue
        order = np.argsort(conf[sel])[::-1]
        hits = tp[sel][order].astype(np.float32)
        cum = np.cumsum(hits)
        precision = cum / np.arange(1, hits.shape[0] + 1)
        recall = cum / n_gt
        ap[label] = float(np.sum((recall[1:] - recall[:-1]) * (precision[1:] + precision[:-1]) / 2.0))
    return ap


def average_precision_object_detection(
        labels_pred_batch, confidences_pred_batch, boxes_pred_batch, iou_threshold: float,
        path_files_labels_boxes: List[str], labels_codes: List[int], label_code_background: int,
) -> Dict[int, float]:
    """Average precision per class at one IoU threshold (evaluators.py:65-187).

    A prediction is a true positive when its best label-matched IoU is >= `iou_threshold`; several
    predictions may hit the same ground-truth box (no one-to-one assignment, as in the reference).
    Raises KeyError for a ground-truth or predicted label that is not in `labels_codes` or is the
    background (the reference's dictionaries do the same)."""
    labels_pred_batch = np.asarray(labels_pred_batch)
    n_pred = labels_pred_batch.shape[0]
    # the reference zips predictions with paths (evaluators.py:106): the shorter of the two decides
    n = min(n_pred, len(path_files_labels_boxes))
    lp = labels_pred_batch.reshape(n_pred, -1).astype(np.int32)[:n]
    conf = np.asarray(confidences_pred_batch, dtype=np.float32).reshape(n_pred, -1)[:n]
    bp = np.asarray(boxes_pred_batch, dtype=np.float32).reshape(n_pred, -1, 4)[:n]
    foreground = [l for l in labels_codes if l != label_code_background]
    if n == 0:
        return {l: 0.0 for l in foreground}
    gl, gb, gc = read_ground_truth(path_files_labels_boxes[:n])
    valid = np.arange(gl.shape[1])[None, :] < gc[:, None]
    for l in np.unique(gl[valid]):
        if int(l) not in foreground:
            raise KeyError(int(l))
    scored = lp != label_code_background
    for l in np.unique(lp[scored]):
        if int(l) not in foreground:
            raise KeyError(int(l))
    # class index = position in the sorted code list (codes need not be 0..C-1)
    codes = sorted(set(foreground) | {label_code_background})
    remap = {c: i for i, c in enumerate(codes)}
    lut = np.vectorize(lambda v: remap.get(int(v), len(codes)))
    lp_i = lut(lp).astype(np.int32) if lp.size else lp
    gl_i = lut(gl).astype(np.int32) if gl.size else gl
    # device: TP flags + per-class counts, then the precision/recall integration (sort, scan, trapezoid); only the
    # per-class AP values come back to the host
    torch = _ssdx.require_cuda()
    num_idx = len(codes) + 1
    bg_idx = remap[label_code_background]
    lp_d, _ = _ssdx.to_device(lp_i, np.int32, None, "ap_lp")
    dev = lp_d.device.index
    bp_d, _ = _ssdx.to_device(bp, np.float32, dev, "ap_bp")
    gl_d, _ = _ssdx.to_device(gl_i, np.int32, dev, "ap_gl")
    gb_d, _ = _ssdx.to_device(gb, np.float32, dev, "ap_gb")
    gc_d, _ = _ssdx.to_device(gc, np.int32, dev, "ap_gc")
    conf_d, _ = _ssdx.to_device(conf, np.float32, dev, "ap_conf")
    tp_d, counts_d = _ssdx.eval_tp(lp_d, bp_d, gb_d, gl_d, gc_d, num_idx, bg_idx, iou_threshold)
    ap_d = _ssdx.average_precision(tp_d.reshape(-1), conf_d.reshape(-1), lp_d.reshape(-1), counts_d, num_idx, bg_idx)
    ap = ap_d.cpu().numpy()
    return {l: float(ap[remap[l]]) for l in foreground}
