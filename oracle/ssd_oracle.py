"""CPU oracle for the SSD target/prediction path (TEST INFRASTRUCTURE, not product).

This file restates, in NumPy float32, what the reference `ssdseglib` computes on
the hot path named by BASELINE.json.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.
The shipped package never does: it fails loudly without the CUDA library.

Pin status (see DESIGN.md "Oracle"):
  * anchors (boxes.py) and detection eval (evaluators.py): PINNED - checked
    against outputs of the reference's own files executed in the build
    container (tests/golden/make_golden.py -> tests/golden/*.npz).
  * encode / decode (datacoder.py, layers.py): the reference needs TensorFlow
    2.13, which is absent.  The reference's own source is executed against a
    NumPy emulation of the ~30 TF primitives it calls
    (tests/golden/tf_numpy_shim.py); the fixtures pin this restatement to the
    reference's literal statement sequence, but the primitives themselves are
    emulated => "parity pinned to reference source over emulated TF ops".
  * NMS: `tf.image.combined_non_max_suppression` is a TF 2.13 C++ kernel that
    is not in /root/reference => PARITY UNPINNED; restated from the published
    algorithm and cross-checked against torchvision.ops.nms.

Every op is done in np.float32 with Python-float scalars (weak under NEP 50) so
nothing is promoted to float64; association follows the reference expressions.
"""
from __future__ import annotations

import csv
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# a1-a4: anchors (reference boxes.py:74-151, :153-177, :179-207, :274-323)
# --------------------------------------------------------------------------
def anchors_generate(
    feature_maps_shapes,
    feature_maps_aspect_ratios=(1, 2, 3, 1 / 2, 1 / 3),
    boxes_scales=(0.2, 0.9),
    centers_padding_from_borders_percentage=0.05,
    additional_square_box=True,
    image_shape=None,
) -> np.ndarray:
    """Anchors in 'ssd' order as (A, 4) float32 xyxy.

    float64 math, cast to float32 on store (boxes.py:139-146), then an fp32
    multiply by a Python-float factor for the rescale (boxes.py:167-174).
    Order: map-major, then row, column, shape index (boxes.py:198-205).
    """
    n_maps = len(feature_maps_shapes)
    if isinstance(centers_padding_from_borders_percentage, float):
        paddings = (centers_padding_from_borders_percentage,) * n_maps
    else:
        paddings = tuple(centers_padding_from_borders_percentage)
    if all(isinstance(r, (int, float)) for r in feature_maps_aspect_ratios):
        ratios = tuple(feature_maps_aspect_ratios for _ in range(n_maps))
    else:
        ratios = feature_maps_aspect_ratios
    scales = np.linspace(boxes_scales[0], boxes_scales[1], n_maps + 1)  # boxes.py:47
    out = []
    for m, ((fh, fw), rs, pad) in enumerate(zip(feature_maps_shapes, ratios, paddings)):
        size = min(fh, fw)
        s_cur, s_next = scales[m], scales[m + 1]
        hw = [[size * s_cur / math.sqrt(r), size * s_cur * math.sqrt(r)] for r in rs]  # boxes.py:98-101
        if additional_square_box:
            e = size * math.sqrt(s_cur * s_next)  # boxes.py:105
            hw.append([e, e])
        hw = np.array(hw)

        def centers(n):
            if n == 1:
                return np.array([0.5])  # boxes.py:112-113
            p = pad * (n - 1.0)
            return np.linspace(p, n - 1.0 - p, n)  # boxes.py:115-120

        cy, cx = centers(fh), centers(fw)
        boxes = np.zeros((fh, fw, len(hw), 4), dtype=F32)
        half_w = (hw[:, 1] - 1.0) / 2.0
        half_h = (hw[:, 0] - 1.0) / 2.0
        boxes[..., 0] = cx[None, :, None] - half_w[None, None, :]
        boxes[..., 1] = cy[:, None, None] - half_h[None, None, :]
        boxes[..., 2] = cx[None, :, None] + half_w[None, None, :]
        boxes[..., 3] = cy[:, None, None] + half_h[None, None, :]
        if image_shape is not None:
            fx = (image_shape[1] - 1) / (fw - 1 if fw > 1 else 1)  # boxes.py:167
            fy = (image_shape[0] - 1) / (fh - 1 if fh > 1 else 1)  # boxes.py:168
            boxes[..., [0, 2]] = boxes[..., [0, 2]] * fx
            boxes[..., [1, 3]] = boxes[..., [1, 3]] * fy
        out.append(boxes.reshape(-1, 4))
    return np.concatenate(out, axis=0)


def corners_to_centroids(xmin, ymin, xmax, ymax):
    """boxes.py:391-417 / datacoder.py:119-146."""
    return (xmax + xmin) / 2.0, (ymax + ymin) / 2.0, xmax - xmin + 1.0, ymax - ymin + 1.0


def centroids_to_corners(cx, cy, w, h):
    """boxes.py:419-445 / datacoder.py:148-175."""
    return cx - (w - 1.0) / 2.0, cy - (h - 1.0) / 2.0, cx + (w - 1.0) / 2.0, cy + (h - 1.0) / 2.0


# --------------------------------------------------------------------------
# a7: ground-truth CSV ingest (datacoder.py:193-203, evaluators.py:109-121)
# --------------------------------------------------------------------------
def parse_gt_csv_text(text: str, sep: str = "\r\n") -> Tuple[np.ndarray, np.ndarray]:
    """label,xmin,ymin,xmax,ymax records.  datacoder.py:194-196 strips the text
    and splits on '\\r\\n' only; the eval path (csv.reader) accepts any newline.
    Pass sep=None for csv.reader behaviour."""
    text = text.strip()
    if sep is None:
        rows = [r for r in csv.reader(text.splitlines()) if r]
    else:
        rows = [r.split(",") for r in text.split(sep)] if text else []
    labels = np.array([int(r[0]) for r in rows], dtype=np.int32)
    boxes = np.array([[float(v) for v in r[1:5]] for r in rows], dtype=F32).reshape(-1, 4)
    return labels, boxes


def flip_boxes_horizontal(boxes_xyxy: np.ndarray, image_width) -> np.ndarray:
    """datacoder.py:202-203: xmin' = W - xmax, xmax' = W - xmin (W, not W-1)."""
    out = boxes_xyxy.copy()
    out[:, 0] = image_width - boxes_xyxy[:, 2]
    out[:, 2] = image_width - boxes_xyxy[:, 0]
    return out.astype(F32)


# --------------------------------------------------------------------------
# a8: IoU matrix (datacoder.py:205-220)  [P1]
# --------------------------------------------------------------------------
def iou_matrix(anchors_xyxy: np.ndarray, gt_xyxy: np.ndarray) -> np.ndarray:
    ax0, ay0, ax1, ay1 = [anchors_xyxy[:, i].astype(F32) for i in range(4)]
    gx0, gy0, gx1, gy1 = [gt_xyxy[:, i].astype(F32) for i in range(4)]
    area_a = ((ay1 - ay0 + 1.0) * (ax1 - ax0 + 1.0))[:, None]  # datacoder.py:111-114
    area_g = (gx1 - gx0 + 1.0) * (gy1 - gy0 + 1.0)  # :206
    ix0 = np.maximum(ax0[:, None], gx0[None, :])  # :210
    iy0 = np.maximum(ay0[:, None], gy0[None, :])
    ix1 = np.minimum(ax1[:, None], gx1[None, :])
    iy1 = np.minimum(ay1[:, None], gy1[None, :])
    inter = np.maximum(0.0, ix1 - ix0 + 1.0) * np.maximum(0.0, iy1 - iy0 + 1.0)  # :216
    iou = inter / (area_a + area_g[None, :] - inter)  # :220
    assert iou.dtype == F32
    return iou


# --------------------------------------------------------------------------
# a9: matching (datacoder.py:222-244)  [P2, P3]
# --------------------------------------------------------------------------
def match_literal(iou: np.ndarray, iou_threshold) -> np.ndarray:
    """Literal three steps: per-GT argmax (max > 0), per-anchor argmax
    (max > thr), concat, unique rows in first-occurrence order.  Returns the
    (M, 2) int32 [anchor, gt] list in the order the scatter consumes it."""
    A, G = iou.shape
    thr = F32(iou_threshold)
    s1 = np.stack([np.argmax(iou, axis=0).astype(np.int32), np.arange(G, dtype=np.int32)], axis=1)
    s1 = s1[np.max(iou, axis=0) > 0.0]  # :230-231
    s2 = np.stack([np.arange(A, dtype=np.int32), np.argmax(iou, axis=1).astype(np.int32)], axis=1)
    s2 = s2[np.max(iou, axis=1) > thr]  # :236-241
    rows = np.concatenate([s1, s2], axis=0)
    seen = set()
    keep = []
    for r in rows:  # UniqueV2(axis=0): first occurrence, input order (:244)
        t = (int(r[0]), int(r[1]))
        if t not in seen:
            seen.add(t)
            keep.append(t)
    return np.array(keep, dtype=np.int32).reshape(-1, 2)


def match_closed_form(iou: np.ndarray, iou_threshold) -> np.ndarray:
    """P3 closed form: matched GT index per anchor, -1 = background.

    S1(a) = {g : argmax_a iou[:, g] == a and max_a iou[:, g] > 0};
    g2(a) = argmax_g iou[a, :] if max_g iou[a, :] > thr.
    g2 exists and g2 not in S1(a) -> g2; else S1(a) non-empty -> max(S1(a));
    else background.  (What the sequential last-writer-wins scatter yields.)"""
    A, G = iou.shape
    thr = F32(iou_threshold)
    best_a = np.argmax(iou, axis=0)
    valid_g = np.max(iou, axis=0) > 0.0
    g2 = np.argmax(iou, axis=1)
    has_g2 = np.max(iou, axis=1) > thr
    s1max = np.full(A, -1, dtype=np.int64)
    for g in range(G):
        if valid_g[g]:
            s1max[best_a[g]] = g  # ascending g => last is the max
    match = np.full(A, -1, dtype=np.int32)
    for a in range(A):
        if has_g2[a]:
            gg = g2[a]
            in_s1 = valid_g[gg] and best_a[gg] == a
            match[a] = s1max[a] if in_s1 else gg
        elif s1max[a] >= 0:
            match[a] = s1max[a]
    return match


# --------------------------------------------------------------------------
# a10-a11: offsets and assembly (datacoder.py:246-300)  [P4]
# --------------------------------------------------------------------------
def encode_offsets(anchors_xyxy, gt_xyxy, a_idx, g_idx, std):
    """tx,ty,tw,th for the matched (anchor, gt) pairs; two successive divisions
    and the +1.0 inside the log exactly as datacoder.py:266-269."""
    sx, sy, sw, sh = std
    dcx, dcy, dw, dh = corners_to_centroids(*[anchors_xyxy[a_idx, i].astype(F32) for i in range(4)])
    gcx, gcy, gw, gh = corners_to_centroids(*[gt_xyxy[g_idx, i].astype(F32) for i in range(4)])
    tx = (gcx - dcx) / dw / sx
    ty = (gcy - dcy) / dh / sy
    tw = np.log(gw / dw + 1.0) / sw
    th = np.log(gh / dh + 1.0) / sh
    out = np.stack([tx, ty, tw, th], axis=1)
    assert out.dtype == F32
    return out


def encode_image_literal(anchors_xyxy, gt_labels, gt_xyxy, num_classes, iou_threshold=0.5,
                         std=(0.1, 0.1, 0.2, 0.2)):
    """datacoder.py:205-300 statement by statement.  Returns (labels (A,C) f32,
    offsets (A,4) f32, match (A,) i32 with -1 = background)."""
    A = anchors_xyxy.shape[0]
    iou = iou_matrix(anchors_xyxy, gt_xyxy)
    idx = match_literal(iou, iou_threshold)
    enc = np.concatenate([np.ones((A, 1), F32), np.zeros((A, num_classes + 3), F32)], axis=1)  # :274-283
    match = np.full(A, -1, np.int32)
    if idx.shape[0]:
        lab = gt_labels[idx[:, 1]]
        onehot = (lab[:, None] == np.arange(num_classes)[None, :]).astype(F32)  # tf.one_hot: all-zero if out of range
        offs = encode_offsets(anchors_xyxy, gt_xyxy, idx[:, 0], idx[:, 1], std)
        upd = np.concatenate([onehot, offs], axis=1)
        for r in range(idx.shape[0]):  # sequential scatter, last writer wins (:286-298)
            enc[idx[r, 0]] = upd[r]
            match[idx[r, 0]] = idx[r, 1]
    return enc[:, :-4].copy(), enc[:, -4:].copy(), match


def encode_image(anchors_xyxy, gt_labels, gt_xyxy, num_classes, iou_threshold=0.5,
                 std=(0.1, 0.1, 0.2, 0.2)):
    """Same result as encode_image_literal through the P3 closed form (the
    formulation the CUDA kernels implement)."""
    A = anchors_xyxy.shape[0]
    iou = iou_matrix(anchors_xyxy, gt_xyxy)
    match = match_closed_form(iou, iou_threshold)
    labels = np.zeros((A, num_classes), F32)
    labels[:, 0] = 1.0
    offsets = np.zeros((A, 4), F32)
    pos = np.nonzero(match >= 0)[0]
    if pos.size:
        g = match[pos]
        labels[pos] = (gt_labels[g][:, None] == np.arange(num_classes)[None, :]).astype(F32)
        offsets[pos] = encode_offsets(anchors_xyxy, gt_xyxy, pos, g, std)
    return labels, offsets, match


def encode_batch(anchors_xyxy, gt_labels, gt_xyxy, gt_counts, num_classes, iou_threshold=0.5,
                 std=(0.1, 0.1, 0.2, 0.2), literal=False):
    """Padded batch front-end: gt_xyxy (B,Gmax,4), gt_labels (B,Gmax), gt_counts (B,).
    An image with zero GT boxes is all background (the reference cannot encode
    it: arg-max over an empty axis)."""
    B = gt_xyxy.shape[0]
    A = anchors_xyxy.shape[0]
    labels = np.zeros((B, A, num_classes), F32)
    offsets = np.zeros((B, A, 4), F32)
    match = np.full((B, A), -1, np.int32)
    fn = encode_image_literal if literal else encode_image
    for b in range(B):
        n = int(gt_counts[b])
        if n == 0:
            labels[b, :, 0] = 1.0
            continue
        labels[b], offsets[b], match[b] = fn(anchors_xyxy, gt_labels[b, :n], gt_xyxy[b, :n], num_classes,
                                             iou_threshold, std)
    return labels, offsets, match


# --------------------------------------------------------------------------
# a13-a15: decode (datacoder.py:349-432, layers.py:45-81)  [P5]
# --------------------------------------------------------------------------
def _decode_centroids_raw(offsets, anchors_cxcywh, std):
    sx, sy, sw, sh = std
    dcx, dcy, dw, dh = [anchors_cxcywh[:, i].astype(F32) for i in range(4)]
    tx, ty, tw, th = [offsets[..., i].astype(F32) for i in range(4)]
    cx = tx * sx * dw + dcx  # ((t*s)*d)+c, datacoder.py:371 / layers.py:67
    cy = ty * sy * dh + dcy
    w = (np.exp(tw * sw) - 1.0) * dw  # datacoder.py:373 / layers.py:69
    h = (np.exp(th * sh) - 1.0) * dh
    return cx, cy, w, h


def decode_to_centroids(offsets, anchors_cxcywh, std=(0.1, 0.1, 0.2, 0.2)):
    """datacoder.py:349-388: single image (A,4) -> (A,4) cx,cy,w,h with rows whose
    sum(|offsets|) == 0 forced to zero."""
    cx, cy, w, h = _decode_centroids_raw(offsets, anchors_cxcywh, std)
    nb = (np.sum(np.abs(offsets.astype(F32)), axis=-1) > 0.0).astype(F32)  # :377-378
    return np.stack([cx * nb, cy * nb, w * nb, h * nb], axis=-1)


def decode_to_corners(offsets, anchors_cxcywh, std=(0.1, 0.1, 0.2, 0.2)):
    """datacoder.py:390-432: masked centroids -> corners -> masked again by
    sum(|centroids|) > 0.  Output xyxy."""
    c = decode_to_centroids(offsets, anchors_cxcywh, std)
    x0, y0, x1, y1 = centroids_to_corners(c[..., 0], c[..., 1], c[..., 2], c[..., 3])
    nb = (np.sum(np.abs(c), axis=-1) > 0.0).astype(F32)  # :421-422
    return np.stack([x0 * nb, y0 * nb, x1 * nb, y1 * nb], axis=-1)


def decode_batched_yxyx(offsets, anchors_cxcywh, std=(0.1, 0.1, 0.2, 0.2)):
    """layers.py:57-79: (B,A,4) -> (B,A,4) (ymin,xmin,ymax,xmax), no masking."""
    cx, cy, w, h = _decode_centroids_raw(offsets, anchors_cxcywh, std)
    x0, y0, x1, y1 = centroids_to_corners(cx, cy, w, h)
    out = np.stack([y0, x0, y1, x1], axis=-1)
    assert out.dtype == F32
    return out


# --------------------------------------------------------------------------
# a16: combined NMS (layers.py:127-168 + TF CombinedNonMaxSuppression)  [P6]
# --------------------------------------------------------------------------
def _nms_iou(bi, bj) -> np.float32:
    """TF non_max_suppression_op.cc IOU(): min/max-normalised corners, no +1,
    0 if either area <= 0."""
    y0i, y1i = min(bi[0], bi[2]), max(bi[0], bi[2])
    x0i, x1i = min(bi[1], bi[3]), max(bi[1], bi[3])
    y0j, y1j = min(bj[0], bj[2]), max(bj[0], bj[2])
    x0j, x1j = min(bj[1], bj[3]), max(bj[1], bj[3])
    ai = F32(F32(y1i - y0i) * F32(x1i - x0i))
    aj = F32(F32(y1j - y0j) * F32(x1j - x0j))
    if ai <= 0 or aj <= 0:
        return F32(0.0)
    ih = max(F32(min(y1i, y1j) - max(y0i, y0j)), F32(0.0))
    iw = max(F32(min(x1i, x1j) - max(x0i, x0j)), F32(0.0))
    inter = F32(ih * iw)
    return F32(inter / F32(F32(ai + aj) - inter))


def nms_per_class(boxes_yxyx: np.ndarray, scores: np.ndarray, max_out: int, iou_thr, score_thr) -> List[int]:
    """Greedy NMS for one (image, class): candidates score > thr, visited by
    (score desc, index asc); suppressed iff IoU > iou_thr with a kept box."""
    iou_thr = F32(iou_thr)
    score_thr = F32(score_thr)
    cand = np.nonzero(scores > score_thr)[0]
    order = cand[np.lexsort((cand, -scores[cand].astype(np.float64)))]
    b = boxes_yxyx.astype(F32)
    # vectorised pre-normalisation (identical arithmetic to _nms_iou)
    y0 = np.minimum(b[:, 0], b[:, 2]); y1 = np.maximum(b[:, 0], b[:, 2])
    x0 = np.minimum(b[:, 1], b[:, 3]); x1 = np.maximum(b[:, 1], b[:, 3])
    area = (y1 - y0) * (x1 - x0)
    keep: List[int] = []
    limit = min(max_out, boxes_yxyx.shape[0])
    for i in order:
        if len(keep) >= limit:
            break
        ok = True
        if keep and area[i] > 0:
            k = np.array(keep)
            ih = np.maximum(np.minimum(y1[i], y1[k]) - np.maximum(y0[i], y0[k]), F32(0.0))
            iw = np.maximum(np.minimum(x1[i], x1[k]) - np.maximum(x0[i], x0[k]), F32(0.0))
            inter = ih * iw
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = inter / ((area[i] + area[k]) - inter)
            iou = np.where(area[k] > 0, iou, F32(0.0))
            ok = not bool(np.any(iou > iou_thr))
        if ok:
            keep.append(int(i))
    return keep


def combined_nms(boxes_yxyx, scores, max_per_class, max_total, iou_thr, score_thr):
    """layers.py:141-162: returns (out (B,K,6) = [label, prob, xmin, ymin, xmax,
    ymax], valid (B,), anchor_idx (B,K) with -1 padding).  Cross-class merge:
    score descending; ties (unspecified in TF) -> lower class, then lower
    anchor index."""
    B, A, C = scores.shape
    K = int(max_total)
    out = np.zeros((B, K, 6), F32)
    valid = np.zeros(B, np.int32)
    aidx = np.full((B, K), -1, np.int32)
    for b in range(B):
        picks = []
        for c in range(C):
            for i in nms_per_class(boxes_yxyx[b], scores[b, :, c], max_per_class, iou_thr, score_thr):
                picks.append((-float(scores[b, i, c]), c, i))
        picks.sort()
        picks = picks[:K]
        valid[b] = len(picks)
        for k, (ns, c, i) in enumerate(picks):
            y0, x0, y1, x1 = boxes_yxyx[b, i]
            out[b, k] = (c, scores[b, i, c], x0, y0, x1, y1)  # gather [1,0,3,2], layers.py:155-162
            aidx[b, k] = i
    return out, valid, aidx


# --------------------------------------------------------------------------
# segmentation suppression (layers.py:189-212)
# --------------------------------------------------------------------------
def seg_class_presence(segmentation_mask: np.ndarray, depth: int = 4) -> np.ndarray:
    """layers.py:203-207: argmax over the class axis (lowest index on ties) ->
    one_hot(depth) -> sum over (batch, height, width) -> clip to [0, 1]."""
    cls = np.argmax(np.asarray(segmentation_mask, F32), axis=-1).reshape(-1)
    present = np.zeros(depth, F32)
    seen = np.unique(cls)
    present[seen[seen < depth]] = F32(1.0)  # one_hot of an index >= depth is an all-zero row
    return present


def segmentation_suppression(segmentation_mask: np.ndarray, labels_probabilities: np.ndarray) -> np.ndarray:
    """layers.py:209: probabilities * is_class_segmented (broadcast over the last axis)."""
    return (np.asarray(labels_probabilities, F32) * seg_class_presence(segmentation_mask, 4)).astype(F32)


# --------------------------------------------------------------------------
# training-time box IoU metric (metrics.py:53-173)
# --------------------------------------------------------------------------
def metric_iou_boxes(y_true, y_pred, anchors_cxcywh, std=(0.1, 0.1, 0.2, 0.2)) -> np.ndarray:
    """metrics.py:125-171 statement by statement: (batch,) mean IoU of the non-background anchors."""
    y_true = np.asarray(y_true, F32)
    y_pred = np.asarray(y_pred, F32)
    a = np.asarray(anchors_cxcywh, F32)
    sx, sy, sw, sh = (F32(s) for s in std)
    nb = (np.sum(np.abs(y_true), axis=-1, dtype=F32) > 0).astype(F32)            # :138-139

    def dec(t):                                                                   # :79-121
        cx = (t[..., 0] * sx * a[:, 2] + a[:, 0]) * nb
        cy = (t[..., 1] * sy * a[:, 3] + a[:, 1]) * nb
        w = np.maximum(F32(0), (np.exp(t[..., 2] * sw) - F32(1)) * a[:, 2]) * nb
        h = np.maximum(F32(0), (np.exp(t[..., 3] * sh) - F32(1)) * a[:, 3]) * nb
        x0 = (cx - (w - F32(1)) / F32(2)) * nb
        y0 = (cy - (h - F32(1)) / F32(2)) * nb
        x1 = (cx + (w - F32(1)) / F32(2)) * nb
        y1 = (cy + (h - F32(1)) / F32(2)) * nb
        return x0, y0, x1, y1, w, h

    px0, py0, px1, py1, pw, ph = dec(y_pred)
    tx0, ty0, tx1, ty1, tw, th = dec(y_true)
    wi = np.maximum(F32(0), np.minimum(tx1, px1) - np.maximum(tx0, px0) + F32(1)) * nb
    hi = np.maximum(F32(0), np.minimum(ty1, py1) - np.maximum(ty0, py0) + F32(1)) * nb
    ai, at, ap = wi * hi, tw * th, pw * ph
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = ai / (ap + at - ai + F32(1e-7))
        return (np.sum(iou, axis=-1, dtype=F32) / np.sum(nb, axis=-1, dtype=F32)).astype(F32)


# --------------------------------------------------------------------------
# 8(f) rank 3: training losses that consume the encoded targets (losses.py:5-172)
# --------------------------------------------------------------------------
def localization_loss(y_true, y_pred, with_gradient=False):
    """losses.py:5-50 statement by statement: smooth-L1 over the non-background anchors of each image, divided by
    their count (at least 1); (batch,) fp32.  with_gradient: also d loss[b] / d y_pred[b] (what TF's autodiff gives:
    -d inside |d| < 1, -sign(d) outside, times the mask over the count)."""
    y_true = np.asarray(y_true, F32)
    y_pred = np.asarray(y_pred, F32)
    nb = (np.sum(np.abs(y_true), axis=-1, dtype=F32) > 0).astype(F32)             # :23-24
    d = y_true - y_pred
    ea, es = np.abs(d), d * d                                                      # :27-28 (pow(x, 2) = x * x)
    quad = ea < F32(1)                                                             # :31
    l = np.where(quad, es * F32(0.5), ea - F32(0.5))                               # :34
    l = np.sum(l, axis=-1, dtype=F32) * nb                                         # :37-40
    den = np.maximum(np.sum(nb, axis=-1, dtype=F32), F32(1))                       # :47
    loss = (np.sum(l, axis=-1, dtype=F32) / den).astype(F32)                       # :43-47
    if not with_gradient:
        return loss
    g = np.where(quad, -d, -np.sign(d)).astype(F32) * nb[..., None] / den[:, None, None]
    return loss, g.astype(F32)


def confidence_loss(y_true, y_pred, epsilon=1e-7, with_gradient=False, return_keep=False):
    """losses.py:52-172 statement by statement: soft-max cross-entropy with hard-negative mining.  The k =
    min(3 * positives, negatives) background anchors with the highest loss are kept, k and the top-k taken over the
    WHOLE batch (:117, :134-138; equal losses: lower flat index first, as tf.math.top_k), every loss divided by the
    image's own count of non-background anchors (at least 1); (batch,) fp32."""
    y_true = np.asarray(y_true, F32)
    y_pred = np.asarray(y_pred, F32)
    B, A, _ = y_true.shape
    is_bg = y_true[:, :, 0]                                                        # :69
    not_bg = np.abs(is_bg - F32(1))                                                # :70
    n_bg, n_nb = int(np.count_nonzero(is_bg)), int(np.count_nonzero(not_bg))       # :73-74
    lo, hi = F32(epsilon), F32(1 - epsilon)                                        # :89-90 (clip bounds cast to fp32)
    clipped = np.clip(y_pred, lo, hi)
    logp = np.log(clipped)
    la = -np.sum(y_true * logp, axis=-1, dtype=F32)                                # :93
    l_nb = np.sum(la * not_bg, axis=-1, dtype=F32)                                 # :100-101
    cnt_nb = np.sum(not_bg, axis=-1, dtype=F32)                                    # :104
    k = min(3 * n_nb, n_bg)                                                        # :117
    v = la * is_bg                                                                 # :120
    keep = np.zeros(B * A, F32)
    if n_bg > 0:                                                                   # :123
        order = np.argsort(-v.reshape(-1), kind="stable")[:k]                      # :134-138
        keep[order] = 1                                                            # :143-152
    keep = keep.reshape(B, A)
    l_bg = np.sum(v * keep, axis=-1, dtype=F32)                                    # :156-159
    den = np.maximum(cnt_nb, F32(1))
    loss = ((l_nb + l_bg) / den).astype(F32)                                       # :169-172
    out = [loss]
    if with_gradient:
        w = not_bg + is_bg * keep                                                  # weight of each anchor's loss
        passes = ((y_pred >= lo) & (y_pred <= hi)).astype(F32)                     # clip_by_value passes inside [lo, hi]
        g = -(w / den[:, None])[..., None] * y_true / clipped * passes
        out.append(g.astype(F32))
    if return_keep:
        out.append(keep.astype(np.uint8))
        out.append(v.astype(F32))
    return out[0] if len(out) == 1 else tuple(out)


# --------------------------------------------------------------------------
# a17-a18: detection evaluation (evaluators.py:6-187)  [P7]
# --------------------------------------------------------------------------
def iou_pred_vs_true(labels_pred, boxes_pred, labels_true, boxes_true) -> np.ndarray:
    """evaluators.py:6-63: +1 pixel convention, eps 1e-7 in the denominator,
    multiplied by the label-equality mask; (P,1) zeros when there is no GT."""
    P = boxes_pred.shape[0]
    if len(labels_true) == 0:
        return np.zeros((P, 1), F32)
    bp = boxes_pred.astype(F32); bt = boxes_true.astype(F32)
    same = (np.asarray(labels_pred)[:, None] == np.asarray(labels_true)[None, :]).astype(F32)
    ix0 = np.maximum(bp[:, 0:1], bt[:, 0][None, :]); iy0 = np.maximum(bp[:, 1:2], bt[:, 1][None, :])
    ix1 = np.minimum(bp[:, 2:3], bt[:, 2][None, :]); iy1 = np.minimum(bp[:, 3:4], bt[:, 3][None, :])
    ap = (bp[:, 2:3] - bp[:, 0:1] + 1.0) * (bp[:, 3:4] - bp[:, 1:2] + 1.0)
    at = ((bt[:, 2] - bt[:, 0] + 1.0) * (bt[:, 3] - bt[:, 1] + 1.0))[None, :]
    inter = np.maximum(0.0, ix1 - ix0 + 1.0) * np.maximum(0.0, iy1 - iy0 + 1.0)
    iou = inter / (ap + at - inter + 1e-7)
    iou = iou * same
    assert iou.dtype == F32
    return iou


def eval_tp_flags(labels_pred, boxes_pred, labels_true, boxes_true, iou_threshold, label_code_background=0):
    """evaluators.py:129-153 for one image: tp flag per prediction (0 for the
    background predictions, which the reference drops before scoring)."""
    thr = F32(iou_threshold)
    tp = np.zeros(len(labels_pred), np.uint8)
    nb = np.asarray(labels_pred) != label_code_background
    if nb.any():
        iou = iou_pred_vs_true(np.asarray(labels_pred)[nb], boxes_pred[nb], labels_true, boxes_true)
        tp[nb] = (np.max(iou, axis=1) >= thr).astype(np.uint8)
    return tp


def average_precision(labels_pred_batch, confidences_pred_batch, boxes_pred_batch, iou_threshold,
                      gt_labels_list: Sequence[np.ndarray], gt_boxes_list: Sequence[np.ndarray],
                      labels_codes, label_code_background) -> Dict[int, float]:
    """evaluators.py:65-187 with the CSV read replaced by in-memory GT arrays."""
    thr = F32(iou_threshold)
    per_label = {l: [] for l in labels_codes if l != label_code_background}
    n_gt = {l: 0 for l in labels_codes if l != label_code_background}
    for lp, cp, bp, lt, bt in zip(labels_pred_batch, confidences_pred_batch, boxes_pred_batch, gt_labels_list,
                                  gt_boxes_list):
        for l in lt:
            n_gt[int(l)] += 1
        nb = lp != label_code_background
        lp, cp, bp = lp[nb], cp[nb], bp[nb]
        if len(lp) == 0:
            continue
        iou = iou_pred_vs_true(lp, bp, np.asarray(lt, np.int32), np.asarray(bt, F32).reshape(-1, 4))
        tp = (np.max(iou, axis=1) >= thr).astype(np.int32)
        for l, c, t in zip(lp, cp, tp):
            per_label[int(l)].append((t, c))
    ap = {}
    for l, lst in per_label.items():
        if n_gt[l] == 0 or len(lst) == 0:
            ap[l] = 0.0
            continue
        arr = np.array(lst, dtype=F32)
        order = np.argsort(arr[:, 1])[::-1]
        t = arr[order, 0]
        prec = np.cumsum(t) / np.arange(1, len(t) + 1)
        rec = np.cumsum(t) / n_gt[l]
        ap[l] = float(np.sum((rec[1:] - rec[:-1]) * (prec[1:] + prec[:-1]) / 2.0))  # np.trapz, evaluators.py:185
    return ap


def average_precision_from_tp(tp, confidences, labels_pred, n_gt_per_label: Dict[int, int]) -> Dict[int, float]:
    """evaluators.py:159-187 on flat arrays of TP flags (the per-image part, :106-157, already done): per class,
    argsort(conf)[::-1], precision = cumsum(tp) / rank, recall = cumsum(tp) / n_gt, np.trapz(precision, recall) with
    no (0, .) sentinel; 0.0 for a class without ground truth or without predictions."""
    tp = np.asarray(tp).reshape(-1)
    conf = np.asarray(confidences, dtype=F32).reshape(-1)
    lab = np.asarray(labels_pred).reshape(-1)
    ap = {}
    for l, n_gt in n_gt_per_label.items():
        sel = lab == l
        if n_gt == 0 or not sel.any():
            ap[l] = 0.0
            continue
        order = np.argsort(conf[sel])[::-1]
        t = tp[sel][order].astype(F32)
        prec = np.cumsum(t) / np.arange(1, len(t) + 1)
        rec = np.cumsum(t) / n_gt
        ap[l] = float(np.sum((rec[1:] - rec[:-1]) * (prec[1:] + prec[:-1]) / 2.0))
    return ap


def synth_eval(rng, B, K=200, G=16, C=4, jitter=8.0):
    """BASELINE config 5 inputs: GT boxes of the config-1 generator; predictions in the NMS output layout - K rows
    per image = a GT box of the image jittered by N(0, jitter) px, labels U{0..C-1} (0 = background rows), distinct
    confidences (the tie order of the reference's argsort is unspecified)."""
    labels, boxes, counts = synth_gt(rng, B, G)
    pl = rng.integers(0, C, size=(B, K)).astype(np.int32)
    src = rng.integers(0, G, size=(B, K))
    pb = (np.take_along_axis(boxes, src[..., None].repeat(4, -1), axis=1)
          + rng.normal(0, jitter, size=(B, K, 4))).astype(F32)
    conf = (rng.permutation(B * K).astype(np.float64) / float(B * K)).astype(F32).reshape(B, K)
    return pl, conf, pb, labels, boxes, counts


# --------------------------------------------------------------------------
# SURVEY.md 8(d) synthetic generators (seeded)
# --------------------------------------------------------------------------
STD_ANCHOR_CFG = dict(
    feature_maps_shapes=((30, 40), (15, 20), (8, 10), (4, 5)),
    centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1),
    boxes_scales=(0.15, 0.95),
    additional_square_box=True,
    image_shape=(480, 640),
)
STRESS_ANCHOR_CFG = dict(
    feature_maps_shapes=((125, 160),),
    feature_maps_aspect_ratios=(1, 2, 3, 1 / 2, 1 / 3),
    additional_square_box=False,
    image_shape=(1000, 1280),
)


def synth_gt(rng, B, G, image_hw=(480, 640), n_fg=3, wh_range=(20.0, 300.0), integer=False):
    """GT boxes [x0,y0,x0+w-1,y0+h-1] with w,h ~ U(range), labels ~ U{1..n_fg}."""
    H, W = image_hw
    w = rng.uniform(wh_range[0], min(wh_range[1], W), size=(B, G))
    h = rng.uniform(wh_range[0], min(wh_range[1], H), size=(B, G))
    x0 = rng.uniform(0, 1, size=(B, G)) * (W - w)
    y0 = rng.uniform(0, 1, size=(B, G)) * (H - h)
    if integer:
        w, h, x0, y0 = np.floor(w), np.floor(h), np.floor(x0), np.floor(y0)
    boxes = np.stack([x0, y0, x0 + w - 1, y0 + h - 1], axis=-1).astype(F32)
    labels = rng.integers(1, n_fg + 1, size=(B, G)).astype(np.int32)
    counts = np.full(B, G, np.int32)
    return labels, boxes, counts


def softmax_f32(logits):
    z = logits.astype(F32)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


def dedup_scores(scores: np.ndarray) -> np.ndarray:
    """Make scores unique per (image, class) by nextafter bumps (NMS parity is
    defined on tie-free scores)."""
    s = scores.copy()
    B, A, C = s.shape
    for b in range(B):
        for c in range(C):
            col = s[b, :, c]
            while True:
                order = np.argsort(col, kind="stable")
                sorted_col = col[order]
                dup = np.nonzero(sorted_col[1:] == sorted_col[:-1])[0]
                if dup.size == 0:
                    break
                col[order[dup + 1]] = np.nextafter(sorted_col[dup + 1], F32(2.0))
            s[b, :, c] = col
    return s


def synth_predictions(rng, B, A, C=4, wh_mean=math.log(2.0) / 0.2, wh_sd=0.5):
    """Config-2 inputs: offsets xy~N(0,1), wh~N(ln2/0.2, .5); scores = softmax of
    N(0,1) logits, de-duplicated per (image, class)."""
    off = np.empty((B, A, 4), F32)
    off[..., :2] = rng.normal(0, 1, size=(B, A, 2))
    off[..., 2:] = rng.normal(wh_mean, wh_sd, size=(B, A, 2))
    scores = dedup_scores(softmax_f32(rng.normal(0, 1, size=(B, A, C))))
    return off, scores
