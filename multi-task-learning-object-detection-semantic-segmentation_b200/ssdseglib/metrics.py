"""Detection side of the reference's `ssdseglib/metrics.py`: the training-time box IoU metric.

`jaccard_iou_bounding_boxes` keeps the reference's factory signature (metrics.py:53-59) and returns a callable
`(y_true, y_pred) -> (batch,)`.  The segmentation metric and the categorical accuracy of the reference are untouched
(not on the SSD box path).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _ssdx
from .boxes import coordinates_centroids_to_corners


def jaccard_iou_bounding_boxes(center_x_boxes_default, center_y_boxes_default, width_boxes_default,
                               height_boxes_default, standard_deviations_centroids_offsets,
                               device: Optional[int] = None) -> Callable:
    """Jaccard IoU metric for object detection regression data of shape (batch, total boxes, 4): predictions and
    targets are standardized centroid offsets; the value per batch item is the average IoU of the non-background
    boxes (metrics.py:53-173).  One fused kernel reads both tensors once."""
    cent = np.stack([np.asarray(v, dtype=np.float32).reshape(-1) for v in (
        center_x_boxes_default, center_y_boxes_default, width_boxes_default, height_boxes_default)], axis=1)
    std = tuple(float(np.float32(s)) for s in standard_deviations_centroids_offsets)
    if len(std) != 4:
        raise ValueError("standard_deviations_centroids_offsets must have 4 entries")
    corners = np.stack(coordinates_centroids_to_corners(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3]), axis=1)
    anchors = _ssdx.Anchors(corners, cent, device)

    def jaccard_iou_bounding_boxes_metric(y_true, y_pred):
        yt, kind = _ssdx.to_device(y_true, np.float32, anchors.device, "metric_true")
        yp, _ = _ssdx.to_device(y_pred, np.float32, anchors.device, "metric_pred")
        if yt.dim() != 3 or tuple(yt.shape) != tuple(yp.shape) or int(yt.shape[1]) != anchors.num_anchors or \
                int(yt.shape[2]) != 4:
            raise ValueError(f"expected y_true and y_pred of shape (batch, {anchors.num_anchors}, 4)")
        return _ssdx.from_device(_ssdx.metric_iou_boxes(anchors, yt, yp, std), kind)

    return jaccard_iou_bounding_boxes_metric
