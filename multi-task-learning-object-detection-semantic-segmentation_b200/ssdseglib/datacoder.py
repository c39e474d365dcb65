"""SSD target encode / decode - mirror of the reference's `ssdseglib/datacoder.py` over libssdx (CUDA).

`DataEncoderDecoder` keeps the reference's constructor, attributes and methods (datacoder.py:5-432); the
arithmetic (IoU matrix, matching, offsets, scatter, decode) runs in the sm_100a kernels behind
`include/ssdx.h`.  One ctypes call per image cannot feed a B200, so the batched entry points
`encode_batch` / `decode_batch_to_corners` are added next to the per-image ones.

Array arguments may be NumPy arrays (host: staged through pinned memory, results come back as NumPy),
torch tensors (CUDA tensors are used in place; results are CUDA tensors) or any `__dlpack__` object.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _ssdx
from .boxes import coordinates_centroids_to_corners, coordinates_corners_to_centroids


def _f32(x) -> np.ndarray:
    if hasattr(x, "detach"):  # torch tensor
        x = x.detach().cpu().numpy()
    elif hasattr(x, "numpy") and not isinstance(x, np.ndarray):  # tf tensor
        x = x.numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32).reshape(-1)


class DataEncoderDecoder:
    def __init__(
        self,
        num_classes: int,
        image_shape: Tuple[int, int],
        xmin_boxes_default=None,
        ymin_boxes_default=None,
        xmax_boxes_default=None,
        ymax_boxes_default=None,
        center_x_boxes_default=None,
        center_y_boxes_default=None,
        width_boxes_default=None,
        height_boxes_default=None,
        iou_threshold: float = 0.5,
        standard_deviations_centroids_offsets: Tuple[float, float, float, float] = (0.1, 0.1, 0.2, 0.2),
        augmentation_horizontal_flip: bool = False,
        device: Optional[int] = None,
    ) -> None:
        """Same arguments as the reference (datacoder.py:6-21) plus `device` (CUDA ordinal, default: current).

        Pass the default boxes as corners OR as centroids.  As in the reference, passing both raises
        (the "both" branch at datacoder.py:92-104 is unreachable and falls through to the error)."""
        self.num_classes = num_classes
        self.image_height, self.image_width = image_shape
        self.iou_threshold = iou_threshold
        (self.standard_deviation_center_x_offsets, self.standard_deviation_center_y_offsets,
         self.standard_deviation_width_offsets, self.standard_deviation_height_offsets) = standard_deviations_centroids_offsets
        self.augmentation_horizontal_flip = augmentation_horizontal_flip

        corners = (xmin_boxes_default, ymin_boxes_default, xmax_boxes_default, ymax_boxes_default)
        centroids = (center_x_boxes_default, center_y_boxes_default, width_boxes_default, height_boxes_default)
        if all(c is None for c in centroids):
            if any(c is None for c in corners):
                raise ValueError('you must pass all default bounding boxes corners coordinates!')
            x0, y0, x1, y1 = (_f32(c) for c in corners)
            cx, cy, w, h = coordinates_corners_to_centroids(x0, y0, x1, y1)  # fp32, datacoder.py:65-70
        elif all(c is None for c in corners):
            if any(c is None for c in centroids):
                raise ValueError('you must pass all default bounding boxes centroids coordinates!')
            cx, cy, w, h = (_f32(c) for c in centroids)
            x0, y0, x1, y1 = coordinates_centroids_to_corners(cx, cy, w, h)  # fp32, datacoder.py:84-89
        else:
            raise ValueError('you must pass all default bounding boxes centroids coordinates, or corners '
                             'coordinates or both!')
        self.xmin_boxes_default, self.ymin_boxes_default = x0, y0
        self.xmax_boxes_default, self.ymax_boxes_default = x1, y1
        self.center_x_boxes_default, self.center_y_boxes_default = cx, cy
        self.width_boxes_default, self.height_boxes_default = w, h
        self.boxes_area_default = ((y1 - y0 + 1.0) * (x1 - x0 + 1.0))[:, None]  # datacoder.py:111-114

        self._std = tuple(float(s) for s in standard_deviations_centroids_offsets)
        self._anchors = _ssdx.Anchors(np.stack([x0, y0, x1, y1], axis=1), np.stack([cx, cy, w, h], axis=1), device)

    # ------------------------------------------------------------------ helpers kept from the reference
    def _coordinates_corners_to_centroids(self, xmin, ymin, xmax, ymax):
        return coordinates_corners_to_centroids(xmin, ymin, xmax, ymax)

    def _coordinates_centroids_to_corners(self, center_x, center_y, width, height):
        return coordinates_centroids_to_corners(center_x, center_y, width, height)

    @property
    def num_anchors(self) -> int:
        return self._anchors.num_anchors

    @property
    def device(self) -> int:
        return self._anchors.device

    # ------------------------------------------------------------------ encode
    def read_labels_boxes(self, path_file_labels_boxes: str, augment_with_horizontal_flip: bool = False):
        """GT ingest of datacoder.py:194-203: strip, split on '\\r\\n', columns label,xmin,ymin,xmax,ymax,
        optional flip xmin' = W - xmax, xmax' = W - xmin.  Returns (labels int32 (G,), boxes float32 (G,4))."""
        if isinstance(path_file_labels_boxes, bytes):
            path_file_labels_boxes = path_file_labels_boxes.decode()
        with open(path_file_labels_boxes, "rb") as f:
            text = f.read()
        return _ssdx.parse_gt_csv(text, crlf_only=True,
                                  flip_width=float(self.image_width) if augment_with_horizontal_flip else 0.0)

    def read_labels_boxes_batch(self, path_files_labels_boxes, augment_with_horizontal_flip=None, threads: int = 0):
        """Batched GT ingest (new entry point next to `read_labels_boxes`): one file per image, read and parsed by C++
        worker threads.  `augment_with_horizontal_flip`: None, or one bool per file (the coin of datacoder.py:337).
        Returns padded (labels (N, G) int32, boxes (N, G, 4) float32, counts (N,) int32) ready for `encode_batch`."""
        return _ssdx.read_gt_csv_batch(list(path_files_labels_boxes), crlf_only=True,
                                       flips=augment_with_horizontal_flip, flip_width=float(self.image_width),
                                       threads=threads)

    def encode_batch(self, gt_boxes, gt_labels, gt_counts=None, return_match: bool = False):
        """Batched encode (new entry point).

        gt_boxes (B, Gmax, 4) xyxy, gt_labels (B, Gmax) int, gt_counts (B,) int or None (= Gmax everywhere).
        Returns labels (B, A, num_classes) one-hot float32 and offsets (B, A, 4) float32
        [+ match (B, A) int32, -1 = background].  Images with gt_counts == 0 encode to all background."""
        dev = self.device
        boxes_d, kind = _ssdx.to_device(gt_boxes, np.float32, dev, "gt_boxes")
        if boxes_d.dim() != 3 or boxes_d.shape[-1] != 4:
            raise ValueError("gt_boxes must have shape (batch, max boxes, 4)")
        B, G = int(boxes_d.shape[0]), int(boxes_d.shape[1])
        labels_d, _ = _ssdx.to_device(gt_labels, np.int32, dev, "gt_labels")
        if tuple(labels_d.shape) != (B, G):
            raise ValueError("gt_labels must have shape (batch, max boxes)")
        if gt_counts is None:
            gt_counts = np.full(B, G, np.int32)
        counts_d, _ = _ssdx.to_device(gt_counts, np.int32, dev, "gt_counts")
        if tuple(counts_d.shape) != (B,):
            raise ValueError("gt_counts must have shape (batch,)")
        lab, off, match = _ssdx.encode(self._anchors, boxes_d, labels_d, counts_d, self.num_classes,
                                       self.iou_threshold, self._std, want_match=return_match)
        out = (_ssdx.from_device(lab, kind), _ssdx.from_device(off, kind))
        if return_match:
            out = out + (_ssdx.from_device(match, kind),)
        return out

    def _encode_ground_truth_labels_boxes(self, path_file_labels_boxes: str, augment_with_horizontal_flip: bool):
        """Per-image encode, as the reference (datacoder.py:177-300): returns (labels (A, num_classes),
        offsets (A, 4)) float32 NumPy arrays.  An empty file raises (the reference's arg-max over an empty
        axis fails too)."""
        labels, boxes = self.read_labels_boxes(path_file_labels_boxes, bool(augment_with_horizontal_flip))
        if labels.shape[0] == 0:
            raise ValueError(f"no ground truth boxes in {path_file_labels_boxes!r}")
        lab, off = self.encode_batch(boxes[None], labels[None])
        return lab[0], off[0]

    def _encode_np(self, path_file_labels_boxes, augment_with_horizontal_flip):
        """Body of the `tf.numpy_function` inside `read_and_encode`: path (bytes / str / 0-d array), flip flag
        (bool / np.bool_) -> (labels (A, num_classes), offsets (A, 4)) float32.  Re-entrant: `tf.data` calls it from
        several threads (03-...ipynb cell 8, num_parallel_calls=AUTOTUNE)."""
        path = path_file_labels_boxes
        if isinstance(path, np.ndarray):
            path = path.item()
        if isinstance(path, bytes):
            path = path.decode()
        lab, off = self._encode_ground_truth_labels_boxes(str(path), bool(augment_with_horizontal_flip))
        return np.ascontiguousarray(lab, np.float32), np.ascontiguousarray(off, np.float32)

    def read_and_encode(self, path_file_image: str, path_file_mask: str, path_file_labels_boxes: str):
        """(image, {'output-mask', 'output-labels', 'output-boxes'}) as datacoder.py:302-347.

        Usable exactly where the reference uses it - as the function of `tf.data.Dataset.map` (graph mode, symbolic
        string tensors) or eagerly with Python strings.  PNG decoding, the one-hot mask and the image flip
        (datacoder.py:325-342) are TensorFlow image ops and stay TensorFlow ops; the flip coin is the reference's
        `tf.random.uniform(...) >= 0.5` (:337); the box side (:344, datacoder.py:177-300) runs on the GPU through a
        `tf.numpy_function` with the static shapes (A, num_classes) / (A, 4) restored.  Without TensorFlow this method
        raises: use `read_labels_boxes(_batch)` + `encode_batch` for the box targets."""
        try:
            import tensorflow as tf  # noqa: F401
        except ImportError as e:
            raise RuntimeError("read_and_encode needs TensorFlow for the PNG/mask part (datacoder.py:325-342); "
                               "use read_labels_boxes + encode_batch for the box targets") from e
        image = tf.cast(tf.image.decode_png(tf.io.read_file(path_file_image), channels=3), tf.float32)
        mask = tf.image.decode_png(tf.io.read_file(path_file_mask), channels=1)
        mask = tf.squeeze(tf.one_hot(mask, depth=self.num_classes, dtype=tf.float32), axis=2)
        if self.augmentation_horizontal_flip:
            flip = tf.random.uniform(shape=[], minval=0, maxval=1) >= 0.5
            image = tf.cond(flip, lambda: tf.image.flip_left_right(image), lambda: image)
            mask = tf.cond(flip, lambda: tf.image.flip_left_right(mask), lambda: mask)
        else:
            flip = tf.constant(False)
        labels, boxes = tf.numpy_function(self._encode_np, [path_file_labels_boxes, flip], [tf.float32, tf.float32],
                                          name="ssdx_encode")
        labels.set_shape((self.num_anchors, self.num_classes))
        boxes.set_shape((self.num_anchors, 4))
        return image, {'output-mask': mask, 'output-labels': labels, 'output-boxes': boxes}

    # ------------------------------------------------------------------ decode
    def _decode(self, offsets_centroids, mode: int):
        off_d, kind = _ssdx.to_device(offsets_centroids, np.float32, self.device, "offsets")
        if off_d.shape[-1] != 4 or off_d.shape[-2] != self.num_anchors:
            raise ValueError(f"offsets must have shape (..., {self.num_anchors}, 4)")
        return _ssdx.from_device(_ssdx.decode(self._anchors, off_d, self._std, mode), kind)

    def decode_to_centroids(self, offsets_centroids, output_decoded_centroids_separately: bool = False):
        """Encoded ground truth -> (center_x, center_y, width, height); rows whose offsets are all zero
        (background) decode to zero (datacoder.py:349-388).  Not for network predictions."""
        out = self._decode(offsets_centroids, _ssdx.DECODE_CENTROIDS_MASKED)
        if output_decoded_centroids_separately:
            return tuple(out[..., i] for i in range(4))
        return out

    def decode_to_corners(self, offsets_centroids, output_decoded_corners_separately: bool = False):
        """Encoded ground truth -> (xmin, ymin, xmax, ymax); background rows are zero (datacoder.py:390-432)."""
        out = self._decode(offsets_centroids, _ssdx.DECODE_CORNERS_MASKED)
        if output_decoded_corners_separately:
            return tuple(out[..., i] for i in range(4))
        return out
