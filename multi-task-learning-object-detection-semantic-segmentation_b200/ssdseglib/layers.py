"""Prediction post-processing - mirror of the reference's `ssdseglib/layers.py` over libssdx.

`DecodeBoxesCentroidsOffsets`, `NonMaximumSuppression` and `SegmentationSuppression` keep the reference's constructor
arguments, `call` signature, output layout and `get_config` (layers.py:4-212).  TensorFlow/Keras is not a dependency
of this package:

  * without TensorFlow the three names are plain callables: `call` accepts NumPy arrays, torch tensors or DLPack
    objects and mirrors the input kind on return;
  * when `tensorflow` is importable they are `tf.keras.layers.Layer` subclasses registered under the reference's
    names (`keras_layers.py`), usable on symbolic tensors inside `tf.keras.Model` exactly where `models.py:378-417`
    puts them; non-TensorFlow inputs still go to the plain implementation, which stays available as
    `PlainDecodeBoxesCentroidsOffsets` / `PlainNonMaximumSuppression` / `PlainSegmentationSuppression`.

`Split` (a `tf.split` wrapper inside the Keras model graph) is not part of this path.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _ssdx
from .boxes import coordinates_centroids_to_corners


def _vec(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    elif hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32).reshape(-1)


class _Layer:
    """Minimal stand-in for the `tf.keras.layers.Layer` call protocol."""

    def __init__(self, name: Optional[str] = None, **kwargs) -> None:
        self.name = name or type(self).__name__

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class DecodeBoxesCentroidsOffsets(_Layer):
    def __init__(self, center_x_boxes_default, center_y_boxes_default, width_boxes_default, height_boxes_default,
                 standard_deviation_center_x_offsets: float, standard_deviation_center_y_offsets: float,
                 standard_deviation_width_offsets: float, standard_deviation_height_offsets: float,
                 device: Optional[int] = None, **kwargs) -> None:
        """Decode the centroid offsets predicted by the network (layers.py:6-43).  The four default-box
        arrays have shape (total boxes,)."""
        super().__init__(**kwargs)
        self.center_x_boxes_default = _vec(center_x_boxes_default)
        self.center_y_boxes_default = _vec(center_y_boxes_default)
        self.width_boxes_default = _vec(width_boxes_default)
        self.height_boxes_default = _vec(height_boxes_default)
        self.standard_deviation_center_x_offsets = np.float32(standard_deviation_center_x_offsets)
        self.standard_deviation_center_y_offsets = np.float32(standard_deviation_center_y_offsets)
        self.standard_deviation_width_offsets = np.float32(standard_deviation_width_offsets)
        self.standard_deviation_height_offsets = np.float32(standard_deviation_height_offsets)
        cent = np.stack([self.center_x_boxes_default, self.center_y_boxes_default, self.width_boxes_default,
                         self.height_boxes_default], axis=1)
        corners = np.stack(coordinates_centroids_to_corners(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3]), axis=1)
        self._std = (float(self.standard_deviation_center_x_offsets), float(self.standard_deviation_center_y_offsets),
                     float(self.standard_deviation_width_offsets), float(self.standard_deviation_height_offsets))
        self._anchors = _ssdx.Anchors(corners, cent, device)

    def call(self, boxes_centroids_offsets):
        """(batch, total boxes, 4) offsets (cx, cy, w, h) -> (batch, total boxes, 4) corners in the order
        (ymin, xmin, ymax, xmax) that the NMS layer expects; no background masking (layers.py:45-81)."""
        off, kind = _ssdx.to_device(boxes_centroids_offsets, np.float32, self._anchors.device, "pred_offsets")
        if off.dim() != 3 or off.shape[1] != self._anchors.num_anchors or off.shape[2] != 4:
            raise ValueError(f"expected offsets of shape (batch, {self._anchors.num_anchors}, 4), got {tuple(off.shape)}")
        return _ssdx.from_device(_ssdx.decode(self._anchors, off, self._std, _ssdx.DECODE_YXYX), kind)

    def get_config(self):
        return {
            'center_x_boxes_default': self.center_x_boxes_default,
            'center_y_boxes_default': self.center_y_boxes_default,
            'width_boxes_default': self.width_boxes_default,
            'height_boxes_default': self.height_boxes_default,
            'standard_deviation_center_x_offsets': self.standard_deviation_center_x_offsets,
            'standard_deviation_center_y_offsets': self.standard_deviation_center_y_offsets,
            'standard_deviation_width_offsets': self.standard_deviation_width_offsets,
            'standard_deviation_height_offsets': self.standard_deviation_height_offsets,
        }


class NonMaximumSuppression(_Layer):
    def __init__(self, max_number_of_boxes_per_class: int, max_number_of_boxes_per_sample: int,
                 boxes_iou_threshold: float, labels_probability_threshold: float, suppress_background_boxes: bool,
                 **kwargs) -> None:
        """Per-class non maximum suppression of the detection outputs (layers.py:97-125)."""
        super().__init__(**kwargs)
        self.max_number_of_boxes_per_class = max_number_of_boxes_per_class
        self.max_number_of_boxes_per_sample = max_number_of_boxes_per_sample
        self.boxes_iou_threshold = boxes_iou_threshold
        self.labels_probability_threshold = labels_probability_threshold
        self.suppress_background_boxes = suppress_background_boxes

    def call(self, boxes_corners_coordinates, labels_probabilities, return_aux: bool = False, class_scale=None):
        """boxes (batch, total boxes, 4) as (ymin, xmin, ymax, xmax), probabilities (batch, total boxes, classes)
        -> (batch, max boxes per sample, 6) rows [label, probability, xmin, ymin, xmax, ymax], zero padded.
        With `suppress_background_boxes` the rows with label > 0 are returned flattened over the batch, as
        the reference does (layers.py:127-168).  `return_aux=True` also returns (valid (batch,), source anchor
        index (batch, max boxes per sample)).  `class_scale` (classes,): per-class factor applied to the
        probabilities as they are read, e.g. `SegmentationSuppression.class_presence(mask)`: same result as calling
        the layer on `SegmentationSuppression()(mask, probabilities)` without materialising that tensor."""
        boxes, kind = _ssdx.to_device(boxes_corners_coordinates, np.float32, None, "nms_boxes")
        scores, _ = _ssdx.to_device(labels_probabilities, np.float32, boxes.device.index, "nms_scores")
        if boxes.dim() != 3 or boxes.shape[2] != 4 or scores.dim() != 3 or scores.shape[:2] != boxes.shape[:2]:
            raise ValueError("expected boxes (batch, boxes, 4) and probabilities (batch, boxes, classes)")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, boxes.device.index, "nms_scale")
            if tuple(class_scale.shape) != (int(scores.shape[2]),):
                raise ValueError("class_scale must have one entry per class")
        out, valid, aidx = _ssdx.nms(boxes, scores, self.max_number_of_boxes_per_class,
                                     self.max_number_of_boxes_per_sample, self.boxes_iou_threshold,
                                     self.labels_probability_threshold, want_aux=return_aux,
                                     class_scale=class_scale)
        if self.suppress_background_boxes:
            out = out[out[..., 0] > 0.0]  # tf.boolean_mask over (batch, boxes): drops the batch axis
        res = _ssdx.from_device(out, kind)
        if return_aux:
            return res, _ssdx.from_device(valid, kind), _ssdx.from_device(aidx, kind)
        return res

    def call_on_logits(self, boxes_corners_coordinates, labels_logits, return_aux: bool = False, class_scale=None):
        """`self.call(boxes, softmax(logits))` as one operation (new entry point): the class head of the reference
        ends in `tf.keras.layers.Softmax` (models.py:259, :694) whose only consumers at inference are SegmentationSuppression and this layer
        (models.py:402-414).  When `max_number_of_boxes_per_class >= max_number_of_boxes_per_sample` and there are 4
        classes, the probabilities are computed inside the score filter as the logits are read and never written to
        memory; otherwise they are materialised in a workspace first.  Same result as `call` on
        `softmax_probabilities(logits)`, bit for bit."""
        boxes, kind = _ssdx.to_device(boxes_corners_coordinates, np.float32, None, "nms_boxes")
        logits, _ = _ssdx.to_device(labels_logits, np.float32, boxes.device.index, "nms_logits")
        if boxes.dim() != 3 or boxes.shape[2] != 4 or logits.dim() != 3 or logits.shape[:2] != boxes.shape[:2]:
            raise ValueError("expected boxes (batch, boxes, 4) and logits (batch, boxes, classes)")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, boxes.device.index, "nms_scale")
            if tuple(class_scale.shape) != (int(logits.shape[2]),):
                raise ValueError("class_scale must have one entry per class")
        out, valid, aidx = _ssdx.nms_logits(boxes, logits, self.max_number_of_boxes_per_class,
                                            self.max_number_of_boxes_per_sample, self.boxes_iou_threshold,
                                            self.labels_probability_threshold, want_aux=return_aux,
                                            class_scale=class_scale)
        if self.suppress_background_boxes:
            out = out[out[..., 0] > 0.0]
        res = _ssdx.from_device(out, kind)
        if return_aux:
            return res, _ssdx.from_device(valid, kind), _ssdx.from_device(aidx, kind)
        return res

    @staticmethod
    def softmax_probabilities(labels_logits, class_scale=None):
        """The soft-max activation of the class head with the arithmetic `call_on_logits` uses (exp(x - max) / sum in
        fp32), materialised: (batch, boxes, classes) -> same shape."""
        logits, kind = _ssdx.to_device(labels_logits, np.float32, None, "softmax_logits")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, logits.device.index, "softmax_scale")
        return _ssdx.from_device(_ssdx.softmax_scores(logits, class_scale), kind)

    def call_on_offsets_and_logits(self, decode_layer, boxes_centroids_offsets, labels_logits, return_aux: bool = False,
                                   class_scale=None):
        """The whole inference tail as one operation (new entry point): `call_on_offsets` with the class head's raw
        logits instead of probabilities, i.e. `self.call(decode_layer.call(offsets), softmax(logits))` without writing
        the decoded boxes or the probabilities.  Same result as that chain with `softmax_probabilities`, bit for bit."""
        off, kind = _ssdx.to_device(boxes_centroids_offsets, np.float32, decode_layer._anchors.device, "nms_off")
        logits, _ = _ssdx.to_device(labels_logits, np.float32, off.device.index, "nms_logits")
        if off.dim() != 3 or off.shape[2] != 4 or logits.dim() != 3 or logits.shape[:2] != off.shape[:2]:
            raise ValueError("expected offsets (batch, boxes, 4) and logits (batch, boxes, classes)")
        if int(off.shape[1]) != decode_layer._anchors.num_anchors:
            raise ValueError("offsets do not match the number of default boxes of the decode layer")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, off.device.index, "nms_scale")
        out, valid, aidx = _ssdx.decode_nms_logits(decode_layer._anchors, off, decode_layer._std, logits,
                                                   self.max_number_of_boxes_per_class,
                                                   self.max_number_of_boxes_per_sample, self.boxes_iou_threshold,
                                                   self.labels_probability_threshold, want_aux=return_aux,
                                                   class_scale=class_scale)
        if self.suppress_background_boxes:
            out = out[out[..., 0] > 0.0]
        res = _ssdx.from_device(out, kind)
        if return_aux:
            return res, _ssdx.from_device(valid, kind), _ssdx.from_device(aidx, kind)
        return res

    def call_on_offsets(self, decode_layer, boxes_centroids_offsets, labels_probabilities, return_aux: bool = False,
                        class_scale=None):
        """`self.call(decode_layer.call(offsets), probabilities)` as one operation (new entry point): the wiring of
        `get_model_for_inference` (models.py:410-414), whose only output is the detections.  The boxes the NMS looks
        at are decoded on the fly; the (batch, boxes, 4) tensor is never written.  Bit-identical results."""
        off, kind = _ssdx.to_device(boxes_centroids_offsets, np.float32, decode_layer._anchors.device, "nms_off")
        scores, _ = _ssdx.to_device(labels_probabilities, np.float32, off.device.index, "nms_scores")
        if off.dim() != 3 or off.shape[2] != 4 or scores.dim() != 3 or scores.shape[:2] != off.shape[:2]:
            raise ValueError("expected offsets (batch, boxes, 4) and probabilities (batch, boxes, classes)")
        if int(off.shape[1]) != decode_layer._anchors.num_anchors:
            raise ValueError("offsets do not match the number of default boxes of the decode layer")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, off.device.index, "nms_scale")
        out, valid, aidx = _ssdx.decode_nms(decode_layer._anchors, off, decode_layer._std, scores,
                                            self.max_number_of_boxes_per_class, self.max_number_of_boxes_per_sample,
                                            self.boxes_iou_threshold, self.labels_probability_threshold,
                                            want_aux=return_aux, class_scale=class_scale)
        if self.suppress_background_boxes:
            out = out[out[..., 0] > 0.0]
        res = _ssdx.from_device(out, kind)
        if return_aux:
            return res, _ssdx.from_device(valid, kind), _ssdx.from_device(aidx, kind)
        return res

    def select(self, labels_probabilities, class_scale=None):
        """Phase 1 of `call` (new entry point): everything that depends only on the probabilities, on the current
        CUDA stream.  In the reference's inference graph the probabilities and the decoded boxes reach this layer
        through independent branches (models.py:402-414), so this can run on a side stream while the boxes are
        still being decoded.  Returns an opaque selection for `finish`."""
        scores, _ = _ssdx.to_device(labels_probabilities, np.float32, None, "nms_scores")
        if scores.dim() != 3:
            raise ValueError("expected probabilities (batch, boxes, classes)")
        if class_scale is not None:
            class_scale, _ = _ssdx.to_device(class_scale, np.float32, scores.device.index, "nms_scale")
        return _ssdx.nms_select(scores, self.max_number_of_boxes_per_class, self.max_number_of_boxes_per_sample,
                                self.labels_probability_threshold, class_scale=class_scale)

    def finish(self, selection, boxes_corners_coordinates, return_aux: bool = False):
        """Phase 2 of `call`: waits (stream-ordered) for `selection`, then suppresses and packs.  Same result as
        `call(boxes, probabilities)`."""
        boxes, kind = _ssdx.to_device(boxes_corners_coordinates, np.float32, selection.scores.device.index, "nms_boxes")
        if boxes.dim() != 3 or boxes.shape[2] != 4 or selection.scores.shape[:2] != boxes.shape[:2]:
            raise ValueError("expected boxes (batch, boxes, 4) matching the selected probabilities")
        out, valid, aidx = _ssdx.nms_finish(selection, boxes, self.boxes_iou_threshold, want_aux=return_aux)
        if self.suppress_background_boxes:
            out = out[out[..., 0] > 0.0]
        res = _ssdx.from_device(out, kind)
        if return_aux:
            return res, _ssdx.from_device(valid, kind), _ssdx.from_device(aidx, kind)
        return res

    def get_config(self):
        return {
            'max_number_of_boxes_per_class': self.max_number_of_boxes_per_class,
            'max_number_of_boxes_per_sample': self.max_number_of_boxes_per_sample,
            'boxes_iou_threshold': self.boxes_iou_threshold,
            'labels_probability_threshold': self.labels_probability_threshold,
            'suppress_background_boxes': self.suppress_background_boxes,
        }


class SegmentationSuppression(_Layer):
    """Process the object detection outputs with the semantic segmentation output: a class that the segmentation
    did not detect anywhere in the batch is ignored by the object detection (layers.py:179-212)."""

    DEPTH = 4  # tf.one_hot(..., depth=4) is hard-coded in the reference (layers.py:204)

    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)

    @classmethod
    def class_presence(cls, segmentation_mask):
        """(batch, height, width, classes) probabilities -> (4,) float32 of 0/1: 1 iff some pixel of the whole batch
        has that class as its arg-max (layers.py:203-207).  Feed it to `NonMaximumSuppression.call(class_scale=...)`
        to fuse the suppression into the NMS score filter."""
        m, kind = _ssdx.to_device(segmentation_mask, np.float32, None, "seg_mask")
        if m.dim() < 2:
            raise ValueError("expected a segmentation mask of shape (batch, height, width, classes)")
        return _ssdx.from_device(_ssdx.seg_class_presence(m, cls.DEPTH), kind)

    def call(self, segmentation_mask, labels_probabilities):
        """-> labels_probabilities * is_class_segmented, same shape as labels_probabilities (layers.py:189-212)."""
        m, _ = _ssdx.to_device(segmentation_mask, np.float32, None, "seg_mask")
        p, kind = _ssdx.to_device(labels_probabilities, np.float32, m.device.index, "seg_probs")
        if int(p.shape[-1]) != self.DEPTH:
            raise ValueError("labels_probabilities must have 4 classes on its last axis (one_hot depth of the "
                             "reference, layers.py:204)")
        presence = _ssdx.seg_class_presence(m, self.DEPTH)
        return _ssdx.from_device(_ssdx.scale_scores(p, presence), kind)

    def get_config(self):
        return {}


# ---- Keras-graph form (only when TensorFlow is importable; see keras_layers.py) ----
PlainDecodeBoxesCentroidsOffsets = DecodeBoxesCentroidsOffsets
PlainNonMaximumSuppression = NonMaximumSuppression
PlainSegmentationSuppression = SegmentationSuppression
try:  # pragma: no cover - TensorFlow is absent from the build and test images
    import tensorflow as _tf
except Exception:
    _tf = None
if _tf is not None:  # pragma: no cover
    from .keras_layers import make_keras_layers

    DecodeBoxesCentroidsOffsets, NonMaximumSuppression, SegmentationSuppression = make_keras_layers(
        _tf, PlainDecodeBoxesCentroidsOffsets, PlainNonMaximumSuppression, PlainSegmentationSuppression)
