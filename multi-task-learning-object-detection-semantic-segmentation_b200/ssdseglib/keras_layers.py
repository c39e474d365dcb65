"""Keras-graph form of the post-processing layers (reference: ssdseglib/layers.py:4, :95, :179 and the wiring of
`get_model_for_inference`, models.py:345-423).

The reference registers its three layers as serialisable `tf.keras.layers.Layer` subclasses and calls them on
symbolic tensors inside `tf.keras.Model`.  TensorFlow is not a dependency of this package; when it IS importable,
`ssdseglib.layers` replaces its plain callables by the classes built here, so that `models.py` builds, saves and
reloads an inference model unchanged:

  * same class names, constructor arguments and `get_config` keys as the reference, registered with
    `tf.keras.saving.register_keras_serializable(name=...)` under the reference's names;
  * `call` on TensorFlow tensors hands the arrays to the C-ABI path through `tf.numpy_function` (host arrays in,
    host arrays out; libssdx stages them through pinned memory) and restores the static output shape the reference's
    ops produce, so downstream Keras layers see `(batch, boxes, 4)` / `(batch, max boxes, 6)`;
  * any other input (NumPy, torch, DLPack) goes straight to the plain implementation and keeps the
    return-kind mirroring and the extra entry points (`call_on_offsets`, `call_on_logits`, ...).

The plain implementation is created lazily on the first call: constructing the layer (as `load_model` does when it
deserialises a graph) needs no CUDA device.
"""
from __future__ import annotations

import numpy as np


def make_keras_layers(tf, plain_decode, plain_nms, plain_seg):
    """Returns (DecodeBoxesCentroidsOffsets, NonMaximumSuppression, SegmentationSuppression) as Keras layers
    wrapping the given plain-callable implementations."""
    register = tf.keras.saving.register_keras_serializable

    def _is_tf(x) -> bool:
        return tf.is_tensor(x) or type(x).__module__.startswith(("tensorflow", "keras"))

    def _np32(x):
        return np.ascontiguousarray(x, dtype=np.float32)

    class _Delegating(tf.keras.layers.Layer):
        """Keras layer whose arithmetic lives in a plain libssdx-backed object (`self._plain()`)."""

        _impl = None

        def _make_impl(self):
            raise NotImplementedError

        def _plain(self):
            if self._impl is None:
                object.__setattr__(self, "_impl", self._make_impl())
            return self._impl

        def __call__(self, *args, **kwargs):
            first = args[0] if args else next(iter(kwargs.values()), None)
            if _is_tf(first):
                return super().__call__(*args, **kwargs)
            return self._plain().call(*args, **kwargs)  # NumPy / torch / DLPack: no graph, no conversion

        def __getattr__(self, name):  # call_on_offsets, call_on_logits, select, finish, class_presence, ...
            if name.startswith("call_on") or name in ("select", "finish", "softmax_probabilities", "class_presence"):
                return getattr(self._plain(), name)
            raise AttributeError(name)

    @register(name="DecodeBoxesCentroidsOffsets")
    class DecodeBoxesCentroidsOffsets(_Delegating):
        def __init__(self, center_x_boxes_default, center_y_boxes_default, width_boxes_default, height_boxes_default,
                     standard_deviation_center_x_offsets, standard_deviation_center_y_offsets,
                     standard_deviation_width_offsets, standard_deviation_height_offsets, **kwargs):
            super().__init__(**kwargs)
            # kept as the reference keeps them (layers.py:33-43): get_config returns these very objects
            self.center_x_boxes_default = center_x_boxes_default
            self.center_y_boxes_default = center_y_boxes_default
            self.width_boxes_default = width_boxes_default
            self.height_boxes_default = height_boxes_default
            self.standard_deviation_center_x_offsets = standard_deviation_center_x_offsets
            self.standard_deviation_center_y_offsets = standard_deviation_center_y_offsets
            self.standard_deviation_width_offsets = standard_deviation_width_offsets
            self.standard_deviation_height_offsets = standard_deviation_height_offsets

        def _make_impl(self):
            return plain_decode(
                np.asarray(self.center_x_boxes_default), np.asarray(self.center_y_boxes_default),
                np.asarray(self.width_boxes_default), np.asarray(self.height_boxes_default),
                float(self.standard_deviation_center_x_offsets), float(self.standard_deviation_center_y_offsets),
                float(self.standard_deviation_width_offsets), float(self.standard_deviation_height_offsets))

        def call(self, boxes_centroids_offsets):
            impl = self._plain()
            out = tf.numpy_function(lambda o: _np32(impl.call(_np32(o))), [boxes_centroids_offsets], tf.float32,
                                    name="ssdx_decode")
            out.set_shape(boxes_centroids_offsets.shape)  # (batch, total boxes, 4), layers.py:79
            return out

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

    @register(name="NonMaximumSuppression")
    class NonMaximumSuppression(_Delegating):
        def __init__(self, max_number_of_boxes_per_class, max_number_of_boxes_per_sample, boxes_iou_threshold,
                     labels_probability_threshold, suppress_background_boxes, **kwargs):
            super().__init__(**kwargs)
            self.max_number_of_boxes_per_class = max_number_of_boxes_per_class
            self.max_number_of_boxes_per_sample = max_number_of_boxes_per_sample
            self.boxes_iou_threshold = boxes_iou_threshold
            self.labels_probability_threshold = labels_probability_threshold
            self.suppress_background_boxes = suppress_background_boxes

        def _make_impl(self):
            return plain_nms(self.max_number_of_boxes_per_class, self.max_number_of_boxes_per_sample,
                             self.boxes_iou_threshold, self.labels_probability_threshold,
                             self.suppress_background_boxes)

        def call(self, boxes_corners_coordinates, labels_probabilities):
            impl = self._plain()
            out = tf.numpy_function(lambda b, p: _np32(impl.call(_np32(b), _np32(p))),
                                    [boxes_corners_coordinates, labels_probabilities], tf.float32, name="ssdx_nms")
            if self.suppress_background_boxes:
                out.set_shape((None, 6))  # tf.boolean_mask drops the batch axis (layers.py:165-166)
            else:
                out.set_shape((boxes_corners_coordinates.shape[0], self.max_number_of_boxes_per_sample, 6))
            return out

        def get_config(self):
            return {
                'max_number_of_boxes_per_class': self.max_number_of_boxes_per_class,
                'max_number_of_boxes_per_sample': self.max_number_of_boxes_per_sample,
                'boxes_iou_threshold': self.boxes_iou_threshold,
                'labels_probability_threshold': self.labels_probability_threshold,
                'suppress_background_boxes': self.suppress_background_boxes
            }

    @register(name="SegmentationSuppression")
    class SegmentationSuppression(_Delegating):
        def __init__(self, **kwargs):
            super().__init__(**kwargs)

        def _make_impl(self):
            return plain_seg()

        def call(self, segmentation_mask, labels_probabilities):
            impl = self._plain()
            out = tf.numpy_function(lambda m, p: _np32(impl.call(_np32(m), _np32(p))),
                                    [segmentation_mask, labels_probabilities], tf.float32, name="ssdx_seg_suppression")
            out.set_shape(labels_probabilities.shape)
            return out

        def get_config(self):
            return {}

    return DecodeBoxesCentroidsOffsets, NonMaximumSuppression, SegmentationSuppression
