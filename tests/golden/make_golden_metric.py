"""Golden vectors for metrics.jaccard_iou_bounding_boxes: the reference's own `ssdseglib/metrics.py` executed UNCHANGED
over the NumPy emulation of the TensorFlow primitives (tf_numpy_shim.py; same assumptions as make_golden.py, plus
tf.reduce_sum = np.sum in fp32 and tf.keras.backend.epsilon() = 1e-7, the Keras default).

    python tests/golden/make_golden_metric.py       (build container only: needs /root/reference)
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf_numpy_shim  # noqa: E402
from oracle import ssd_oracle as orc  # noqa: E402  (input generators only)

tf = tf_numpy_shim.install()
tf.reduce_sum = tf.math.reduce_sum
tf.keras.backend = types.SimpleNamespace(epsilon=lambda: 1e-7)
spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "ssdseglib", "metrics.py"))
ref_metrics = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_metrics)

STD = (0.1, 0.1, 0.2, 0.2)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1).astype(np.float32)
rng = np.random.default_rng(2002)
labels, boxes, counts = orc.synth_gt(rng, 6, 16)
counts = np.array([16, 3, 16, 9, 16, 1], np.int32)
_, y_true, _ = orc.encode_batch(anchors, labels, boxes, counts, 4, 0.5, STD)
# predictions: the targets plus noise on every row (the metric must ignore the background rows), a few wild rows
y_pred = (y_true + rng.normal(0, 0.6, y_true.shape)).astype(np.float32)
y_pred[0, ::50, 2:] = -40.0          # exp underflow -> width/height clamp at 0
y_pred[2] = y_true[2]                # perfect prediction -> IoU 1 on every positive
metric = ref_metrics.jaccard_iou_bounding_boxes(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], STD)
with np.errstate(all="ignore"):
    value = np.asarray(metric(y_true, y_pred), dtype=np.float32)
np.savez_compressed(os.path.join(HERE, "metric.npz"), y_true=y_true, y_pred=y_pred, value=value)
print("metric.npz written:", value)
