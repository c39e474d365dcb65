"""Golden vectors for losses.localization_loss / losses.confidence_loss: the reference's own `ssdseglib/losses.py`
executed UNCHANGED over the NumPy emulation of the TensorFlow primitives (tf_numpy_shim.py plus the few ops below:
tf.math.top_k = stable descending order (equal values: lower index first, as TensorFlow documents), tf.scatter_nd
into zeros, tf.keras.backend.epsilon() = 1e-7, the Keras default).

    python tests/golden/make_golden_losses.py       (build container only: needs /root/reference)
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

F32 = np.float32
tf = tf_numpy_shim.install()


def _top_k(input, k, sorted=True):
    x = np.asarray(input)
    idx = np.argsort(-x, kind="stable")[: int(k)].astype(np.int32)
    return x[idx], idx


def _scatter_nd(indices, updates, shape):
    out = np.zeros(tuple(int(v) for v in np.asarray(shape).reshape(-1)), dtype=np.asarray(updates).dtype)
    out[tuple(np.asarray(indices).T)] = updates
    return out


tf.reduce_sum = tf.math.reduce_sum
tf.where = lambda c, a, b: np.where(c, a, b)
tf.zeros_like = lambda x, dtype=None: np.zeros_like(x, dtype=dtype)
tf.ones_like = lambda x, dtype=None: np.ones_like(x, dtype=dtype)
tf.reshape = lambda x, shape: np.reshape(np.asarray(x), tuple(int(v) for v in np.asarray(shape).reshape(-1)))
tf.shape = lambda x: np.asarray(np.asarray(x).shape, np.int32)
tf.scatter_nd = _scatter_nd
tf.clip_by_value = lambda t, clip_value_min, clip_value_max: np.clip(np.asarray(t, F32), F32(clip_value_min),
                                                                     F32(clip_value_max))
tf.math.pow = lambda x, y: np.power(x, F32(y))
tf.math.less = lambda a, b: np.less(a, b)
tf.math.equal = lambda a, b: np.equal(a, b)
tf.math.count_nonzero = lambda x, dtype=np.int64: np.asarray(np.count_nonzero(x), dtype=dtype)
tf.math.top_k = _top_k
tf.keras.backend = types.SimpleNamespace(epsilon=lambda: 1e-7)
spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(REF, "ssdseglib", "losses.py"))
ref_losses = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_losses)

STD = (0.1, 0.1, 0.2, 0.2)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
rng = np.random.default_rng(2003)
labels, boxes, counts = orc.synth_gt(rng, 6, 16)
counts = np.array([16, 3, 0, 9, 16, 1], np.int32)          # image 2 is all background
t_lab, t_off, _ = orc.encode_batch(anchors, labels, boxes, counts, 4, 0.5, STD)
out = {}
# localization: targets + noise, some errors well above 1 (linear branch), one exact |d| == 1, one perfect image
p_off = (t_off + rng.normal(0, 0.7, t_off.shape)).astype(F32)
p_off[4] = t_off[4]
pos = np.argwhere(np.abs(t_off[0]).sum(-1) > 0)[:, 0]
p_off[0, pos[0]] = t_off[0, pos[0]] + F32(1.0)
out["loc_y_true"], out["loc_y_pred"] = t_off, p_off
out["loc_value"] = np.asarray(ref_losses.localization_loss(t_off, p_off), F32)
# confidence: soft-max of noisy logits pulled towards the target, with saturated entries (exact 0 / 1: both clip bounds)
logits = rng.normal(0, 1.5, t_lab.shape).astype(F32) + 2.0 * t_lab
p_lab = orc.softmax_f32(logits)
p_lab[1, ::40] = np.array([1, 0, 0, 0], F32)
p_lab[3, 5::60] = np.array([0, 1, 0, 0], F32)
out["conf_y_true"], out["conf_y_pred"] = t_lab, p_lab
with np.errstate(all="ignore"):
    out["conf_value"] = np.asarray(ref_losses.confidence_loss(t_lab, p_lab), F32)
    # a batch with no positive at all (k = 0) and a batch with more positives than a third of the negatives
    out["conf_value_no_positive"] = np.asarray(ref_losses.confidence_loss(t_lab[2:3], p_lab[2:3]), F32)
    few = t_lab[:2, :400].copy()
    few[:, ::3] = np.array([0, 0, 1, 0], F32)
    out["conf_few_y_true"], out["conf_few_y_pred"] = few, p_lab[:2, :400]
    out["conf_few_value"] = np.asarray(ref_losses.confidence_loss(few, p_lab[:2, :400]), F32)
np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
print({k: v for k, v in out.items() if k.endswith("value") or "value_" in k})
for k in ("loc", "conf"):
    o = (orc.localization_loss if k == "loc" else orc.confidence_loss)(out[k + "_y_true"], out[k + "_y_pred"])
    print(k, "oracle max rel diff", np.max(np.abs(o - out[k + "_value"]) / np.maximum(np.abs(out[k + "_value"]), 1e-12)))
print("few", orc.confidence_loss(few, p_lab[:2, :400]), "no positive", orc.confidence_loss(t_lab[2:3], p_lab[2:3]))
