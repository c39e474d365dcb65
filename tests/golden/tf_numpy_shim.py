"""NumPy emulation of the TensorFlow primitives that the reference's
`ssdseglib/datacoder.py` and `ssdseglib/layers.py` call.

Purpose: TensorFlow 2.13 is not installable in the build container, so the
reference's own source cannot run as-is.  `make_golden.py` installs this module
as `tensorflow` in `sys.modules` and then executes the reference's files
UNCHANGED, so the committed fixtures follow the reference's literal statement
sequence.  What is emulated (and therefore an assumption, listed in DESIGN.md):

  * every elementwise op is IEEE fp32, one rounding per op (Eigen CPU kernels
    are separate ops, no cross-op FMA contraction);
  * argmax returns the lowest index on ties;
  * UniqueV2(axis=0) returns rows in first-occurrence order;
  * tensor_scatter_nd_update applies updates sequentially (CPU kernel), so the
    last duplicate wins;
  * one_hot of an out-of-range index is an all-zero row;
  * tf.image.combined_non_max_suppression is NOT available here: it is routed
    to the oracle's restatement (oracle/ssd_oracle.py), so fixtures that pass
    through it only pin the layer's wrapper logic, not the NMS core.

This is test tooling; nothing in the product imports it.
"""
from __future__ import annotations

import sys
import types

import numpy as np

F32 = np.float32


def _ns(**kw):
    m = types.SimpleNamespace()
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def _to_f32(x, dtype=None):
    if dtype is None:
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(F32)
        return a
    return np.asarray(x).astype(dtype)


def _argmax(x, axis=None, output_type=np.int64):
    return np.argmax(x, axis=axis).astype(output_type)


def _boolean_mask(tensor, mask, axis=None):
    tensor = np.asarray(tensor)
    mask = np.asarray(mask).astype(bool)
    if mask.ndim == 1:
        return tensor[mask]
    # mask spans the leading mask.ndim dims (flattens them)
    lead = int(np.prod(mask.shape))
    return tensor.reshape((lead,) + tensor.shape[mask.ndim:])[mask.reshape(-1)]


def _unique_v2(x, axis):
    x = np.asarray(x)
    assert list(axis) == [0]
    seen, rows, idx = {}, [], []
    for r in x:
        key = tuple(int(v) for v in np.atleast_1d(r))
        if key not in seen:
            seen[key] = len(rows)
            rows.append(r)
        idx.append(seen[key])
    y = np.array(rows, dtype=x.dtype).reshape((-1,) + x.shape[1:])
    return y, np.array(idx, dtype=np.int32)


def _one_hot(indices, depth, dtype=F32, axis=-1):
    indices = np.asarray(indices)
    out = (indices[..., None] == np.arange(depth)).astype(dtype)
    if axis not in (-1, out.ndim - 1):
        out = np.moveaxis(out, -1, axis)
    return out


def _tensor_scatter_nd_update(tensor, indices, updates):
    out = np.array(tensor, copy=True)
    indices = np.asarray(indices)
    for i in range(indices.shape[0]):  # sequential, last writer wins
        out[tuple(indices[i])] = updates[i]
    return out


def _split(value, num_or_size_splits, axis=0):
    return np.split(np.asarray(value), num_or_size_splits, axis=axis)


def _read_file(path):
    if isinstance(path, bytes):
        path = path.decode()
    with open(path, "rb") as f:
        return f.read().decode()


def _strings_split(s, sep):
    return s.split(sep)


def _decode_csv(records, record_defaults):
    cols = [[] for _ in record_defaults]
    for rec in records:
        fields = rec.split(",")
        assert len(fields) == len(record_defaults), f"bad CSV record {rec!r}"
        for c, (f, d) in enumerate(zip(fields, record_defaults)):
            cols[c].append(type(d)(f))
    return [np.array(c, dtype=np.int32 if isinstance(d, int) else F32) for c, d in zip(cols, record_defaults)]


def _combined_nms(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold, score_threshold,
                  clip_boxes=False):
    from oracle import ssd_oracle

    assert boxes.shape[2] == 1 and not clip_boxes
    out, valid, _ = ssd_oracle.combined_nms(np.asarray(boxes)[:, :, 0, :], np.asarray(scores),
                                            max_output_size_per_class, max_total_size, iou_threshold,
                                            score_threshold)
    nmsed_boxes = out[:, :, [3, 2, 5, 4]]  # back to TF's (ymin, xmin, ymax, xmax)
    return nmsed_boxes, out[:, :, 1], out[:, :, 0], valid


class _Layer:
    def __init__(self, **kwargs):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


def make_module() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.float32 = F32
    tf.int32 = np.int32
    tf.Tensor = np.ndarray
    tf.dtypes = _ns(int32=np.int32, float32=F32)
    tf.convert_to_tensor = lambda x, dtype=None: _to_f32(x, dtype)
    tf.constant = lambda x, dtype=None: _to_f32(x, dtype)
    tf.expand_dims = lambda input, axis: np.expand_dims(np.asarray(input), axis)
    tf.transpose = lambda x: np.transpose(np.asarray(x))
    tf.stack = lambda values, axis=0: np.stack([np.asarray(v) for v in values], axis=axis)
    tf.concat = lambda values, axis: np.concatenate([np.asarray(v) for v in values], axis=axis)
    tf.range = lambda n: np.arange(n, dtype=np.int32)
    tf.boolean_mask = _boolean_mask
    tf.gather = lambda params, indices, axis=0: np.take(np.asarray(params), np.asarray(indices), axis=axis)
    tf.one_hot = _one_hot
    tf.ones = lambda shape, dtype=F32: np.ones(shape, dtype)
    tf.zeros = lambda shape, dtype=F32: np.zeros(shape, dtype)
    tf.tensor_scatter_nd_update = _tensor_scatter_nd_update
    tf.split = lambda value, num_or_size_splits, axis=0: _split(value, num_or_size_splits, axis)
    tf.squeeze = lambda input, axis=None: np.squeeze(np.asarray(input), axis=axis)
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.clip_by_value = lambda t, lo, hi: np.clip(t, lo, hi)
    tf.math = _ns(
        maximum=lambda a, b: np.maximum(a, b),
        minimum=lambda a, b: np.minimum(a, b),
        argmax=_argmax,
        greater=lambda a, b: np.greater(a, b),
        reduce_max=lambda x, axis=None: np.max(x, axis=axis),
        reduce_sum=lambda x, axis=None: np.sum(x, axis=axis, dtype=np.asarray(x).dtype),
        log=lambda x: np.log(np.asarray(x, dtype=F32)),
        exp=lambda x: np.exp(np.asarray(x, dtype=F32)),
        abs=lambda x: np.abs(x),
        multiply=lambda a, b: np.multiply(a, b),
    )
    tf.raw_ops = _ns(UniqueV2=_unique_v2)
    tf.strings = _ns(strip=lambda s: s.strip(), split=_strings_split)
    tf.io = _ns(read_file=_read_file, decode_csv=_decode_csv)
    tf.image = _ns(combined_non_max_suppression=_combined_nms)
    tf.random = _ns(uniform=lambda shape, minval=0, maxval=1: np.random.uniform(minval, maxval))
    tf.keras = _ns(
        layers=_ns(Layer=_Layer),
        saving=_ns(register_keras_serializable=lambda name=None, package=None: (lambda cls: cls)),
    )
    return tf


def install():
    mod = make_module()
    sys.modules["tensorflow"] = mod
    return mod
