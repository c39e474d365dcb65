"""Detection side of the reference's `ssdseglib/losses.py`: the two SSD losses that consume the encoded targets.

`localization_loss(y_true, y_pred)` and `confidence_loss(y_true, y_pred)` keep the reference's names, argument order
and result (one value per batch item, losses.py:5-50 and :52-172).  The reference relies on TensorFlow's autodiff for
training; here the gradient with respect to the predictions comes from the same pass when asked for
(`with_gradient=True`), so a training loop outside TensorFlow can use both.  The segmentation losses of the reference
(dice, cross-entropy, focal) are not on the SSD box path and are untouched.

Inputs: NumPy arrays (host; staged through pinned memory) or anything with `__dlpack__` (torch CUDA tensors stay on
the device, zero copy); the result mirrors the input kind.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _ssdx

KERAS_EPSILON = 1e-7  # tf.keras.backend.epsilon() default, losses.py:89


def _pair(y_true, y_pred, last: Optional[int], device: Optional[int], tag: str):
    yt, kind = _ssdx.to_device(y_true, np.float32, device, tag + "_true")
    yp, _ = _ssdx.to_device(y_pred, np.float32, yt.device.index, tag + "_pred")
    if yt.dim() != 3 or tuple(yt.shape) != tuple(yp.shape) or (last is not None and int(yt.shape[2]) != last):
        want = "(batch, total boxes, 4)" if last == 4 else "(batch, total boxes, classes)"
        raise ValueError(f"expected y_true and y_pred of the same shape {want}")
    return yt.contiguous(), yp.contiguous(), kind


def localization_loss(y_true, y_pred, with_gradient: bool = False, device: Optional[int] = None):
    """Smooth-L1 localization loss (losses.py:5-50): computed only on the boxes whose target is not background
    (encoded offsets of background boxes are all zero), averaged by their number (1 if there is none); output
    shape (batch,).  with_gradient=True also returns d loss[b] / d y_pred[b], shape (batch, total boxes, 4)."""
    yt, yp, kind = _pair(y_true, y_pred, 4, device, "loc")
    res = _ssdx.localization_loss(yt, yp, with_gradient)
    if with_gradient:
        return _ssdx.from_device(res[0], kind), _ssdx.from_device(res[1], kind)
    return _ssdx.from_device(res, kind)


def confidence_loss(y_true, y_pred, with_gradient: bool = False, return_kept_background: bool = False,
                    device: Optional[int] = None):
    """Soft-max confidence loss with hard-negative mining (losses.py:52-172): `y_true` one-hot targets whose column 0
    is the background indicator, `y_pred` probabilities.  Over the whole batch the `min(3 * positives, negatives)`
    background boxes with the highest loss are kept (ratio 3:1, as the SSD paper); each item's loss is divided by its
    own number of non-background boxes (1 if there is none); output shape (batch,).
    with_gradient=True also returns d loss[b] / d y_pred[b]; return_kept_background=True the (batch, total boxes)
    uint8 mask of the background boxes that entered the loss."""
    yt, yp, kind = _pair(y_true, y_pred, None, device, "conf")
    res = _ssdx.confidence_loss(yt, yp, KERAS_EPSILON, with_gradient, return_kept_background)
    if isinstance(res, tuple):
        return tuple(_ssdx.from_device(r, kind) for r in res)
    return _ssdx.from_device(res, kind)
