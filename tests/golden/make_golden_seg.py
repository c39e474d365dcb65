"""Golden vectors for layers.SegmentationSuppression: the reference's own `ssdseglib/layers.py` executed UNCHANGED over
the NumPy emulation of the TensorFlow primitives (tf_numpy_shim.py, same assumptions as make_golden.py).

    python tests/golden/make_golden_seg.py       (build container only: needs /root/reference)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
spec = importlib.util.spec_from_file_location("ref_layers", os.path.join(REF, "ssdseglib", "layers.py"))
ref_layers = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_layers)

rng = np.random.default_rng(2001)
layer = ref_layers.SegmentationSuppression()
out = {}
# case a: every class present; case b: class 2 never wins; case c: only the background wins; case d: exact ties
for name, kill in (("a", []), ("b", [2]), ("c", [1, 2, 3]), ("d", [])):
    mask = rng.uniform(0, 1, size=(2, 24, 32, 4)).astype(np.float32)
    for k in kill:
        mask[..., k] = 0.0
    if name == "d":
        mask[...] = 0.25          # all classes tie everywhere -> arg-max is class 0 only
        mask[1, 3, 4, 3] = 0.5    # ... except one pixel for class 3
    probs = rng.uniform(0, 1, size=(2, 300, 4)).astype(np.float32)
    out[f"{name}_mask"] = mask
    out[f"{name}_probs"] = probs
    out[f"{name}_suppressed"] = np.asarray(layer.call(mask, probs), dtype=np.float32)
np.savez_compressed(os.path.join(HERE, "seg.npz"), **out)
print("seg.npz written:", {k: v.shape for k, v in out.items() if k.endswith("suppressed")})
