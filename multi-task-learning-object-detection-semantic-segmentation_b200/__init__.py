"""B200-native SSD target/prediction path (encode / decode / NMS / detection eval).

The directory name is not a Python identifier; load it with
`importlib.import_module("multi-task-learning-object-detection-semantic-segmentation_b200")` or put this
directory on `sys.path` and `import ssdseglib` (the drop-in name of the reference package).
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)

import ssdseglib  # noqa: E402,F401
