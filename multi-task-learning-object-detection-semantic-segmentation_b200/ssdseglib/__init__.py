"""ssdseglib (B200 build): the SSD target/prediction path of the reference package behind the same names.

Sub-modules mirror the reference's: `boxes`, `datacoder`, `layers` (decode + NMS), `evaluators` (detection
side).  The model / loss / metric / plot modules of the reference are untouched and not shipped here.
Importing this package loads libssdx.so; it raises ImportError if the library has not been built.
"""
from . import _ssdx  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import boxes, datacoder, evaluators, layers, losses, metrics  # noqa: F401

__all__ = ["boxes", "datacoder", "layers", "evaluators", "metrics", "losses"]
