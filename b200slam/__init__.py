"""Import name for the product package.  The source lives in `opencv-simpleslam_b200/`
(the layout the build contract names); a hyphen is not importable, so this stub points
the `b200slam` package path at that directory.

`from b200slam import ALIKED, LightGlue, rbd` mirrors `from lightglue import ...` at
/root/reference/slam/core/features_utils.py:8-9 (resolved lazily: importing them loads
libb200slam.so and fails loudly if it has not been built)."""
import os as _os

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
__path__ = [_os.path.join(_root, "opencv-simpleslam_b200")]
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("ALIKED", "LightGlue", "rbd"):
        from . import frontend
        return getattr(frontend, name)
    raise AttributeError(f"module 'b200slam' has no attribute {name!r}")
