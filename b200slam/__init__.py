"""Import name for the product package.  The source lives in `opencv-simpleslam_b200/`
(the layout the build contract names); a hyphen is not importable, so this stub points
the `b200slam` package path at that directory."""
import os as _os

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
__path__ = [_os.path.join(_root, "opencv-simpleslam_b200")]
__version__ = "0.1.0"
