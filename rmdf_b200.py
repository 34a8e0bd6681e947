"""Import alias: the package directory is named after the reference
(`ray-marching-distance-fields_b200`), which is not a valid Python identifier."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("ray-marching-distance-fields_b200")
sys.modules[__name__] = _pkg
