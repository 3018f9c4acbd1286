"""Importable alias for the package directory ``3d-point-cloud-multiday-imagery_b200/``.

The directory name is fixed by the project layout but is not a Python identifier, so
``import mdkm_b200`` loads it through importlib and re-exports its public names.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
package = _pkg
__all__ = list(_pkg.__all__) + ["package"]
