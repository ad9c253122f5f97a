"""Importable alias for the package directory ``buffer-x_b200/``.

The task fixes the package directory name (it contains a hyphen, so a plain
``import`` statement cannot name it).  This shim imports it through importlib
and registers every sub-module under the ``bufferx_b200.*`` prefix so that both
spellings refer to the SAME module objects (one CUDA library handle).
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)

_REAL = "buffer-x_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["bufferx_b200" + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
