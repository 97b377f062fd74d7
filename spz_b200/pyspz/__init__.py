"""The reference's Python module surface (`import spz`) over the B200-native library.

    from spz_b200.pyspz import spz          # or: sys.path.insert(0, ".../spz_b200/pyspz"); import spz

The extension (spz.cpython-*.so, built from spz_b200/csrc/py_spz.cc by spz_b200.build) sits next to
this file.  It is built on first import if missing.
"""
import importlib
import os

from .. import build as _build

if not os.path.exists(_build.python_module_path()):
    _build.build_python_module()
spz = importlib.import_module(".spz", __name__)
