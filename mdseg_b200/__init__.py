"""Importable alias of the ``mul-datasets-semantic-segmentation_b200/`` package directory.

The product directory is named after the repository and contains a hyphen, which
Python's ``import`` statement cannot spell.  This package has no code of its own:
its search path IS that directory, so ``import mdseg_b200.ops`` loads
``mul-datasets-semantic-segmentation_b200/ops.py``.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_ROOT, "mul-datasets-semantic-segmentation_b200")
__path__ = [PACKAGE_DIR]
__version__ = "0.1.0"
