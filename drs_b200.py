"""Importable alias of the package directory ``information-retrieval-with-contrastive-learning_b200``
(its name has hyphens, so a plain ``import`` statement cannot spell it)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("information-retrieval-with-contrastive-learning_b200")
sys.modules[__name__] = _pkg
