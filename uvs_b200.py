"""Import alias: ``import uvs_b200`` == the package in ``uncalibrated-visual-servoing_b200/``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("uncalibrated-visual-servoing_b200")
