"""Importable alias of the package directory ``physical-interaction-video-prediction_b200`` (hyphenated names cannot
be written in an ``import`` statement)."""
import importlib
import sys

_real = importlib.import_module("physical-interaction-video-prediction_b200")
sys.modules[__name__] = _real
