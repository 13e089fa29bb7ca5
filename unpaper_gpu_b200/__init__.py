"""Importable alias of the `unpaper-gpu_b200/` package directory.

The package directory carries the reference's name (with a hyphen, which
Python cannot import); this stub extends ``__path__`` so that
``import unpaper_gpu_b200.abi`` resolves into it.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                              "unpaper-gpu_b200"))

from .abi import *  # noqa: F401,F403,E402
