# -*- coding: utf-8 -*-
"""mcsas_b200 -- B200-native Monte-Carlo fitting core with McSAS's plug-in surface."""
from ._lib import McSASB200Error, FP32, FP64, QUAD_REFERENCE, QUAD_GAUSS_LEGENDRE  # noqa: F401
from .engine import Engine  # noqa: F401

__version__ = "0.1.0"
