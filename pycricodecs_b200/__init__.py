"""B200-native batch engine for the CRI ADX / HCA codec hot paths of PyCriCodecs.

    from pycricodecs_b200 import ADX, AWB, HCA, CriHcaQuality
"""
from .adx import ADX
from .awb import AWB
from .chunk import CriHcaQuality, HCAType
from .hca import HCA

__all__ = ["ADX", "AWB", "HCA", "CriHcaQuality", "HCAType"]
