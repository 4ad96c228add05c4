"""B200-native batch engine for the CRI ADX / HCA codec hot paths of PyCriCodecs.

    from pycricodecs_b200 import ACB, ADX, AWB, HCA, CriHcaQuality
"""
from .acb import ACB
from .adx import ADX
from .awb import AWB
from .chunk import CriHcaQuality, HCAType
from .hca import HCA

__all__ = ["ACB", "ADX", "AWB", "HCA", "CriHcaQuality", "HCAType"]
