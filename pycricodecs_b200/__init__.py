"""B200-native batch engine for the CRI ADX / HCA codec hot paths of PyCriCodecs.

    from pycricodecs_b200 import ADX, HCA, CriHcaQuality
"""
from .adx import ADX
from .chunk import CriHcaQuality, HCAType
from .hca import HCA

__all__ = ["ADX", "HCA", "CriHcaQuality", "HCAType"]
