"""B200-native read -> splice-graph alignment + EM for Whippet's `whippet-quant` path.

Host-side mirror of the reference interface (names follow src/Whippet.jl:50-120); the compute
lives in csrc/ behind the C ABI of include/whippet_b200.h.
"""
from . import abi, jls, flat
from .flat import FlatIndex, AlignParam, ReadBatch
from .quant import GraphLibQuant, AlignResult, MapStats, write_index

__all__ = ["abi", "jls", "flat", "FlatIndex", "AlignParam", "ReadBatch", "GraphLibQuant", "AlignResult", "MapStats", "write_index"]
