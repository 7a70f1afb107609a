"""genmap_b200 — B200-native (K,E)-mappability: the `genmap map` hot path of cpockrandt/genmap as
hand-written sm_100a CUDA behind a C ABI (include/genmap_b200.h)."""
from ._lib import GenmapError, LIB_PATH  # noqa: F401
from .api import Index, SearchParams, compute_mappability  # noqa: F401
from .synth import synth_genome, synth_pangenome  # noqa: F401

__all__ = ["Index", "SearchParams", "compute_mappability", "synth_genome", "synth_pangenome", "GenmapError", "LIB_PATH"]
