"""medvill_b200 — B200-native (sm_100a) implementation of the MedViLL joint vision-language pre-training step.

The directory is named `multi-modality-self-supervision_b200/` after the reference repository; because that is not a
valid Python identifier the repo root ships `medvill_b200.py`, an import shim that loads this package under the
name `medvill_b200`.  Layout:
  csrc/            hand-written CUDA (tcgen05 / TMA / TMEM) + the C ABI  -> libmedvill_sm100.so
  _lib.py          ctypes binding (mirrors include/medvill_sm100.h)
  engine.py        flat arenas + step drivers on top of the C ABI
  models/ data/ utils/ main_origin.py   drop-in mirror of the reference's Python surface for this path
"""
from ._lib import MedvillError, LIB_PATH, build  # noqa: F401
from .engine import EngineDims, PretrainEngine, bucket_plan, param_map, query_layout  # noqa: F401

__all__ = ["MedvillError", "EngineDims", "PretrainEngine", "bucket_plan", "param_map", "query_layout", "build", "LIB_PATH"]
