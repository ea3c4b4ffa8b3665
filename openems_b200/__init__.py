"""openems_b200 -- B200-native FDTD engine for openEMS (Engine_CUDA / Operator_CUDA).

The product is libopenems_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/openems_b200.h).  This package is the Python host side above that ABI: a ctypes
binding (`_lib`) and `Operator_CUDA` / `Engine_CUDA` classes that mirror the reference's
Operator / Engine interface for the hot path (same method names, argument meaning and error
behaviour).  Nothing here computes fields on the CPU and nothing imports oracle/.
"""
from ._lib import load_library, library_path, LibraryNotBuilt  # noqa: F401
from .engine import Engine_CUDA, Operator_CUDA, Engine_Interface_CUDA, EngineError  # noqa: F401
from .synthetic import SyntheticOperator  # noqa: F401

__all__ = ["load_library", "library_path", "LibraryNotBuilt", "Engine_CUDA", "Operator_CUDA",
           "Engine_Interface_CUDA", "EngineError", "SyntheticOperator"]
