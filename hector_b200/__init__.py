"""hector_b200 -- B200-native ensemble engine for Hector's per-year coupled hot path.

Host-side mirror (Python) of the reference's user-facing surface for this path
(newcore / setvar / run / reset / fetchvars, R/hector.R:57-87, R/messages.R:46-140) on top of
the C ABI in include/hector_b200.h.  The compute lives in libhector_b200.so (hand-written
sm_100a CUDA); there is no CPU fallback.
"""
from ._capi import HxError, lib, lib_path  # noqa: F401
from .ensemble import (BIOME_OUTPUTS, BIOME_PARAMETERS, DERIVED_VARIABLES, FUNCTION_VARIABLES, Ensemble, OUTPUT_VARIABLES, PARAMETERS, RAW_SERIES, STASH_OUTPUTS,  # noqa: F401
                       TRACK_POOLS, TRACK_POOL_OUTPUT, TRACK_SOURCES, load_scenario_tables)

__all__ = ["BIOME_OUTPUTS", "BIOME_PARAMETERS", "DERIVED_VARIABLES", "FUNCTION_VARIABLES", "Ensemble", "HxError", "OUTPUT_VARIABLES", "PARAMETERS", "RAW_SERIES", "STASH_OUTPUTS",
           "TRACK_POOLS", "TRACK_POOL_OUTPUT", "TRACK_SOURCES", "load_scenario_tables", "lib", "lib_path"]
