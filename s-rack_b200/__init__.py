"""srack_b200: B200-native (sm_100a) offline voice renderer for s-rack patch graphs.

Only the module-graph tick of sharph/s-rack (src/synth.rs + src/synth/*) lives here,
behind the C ABI in include/srack_b200.h.  Importing this package loads the CUDA
library and fails loudly if it has not been built; there is no CPU path.
"""
from ._ffi import KIND, PARAM, SEQ_NONE, STATUS, LIB_PATH, grid_cell, lib  # noqa: F401
from .synth import (AudioConfig, Patch, PortError, SrackError, SynthModule, execute, get_catalog, get_inputs,  # noqa: F401
                    plan_execution, write_wav)
from . import patches, shard  # noqa: F401

__version__ = lib.srk_version().decode()
