"""tortoisesat.jl_b200 -- B200-native batched Monte-Carlo slew engine.

Host-side mirror (Python, because this image has no Julia) of the reference's
function-level interface for the Monte-Carlo hot path, over the C ABI of
``libtortoise_b200.so`` (include/tortoise_b200.h).  The Julia drop-in files that
bind the same ABI through ``ccall`` are under ``julia/``.

There is no CPU fallback: every compute entry point raises if the CUDA
library is missing or no GPU is usable.
"""
from .host import (  # noqa: F401
    Engine,
    MultiEngine,
    TortoiseError,
    lib_path,
    load_library,
)
from . import host  # noqa: F401
