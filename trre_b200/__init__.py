"""trre_b200 - B200-native scan engine for trre (transductive regular expressions).

Thin host-side mirror of the C ABI in include/trre_b200.h.  The scan path runs in hand-written
sm_100a CUDA kernels inside trre_b200/lib/libtrre_b200.so; there is no Python or CPU implementation
of it here, and importing the scan entry points without the library raises.
"""
from ._lib import (  # noqa: F401
    FLAVOR_NFT, FLAVOR_DFT, EXEC_BYTEMAP, EXEC_TABLE, EXEC_BACKTRACK, EXEC_LAZY,
    TrreError, Program, compile, from_blob, shard_lines, synth_text, lib_path, version,
)

__all__ = ["FLAVOR_NFT", "FLAVOR_DFT", "EXEC_BYTEMAP", "EXEC_TABLE", "EXEC_BACKTRACK", "EXEC_LAZY", "TrreError",
           "Program", "compile", "from_blob", "shard_lines", "synth_text", "lib_path", "version"]
