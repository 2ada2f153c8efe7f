"""rustronomy-watershed_b200: B200-native (sm_100a) watershed engine behind the
public API of smups/rustronomy-watershed.

The directory name carries a hyphen (it mirrors the crate name), so import it
through the loader at the repo root:

    from wsb200_loader import load
    ws = load()                      # -> module `rustronomy_watershed_b200`
    t = ws.TransformBuilder.default().build_segmenting()
"""
from . import _native
from ._native import Context, Plan, WatershedError, default_context, load_library
from .api import (ALWAYS_FILL, NEVER_FILL, NORMAL_MAX, UNCOLOURED, BuildErr, HookCtx, MergingWatershed,
                  SegmentingWatershed, TransformBuilder, Watershed, WatershedUtils)

# lib.rs:144-154 (`prelude` re-exports exactly these four)
prelude = ("MergingWatershed", "TransformBuilder", "Watershed", "WatershedUtils")

__all__ = [
    "ALWAYS_FILL", "NEVER_FILL", "NORMAL_MAX", "UNCOLOURED", "BuildErr", "HookCtx", "MergingWatershed",
    "SegmentingWatershed", "TransformBuilder", "Watershed", "WatershedUtils", "Context", "Plan",
    "WatershedError", "default_context", "load_library", "prelude",
]
# `strips` (row-strip decomposition over several GPUs) imports torch; load it on demand:
#   from rustronomy_watershed_b200 import strips
