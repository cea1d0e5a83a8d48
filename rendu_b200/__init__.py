"""rendu_b200 — B200-native implementation of Rendu's offline path-tracer hot path.

The product is the CUDA shared library ``rendu_b200/librendu_b200.so`` (C ABI in ``include/rendu_b200.h``).
This package is the thin Python host layer over that ABI used by the tests and by ``bench.py``: ``Raycaster`` and
``PathTracer`` mirror the reference classes (src/engine/raycaster/Raycaster.hpp, src/apps/pathtracer/PathTracer.hpp).
There is no CPU fallback anywhere in this package: every compute call goes through the CUDA library and raises
``RenduError`` if the library or a GPU is missing.
"""
from ._capi import RenduError, lib, lib_path, device_count, HIT_DTYPE, NODE_DTYPE, TRI_DTYPE  # noqa: F401
from .raycaster import Raycaster, BUILD_REFERENCE, BUILD_LBVH, BUILD_SAH, TRACE_EXACT, TRACE_ORDERED, TRACE_WIDE  # noqa: F401

__all__ = ["RenduError", "lib", "lib_path", "device_count", "Raycaster", "HIT_DTYPE", "NODE_DTYPE", "TRI_DTYPE",
           "BUILD_REFERENCE", "BUILD_LBVH", "BUILD_SAH", "TRACE_EXACT", "TRACE_ORDERED", "TRACE_WIDE"]
