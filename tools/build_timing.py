#!/usr/bin/env python
"""Developer probe (GPU box): time-to-first-ray of both build paths (1M and 10M triangles)."""
import os, sys, time
os.environ["RB_BUILD_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb
from rendu_b200 import workloads as wl
for n, modes in ((708, (rb.BUILD_REFERENCE, rb.BUILD_LBVH)), (2236, (rb.BUILD_LBVH,))):
    pos, idx = wl.displaced_grid(n)
    for mode in modes:
        for k in range(2):
            r = rb.Raycaster()
            t0 = time.perf_counter(); r.addMesh(pos, idx); t1 = time.perf_counter()
            r.updateHierarchy(mode); t2 = time.perf_counter()
            print("N=%d mode=%d addMesh %.1f ms, updateHierarchy %.1f ms" % (n, mode, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
            del r
