#!/usr/bin/env python
"""Developer probe (GPU box): host-side phases of rb_raycaster_build(LBVH) at 10M triangles (RB_BUILD_TIMING=1)."""
import os, sys, time
os.environ["RB_BUILD_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb
from rendu_b200 import workloads as wl
pos, idx = wl.displaced_grid(2236)
for k in range(3):
    r = rb.Raycaster()
    t0 = time.perf_counter(); r.addMesh(pos, idx); t1 = time.perf_counter()
    r.updateHierarchy(rb.BUILD_LBVH); t2 = time.perf_counter()
    print("addMesh %.1f ms, updateHierarchy(LBVH) %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    del r
