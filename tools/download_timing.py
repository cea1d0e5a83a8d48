#!/usr/bin/env python
"""Developer probe (GPU box): cost of reading an LBVH tree back in the reference's format (the C++ drop-in does this
after every build to refresh the members RaycasterVisualisation reads)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb
from rendu_b200 import workloads as wl
for n in (708, 2236):
    pos, idx = wl.displaced_grid(n)
    r = rb.Raycaster(); r.addMesh(pos, idx); r.updateHierarchy(rb.BUILD_LBVH)
    for k in range(2):
        t0 = time.perf_counter(); nodes, tris, verts = r.download_bvh(); t1 = time.perf_counter()
        print("N=%d LBVH download_bvh: %.1f ms (%d nodes)" % (n, 1e3 * (t1 - t0), nodes.size), flush=True)
