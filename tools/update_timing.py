#!/usr/bin/env python
"""Developer probe (GPU box): cost of an incremental transform update on the C2 mesh (1,002,528 triangles) and what a
refitted tree costs in traversal — rb_raycaster_update_mesh + rb_raycaster_refit(NULL) against a full rebuild."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb
from rendu_b200 import workloads as wl


def model(angle, ty):
    c, s = np.float32(np.cos(angle)), np.float32(np.sin(angle))
    return np.array([c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0, 0, ty, 0, 1], np.float32)


def trace_ms(r, org, dirs, reps=3):
    r.intersects(org[:4096], dirs[:4096])
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        r.intersects(org, dirs)
        best = min(best, 1e3 * (time.perf_counter() - t0))
    return best


pos, idx = wl.displaced_grid(708)
org, dirs = wl.coherent_rays(1024)
out = {}
for mode, name in ((rb.BUILD_REFERENCE, "reference"), (rb.BUILD_LBVH, "lbvh")):
    r = rb.Raycaster()
    r.addMesh(pos, idx)
    t0 = time.perf_counter(); r.updateHierarchy(mode); build = 1e3 * (time.perf_counter() - t0)
    base = trace_ms(r, org, dirs)
    upd, fit = [], []
    for k in range(1, 6):
        t0 = time.perf_counter(); r.update_mesh(0, pos, model(0.1 * k, 0.01 * k)); t1 = time.perf_counter()
        r.refit(); t2 = time.perf_counter()
        upd.append(1e3 * (t1 - t0)); fit.append(1e3 * (t2 - t1))
    moved = trace_ms(r, org, dirs)
    out[name] = dict(build_ms=build, update_mesh_ms=float(np.median(upd)), refit_ms=float(np.median(fit)), refit_device_ms=r.build_times()["refit_device_ms"],
                     trace_1M_rays_e2e_ms_before=base, after_rotation_0p5rad=moved)
print(json.dumps(out))
