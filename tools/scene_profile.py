#!/usr/bin/env python
"""developer aid (GPU box): one bundled scene taken apart with rb_pathtracer_profile_frame — device ms per kernel class of a
single-stream 1080p frame, rays, and the nodes / triangles / roots the traversal wavefronts test per ray.
usage: python tools/scene_profile.py gallery-materials [spp]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import ref  # noqa: E402  (the reference's loader flattens the .scene; test infrastructure, a developer aid may use it)
from rendu_b200 import pathtracer as rpt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "gallery-materials"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rs = ref.RefScene(name)
desc, cam = rs.flatten(), rs.camera(1920 / 1080)
sc = rpt.Scene(desc)
pt = rpt.PathTracer(sc)
img = np.zeros((1080, 1920, 3), np.float32)
pt.render(cam, spp, 8, img, rpt.make_opts(seed=1))
t0 = time.perf_counter()
pt.render(cam, spp, 8, img, rpt.make_opts(seed=1))
dt = time.perf_counter() - t0
p = pt.profile_frame(cam, 1920, 1080, spp, 8, rpt.make_opts(seed=1))
rc, rsh = max(p["rays_closest"], 1), max(p["rays_shadow"], 1)
print(json.dumps({"scene": name, "meshes": len(desc["objects"]), "spp": spp, "frame_ms_e2e": 1e3 * dt, "spp_per_s": spp / dt, "ms": p["ms"],
                  "closest_per_sample": p["rays_closest"] / p["samples"], "shadow_per_sample": p["rays_shadow"] / p["samples"],
                  "per_closest_ray": {"interior": p["trace_interior"] / rc, "tris": p["trace_tris"] / rc, "roots": p["trace_roots"] / rc},
                  "per_shadow_ray": {"interior": p["shadow_interior"] / rsh, "tris": p["shadow_tris"] / rsh, "roots": p["shadow_roots"] / rsh}}))
