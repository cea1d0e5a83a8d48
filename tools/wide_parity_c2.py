#!/usr/bin/env python
"""Parity of the eight-wide traversal at BASELINE config 2's full size, without a GPU: the kernels' own per-ray source
(rb_wide.cuh compiled by g++, tests/host_check) against the C oracle on all 4,194,304 coherent and 4,194,304 incoherent rays
of C2 (1,002,528 triangles), closest hit and any hit. Test-side tooling (it drives the oracle); one host thread for the
wide side, about two minutes.

    python tools/wide_parity_c2.py [--res 2048]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_wide as tw
import wide_model as wm
from rendu_b200 import workloads as wl
from rendu_b200._capi import ptr
from oracle import oracle as orc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=2048)
    ap.add_argument("--grid", type=int, default=708)
    args = ap.parse_args()
    L = tw._host_check()
    pos, idx = wl.displaced_grid(args.grid)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    r, T = wm.build([(pos, idx)])
    roots = np.zeros((1, 8), np.float32)
    roots[0, 0:3], roots[0, 4:7] = T.root_boxes[0]
    nodes, tris, wroots = np.ascontiguousarray(T.nodes), np.ascontiguousarray(T.tris), np.ascontiguousarray(T.roots)
    co, cd = wl.coherent_rays(args.res)
    h = o.intersect(co, cd, threads=os.cpu_count())
    nrm = wl.triangle_normals(pos, idx, np.where(h["hit"] == 1, h["localId"], 0))
    io, idr = wl.incoherent_rays(co, cd, h["hit"], h["dist"], nrm, seed=1234)
    out = {"triangles": int(idx.size // 3), "wide_nodes": int(nodes.shape[0]), "rays_per_set": int(co.shape[0]), "sets": {}}
    for name, org, dirs in (("coherent", co, cd), ("incoherent", io, idr)):
        org, dirs = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(dirs, np.float32)
        n = org.shape[0]
        ho = o.intersect(org, dirs, threads=os.cpu_count())
        ha = o.intersect_any(org, dirs, threads=os.cpu_count())
        rec = {}
        for mode, label in ((0, "rb_wide_step"), (1, "phase-split")):
            tri, tuv = np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
            t0 = time.time()
            skipped = L.rb_hc_wide_trace(ptr(roots), 1, ptr(nodes), ptr(tris), ptr(wroots), r._wide_sizes["depth"], ptr(org), ptr(dirs), n, 1e-4, 1e8, 0, mode, 0, ptr(tri), ptr(tuv), None)
            took, hit = tri != -2, tri >= 0
            differ = took & ((hit != ho["hit"].astype(bool)) | (hit & ((tri != ho["internalId"].astype(np.int64)) | (tuv[:, 0].view(np.uint32) != ho["dist"].view(np.uint32)))))
            ta, tuva = np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
            L.rb_hc_wide_trace(ptr(roots), 1, ptr(nodes), ptr(tris), ptr(wroots), r._wide_sizes["depth"], ptr(org), ptr(dirs), n, 1e-4, 1e8, 1, mode, 0, ptr(ta), ptr(tuva), None)
            rec[label] = {"rays_on_the_wide_tree": int(took.sum()), "irregular_rays_left_to_the_binary_tree": int(skipped), "closest_hit_records_that_differ": int(differ.sum()),
                          "any_hit_flags_that_differ": int((took & ((ta >= 0) != ha.astype(bool))).sum()), "seconds": round(time.time() - t0, 1)}
        out["sets"][name] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
