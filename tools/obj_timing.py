#!/usr/bin/env python
"""OBJ ingestion: rb_obj_parse against the reference's Mesh(std::istream &, Load::Indexed) on the C2 mesh written as an
OBJ text (1,002,528 triangles). Host code on both sides — runs anywhere oracle/_ref is built; no GPU involved.

    python tools/obj_timing.py [--grid 708]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendu_b200 import mesh as rm
from rendu_b200 import workloads as wl
from oracle import ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=708)
    args = ap.parse_args()
    pos, idx = wl.displaced_grid(args.grid)
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (pos.shape[0], 1))
    lines = ["v %r %r %r" % (float(p[0]), float(p[1]), float(p[2])) for p in pos]
    lines += ["vn %r %r %r" % (float(n[0]), float(n[1]), float(n[2])) for n in nrm]
    lines += ["vt %r %r" % (float(p[0]), float(p[2])) for p in pos]
    lines += ["f %d/%d/%d %d/%d/%d %d/%d/%d" % (a, a, a, b, b, b, c, c, c) for a, b, c in (idx.reshape(-1, 3) + 1)]
    text = ("\n".join(lines) + "\n").encode()
    out = {"triangles": int(idx.size // 3), "vertices": int(pos.shape[0]), "bytes": len(text), "host_threads": os.cpu_count(), "mode": "Indexed (what Resources::getMesh uses)"}
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        m = rm.Mesh(text, rm.INDEXED)
        t.append(time.perf_counter() - t0)
    out["rb_obj_parse_ms"] = 1e3 * min(t)
    t0 = time.perf_counter()
    r = ref.RefMesh(text, rm.INDEXED)
    out["reference_mesh_ms"] = 1e3 * (time.perf_counter() - t0)
    out["identical"] = bool(np.array_equal(m.positions.view(np.uint32), r.positions.view(np.uint32)) and np.array_equal(m.indices, r.indices) and np.array_equal(m.texcoords.view(np.uint32), r.texcoords.view(np.uint32)) and np.array_equal(m.normals.view(np.uint32), r.normals.view(np.uint32)))
    out["speedup"] = out["reference_mesh_ms"] / out["rb_obj_parse_ms"]
    out["note"] = "both include copying the arrays out into numpy"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
