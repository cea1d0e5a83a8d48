#!/usr/bin/env python
"""Developer probe (GPU box): many-mesh scenes. Closest-hit throughput with and without the root-block skipping
(RB_NO_ROOT_BLOCKS=1) and, for 9..64 meshes, with and without the tree over the root boxes (RB_NO_ROOT_TREE=1), results
compared bit for bit."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb
from rendu_b200 import workloads as wl

def build(nmesh, ordered):
    rng = np.random.RandomState(7)
    r = rb.Raycaster()
    p, i = wl.displaced_grid(8)
    side = int(np.ceil(np.sqrt(nmesh)))
    for k in range(nmesh):
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[1, 1] = m[2, 2] = 0.8 / side
        if ordered:  # objects listed row by row, as a scene file or an instancing loop does
            m[3, :3] = [-0.9 + 1.8 * (k % side) / side, 0.1 * rng.uniform(), -0.9 + 1.8 * (k // side) / side]
        else:
            m[3, :3] = rng.uniform(-1, 1, 3) * np.array([1, 0.3, 1], np.float32)
        r.addMesh(p, i, m.reshape(16))
    r.updateHierarchy()
    return r

co, cd = wl.coherent_rays(1024)
n = co.shape[0]
dev = torch.device("cuda", 0)
d_o, d_d = torch.from_numpy(co).to(dev), torch.from_numpy(cd).to(dev)
sh = torch.cuda.current_stream().cuda_stream
for nmesh, ordered in ((64, True), (443, True), (2000, True), (443, False)):
    outs = {}
    for name, env in (("blocks", None), ("linear", "1")):
        if env: os.environ["RB_NO_ROOT_BLOCKS"] = env
        else: os.environ.pop("RB_NO_ROOT_BLOCKS", None)
        r = build(nmesh, ordered)
        d_h = torch.empty(n * 48, dtype=torch.uint8, device=dev)
        for _ in range(3):
            r.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
        e1.record(); torch.cuda.synchronize()
        outs[name] = (e0.elapsed_time(e1) / 5, d_h.clone())
    same = bool((outs["blocks"][1] == outs["linear"][1]).all())
    print("%s order, %5d meshes: with root blocks %.3f ms, linear root loop %.3f ms (%.1fx), identical hit records: %s" % ("row" if ordered else "random", nmesh, outs["blocks"][0], outs["linear"][0], outs["linear"][0] / outs["blocks"][0], same), flush=True)

# 9..512 meshes: the root tree (rb_roots_tree_pretest) against the plain root loop; random mesh order (no blocks either way)
for nmesh in (12, 24, 52, 64, 65, 128, 443, 512):
    outs = {}
    for name, env in (("tree", None), ("linear", "1")):
        if env: os.environ["RB_NO_ROOT_TREE"] = env
        else: os.environ.pop("RB_NO_ROOT_TREE", None)
        os.environ["RB_NO_ROOT_BLOCKS"] = "1"
        r = build(nmesh, False)
        d_h = torch.empty(n * 48, dtype=torch.uint8, device=dev)
        d_a = torch.empty(n, dtype=torch.uint8, device=dev)
        for _ in range(3):
            r.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(5):
            r.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
        e1.record()
        for _ in range(5):
            r.intersectsAny_device(d_o.data_ptr(), d_d.data_ptr(), n, d_a.data_ptr(), 1e-4, 1e8, sh)
        e2.record(); torch.cuda.synchronize()
        outs[name] = (e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5, d_h.clone(), d_a.clone())
    same = bool((outs["tree"][2] == outs["linear"][2]).all()) and bool((outs["tree"][3] == outs["linear"][3]).all())
    print("random order, %5d meshes: closest-hit with the root tree %.3f ms, root loop %.3f ms (%.2fx); any-hit %.3f / %.3f ms (%.2fx); identical records: %s"
          % (nmesh, outs["tree"][0], outs["linear"][0], outs["linear"][0] / outs["tree"][0], outs["tree"][1], outs["linear"][1], outs["linear"][1] / outs["tree"][1], same), flush=True)
os.environ.pop("RB_NO_ROOT_TREE", None); os.environ.pop("RB_NO_ROOT_BLOCKS", None)
