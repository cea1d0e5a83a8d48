#!/usr/bin/env python
"""Developer probe (GPU box): how much does the ORDER of the rays in a batch matter for the traversal kernels?
Times closest-hit on the C2 incoherent set in its natural order, shuffled, and sorted by (direction octant, origin
Morton code) / (origin Morton code, octant). Decides whether a ray-sorting pass could pay for itself."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendu_b200 as rb  # noqa: E402
from rendu_b200 import workloads as wl  # noqa: E402

dev = torch.device("cuda", 0)
pos, idx = wl.displaced_grid(708)
rc = rb.Raycaster()
rc.addMesh(pos, idx)
rc.updateHierarchy()
co, cd = wl.coherent_rays(2048)
n = co.shape[0]
sh = torch.cuda.current_stream().cuda_stream
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
d_co, d_cd = to(co), to(cd)
d_h = torch.empty(n * 48, dtype=torch.uint8, device=dev)
rc.intersects_device(d_co.data_ptr(), d_cd.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
torch.cuda.synchronize()
hits = d_h.cpu().numpy().view(rb.HIT_DTYPE)
nrm = wl.triangle_normals(pos, idx, np.where(hits["hit"] == 1, hits["localId"], 0))
io, idr = wl.incoherent_rays(co, cd, hits["hit"], hits["dist"], nrm, seed=1234)


def part1by2(x):
    x = x.astype(np.uint64) & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton(p, bits):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


octant = ((idr[:, 0] < 0).astype(np.uint64) | ((idr[:, 1] < 0).astype(np.uint64) << 1) | ((idr[:, 2] < 0).astype(np.uint64) << 2))
orders = {"natural": np.arange(n), "shuffled": np.random.RandomState(1).permutation(n),
          "octant_then_morton10": np.argsort((octant << 30) | morton(io, 10), kind="stable"),
          "morton5_then_octant": np.argsort((morton(io, 5) << 3) | octant, kind="stable"),
          "morton10": np.argsort(morton(io, 10), kind="stable"),
          "octant_only": np.argsort(octant, kind="stable")}
for name, order in orders.items():
    d_o, d_d = to(io[order]), to(idr[order])
    for _ in range(5):
        rc.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), 1e-4, 1e8, sh)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-24s %.3f ms  %.0f Mrays/s" % (name, np.median(ts), n / np.median(ts) / 1e3), flush=True)
