#!/usr/bin/env python
"""Developer probe (one GPU): time of each rank's share of the C3 frame under the tile split, for several band heights.
The slowest share bounds the multi-GPU frame time; max/mean is the load imbalance."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rendu_b200 import pathtracer as rpt, workloads as wl
W, H, SPP, DEPTH, WORLD = 1920, 1080, 64, 8, 8
pt = rpt.PathTracer(rpt.Scene(wl.cornell_synthetic(708, 512, 256)))
cam = wl.cornell_camera(W / H)
d_sum = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
sh = torch.cuda.current_stream().cuda_stream
for rows in (1, 2, 4, 8, 16, 32):
    times = []
    for r in range(WORLD):
        o = rpt.make_opts(seed=1, tile=(rows, r, WORLD))
        pt.render_device(cam, W, H, SPP, DEPTH, d_sum.data_ptr(), o, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pt.render_device(cam, W, H, SPP, DEPTH, d_sum.data_ptr(), o, sh)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = np.array(times)
    print("band rows %2d: share times ms min %.2f mean %.2f max %.2f  imbalance max/mean %.3f  sum %.1f" % (rows, t.min(), t.mean(), t.max(), t.max() / t.mean(), t.sum()), flush=True)
o = rpt.make_opts(seed=1)
pt.render_device(cam, W, H, SPP, DEPTH, d_sum.data_ptr(), o, sh); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pt.render_device(cam, W, H, SPP, DEPTH, d_sum.data_ptr(), o, sh); e1.record(); torch.cuda.synchronize()
print("whole frame on one GPU: %.1f ms" % e0.elapsed_time(e1))
