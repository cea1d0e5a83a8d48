"""The scheduler model (tests/sched_model.py) stays runnable: on a small mesh its step sequences agree with the oracle's
own traversal, and the simulated warp never beats perfect SIMD packing."""
import numpy as np

import sched_model as sm
from rendu_b200 import workloads as wl
from oracle import oracle as orc


def test_step_sequences_and_cost_bounds():
    pos, idx = wl.displaced_grid(24)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    org, dirs = wl.coherent_rays(12)  # 144 rays
    rays = sm.ray_tokens(o, org, dirs)
    assert len(rays) == org.shape[0]
    # every visited interior node is one N, every triangle of every visited leaf one L
    leaf = o.nodes["leaf"].astype(bool)
    for k in (0, 77, 143):
        ids = o.ray_levels(org[k], dirs[k])[0].astype(np.int64)
        assert (rays[k] == 0).sum() == (~leaf[ids]).sum()
        assert (rays[k] == 2).sum() == o.nodes["right"][ids][leaf[ids]].sum()
    ideal = sum((t != 2).sum() * sm.C_NODE + (t == 2).sum() * sm.C_LEAF for t in rays) / 32.0 / len(rays)
    for threshold in (8, 12, 32):
        cost, lanes = sm.simulate(rays, threshold, 6, 6, 150, 45)
        assert cost > ideal and 1.0 <= lanes <= 32.0
