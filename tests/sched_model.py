#!/usr/bin/env python
"""Analysis aid on the test side (CPU only; it drives the oracle, which only tests/ may do): a warp-instruction model of the phase-voting traversal scheduler (rb_k_trace_phased).

The oracle replays the reference traversal of a block of C2 rays and hands back, per ray, the sequence of steps the
kernel executes for it (N = node step on an interior node, P = node step that only pops a leaf, L = one triangle test).
A 32-lane warp is then simulated under a scheduling policy and every warp-level instruction block is charged once per
execution, whatever the number of lanes that take part — the quantity the kernel is bound by (issue /
divergence bound). The model ignores latency, so it ranks policies; it does not predict milliseconds.

    python tests/sched_model.py [--rows 8] [--set incoherent|coherent]
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendu_b200 import workloads as wl
from oracle import oracle as orc

# warp-instruction costs read off the SASS of rb_k_trace_phased<false,false,false> (cuobjdump -sass)
C_NODE, C_LEAF, C_VOTE = 61, 80, 28


def ray_tokens(o, org, dirs):
    """Per ray: uint8 array of steps (0 = N, 1 = P, 2 = L) in execution order."""
    leaf = o.nodes["leaf"].astype(bool)
    count = o.nodes["right"].astype(np.int64)
    out = []
    for k in range(org.shape[0]):
        ids, depths, _, total, _ = o.ray_levels(org[k], dirs[k], cap=8192)
        assert total == ids.size
        toks = []
        prev_descends = False  # previous visited node was an interior node one level up: this one is followed directly
        prev_depth = -1
        for i, d in zip(ids.astype(np.int64), depths.astype(np.int64)):
            direct = prev_descends and d == prev_depth + 1
            if leaf[i]:
                if not direct and toks:
                    toks.append(1)
                toks.extend([2] * int(count[i]))
                prev_descends = False
            else:
                toks.append(0)
                prev_descends = True
            prev_depth = d
        out.append(np.array(toks, np.uint8))
    return out


def simulate_postponed(rays, threshold, votes, steps, c_refill, c_finish, c_park=4):
    """The same warp when a lane may PARK one leaf and keep descending (rb_k_trace_phased with POSTPONE):
    a lane waits for the leaf phase only when it reaches a second leaf while one is parked, or when it has run out of
    nodes with a leaf still parked; a leaf phase serves every lane that has a parked triangle, waiting or not. The step
    sequences are the reference's (the few extra nodes a later-shrinking `maxi` lets through are not modelled).
    Returns (warp-instructions per ray, mean lanes busy in node steps, mean lanes busy in leaf steps)."""
    nxt, n = 0, len(rays)
    seq = [None] * 32
    pos = np.zeros(32, np.int64)
    parked = np.zeros(32, np.int64)  # triangles left in the parked leaf
    live = np.zeros(32, bool)
    cost = 0.0
    node_steps = node_lanes = leaf_steps = leaf_lanes = 0

    def at_leaf(l):
        return pos[l] < seq[l].size and seq[l][pos[l]] == 2

    def park(l):  # the leaf the lane stands on becomes its parked leaf (all its triangles)
        k = 0
        while at_leaf(l):
            pos[l] += 1
            k += 1
        parked[l] = k

    while True:
        exhausted = nxt >= n
        if not exhausted and (~live).any():
            for l in np.flatnonzero(~live):
                if nxt < n:
                    seq[l], pos[l], live[l], parked[l] = rays[nxt], 0, True, 0
                    nxt += 1
            cost += c_refill
            exhausted = nxt >= n
        if not live.any():
            break
        while True:
            for l in np.flatnonzero(live):
                if parked[l] == 0 and at_leaf(l):
                    park(l)
            blocked = np.array([live[l] and parked[l] > 0 and (at_leaf(l) or pos[l] >= seq[l].size) for l in range(32)])
            wants_node = live & ~blocked
            cost += C_VOTE
            finished = np.zeros(32, bool)
            if blocked.sum() >= votes or not wants_node.any():
                cost += C_LEAF
                leaf_steps += 1
                for l in np.flatnonzero(live & (parked > 0)):
                    parked[l] -= 1
                    leaf_lanes += 1
            else:
                act = wants_node.copy()
                for _ in range(steps):
                    if not act.any():
                        break
                    cost += C_NODE + c_park
                    node_steps += 1
                    node_lanes += int(act.sum())
                    for l in np.flatnonzero(act):
                        if pos[l] >= seq[l].size:
                            if parked[l] > 0:  # out of nodes, but a parked leaf is left: wait for the leaf phase
                                act[l] = False
                            else:
                                finished[l] = True
                                act[l] = False
                            continue
                        pos[l] += 1
                        if at_leaf(l):
                            if parked[l] == 0:
                                park(l)
                            else:
                                act[l] = False
            if finished.any():
                cost += c_finish
                live &= ~finished
            nlive = int(live.sum())
            if not (nlive > 0 if exhausted else nlive >= threshold):
                break
    return cost / n, node_lanes / max(node_steps, 1), leaf_lanes / max(leaf_steps, 1)


def simulate(rays, threshold, votes, steps, c_refill, c_finish, per_ray_extra=0.0):
    """One persistent warp consuming `rays` in order. Returns (warp-instructions per ray, mean lanes busy in node steps)."""
    nxt = 0
    seq = [None] * 32
    pos = np.zeros(32, np.int64)
    live = np.zeros(32, bool)
    cost = 0.0
    node_steps = node_lanes = 0
    n = len(rays)
    while True:
        exhausted = nxt >= n
        if not exhausted and (~live).any():
            for l in np.flatnonzero(~live):
                if nxt < n:
                    seq[l], pos[l], live[l] = rays[nxt], 0, True
                    nxt += 1
            cost += c_refill
            exhausted = nxt >= n
        if not live.any():
            break
        while True:
            wants_leaf = np.array([live[l] and pos[l] < seq[l].size and seq[l][pos[l]] == 2 for l in range(32)])
            wants_node = live & ~wants_leaf
            cost += C_VOTE
            finished = np.zeros(32, bool)
            if wants_leaf.sum() >= votes or not wants_node.any():
                cost += C_LEAF
                for l in np.flatnonzero(wants_leaf):
                    pos[l] += 1
                    # a ray ends in the node phase (stack and roots empty), never here
            else:
                act = wants_node.copy()
                for _ in range(steps):
                    if not act.any():
                        break
                    cost += C_NODE
                    node_steps += 1
                    node_lanes += int(act.sum())
                    for l in np.flatnonzero(act):
                        if pos[l] >= seq[l].size:  # nothing left: this step finds the stack and the roots empty
                            finished[l] = True
                            act[l] = False
                            continue
                        pos[l] += 1
                        if pos[l] < seq[l].size and seq[l][pos[l]] == 2:
                            act[l] = False
            if finished.any():
                cost += c_finish
                live &= ~finished
            nlive = int(live.sum())
            if not (nlive > 0 if exhausted else nlive >= threshold):
                break
    return cost / n + per_ray_extra, node_lanes / max(node_steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4, help="rows of the 2048-wide coherent grid to replay")
    ap.add_argument("--set", default="incoherent")
    args = ap.parse_args()
    pos, idx = wl.displaced_grid(708)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    co, cd = wl.coherent_rays(2048)
    r0 = 1000 * 2048
    co, cd = co[r0:r0 + args.rows * 2048], cd[r0:r0 + args.rows * 2048]
    if args.set == "incoherent":
        h = o.intersect(co, cd, threads=8)
        nrm = wl.triangle_normals(pos, idx, np.where(h["hit"] == 1, h["localId"], 0))
        org, dirs = wl.incoherent_rays(co, cd, h["hit"], h["dist"], nrm, seed=1234, index_base=r0)
    else:
        org, dirs = co, cd
    rays = ray_tokens(o, org, dirs)
    nn = np.mean([(t == 0).sum() for t in rays]); npop = np.mean([(t == 1).sum() for t in rays]); nl = np.mean([(t == 2).sum() for t in rays])
    ideal = nn * C_NODE + npop * C_NODE + nl * C_LEAF
    print("%d %s rays: %.1f node steps, %.1f pop-only steps, %.1f triangle tests per ray; %.0f thread-instructions per ray" % (len(rays), args.set, nn, npop, nl, ideal))
    print("perfect SIMD: %.0f warp-instructions per ray" % (ideal / 32))
    print("%-34s %s" % ("policy", "  ".join("T=%-2d" % t for t in (4, 8, 12, 16, 20, 24, 28, 32))))
    for name, votes, steps, cr, cf, extra in (("kernel today  v6 s6 (R150 F45)", 6, 6, 150, 45, 0.0),
                                             ("kernel today  v8 s1 (R150 F45)", 8, 1, 150, 45, 0.0),
                                             ("kernel today  v4 s4 (R150 F45)", 4, 4, 150, 45, 0.0),
                                             ("slim refill   v6 s6 (R30 F45)", 6, 6, 30, 45, 150 / 32),
                                             ("slim both     v6 s6 (R30 F12)", 6, 6, 30, 12, (150 + 45) / 32),
                                             ("slim both     v6 s3 (R30 F12)", 6, 3, 30, 12, (150 + 45) / 32),
                                             ("slim both     v8 s6 (R30 F12)", 8, 6, 30, 12, (150 + 45) / 32)):
        row, occ = [], []
        for t in (4, 8, 12, 16, 20, 24, 28, 32):
            c, a = simulate(rays, t, votes, steps, cr, cf, extra)
            row.append("%4.0f" % c)
            occ.append("%4.1f" % a)
        print("%-34s %s   | lanes/node step: %s" % (name, "  ".join(row), " ".join(occ)))
    for name, votes, steps in (("parked leaf   v6 s6 (R150 F45)", 6, 6), ("parked leaf   v12 s6", 12, 6), ("parked leaf   v16 s8", 16, 8), ("parked leaf   v24 s8", 24, 8)):
        row, occ, locc = [], [], []
        for t in (4, 8, 12, 16, 20, 24, 28, 32):
            c, a, b = simulate_postponed(rays, t, votes, steps, 150, 45)
            row.append("%4.0f" % c)
            occ.append("%4.1f" % a)
            locc.append("%4.1f" % b)
        print("%-34s %s   | lanes/node step: %s | lanes/leaf step: %s" % (name, "  ".join(row), " ".join(occ), " ".join(locc)))


if __name__ == "__main__":
    main()
