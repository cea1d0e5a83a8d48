#!/usr/bin/env python
"""Analysis aid on the test side (CPU only): a restatement of the eight-wide traversal (rendu_b200/csrc/rb_wide.cuh)
in numpy fp32, run on the wide tree the library's host collapse produces (rb_raycaster_wide_build needs no GPU).

It answers, before any GPU time is spent, the questions the layout stands or falls with:
  * is every stored child box a superset of the binary box it replaces, by the margin the kernel's arithmetic needs;
  * does the traversal return the oracle's closest hit (same triangle, same distance bits);
  * how many wide nodes / triangle tests / L1 wavefronts a ray costs against the binary tree's figures.

    python tests/wide_model.py [--rows 1] [--set incoherent|coherent]
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

F = np.float32


class WideTree:
    def __init__(self, nodes, tris, roots, root_boxes):
        self.nodes, self.tris, self.roots, self.root_boxes = nodes, tris, roots, root_boxes
        self.p = nodes[:, 0:3].copy().view(np.float32)
        self.e = np.stack([(nodes[:, 3] >> (8 * a)) & 0xff for a in range(3)], 1).astype(np.int32) - 127
        self.child_base = nodes[:, 4].astype(np.int64)
        self.tri_base = nodes[:, 5].astype(np.int64)
        self.meta = np.ascontiguousarray(nodes[:, 6:8]).view(np.uint8).reshape(-1, 8)
        q = np.ascontiguousarray(nodes[:, 8:20]).view(np.uint8).reshape(-1, 6, 8)
        self.qlo, self.qhi = q[:, 0:3, :], q[:, 3:6, :]
        self.hier = np.ascontiguousarray(tris[:, 11]).view(np.int32)

    def child_boxes(self, n):
        """Decoded boxes of node n in float64: (lo (3, 8), hi (3, 8))."""
        s = np.ldexp(1.0, self.e[n])[:, None]
        p = self.p[n].astype(np.float64)[:, None]
        return p + self.qlo[n].astype(np.float64) * s, p + self.qhi[n].astype(np.float64) * s


def make_ray(o, d):
    """Ray::Ray (Intersection.cpp:3-5) in fp32: normalised direction and its reciprocal."""
    o, d = o.astype(F), d.astype(F)
    n2 = F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
    d = (d * F(F(1.0) / np.sqrt(n2))).astype(F)
    with np.errstate(divide="ignore"):
        inv = (F(1.0) / d).astype(F)
    return o, d, inv


def _cross(x, y):
    return np.array([F(F(x[1] * y[2]) - F(y[1] * x[2])), F(F(x[2] * y[0]) - F(y[2] * x[0])), F(F(x[0] * y[1]) - F(y[0] * x[1]))], F)


def _dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def tri_test(o, d, rec, mini, maxi):
    """rb_tri_test (= Raycaster.cpp:232-262), operation for operation in fp32."""
    v0, v01, v02 = rec[0:3], rec[4:7], rec[8:11]
    p = _cross(d, v02)
    det = _dot(v01, p)
    if abs(det) < np.finfo(F).eps:
        return None
    inv_det = F(F(1.0) / det)
    q = (o - v0).astype(F)
    u = F(inv_det * _dot(q, p))
    if u < 0 or u > 1:
        return None
    r = _cross(q, v01)
    v = F(inv_det * _dot(d, r))
    if v < 0 or F(u + v) > 1:
        return None
    t = F(inv_det * _dot(v02, r))
    return (t, u, v) if (t > mini and t < maxi) else None


def box_test(o, inv, lo, hi, mini, maxi):
    """Intersection::box (Intersection.cpp:22-32) for a regular ray."""
    a, b = ((lo - o) * inv).astype(F), ((hi - o) * inv).astype(F)
    return max(np.minimum(a, b).max(), mini) <= min(np.maximum(a, b).min(), maxi)


def enter(T, n, o, inv, mini, maxi):
    """rb_wide_enter: the 24-bit pending list of node n (three bits per slot)."""
    A = (np.ldexp(F(1.0), T.e[n] + 15).astype(F) * inv).astype(F)
    B = (((T.p[n] - o).astype(F) * inv).astype(F) - A).astype(F)
    neg = inv < 0
    near = np.where(neg[:, None], T.qhi[n], T.qlo[n]).astype(np.uint32)
    far = np.where(neg[:, None], T.qlo[n], T.qhi[n]).astype(np.uint32)
    fn = ((near << 8) | 0x3f800000).astype(np.uint32).view(F)
    ff = ((far << 8) | 0x3f800000).astype(np.uint32).view(F)
    tn = (fn.astype(np.float64) * A.astype(np.float64)[:, None] + B.astype(np.float64)[:, None]).astype(F)
    tf = (ff.astype(np.float64) * A.astype(np.float64)[:, None] + B.astype(np.float64)[:, None]).astype(F)
    hit = np.maximum(tn.max(0), F(mini)) <= np.minimum(tf.min(0), F(maxi))
    # the kernel's bit arithmetic, literally: three bits per hit slot, masked by the unary counts of the meta bytes
    bits = 0
    for s in range(8):
        if hit[s]:
            bits |= 7 << (3 * s)
    meta_lo, meta_hi = int(T.nodes[n, 6]), int(T.nodes[n, 7])
    return bits & (_pack3((meta_lo >> 5) & 0x07070707) | (_pack3((meta_hi >> 5) & 0x07070707) << 12))


def _pack3(c):
    return (c & 0x7) | ((c >> 5) & 0x38) | ((c >> 10) & 0x1c0) | ((c >> 15) & 0xe00)


def _byte_perm_low(lo, hi, slot):
    """__byte_perm(lo, hi, slot) & 0xff: byte `slot` of the eight bytes lo (0-3), hi (4-7)."""
    return ((lo | (hi << 32)) >> (8 * slot)) & 0xff


def trace(T, o, d, mini=1e-4, maxi=1e8, any_hit=False, steps=None):
    """rb_wide_step until the ray finishes. Returns (hierarchy triangle or -1, t, u, v, nodes entered, triangle tests,
    pops)."""
    o, d, inv = make_ray(np.asarray(o), np.asarray(d))
    mini, maxi0 = F(mini), F(maxi)
    maxi = maxi0
    best = (-1, F(np.finfo(F).max), F(0), F(0))
    n_nodes = n_tris = n_pops = 0
    stack = []
    for r in range(len(T.roots) - 1, -1, -1):
        lo, hi = T.root_boxes[r]
        if T.roots[r] < 0 or not box_test(o, inv, lo, hi, mini, maxi0):
            continue
        cur = int(T.roots[r])
        pending = enter(T, cur, o, inv, mini, maxi)
        n_nodes += 1
        if steps is not None:
            steps.append(0)
        while True:
            if pending == 0:
                if not stack:
                    break
                cur, pending = stack.pop()
                n_pops += 1
            bit = (pending & -pending).bit_length() - 1
            pending &= pending - 1
            slot = (bit * 11) >> 5  # the kernel's bit / 3
            assert slot == bit // 3
            low = _byte_perm_low(int(T.nodes[cur, 6]), int(T.nodes[cur, 7]), slot) & 31
            assert low == int(T.meta[cur, slot]) & 31
            if low >= 24:
                if pending:
                    stack.append((cur, pending))
                cur = int(T.child_base[cur]) + low - 24
                pending = enter(T, cur, o, inv, mini, maxi)
                n_nodes += 1
                if steps is not None:
                    steps.append(0)
                continue
            tri = int(T.tri_base[cur]) + low + (bit - 3 * slot)
            n_tris += 1
            if steps is not None:
                steps.append(2)
            h = tri_test(o, d, T.tris[tri], mini, maxi)
            if h is not None:
                if any_hit:
                    return int(T.hier[tri]), h[0], h[1], h[2], n_nodes, n_tris, n_pops
                if h[0] < best[1]:
                    best = (int(T.hier[tri]), h[0], h[1], h[2])
                    maxi = h[0]
    return best[0], best[1], best[2], best[3], n_nodes, n_tris, n_pops


# warp-instruction costs read off the SASS of rb_k_trace_wide / rb_k_trace_wide_phased (cuobjdump -sass): entering a
# node (with the pop / push around it), one triangle test, the vote of the phased kernel, loop control of the plain one
C_ENTER, C_TRI, C_VOTE, C_LOOP = 200, 85, 30, 12


def simulate_warp(rays, threshold, phased, votes=6, c_refill=150, c_finish=45):
    """One persistent warp consuming `rays` (lists of 0 = node entry, 2 = triangle test) in order, as tests/sched_model.py
    does for the binary kernels: every instruction block is charged once per execution, whatever the number of lanes in it.
    Plain kernel: one rb_wide_step per lane and iteration, so an iteration costs a node entry if any lane enters a node plus
    a triangle test if any lane tests one. Phased: one kind per iteration, chosen by vote. Returns warp-instructions per ray
    and the mean number of lanes busy in node entries."""
    nxt, n = 0, len(rays)
    seq, pos = [None] * 32, np.zeros(32, np.int64)
    live = np.zeros(32, bool)
    cost, node_steps, node_lanes = 0.0, 0, 0
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
            kind = np.array([seq[l][pos[l]] if live[l] and pos[l] < len(seq[l]) else (1 if live[l] else -1) for l in range(32)])  # 1: out of items -> finishes in a node step
            wants_tri, wants_node = kind == 2, (kind == 0) | (kind == 1)
            if phased:
                cost += C_VOTE
                run_tri = wants_tri.sum() >= votes or not wants_node.any()
                run_node = not run_tri
            else:
                cost += C_LOOP
                run_tri, run_node = wants_tri.any(), wants_node.any()
            finished = np.zeros(32, bool)
            if run_tri:
                cost += C_TRI
                pos[wants_tri] += 1
            if run_node:
                enters = wants_node & (kind == 0)
                if enters.any():
                    cost += C_ENTER
                    node_steps += 1
                    node_lanes += int(enters.sum())
                    pos[enters] += 1
                finished = wants_node & (kind == 1)
            if finished.any():
                cost += c_finish
                live &= ~finished
            nlive = int(live.sum())
            if not (nlive > 0 if exhausted else nlive >= threshold):
                break
    return cost / n, node_lanes / max(node_steps, 1)


def build(pos_idx_list):
    """(library raycaster, WideTree) for a list of (positions, indices[, model]) meshes — host only, no GPU needed."""
    import rendu_b200 as rb
    r = rb.Raycaster()
    for m in pos_idx_list:
        r.addMesh(*m)
    r.build_host_only()
    r.build_wide()
    nodes, tris, roots = r.download_wide()
    bn, _, _ = r.download_bvh()
    nm = r.sizes()["meshes"]
    root_boxes = [(bn["bmin"][m].astype(F), bn["bmax"][m].astype(F)) for m in range(nm)]
    return r, WideTree(nodes, tris, roots, root_boxes)


def main():
    from rendu_b200 import workloads as wl
    from oracle import oracle as orc
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1)
    ap.add_argument("--set", default="incoherent")
    ap.add_argument("--grid", type=int, default=708)
    args = ap.parse_args()
    pos, idx = wl.displaced_grid(args.grid)
    o = orc.OracleRaycaster()
    o.add_mesh(pos, idx)
    o.build()
    r, T = build([(pos, idx)])
    kids = (T.meta != 0).sum(1)
    print("wide tree: %d nodes (%.1f MB), %d triangles, depth %d; children per node: %s" % (T.nodes.shape[0], T.nodes.nbytes / 1e6, T.tris.shape[0], r._wide_sizes["depth"], np.bincount(kids, minlength=9).tolist()))
    co, cd = wl.coherent_rays(2048)
    r0 = 1000 * 2048
    co, cd = co[r0:r0 + args.rows * 2048:4], cd[r0:r0 + args.rows * 2048:4]
    h = o.intersect(co, cd, threads=8)
    if args.set == "incoherent":
        nrm = wl.triangle_normals(pos, idx, np.where(h["hit"] == 1, h["localId"], 0))
        org, dirs = wl.incoherent_rays(co, cd, h["hit"], h["dist"], nrm, seed=1234, index_base=r0)
        h = o.intersect(org, dirs, threads=8)
    else:
        org, dirs = co, cd
    nn = nt = npop = bad = 0
    for k in range(org.shape[0]):
        tri, t, u, v, a, b, c = trace(T, org[k], dirs[k])
        nn, nt, npop = nn + a, nt + b, npop + c
        ok = (tri >= 0) == bool(h["hit"][k]) and (tri < 0 or (tri == int(h["internalId"][k]) and F(t).view(np.uint32) == h["dist"][k].view(np.uint32)))
        bad += not ok
    n = org.shape[0]
    print("%d %s rays: %.2f wide nodes, %.2f triangle tests, %.2f pops per ray; %d differ from the oracle" % (n, args.set, nn / n, nt / n, npop / n, bad))
    print("L1 wavefronts per ray if every lane is alone on its lines: wide %.0f (5 per node, 3 per triangle) — binary tree 4 per node, 3 per triangle" % (5 * nn / n + 3 * nt / n))
    seqs = []
    for k in range(org.shape[0]):
        st = []
        trace(T, org[k], dirs[k], steps=st)
        seqs.append(st)
    print("warp-instructions per ray (tests/sched_model.py gives the binary kernels: default scheduler 279 incoherent at threshold 12)")
    print("%-28s %s" % ("policy", "  ".join("T=%-2d" % t for t in (8, 12, 16, 20, 24, 28))))
    for name, phased, votes in (("wide, plain persistent", False, 0), ("wide, phased v6", True, 6), ("wide, phased v12", True, 12), ("wide, phased v16", True, 16)):
        row, occ = [], []
        for t in (8, 12, 16, 20, 24, 28):
            c, a = simulate_warp(seqs, t, phased, votes)
            row.append("%4.0f" % c)
            occ.append("%4.1f" % a)
        print("%-28s %s   | lanes per node entry: %s" % (name, "  ".join(row), " ".join(occ)))


if __name__ == "__main__":
    main()
