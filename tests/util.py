"""Shared helpers for the parity tests."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def case_meshes(case):
    out = []
    i = 0
    while "mesh%d_pos" % i in case:
        out.append((case["mesh%d_pos" % i], case["mesh%d_idx" % i], case["mesh%d_model" % i]))
        i += 1
    return out


def header_symbols():
    """Every RB_API entry point declared in include/rendu_b200.h."""
    text = open(os.path.join(ROOT, "include", "rendu_b200.h")).read()
    return sorted(set(re.findall(r"RB_API[^;(]*?\b(rb_[a-z0-9_]+)\s*\(", text)))


def assert_hits_equal(got, exp_hit, exp_dist, exp_u, exp_v, exp_w, exp_local, exp_mesh, what=""):
    """Bit-exact comparison of rb_hit records with reference / oracle arrays."""
    hit = got["hit"].astype(np.uint32)
    assert np.array_equal(hit, np.asarray(exp_hit, np.uint32)), what + ": hit flags differ at %s" % np.flatnonzero(hit != exp_hit)[:8]
    for name, exp in (("dist", exp_dist), ("u", exp_u), ("v", exp_v), ("w", exp_w)):
        a = got[name].view(np.uint32)
        b = np.asarray(exp, np.float32).view(np.uint32)
        bad = np.flatnonzero(a != b)
        assert bad.size == 0, "%s: %s differs in %d rays, first %s: got %r want %r" % (what, name, bad.size, bad[:4], got[name][bad[:4]], np.asarray(exp)[bad[:4]])
    assert np.array_equal(got["localId"], np.asarray(exp_local, np.uint64)), what + ": localId differs"
    assert np.array_equal(got["meshId"], np.asarray(exp_mesh, np.uint64)), what + ": meshId differs"
