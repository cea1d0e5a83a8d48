"""Multi-GPU rendering INSIDE the C ABI (rb_render_opts.gpus / split; rendu_b200/csrc/rb_multi.cu): one process, one host
thread and one scene replica per GPU — what replaces the reference's System::forParallel fan-out
(src/engine/system/System.hpp:72-104, PathTracer.cpp:186) on an 8-GPU box.

Gates (no reference analogue, so the 1-GPU image of the same counter-based RNG is the oracle):
  tile split   : bit-identical to the 1-GPU image;
  sample split : equal up to fp32 summation order (<= 2e-6 relative), with the NCCL reduce and with the peer-memory resolve.
On a box with a single GPU the same test checks that asking for more devices than exist fails loudly instead.
"""
import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200 import pathtracer as rpt
from rendu_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _render(pt, cam, w, h, spp, depth, **kw):
    img = np.zeros((h, w, 3), np.float32)
    pt.render(cam, spp, depth, img, rpt.make_opts(seed=7, **kw))
    return img


def test_multi_gpu_render_through_the_abi(monkeypatch):
    n = rb.device_count()
    desc, cam = wl.cornell_synthetic(48, 64, 32), wl.cornell_camera(160 / 90)
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    one = _render(pt, cam, 160, 90, 16, 5)
    assert np.isfinite(one).all() and one.mean() > 0.01
    if n < 2:
        with pytest.raises(rb.RenduError):
            _render(pt, cam, 160, 90, 16, 5, gpus=2)
        return
    for g in sorted({2, min(n, 4), n}):
        tile = _render(pt, cam, 160, 90, 16, 5, gpus=g, split=rpt.SPLIT_TILE)
        assert np.array_equal(tile, one), "tile split over %d GPUs differs from one GPU" % g
        ms = pt.multi_stats()
        assert ms["gpus"] == g and ms["split"] == rpt.SPLIT_TILE and not ms["used_nccl"]
        odd = _render(pt, cam, 160, 90, 16, 5, gpus=g, split=rpt.SPLIT_TILE, tile_rows=7)  # 90 rows: partial last band
        assert np.array_equal(odd, one)
        smp = _render(pt, cam, 160, 90, 16, 5, gpus=g, split=rpt.SPLIT_SAMPLE)
        assert pt.multi_stats()["used_nccl"]
        assert np.abs(smp - one).max() <= 2e-6 * max(1.0, float(one.max())), np.abs(smp - one).max()
        monkeypatch.setenv("RB_SAMPLE_REDUCE", "p2p")
        p2p = _render(pt, cam, 160, 90, 16, 5, gpus=g, split=rpt.SPLIT_SAMPLE)
        monkeypatch.delenv("RB_SAMPLE_REDUCE")
        assert not pt.multi_stats()["used_nccl"]
        assert np.abs(p2p - one).max() <= 2e-6 * max(1.0, float(one.max()))
    # A sample count that is not a power of two is rounded like the reference does (12 -> 16, PathTracer.cpp:169-170); every
    # combine divides by the rounded count (the peer-memory resolve once divided by the caller's).
    one12 = _render(pt, cam, 160, 90, 12, 5)
    assert np.array_equal(one12, one)
    monkeypatch.setenv("RB_SAMPLE_REDUCE", "p2p")
    p2p12 = _render(pt, cam, 160, 90, 12, 5, gpus=2, split=rpt.SPLIT_SAMPLE)
    monkeypatch.delenv("RB_SAMPLE_REDUCE")
    assert np.abs(p2p12 - one).max() <= 2e-6 * max(1.0, float(one.max()))
    assert np.array_equal(_render(pt, cam, 160, 90, 12, 5, gpus=2, split=rpt.SPLIT_TILE), one)
    st = pt.stats()
    assert st["samples"] == 160 * 90 * 16
    # a scene edit invalidates the replicas: the next multi-GPU frame sees the new light on every GPU
    lights = [dict(l) for l in desc["lights"]]
    lights[0]["intensity"] = np.array([1.0, 9.0, 1.0], np.float32)
    sc.set_lights(lights)
    a = _render(pt, cam, 160, 90, 16, 5)
    b = _render(pt, cam, 160, 90, 16, 5, gpus=2)
    assert np.array_equal(a, b) and not np.array_equal(a, one)
    # the caller's image is accumulated into, as in the reference (PathTracer.cpp:288-292)
    img = np.full((90, 160, 3), 16.0, np.float32)
    pt.render(cam, 16, 5, img, rpt.make_opts(seed=7, gpus=2))
    assert np.allclose(img, a + 1.0, rtol=0, atol=1e-6)
    with pytest.raises(rb.RenduError):
        _render(pt, cam, 160, 90, 16, 5, gpus=n + 1)


def test_profile_frame_times_and_counts():
    """rb_pathtracer_profile_frame (bench.py's roofline source): every kernel class reports time and launches, the counting
    pass sees the same rays as an ordinary render, and the per-ray node counts are in a sane range."""
    desc, cam = wl.cornell_synthetic(64, 64, 32), wl.cornell_camera(1.0)
    pt = rpt.PathTracer(rpt.Scene(desc))
    img = np.zeros((128, 128, 3), np.float32)
    pt.render(cam, 16, 6, img, rpt.make_opts(seed=3))
    st = pt.stats()
    p = pt.profile_frame(cam, 128, 128, 16, 6, rpt.make_opts(seed=3))
    assert p["rays_closest"] == st["rays_closest"] and p["rays_shadow"] == st["rays_shadow"] and p["samples"] == 128 * 128 * 16
    for k in ("raygen", "trace", "shade", "shadow", "accumulate"):
        assert p["ms"][k] > 0 and p["launches"][k] > 0, k
    assert p["launches"]["trace"] == p["launches"]["shade"] == 6
    assert 2 < p["trace_interior"] / p["rays_closest"] < 200 and p["trace_tris"] > 0 and p["trace_roots"] >= p["rays_closest"]
    assert p["shadow_interior"] > 0
    again = np.zeros((128, 128, 3), np.float32)
    pt.render(cam, 16, 6, again, rpt.make_opts(seed=3))  # the handle is back to its ordinary two-lane mode
    assert np.array_equal(img, again)
