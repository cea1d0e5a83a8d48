"""GPU parity tests of the CUDA path tracer (through the C ABI) against the oracle.

Two kinds of gates:
 * per-pixel: the oracle is run with the SAME counter-based generator as the kernels, so both trace the same paths;
   differences come only from libm (powf/expf/sincos) and are bounded per pixel (tolerances stated in each test);
 * statistical: against the reference's own converged image (golden, 2048 spp), relative L2 within
   1.5 x 0.43 x sqrt(1/N + 1/M) and image means within 1 %.
"""
import ctypes as C
import os

import numpy as np
import pytest

import rendu_b200 as rb
from rendu_b200 import pathtracer as rpt
from rendu_b200 import workloads as wl
from rendu_b200._capi import LightDesc, check, lib, ptr
from oracle import oracle as orc
from oracle import ref
from util import GOLDEN

pytestmark = pytest.mark.gpu


def kat():
    return dict(np.load(os.path.join(GOLDEN, "shading_kat.npz")))


def golden_scene(name):
    desc = rpt.load_scene(os.path.join(GOLDEN, "scene_%s.npz" % name))
    g = dict(np.load(os.path.join(GOLDEN, "render_%s.npz" % name)))
    cam = dict(eye=g["cam_eye"], corner=g["cam_corner"], dx=g["cam_dx"], dy=g["cam_dy"])
    return desc, cam, g


def gpu_render(desc, cam, w, h, spp, depth, seed=0, components=3, image=None, **kw):
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    img = np.zeros((h, w, components), np.float32) if image is None else image
    pt.render(cam, spp, depth, img, rpt.make_opts(seed=seed, **kw))
    return img, pt


def assert_same_paths(gpu, cpu, frac_tol=0.01, abs_tol=2e-3):
    """Same RNG, same estimator: almost every pixel agrees to libm precision; a handful may differ where a 1-ulp
    difference flips a discrete decision (lobe pick, a hit at a grazing angle)."""
    assert np.isfinite(gpu).all()
    diff = np.abs(gpu - cpu).max(axis=2)
    bad = (diff > abs_tol).mean()
    assert bad <= frac_tol, "fraction of pixels off by > %g: %g (max %g)" % (abs_tol, bad, diff.max())
    assert abs(gpu.mean() - cpu.mean()) <= 2e-3 * max(cpu.mean(), 1e-3)


# ------------------------------------------------------------------------------------------------ scalar KATs
def test_kat_ggx_eval_and_sample():
    k = kat()
    n = k["wo"].shape[0]
    out = np.zeros((n, 3), np.float32)
    arrs = [np.ascontiguousarray(k[x], np.float32) for x in ("wo", "base", "rough", "metal", "wi")]
    check(lib().rb_kat_ggx_eval(*[ptr(a) for a in arrs], n, ptr(out)))
    np.testing.assert_allclose(out, k["ggx_eval"], rtol=2e-5, atol=1e-7)  # powf(1-VdotH,5), sqrtf
    wi, ev = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    u3 = np.ascontiguousarray(k["sample_u3"], np.float32)
    check(lib().rb_kat_ggx_sample(ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]), ptr(u3), n, ptr(wi), ptr(ev)))
    np.testing.assert_allclose(wi, k["sample_wi"], rtol=0, atol=2e-6)  # sincos / sqrtf
    same_lobe = np.isclose(ev, k["sample_eval"], rtol=2e-3, atol=1e-5).all(axis=1)
    assert same_lobe.mean() > 0.999


def test_kat_image_cubemap_lights_exact():
    k = kat()
    n = k["uv"].shape[0]
    img = np.ascontiguousarray(k["img"], np.float32)
    out4 = np.zeros((n, 4), np.float32)
    check(lib().rb_kat_image_sample(ptr(img), img.shape[1], img.shape[0], ptr(np.ascontiguousarray(k["uv"])), n, ptr(out4)))
    assert np.array_equal(out4.view(np.uint32), k["img_rgba"].view(np.uint32))  # no transcendental: bit-exact
    faces = np.ascontiguousarray(k["faces"], np.float32)
    out3 = np.zeros((n, 3), np.float32)
    check(lib().rb_kat_cubemap_sample(ptr(faces), faces.shape[2], faces.shape[1], ptr(np.ascontiguousarray(k["dirs"])), n, ptr(out3)))
    assert np.array_equal(out3.view(np.uint32), k["cube"].view(np.uint32))
    for name in [x[:-5] for x in k if x.endswith("_desc")]:
        a = k[name + "_desc"]
        d = LightDesc()
        d.kind, d.casts_shadow = int(a[0]), int(a[1])
        d.intensity, d.position, d.direction = [(C.c_float * 3)(*[float(np.float32(v)) for v in a[i:i + 3]]) for i in (2, 5, 8)]
        d.radius, d.inner_angle, d.outer_angle = [float(np.float32(v)) for v in a[11:14]]
        pos = np.ascontiguousarray(k[name + "_pos"], np.float32)
        out5 = np.zeros((pos.shape[0], 5), np.float32)
        check(lib().rb_kat_light_sample(C.byref(d), ptr(pos), pos.shape[0], ptr(out5)))
        assert np.array_equal(out5.view(np.uint32), k[name + "_out"].view(np.uint32)), name


def test_kat_sky_eval():
    rng = np.random.RandomState(8)
    lut = rng.uniform(0.2, 1.0, (64, 64, 4)).astype(np.float32)
    d = rng.normal(size=(512, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = rng.uniform(-2, 2, (512, 3)).astype(np.float32)
    sun = np.array([0.3, 0.8, 0.52], np.float32)
    sun /= np.linalg.norm(sun)
    want = orc.sky_eval(lut, o, d, sun)
    got = np.zeros_like(want)
    check(lib().rb_kat_sky_eval(ptr(lut), 64, 64, ptr(o), ptr(d), ptr(sun), 512, ptr(got)))
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6)  # 48 expf + powf per sample
    assert want.max() > 0


# ------------------------------------------------------------------------------------------------ renders
def test_kat_exact_division():
    """rb_div (slot -> pixel / sample in every path-tracer kernel) equals the u32 division for every operand: edge values
    (1, powers of two and their neighbours, 2^32 - 1, equal operands) against each other, and a million random pairs."""
    edge = np.array([1, 2, 3, 5, 7, 10, 255, 256, 257, 1080, 1920, 4095, 4096, 4097, 65535, 65536, 65537, 2**24 - 1, 2**24, 2**25, 2**31 - 1, 2**31,
                     2**31 + 1, 2**32 - 2, 2**32 - 1, 33554432, 2073600, 530841600 % 2**32], np.uint64)
    nn, dd = np.meshgrid(np.concatenate([edge, [0]]), edge)
    rng = np.random.RandomState(4)
    num = np.concatenate([nn.ravel(), rng.randint(0, 2**32, 1 << 20, dtype=np.uint64), rng.randint(0, 2**32, 1 << 18, dtype=np.uint64)]).astype(np.uint32)
    den = np.concatenate([dd.ravel(), rng.randint(1, 2**32, 1 << 20, dtype=np.uint64), rng.randint(1, 5000, 1 << 18, dtype=np.uint64)]).astype(np.uint32)
    out = np.zeros((num.size, 2), np.uint32)
    check(lib().rb_kat_div(ptr(num), ptr(den), num.size, ptr(out)))
    assert np.array_equal(out[:, 0], (num.astype(np.uint64) // den.astype(np.uint64)).astype(np.uint32))
    assert np.array_equal(out[:, 1], out[:, 0])


@pytest.mark.parametrize("name", ["cornellbox", "cornellbox-emissive"])
def test_render_matches_oracle_path_for_path(name):
    desc, cam, _ = golden_scene(name)
    gpu, pt = gpu_render(desc, cam, 64, 64, 64, 5, seed=9)
    cpu, rays = orc.OracleScene(desc).render(cam, 64, 64, 64, 5, seed=9, rng="counter", threads=8)
    assert_same_paths(gpu, cpu)
    st = pt.stats()
    # The CUDA path drops paths whose attenuation reached exactly zero, so it never casts more rays than the reference.
    assert 0 < st["rays_closest"] + st["rays_shadow"] <= rays
    assert st["samples"] == 64 * 64 * 64 and st["kernel_launches"] > 0


@pytest.mark.parametrize("name", ["cornellbox", "cornellbox-emissive"])
def test_render_converges_to_reference_image(name):
    desc, cam, g = golden_scene(name)
    n = 256
    gpu, _ = gpu_render(desc, cam, 64, 64, n, 5, seed=21)
    conv, m = g["converged"], int(g["converged_spp"])
    assert abs(gpu.mean() - conv.mean()) / conv.mean() < 0.01
    rel = np.linalg.norm(gpu - conv) / np.linalg.norm(conv)
    assert rel < 1.5 * 0.43 * np.sqrt(1.0 / n + 1.0 / m), rel


def test_synthetic_c3_scene_small(monkeypatch):
    """The C3/C4 benchmark scene at reduced size: textures with UVs, normal map, cubemap environment, emissive panel,
    shadowed point light, displaced-grid floor."""
    desc = wl.cornell_synthetic(24, 64, 16)
    cam = wl.cornell_camera(16 / 9)
    gpu, _ = gpu_render(desc, cam, 96, 54, 32, 6, seed=4)
    cpu, _ = orc.OracleScene(desc).render(cam, 96, 54, 32, 6, seed=4, rng="counter", threads=8)
    assert_same_paths(gpu, cpu, frac_tol=0.02)
    assert gpu.mean() > 0.05


@pytest.mark.skipif(not ref.available(), reason="bundled scenes need oracle/_ref/resources")
@pytest.mark.parametrize("name", ["cornellbox-textured", "cornellbox-cutout", "cornellbox-reflections", "boxes", "parameter-spheres", "gallery-materials"])
def test_bundled_scenes_match_oracle(name):
    """Cubemap + normal maps; masked leaves (alpha pass-through in render and in the shadow walk) + spot lights;
    atmosphere sky; 443 roots + 25 lights; roughness/metalness sweep under a cubemap with no light; 483,584 triangles
    in 52 meshes (51 transformed instances) with a directional light (FLT_MAX shadow distance) and all material types."""
    sc = ref.RefScene(name)
    desc, cam = sc.flatten(), sc.camera(1.0)
    osc = orc.OracleScene(desc)
    gpu, _ = gpu_render(desc, cam, 48, 48, 32, 5, seed=2)
    cpu, _ = osc.render(cam, 48, 48, 32, 5, seed=2, rng="counter", threads=8)
    # cornellbox-cutout places its two spot lights exactly in the planes of the side walls (x = -1 and x = +1, z = 1),
    # so from the second bounce on "does the shadow segment end before or on the wall" is decided by the last ulp of
    # the bounce direction (sinf/cosf differ by an ulp between CUDA and glibc). Measured: depth 1 is identical,
    # depth 5 differs in ~3 % of the pixels by one flipped sample while the image means agree to 5e-4.
    assert_same_paths(gpu, cpu, frac_tol=0.06 if name == "cornellbox-cutout" else 0.02)
    assert gpu.mean() > 0.01
    if name == "cornellbox-cutout":
        gpu1, _ = gpu_render(desc, cam, 48, 48, 32, 1, seed=2)
        cpu1, _ = osc.render(cam, 48, 48, 32, 1, seed=2, rng="counter", threads=8)
        assert_same_paths(gpu1, cpu1, frac_tol=0.001)  # masked leaves + spot cones + alpha shadow walk, first bounce
        # ... and the explanation itself, at full depth: with the two spot lights pulled 0.5 % towards the room's centre, off the
        # wall planes, the same scene (masked leaves, alpha walk, spot cones, five bounces) agrees like every other one.
        import copy
        moved = copy.deepcopy(desc)
        for l in moved["lights"]:
            l["position"] = (np.asarray(l["position"], np.float32) * np.float32(0.995) + np.float32(0.005) * np.array([0.0, 1.0, 0.0], np.float32)).astype(np.float32)
        gpu5, _ = gpu_render(moved, cam, 48, 48, 32, 5, seed=2)
        cpu5, _ = orc.OracleScene(moved).render(cam, 48, 48, 32, 5, seed=2, rng="counter", threads=8)
        assert_same_paths(gpu5, cpu5, frac_tol=0.02)


def test_seed_determinism_tiles_and_sample_split():
    desc, cam, _ = golden_scene("cornellbox")
    a, _ = gpu_render(desc, cam, 64, 40, 16, 5, seed=5)
    b, _ = gpu_render(desc, cam, 64, 40, 16, 5, seed=5, max_paths=4096)  # different batching, same image
    c, _ = gpu_render(desc, cam, 64, 40, 16, 5, seed=6)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert not np.array_equal(a, c)
    # Tile split: three ranks render interleaved bands of 8 rows into one image -> bit-identical to the full render.
    tiled = np.zeros_like(a)
    for r in range(3):
        gpu_render(desc, cam, 64, 40, 16, 5, seed=5, image=tiled, tile=(8, r, 3))
    assert np.array_equal(tiled.view(np.uint32), a.view(np.uint32))
    # Sample split: per-rank partial images (each normalised by the FULL sample count) add up to the full image.
    parts = [gpu_render(desc, cam, 64, 40, 16, 5, seed=5, samples=(r, 4))[0] for r in range(4)]
    np.testing.assert_allclose(sum(parts), a, rtol=0, atol=2e-6)


@pytest.mark.parametrize("w,h,spp,depth", [(1, 1, 1, 1), (3, 2, 2, 1), (5, 7, 1, 3), (2, 9, 8, 2), (33, 1, 4, 5)])
def test_tiny_frames_match_oracle(w, h, spp, depth):
    """Ragged / minimal inputs: single pixel, single sample (1x1 stratification grid), depth 1 (direct light only), odd
    sizes that do not fill a warp or a tile band."""
    desc, cam, _ = golden_scene("cornellbox")
    gpu, pt = gpu_render(desc, cam, w, h, spp, depth, seed=4)
    cpu, _ = orc.OracleScene(desc).render(cam, w, h, spp, depth, seed=4, rng="counter", threads=1)
    assert np.isfinite(gpu).all() and pt.stats()["samples"] == w * h * spp
    np.testing.assert_allclose(gpu, cpu, rtol=2e-3, atol=2e-3)


def test_empty_shares_and_zero_depth():
    """A rank whose tile share is empty (more ranks than 8-row bands) or whose sample share is empty (more ranks than
    samples) renders nothing and leaves the image as it was; depth 0 traces nothing, as the reference's bounce loop
    `for(did = 0; did < depth; ...)` (PathTracer.cpp:202) does."""
    desc, cam, _ = golden_scene("cornellbox")
    img = np.full((6, 16, 3), 0.25, np.float32)
    out, pt = gpu_render(desc, cam, 16, 6, 4, 3, seed=2, image=img.copy(), tile=(8, 1, 4))  # 6 rows: only band 0 exists
    assert np.array_equal(out, img) and pt.stats()["samples"] == 0
    out, pt = gpu_render(desc, cam, 16, 6, 2, 3, seed=2, image=img.copy(), samples=(3, 4))  # 2 samples, rank 3 of 4
    np.testing.assert_allclose(out, img / 2, rtol=0, atol=1e-7)  # (pixels + 0) / samples
    assert pt.stats()["rays_closest"] == 0
    out, pt = gpu_render(desc, cam, 16, 6, 4, 0, seed=2)
    assert (out == 0).all() and pt.stats()["rays_closest"] == 0


def test_image_contract_components_accumulation_and_rounding():
    desc, cam, _ = golden_scene("cornellbox-emissive")
    base, _ = gpu_render(desc, cam, 32, 32, 16, 4, seed=1)
    # 4-component image (PathTracerApp passes RGBA, Image.cpp:34-36): alpha is left alone.
    rgba = np.full((32, 32, 4), 0.0, np.float32)
    rgba[..., 3] = 7.0
    gpu_render(desc, cam, 32, 32, 16, 4, seed=1, components=4, image=rgba)
    assert np.array_equal(rgba[..., :3].view(np.uint32), base.view(np.uint32)) and (rgba[..., 3] == 7.0).all()
    # The result is accumulated into the caller's image and the SUM is divided (PathTracer.cpp:288-292).
    pre = np.full((32, 32, 3), 1.6, np.float32)
    gpu_render(desc, cam, 32, 32, 16, 4, seed=1, image=pre)
    np.testing.assert_allclose(pre, base + 1.6 / 16, rtol=0, atol=1e-6)
    # Non power-of-two sample counts are rounded like the reference: 12 -> 2^round(log2 12) = 16 (PathTracer.cpp:169-173).
    r12, pt = gpu_render(desc, cam, 32, 32, 12, 4, seed=1)
    assert np.array_equal(r12.view(np.uint32), base.view(np.uint32)) and pt.stats()["samples"] == 32 * 32 * 16
    with pytest.raises(rb.RenduError):
        gpu_render(desc, cam, 32, 32, 16, 4, components=2, image=np.zeros((32, 32, 2), np.float32))


def test_batches_and_lanes_do_not_change_the_image(monkeypatch):
    """The wavefront is cut into batches that alternate between stream lanes; every variate is keyed by (seed, pixel,
    cell, bounce, dimension) and batches own disjoint pixels, so lane count and batch size must not change one bit."""
    desc, cam, _ = golden_scene("cornellbox")
    w, h, spp, depth = 320, 256, 16, 5  # 1.3 M paths: enough for the lanes to engage
    imgs = []
    for lanes, max_paths in ((1, 0), (2, 0), (4, 0), (2, 200000), (1, 65536), (3, 300000)):
        monkeypatch.setenv("RB_PT_LANES", str(lanes))
        img, pt = gpu_render(desc, cam, w, h, spp, depth, seed=11, max_paths=max_paths)
        assert pt.stats()["samples"] == w * h * spp
        imgs.append(img)
    for img in imgs[1:]:
        assert np.array_equal(img.view(np.uint32), imgs[0].view(np.uint32))
    assert np.isfinite(imgs[0]).all() and imgs[0].mean() > 0.01
    # More samples per pixel than paths per batch: a pixel's samples are split over several batches, which then run on one
    # lane in sample order (same per-pixel summation order is not guaranteed across chunkings: compare to fp32 rounding).
    monkeypatch.setenv("RB_PT_LANES", "2")
    whole, _ = gpu_render(desc, cam, 12, 10, 2048, 4, seed=5)
    chunked, pt = gpu_render(desc, cam, 12, 10, 2048, 4, seed=5, max_paths=1024)
    assert pt.stats()["samples"] == 12 * 10 * 2048
    np.testing.assert_allclose(chunked, whole, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["cornellbox", "cornellbox-cutout"])
def test_cp_async_fed_wavefronts_return_the_same_image(monkeypatch, name):
    """RB_PT_FEED=1 selects the wavefront kernels that stage their rays through shared memory with cp.async (an
    experiment that did not pay, kept opt-in): per ray nothing changes, so the image must not change by one bit —
    with the any-hit shadow walk (cornellbox) and the masked one (cutout), small and ragged queue sizes included."""
    if name == "cornellbox-cutout" and not ref.available():
        pytest.skip("bundled scenes need oracle/_ref/resources")
    if name == "cornellbox":
        desc, cam, _ = golden_scene(name)
    else:
        rs = ref.RefScene(name)
        desc, cam = rs.flatten(), rs.camera(1.0)
    for w, h, spp, depth in ((200, 150, 8, 5), (7, 5, 2, 3), (1, 1, 1, 2)):
        monkeypatch.delenv("RB_PT_FEED", raising=False)
        plain, _ = gpu_render(desc, cam, w, h, spp, depth, seed=3)
        monkeypatch.setenv("RB_PT_FEED", "1")
        fed, pt = gpu_render(desc, cam, w, h, spp, depth, seed=3)
        assert pt.stats()["samples"] == w * h * spp
        assert np.array_equal(fed.view(np.uint32), plain.view(np.uint32))


@pytest.mark.parametrize("extra", [2, 5, 33, 57, 58, 250, 505])
def test_root_tree_returns_the_same_image_as_the_root_loop(monkeypatch, extra):
    """Scenes of 9..64 meshes find the roots a ray enters with one walk of a tree over the root boxes instead of one test
    per root (RB_ROOTS_TREE, rb_roots_tree_pretest) — closest-hit and shadow wavefronts. The set of entered roots and the
    order they are visited in are the reference's, so the image equals the one of the lazy root loop (RB_PT_LAZY_ROOTS=1)
    bit for bit: 7 + `extra` meshes = 9 (smallest tree), 12, 40 (both mask registers), 64 (largest register mask), 65, 257 and
    512 (the mask in the ray's stack column: 3, 9 and 16 words); overlapping boxes, meshes
    that share a position; also after a transform edit re-fits the roots."""
    import copy
    desc = wl.cornell_synthetic(24, 16, 16)
    cam = wl.cornell_camera(4 / 3)
    rng = np.random.RandomState(extra)
    proto = desc["objects"][0]
    for k in range(extra):
        o = copy.deepcopy(desc["objects"][1 + k % (len(desc["objects"]) - 1)])
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[1, 1] = m[2, 2] = np.float32(rng.uniform(0.08, 0.3))
        m[3, :3] = rng.uniform(-0.7, 0.7, 3).astype(np.float32) + np.array([0, 1.0, 0], np.float32)
        if k % 7 == 3:
            m[3, :3] = desc_shift  # two meshes at the same place
        desc_shift = m[3, :3].copy()
        o["model"] = m.reshape(16)
        desc["objects"].append(o)
    assert len(desc["objects"]) == 7 + extra and proto is desc["objects"][0]
    imgs = {}
    for mode in ("tree", "lazy"):
        if mode == "lazy":
            monkeypatch.setenv("RB_PT_LAZY_ROOTS", "1")
        else:
            monkeypatch.delenv("RB_PT_LAZY_ROOTS", raising=False)
        sc = rpt.Scene(desc)
        pt = rpt.PathTracer(sc)
        img = np.zeros((90, 120, 3), np.float32)
        pt.render(cam, 16, 6, img, rpt.make_opts(seed=5))
        moved = np.array(desc["objects"][-1]["model"], np.float32).copy()
        moved[12:15] += np.array([0.2, -0.1, 0.15], np.float32)
        sc.set_object_transform(len(desc["objects"]) - 1, moved)
        img2 = np.zeros((90, 120, 3), np.float32)
        pt.render(cam, 16, 6, img2, rpt.make_opts(seed=5))
        imgs[mode] = (img, img2)
        assert np.isfinite(img).all() and img.mean() > 0.01 and not np.array_equal(img, img2)
    assert np.array_equal(imgs["tree"][0].view(np.uint32), imgs["lazy"][0].view(np.uint32))
    assert np.array_equal(imgs["tree"][1].view(np.uint32), imgs["lazy"][1].view(np.uint32))


def test_lights_and_background_follow_edits_between_renders():
    """The reference reads scene->lights and the background through the shared Scene inside render(), so an edit between
    two frames shows in the next one. Here: rb_scene_clear_lights / add_light / set_background + rb_scene_sync on a
    committed scene give exactly the image of a scene built with the new values; geometry is not rebuilt."""
    import copy
    desc, cam, _ = golden_scene("cornellbox")
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    img0 = np.zeros((40, 48, 3), np.float32)
    pt.render(cam, 16, 4, img0, rpt.make_opts(seed=6))
    # brighter, moved light; another background colour
    d2 = copy.deepcopy(desc)
    for l in d2["lights"]:
        l["intensity"] = (np.asarray(l["intensity"], np.float32) * np.float32(2.5)).astype(np.float32)
        l["position"] = (np.asarray(l["position"], np.float32) + np.array([0.1, -0.05, 0.0], np.float32)).astype(np.float32)
    d2["background"]["color"] = np.array([0.2, 0.4, 0.9], np.float32)
    sc.set_lights(d2["lights"])
    sc.set_background(d2["background"])
    img1 = np.zeros_like(img0)
    pt.render(cam, 16, 4, img1, rpt.make_opts(seed=6))
    fresh, _ = gpu_render(d2, cam, 48, 40, 16, 4, seed=6)
    assert np.array_equal(img1.view(np.uint32), fresh.view(np.uint32)) and not np.array_equal(img1, img0)
    # no lights at all: next-event estimation and the shadow wavefront switch off
    sc.set_lights([])
    img2 = np.zeros_like(img0)
    pt.render(cam, 16, 4, img2, rpt.make_opts(seed=6))
    assert pt.stats()["rays_shadow"] == 0 and img2.mean() < img1.mean()
    # and back
    sc.set_lights(desc["lights"])
    sc.set_background(desc["background"])
    img3 = np.zeros_like(img0)
    pt.render(cam, 16, 4, img3, rpt.make_opts(seed=6))
    assert np.array_equal(img3.view(np.uint32), img0.view(np.uint32))


def test_object_transform_follows_edits_between_renders():
    """Incremental geometry: a new model matrix for one object of a committed scene re-transforms
    only that object's vertices and refits the hierarchy on the device. The image equals the one of a scene built with
    that matrix, except where the refitted tree (old topology) and the fresh tree break an exact-distance tie the other
    way; putting the matrix back reproduces the first image bit for bit (refit with the original vertices gives the
    original boxes)."""
    import copy
    desc, cam, _ = golden_scene("cornellbox")
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    img0 = np.zeros((40, 48, 3), np.float32)
    pt.render(cam, 16, 4, img0, rpt.make_opts(seed=9))
    k = 5  # one of the two boxes
    m0 = np.asarray(desc["objects"][k]["model"], np.float32).reshape(4, 4).copy()  # [col][row]
    ang = np.float32(0.4)
    rot = np.eye(4, dtype=np.float32)
    rot[0, 0], rot[0, 2], rot[2, 0], rot[2, 2] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    m1 = (m0.T @ rot.T).T.astype(np.float32)  # model * rotation, kept in [col][row] storage
    m1[3, :3] += np.array([0.05, 0.1, -0.05], np.float32)
    sc.set_object_transform(k, m1)
    img1 = np.zeros_like(img0)
    pt.render(cam, 16, 4, img1, rpt.make_opts(seed=9))
    d2 = copy.deepcopy(desc)
    d2["objects"][k]["model"] = m1
    fresh, _ = gpu_render(d2, cam, 48, 40, 16, 4, seed=9)
    same = (img1.view(np.uint32) == fresh.view(np.uint32)).all(axis=2).mean()
    assert same >= 0.995 and not np.array_equal(img1, img0), same
    assert abs(img1.mean() - fresh.mean()) <= 2e-3 * fresh.mean()
    # without the sync the edit waits for the next render, which performs it
    sc.set_object_transform(k, m0, sync=False)
    img2 = np.zeros_like(img0)
    pt.render(cam, 16, 4, img2, rpt.make_opts(seed=9))
    assert np.array_equal(img2.view(np.uint32), img0.view(np.uint32))
    with pytest.raises(rb.RenduError):
        sc.set_object_transform(len(desc["objects"]), m0)


def test_output_stage_on_device():
    """Fused tonemap + sRGB + 8-bit conversion kernel vs the reference's bytes. expf/powf may differ from glibc in the last
    ulp, which can move a value across a truncation boundary: at most one code value, in a vanishing fraction of bytes."""
    g = dict(np.load(os.path.join(GOLDEN, "output_kat.npz")))
    for img, want, tm in ((g["img"], g["tonemapped"], True), (np.clip(g["img"], 0, 1), g["clipped"], False), (g["rgba"], g["rgba_out"], True)):
        got = rpt.output_ldr(img, tonemap=tm, srgb=True)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3, (diff.max(), (diff > 0).mean())
    # whole pipeline: render -> bytes
    desc, cam, _ = golden_scene("cornellbox")
    img, _ = gpu_render(desc, cam, 48, 48, 16, 4, seed=3)
    ldr = rpt.output_ldr(img)
    want = orc.output_ldr(img, True, True)
    assert np.abs(ldr.astype(np.int32) - want.astype(np.int32)).max() <= 1 and ldr.max() > 100


def test_output_stage_hdr_on_device():
    """Layer split + the EXR writer's float -> half rule on the GPU: integer work, so bit-exact against the oracle and
    against what the reference stored (golden)."""
    g = dict(np.load(os.path.join(GOLDEN, "output_hdr_kat.npz")))
    for img, stored in ((g["img"], g["stored"]), (g["rgba"], g["rgba_stored"])):
        planes = rpt.output_hdr(img)
        assert np.array_equal(planes, orc.output_hdr(img))
        assert np.array_equal(orc.half_planes_to_image(planes, img.shape[2]).view(np.uint32), stored.view(np.uint32))
        assert np.array_equal(rpt.output_hdr(img, flip=True), orc.output_hdr(img, flip=True))
        assert np.array_equal(rpt.output_hdr(img, ignore_alpha=False), orc.output_hdr(img, ignore_alpha=False))
    # every finite float exponent/rounding class: sweep 1M random bit patterns
    rng = np.random.RandomState(4)
    bits = rng.randint(0, 2 ** 32, size=(512, 683, 3), dtype=np.uint64).astype(np.uint32)
    img = bits.view(np.float32)
    assert np.array_equal(rpt.output_hdr(img), orc.output_hdr(img))
    with pytest.raises(rb.RenduError):
        rpt.output_hdr(np.zeros((8, 8, 2), np.float32))


def test_progressive_accumulation_on_device():
    """sums stay on the device; four progressive passes over disjoint slices of the stratification grid
    add up to the one-shot render (fp32 summation order aside)."""
    import torch
    desc, cam, _ = golden_scene("cornellbox")
    pt = rpt.PathTracer(rpt.Scene(desc))
    w, h, spp = 64, 48, 16
    d_sum = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    d_img = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    for k in range(4):
        pt.render_device(cam, w, h, spp, 5, d_sum.data_ptr(), rpt.make_opts(seed=5, samples=(k, 4), accumulate=(k > 0)), sh)
    pt.resolve_device(d_sum.data_ptr(), w, h, spp, 3, d_img.data_ptr(), None, sh)
    torch.cuda.synchronize()
    full, _ = gpu_render(desc, cam, w, h, spp, 5, seed=5)
    np.testing.assert_allclose(d_img.cpu().numpy(), full, rtol=0, atol=2e-6)


def test_scene_changes_after_the_commit_are_refused_until_recommitted():
    """a texture added after the commit lives on the host only (its id would index past the device pool), and
    a rebuild of the scene's raycaster reorders the triangles under the per-triangle shading records: both must make the
    next render fail cleanly; a new commit repairs both."""
    desc, cam = wl.tiny_scene(), wl.tiny_camera()
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    a = np.zeros((16, 16, 3), np.float32)
    pt.render(cam, 4, 3, a, rpt.make_opts(seed=1))
    tex = np.full((4, 4, 4), 0.5, np.float32)
    out = C.c_int32(-1)
    check(lib().rb_scene_add_texture(sc._h, ptr(tex), 4, 4, 4, C.byref(out)))
    with pytest.raises(rb.RenduError):
        pt.render(cam, 4, 3, np.zeros((16, 16, 3), np.float32), rpt.make_opts(seed=1))
    with pytest.raises(rb.RenduError):
        check(lib().rb_scene_sync(sc._h))
    check(lib().rb_scene_commit(sc._h, rb.BUILD_REFERENCE))
    b = np.zeros((16, 16, 3), np.float32)
    pt.render(cam, 4, 3, b, rpt.make_opts(seed=1))
    assert np.array_equal(a, b)
    sc.raycaster().updateHierarchy(rb.BUILD_LBVH)  # behind the scene's back
    with pytest.raises(rb.RenduError):
        pt.render(cam, 4, 3, np.zeros((16, 16, 3), np.float32), rpt.make_opts(seed=1))
    check(lib().rb_scene_commit(sc._h, rb.BUILD_REFERENCE))
    c = np.zeros((16, 16, 3), np.float32)
    pt.render(cam, 4, 3, c, rpt.make_opts(seed=1))
    assert np.array_equal(a, c)


def test_live_session_refines_on_the_device_and_ends_at_the_one_shot_frame():
    """rb_pathtracer_live_* (PathTracerApp's live render, PathTracerApp.cpp:57-98,127-136): passes add slices of the
    stratification grid to sums that stay on the device; what comes back is 8-bit pixels. After all passes the bytes equal the
    output stage applied to one render() of the same samples and seed; earlier displays are valid previews (mean over the
    cells so far); a moved camera starts the frame over; extra passes on a finished frame change nothing."""
    desc, cam = wl.cornell_synthetic(32, 64, 32), wl.cornell_camera(96 / 64)
    pt = rpt.PathTracer(rpt.Scene(desc))
    one = np.zeros((64, 96, 3), np.float32)
    pt.render(cam, 64, 4, one, rpt.make_opts(seed=9))
    want = rpt.output_ldr(np.concatenate([one, np.ones((64, 96, 1), np.float32)], axis=2))
    pt.live_begin(96, 64)
    previews = []
    for p in range(8):
        pt.live_pass(cam, 64, 4, 8, rpt.make_opts(seed=9))
        img, done = pt.live_display(4)
        assert done == p + 1
        previews.append(img)
    # the passes add the same clamped samples in another order: fp32 sums may differ in the last bit, a byte by one step
    d = np.abs(previews[-1].astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.005
    err = [np.abs(x[..., :3].astype(np.int32) - want[..., :3].astype(np.int32)).mean() for x in previews]
    assert err[0] > err[3] > err[7] and err[7] < 0.01  # the preview converges to the final frame
    pt.live_pass(cam, 64, 4, 8, rpt.make_opts(seed=9))  # finished: a no-op
    img, done = pt.live_display(3)
    assert done == 8 and np.array_equal(img, previews[-1][..., :3])
    cam2 = wl.cornell_camera(96 / 64)
    cam2["eye"] = cam2["eye"] + np.float32(0.05)
    pt.live_pass(cam2, 64, 4, 8, rpt.make_opts(seed=9))
    img2, done = pt.live_display(4)
    assert done == 1 and not np.array_equal(img2, previews[0])
    with pytest.raises(rb.RenduError):
        pt.live_pass(cam, 64, 4, 128)


def test_scene_sync_rebuilds_instead_of_refitting_when_the_tree_cost_doubles(monkeypatch):
    """rb_raycaster_tree_cost + the policy in rb_scene_sync ('nothing decides refit-vs-rebuild'): a small
    motion re-fits (cost ratio stays near 1, the hierarchy keeps its build), a large one — objects thrown far apart from where
    the tree put them — makes the re-fitted boxes more than twice as large and the scene is committed again: the cost is back
    at the baseline of a fresh build and the frame equals that of a scene created with the new transforms."""
    desc, cam = wl.cornell_synthetic(32, 64, 32), wl.cornell_camera(1.0)
    sc = rpt.Scene(desc)
    pt = rpt.PathTracer(sc)
    c0 = sc.raycaster().tree_cost()
    assert c0["at_build"] > 0 and c0["ratio"] == 1.0

    def moved_model(dx, dy, dz, s=1.0):
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[1, 1] = m[2, 2] = s
        m[3, :3] = [dx, dy, dz]
        return m.reshape(16)

    sc.set_object_transform(5, moved_model(0.02, 0.0, 0.01))  # the metal panel, nudged: a refit
    c1 = sc.raycaster().tree_cost()
    assert c1["at_build"] == c0["at_build"] and 0.9 < c1["ratio"] < 1.5
    monkeypatch.setenv("RB_REBUILD_RATIO", "0")  # policy off: the refit has to live with the stretched boxes
    sc.set_object_transform(0, moved_model(0.0, 0.0, 0.0, 6.0))  # the 2k-triangle floor blown up six-fold
    stretched = sc.raycaster().tree_cost()
    assert stretched["ratio"] > 2.0
    a = np.zeros((48, 48, 3), np.float32)
    pt.render(cam, 16, 4, a, rpt.make_opts(seed=3))
    monkeypatch.delenv("RB_REBUILD_RATIO")
    sc.set_object_transform(0, moved_model(0.0, 0.0, 0.0, 6.000001))  # any further edit: the sync now rebuilds
    rebuilt = sc.raycaster().tree_cost()
    assert rebuilt["ratio"] == 1.0 and rebuilt["at_build"] != c0["at_build"]
    b = np.zeros((48, 48, 3), np.float32)
    pt.render(cam, 16, 4, b, rpt.make_opts(seed=3))
    fresh_desc = wl.cornell_synthetic(32, 64, 32)
    fresh_desc["objects"][0]["model"] = moved_model(0.0, 0.0, 0.0, 6.000001)
    fresh_desc["objects"][5]["model"] = moved_model(0.02, 0.0, 0.01)
    c = np.zeros((48, 48, 3), np.float32)
    rpt.PathTracer(rpt.Scene(fresh_desc)).render(cam, 16, 4, c, rpt.make_opts(seed=3))
    assert np.array_equal(b, c)  # the rebuilt scene IS the fresh scene
    assert np.isfinite(a).all() and np.abs(a - b).mean() < 0.05  # and the re-fitted tree rendered (almost) the same frame, only slower
