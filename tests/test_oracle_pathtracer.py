"""CPU tests (no GPU): the path-tracer half of the C oracle against golden vectors generated from the real reference,
and against the live reference (oracle/_ref) when present. The strongest pin: with the reference's own mt19937 stream
and a single forParallel worker, the oracle's image equals the reference's image BIT FOR BIT."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from rendu_b200 import pathtracer as rpt
from rendu_b200 import workloads as wl
from util import GOLDEN


def kat():
    return dict(np.load(os.path.join(GOLDEN, "shading_kat.npz")))


def light_from_desc(a):
    return dict(kind=int(a[0]), casts_shadow=bool(a[1]), intensity=a[2:5].astype(np.float32), position=a[5:8].astype(np.float32),
                direction=a[8:11].astype(np.float32), radius=float(np.float32(a[11])), inner=float(np.float32(a[12])), outer=float(np.float32(a[13])))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_random_stream_matches_reference():
    k = kat()
    assert np.array_equal(bits(orc.mt_floats(42, 256)), bits(k["mt_floats"]))  # Random::seed(42); Random::Float() x256


def test_ggx_eval_and_sample_bit_exact():
    k = kat()
    ev = orc.ggx_eval(k["wo"], k["base"], k["rough"], k["metal"], k["wi"])
    assert np.array_equal(bits(ev), bits(k["ggx_eval"]))
    wi, se = orc.ggx_sample(k["wo"], k["base"], k["rough"], k["metal"], k["sample_u3"])
    assert np.array_equal(bits(wi), bits(k["sample_wi"]))
    assert np.array_equal(bits(se), bits(k["sample_eval"]))
    assert (se.sum(1) > 0).mean() > 0.5


def test_image_cubemap_lights_bit_exact():
    k = kat()
    assert np.array_equal(bits(orc.image_rgbal(k["img"], k["uv"])), bits(k["img_rgba"]))
    assert np.array_equal(bits(orc.cubemap(k["faces"], k["dirs"])), bits(k["cube"]))
    names = [n[:-5] for n in k if n.endswith("_desc")]
    assert len(names) == 4  # point (cornellbox), 2 spots (cutout), directional (gallery-materials)
    kinds = set()
    for n in names:
        l = light_from_desc(k[n + "_desc"])
        kinds.add(l["kind"])
        out = orc.light_sample(l, k[n + "_pos"])
        assert np.array_equal(bits(out), bits(k[n + "_out"])), n
    assert kinds == {0, 1, 2}


@pytest.mark.parametrize("name", ["cornellbox", "cornellbox-emissive"])
def test_render_bit_exact_against_golden_reference_image(name):
    desc = rpt.load_scene(os.path.join(GOLDEN, "scene_%s.npz" % name))
    g = dict(np.load(os.path.join(GOLDEN, "render_%s.npz" % name)))
    cam = dict(eye=g["cam_eye"], corner=g["cam_corner"], dx=g["cam_dx"], dy=g["cam_dy"])
    sc = orc.OracleScene(desc)
    img, rays = sc.render(cam, 48, 48, 16, 5, seed=int(g["mt_seed"]), rng="mt")
    assert rays == int(g["mt_rays"])
    assert np.array_equal(bits(img), bits(g["mt_image"]))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["cornellbox-textured", "cornellbox-cutout", "cornellbox-reflections", "boxes"])
def test_render_bit_exact_against_live_reference(name):
    """Cubemap + normal maps, masked leaves + spot lights, atmosphere sky, 443 roots + 25 lights."""
    sc = ref.RefScene(name)
    desc = sc.flatten()
    cam = sc.camera(1.0)
    try:
        ref.set_concurrency(1)
        ref.seed(3)
        img_ref, _, rays_ref = sc.render(24, 24, 8, 5)
    finally:
        ref.set_concurrency(0)
    img, rays = orc.OracleScene(desc).render(cam, 24, 24, 8, 5, seed=3, rng="mt")
    assert rays == rays_ref
    assert np.array_equal(bits(img), bits(img_ref))
    assert img.mean() > 0.01


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_sky_eval_bit_exact_against_live_reference():
    rng = np.random.RandomState(4)
    d = rng.normal(size=(256, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = rng.uniform(-2, 2, (256, 3)).astype(np.float32)
    sun = np.array([0.3, 0.8, 0.52], np.float32)
    sun /= np.linalg.norm(sun)
    want = ref.sky_eval(o, d, sun)
    got = orc.sky_eval(ref.sky_lut(), o, d, sun)
    assert np.array_equal(bits(got), bits(want))
    assert want.max() > 0


def test_counter_rng_mode_is_thread_count_independent_and_unbiased():
    desc = rpt.load_scene(os.path.join(GOLDEN, "scene_cornellbox.npz"))
    g = dict(np.load(os.path.join(GOLDEN, "render_cornellbox.npz")))
    cam = dict(eye=g["cam_eye"], corner=g["cam_corner"], dx=g["cam_dx"], dy=g["cam_dy"])
    sc = orc.OracleScene(desc)
    a, _ = sc.render(cam, 64, 64, 64, 5, seed=9, rng="counter", threads=1)
    b, _ = sc.render(cam, 64, 64, 64, 5, seed=9, rng="counter", threads=6)
    assert np.array_equal(bits(a), bits(b))
    # Against the reference's converged image (2048 spp): same estimator => same mean; error at the noise floor
    # sigma_rel ~ 0.43/sqrt(spp) measured on this scene.
    conv = g["converged"]
    assert abs(a.mean() - conv.mean()) / conv.mean() < 0.01
    rel = np.linalg.norm(a - conv) / np.linalg.norm(conv)
    assert rel < 1.5 * 0.43 * np.sqrt(1 / 64 + 1 / 2048)
    # Sample split: the two halves of the stratification grid add up to the full image (fp32 summation order aside).
    h0, _ = sc.render(cam, 64, 64, 64, 5, seed=9, rng="counter", threads=4, sample_split=(0, 2))
    h1, _ = sc.render(cam, 64, 64, 64, 5, seed=9, rng="counter", threads=4, sample_split=(1, 2))
    assert np.allclose(h0 + h1, a, rtol=0, atol=2e-6)


def test_synthetic_scene_builders():
    d = wl.cornell_synthetic(16, 32, 8)
    assert len(d["objects"]) == 7 and d["objects"][-1]["material_type"] == 8 and d["background"]["mode"] == 2
    sc = orc.OracleScene(d)
    img, rays = sc.render(wl.cornell_camera(16 / 9), 32, 18, 4, 4, seed=1, threads=4)
    assert np.isfinite(img).all() and img.mean() > 0.05 and rays > 32 * 18 * 4


def test_output_stage_bytes_match_reference_png():
    """Tonemap + sRGB + quantisation (main.cpp:109-119, Image.cpp:138-152): the oracle's bytes equal what the reference's
    own Image::save wrote into a PNG (golden; re-checked live when oracle/_ref is present)."""
    g = dict(np.load(os.path.join(GOLDEN, "output_kat.npz")))
    assert np.array_equal(orc.output_ldr(g["img"], True, True), g["tonemapped"])
    assert np.array_equal(orc.output_ldr(np.clip(g["img"], 0, 1), False, True), g["clipped"])
    assert np.array_equal(orc.output_ldr(g["rgba"], True, True), g["rgba_out"])
    assert (g["rgba_out"][..., 3] == 255).all()
    if ref.available():
        rng = np.random.RandomState(2)
        img = rng.uniform(0, 4, (17, 23, 3)).astype(np.float32)
        assert np.array_equal(orc.output_ldr(img, True, True), ref.output_ldr(img, True))


def test_output_stage_hdr_matches_reference_exr():
    """.exr branch (Image.cpp:160-289 + tinyexr float_to_half_full): the oracle's half layers decode to exactly the floats
    the reference's own loader reads back from the EXR its own writer produced (golden; live when oracle/_ref is there).
    The vectors hold the corner cases of that conversion: subnormal floats, the half subnormal range, ties (round half up),
    overflow to infinity, infinities, negative values."""
    g = dict(np.load(os.path.join(GOLDEN, "output_hdr_kat.npz")))
    for img, stored in ((g["img"], g["stored"]), (g["rgba"], g["rgba_stored"])):
        planes = orc.output_hdr(img, ignore_alpha=True)
        assert planes.shape == (img.shape[2],) + img.shape[:2]
        back = orc.half_planes_to_image(planes, img.shape[2])
        assert np.array_equal(back.view(np.uint32), stored.view(np.uint32))
    assert (g["rgba_stored"][..., 3] == 1.0).all()  # IGNORE_ALPHA
    # round half UP, unlike IEEE round-to-nearest-even: 1 + 2^-11 (a tie) goes up to 1 + 2^-10
    tie = np.full((16, 16, 3), 1.00048828125, np.float32)
    assert (orc.half_planes_to_image(orc.output_hdr(tie), 3) == np.float32(1.0009765625)).all()
    assert (tie.astype(np.float16).astype(np.float32) == 1.0).all()
    # flip reverses rows
    assert np.array_equal(orc.output_hdr(g["img"], flip=True), orc.output_hdr(g["img"][::-1]))
    if ref.available():
        rng = np.random.RandomState(8)
        img = np.exp(rng.uniform(-20, 12, (19, 33, 3))).astype(np.float32)
        assert np.array_equal(orc.half_planes_to_image(orc.output_hdr(img), 3).view(np.uint32), ref.output_hdr(img).view(np.uint32))
