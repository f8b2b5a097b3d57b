"""Live cross-check of the C restatement against the reference compiled here (oracle/_ref/libref.so).
Skipped where libref.so is absent; tests/test_oracle_golden.py covers the same ground from fixtures."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libref.so not built")


def _rays(centres, radii, rng, n):
    k = rng.integers(0, len(radii), size=n)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    scale = rng.choice([0.0, 0.5, 1.0, 1.0, 3.0], size=(n, 1))
    org = centres[k] + scale * radii[k][:, None] * u
    d = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1.0, 1.0, 50.0], size=(n, 1))
    return org, d


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_hit_batch_random(book, seed):
    c, r = book
    rng = np.random.default_rng(seed)
    org, d = _rays(c, r, rng, 1500)
    for tmin, tmax in ((0.0, np.inf), (0.001, np.inf), (0.0, 7.5), (-1.0, 2.0)):
        ia, ra = ol.hit_batch("ref", c, r, org, d, tmin, tmax)
        ib, rb = ol.hit_batch("orc", c, r, org, d, tmin, tmax)
        assert np.array_equal(ia, ib)
        assert np.array_equal(ra.view(np.uint64), rb.view(np.uint64))


def test_tie_goes_to_later_object():
    """Duplicate spheres: hittable_list.cc:11-15 keeps the LAST of equal-t hits."""
    c = np.array([[0, 0, -3.0], [0, 0, -3.0], [5, 5, 5.0], [0, 0, -3.0]])
    r = np.array([1.0, 1.0, 0.5, 1.0])
    org = np.zeros((4, 3))
    d = np.array([[0, 0, -1.0], [0.1, 0, -1.0], [0, 0.2, -1.0], [0, 0, 1.0]])
    for which in ("ref", "orc"):
        idx, _ = ol.hit_batch(which, c, r, org, d)
        assert list(idx) == [3, 3, 3, -1]


def test_ray_color_random(book, default_scene):
    rng = np.random.default_rng(7)
    for c, r in (book, default_scene):
        org, d = _rays(c, r, rng, 300)
        seeds = rng.integers(1, 2**63, size=300, dtype=np.uint64)
        for depth in (0, 1, 50):
            a, _ = ol.ray_color_batch("ref", c, r, org, d, seeds, depth)
            b, _ = ol.ray_color_batch("orc", c, r, org, d, seeds, depth)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_render_rows_random_small(book):
    c, r = book
    from petershirleyraytracer_b200 import scenes
    cam = scenes.book_camera(36, 24).as12()
    a, _, sa = ol.render("ref", c, r, cam, 36, 24, 3, 50, seed=99, nthreads=2)
    b, _, sb = ol.render("orc", c, r, cam, 36, 24, 3, 50, seed=99, nthreads=3)
    assert np.array_equal(a, b)
    assert sa["casts"] == sb["casts"] and sa["black"] == sb["black"]


def test_write_color_edges():
    sums = np.array([[0.0, 1e-300, 1.0], [100.0, 99.99, 100.01], [np.inf, 50.0, 25.0], [400.0, 1.0, 4.0]])
    assert np.array_equal(ol.write_color_batch("ref", sums, 100), ol.write_color_batch("orc", sums, 100))


# ------------------------------------------------------------------ parameterised shading
SHADINGS = {
    "tmin_book": dict(tmin=0.001),
    "albedo_0.25": dict(albedo=0.25),                      # a power of two: products stay exact
    "albedo_0.8": dict(albedo=0.8),                        # every bounce rounds
    "sunset_sky": dict(sky_a=(1.0, 0.6, 0.3), sky_b=(0.1, 0.2, 0.55)),
    "lambertian": dict(scatter_mode=ol.SCATTER_LAMBERTIAN),
    "book_next_chapter": dict(tmin=0.001, albedo=0.7, scatter_mode=ol.SCATTER_LAMBERTIAN, sky_b=(0.4, 0.6, 0.9)),
}


def test_param_harness_with_reference_constants_is_the_reference(book, default_scene):
    """ray_color_param (the harness's parameterised ray_color over the reference classes) with main.cc's own
    constants == the reference's ray_color, bit for bit: the parameterised pin is anchored on the unmodified code."""
    rng = np.random.default_rng(17)
    for c, r in (book, default_scene):
        org, d = _rays(c, r, rng, 300)
        seeds = rng.integers(1, 2**63, size=300, dtype=np.uint64)
        for depth in (0, 3, 50):
            a, _ = ol.ray_color_batch("ref", c, r, org, d, seeds, depth)
            b, _ = ol.ray_color_batch("ref", c, r, org, d, seeds, depth, shading=ol.shading())
            o, _ = ol.ray_color_batch("orc", c, r, org, d, seeds, depth, shading=ol.shading())
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
            assert np.array_equal(a.view(np.uint64), o.view(np.uint64))


@pytest.mark.parametrize("name", sorted(SHADINGS))
def test_ray_color_with_shading_parameters(book, default_scene, name):
    """tmin / albedo / sky colours / Lambertian scatter: the C restatement against the reference classes."""
    sh = ol.shading(**SHADINGS[name])
    rng = np.random.default_rng(len(name))
    for c, r in (book, default_scene):
        org, d = _rays(c, r, rng, 250)
        seeds = rng.integers(1, 2**63, size=250, dtype=np.uint64)
        for depth in (1, 50):
            a, _ = ol.ray_color_batch("ref", c, r, org, d, seeds, depth, shading=sh)
            b, _ = ol.ray_color_batch("orc", c, r, org, d, seeds, depth, shading=sh)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (name, depth)


@pytest.mark.parametrize("name", ["tmin_book", "albedo_0.8", "book_next_chapter"])
def test_render_rows_with_shading_parameters(book, name):
    from petershirleyraytracer_b200 import scenes
    c, r = book
    sh = ol.shading(**SHADINGS[name])
    cam = scenes.book_camera(30, 20).as12()
    a, _, sa = ol.render("ref", c, r, cam, 30, 20, 3, 50, seed=5, nthreads=2, shading=sh)
    b, _, sb = ol.render("orc", c, r, cam, 30, 20, 3, 50, seed=5, nthreads=3, shading=sh)
    assert np.array_equal(a, b)
    assert sa["casts"] == sb["casts"] and sa["black"] == sb["black"]
