"""Generates the shading-parameter fixtures (tmin / albedo / sky colours / Lambertian scatter) from the
reference classes compiled here (oracle/_ref/libref.so: ref_ray_color_param_batch / ref_render_rows_param, i.e.
ray_color of programs/main.cc:34-49 with its literals as arguments, every operation done by the reference's own
vec3 / hittable code; with the reference's constants it equals the unmodified ray_color bit for bit, see
tests/test_oracle_vs_ref.py).

    make -C oracle ref && python tests/golden/make_golden_shading.py

Separate from make_golden.py so that the round-1 fixtures stay byte-identical.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import oracle_lib as ol  # noqa: E402
from make_golden import edge_rays  # noqa: E402
from petershirleyraytracer_b200 import scenes  # noqa: E402

# name -> keyword arguments of oracle_lib.shading()
CONFIGS = {
    "tmin_book": dict(tmin=0.001),
    "albedo_0.25": dict(albedo=0.25),
    "albedo_0.8": dict(albedo=0.8),
    "sunset_sky": dict(sky_a=(1.0, 0.6, 0.3), sky_b=(0.1, 0.2, 0.55)),
    "lambertian": dict(scatter_mode=ol.SCATTER_LAMBERTIAN),
    "book_next_chapter": dict(tmin=0.001, albedo=0.7, scatter_mode=ol.SCATTER_LAMBERTIAN, sky_b=(0.4, 0.6, 0.9)),
}


def pack(kw):
    sh = ol.shading(**kw)
    return np.array([sh.tmin, sh.albedo, *sh.sky_a, *sh.sky_b, float(sh.scatter_mode)])


def main():
    assert ol.have_ref(), "build oracle/_ref/libref.so first (make -C oracle ref)"
    rng = np.random.default_rng(20261019)
    bc, br = scenes.book_scene(11)
    bcam = scenes.book_camera(300, 200).as12()
    org, d = edge_rays(bc, br, bcam, rng, 400)
    seeds = rng.integers(1, 2**63, size=400, dtype=np.uint64)
    out = dict(centres=bc, radii=br, org=org, dir=d, seeds=seeds, depth=50, names=np.array(sorted(CONFIGS)))
    for name in sorted(CONFIGS):
        rgb, _ = ol.ray_color_batch("ref", bc, br, org, d, seeds, 50, shading=ol.shading(**CONFIGS[name]))
        out["shading_" + name] = pack(CONFIGS[name])
        out["rgb_" + name] = rgb
    np.savez_compressed(os.path.join(HERE, "ref_ray_color_shading.npz"), **out)

    W, H, spp = 36, 24, 3
    cam = scenes.book_camera(W, H).as12()
    out = dict(centres=bc, radii=br, cam12=cam, W=W, H=H, spp=spp, max_depth=50, seed=5, names=np.array(sorted(CONFIGS)))
    for name in sorted(CONFIGS):
        rgb, _, st = ol.render("ref", bc, br, cam, W, H, spp, 50, seed=5, nthreads=0, shading=ol.shading(**CONFIGS[name]))
        out["shading_" + name] = pack(CONFIGS[name])
        out["rgb_" + name] = rgb
        out["stats_" + name] = np.array([st["samples"], st["casts"], st["black"]])
    np.savez_compressed(os.path.join(HERE, "ref_render_shading.npz"), **out)

    # a converged render in the book's next-chapter configuration (tmin 0.001, Lambertian, albedo 0.7): PSNR gate for
    # the non-default path of the CUDA renderer
    W, H, spp = 120, 80, 2048
    cam = scenes.book_camera(W, H).as12()
    kw = CONFIGS["book_next_chapter"]
    rgb, _, st = ol.render("ref", bc, br, cam, W, H, spp, 50, seed=13, shading=ol.shading(**kw))
    np.savez_compressed(os.path.join(HERE, "ref_converged_book_next_chapter_120x80.npz"), centres=bc, radii=br, cam12=cam, W=W,
                        H=H, spp=spp, max_depth=50, seed=13, rgb=rgb, shading=pack(kw),
                        stats=np.array([st["samples"], st["casts"], st["black"]]))
    print("done", st)


if __name__ == "__main__":
    main()
