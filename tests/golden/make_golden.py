"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled here (oracle/_ref/libref.so).

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py [--skip-slow]

Every fixture stores its inputs next to the reference's outputs, so the tests that consume them need
neither /root/reference nor libref.so.  Slow part: the two high-spp renders used for the PSNR gate
(~12 min on 8 cores).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import oracle_lib as ol  # noqa: E402
from petershirleyraytracer_b200 import scenes  # noqa: E402

MAIN_SEED = 0x9E3779B97F4A7C15


def edge_rays(centres, radii, cam12, rng, n):
    """Rays that stress the tmin=0 behaviour: origins ON spheres (first-hit points), inside spheres,
    tangent rays, rays pointing away, plus random ones."""
    k = rng.integers(0, len(radii), size=n)
    c, r = centres[k], radii[k][:, None]
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    d = rng.normal(size=(n, 3))
    org = np.empty((n, 3))
    kind = np.arange(n) % 5
    org[kind == 0] = (c + r * u)[kind == 0]                      # on the surface (rounded)
    org[kind == 1] = (c + 0.5 * r * u)[kind == 1]                # inside
    org[kind == 2] = (c + 3.0 * r * u + rng.normal(size=(n, 3)))[kind == 2]  # outside, random direction
    # tangent: origin at c + r*u + s*t with t perpendicular to u, direction = -t (grazes the sphere)
    t = np.cross(u, rng.normal(size=(n, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    m = kind == 3
    org[m] = (c + r * u + 2.0 * t)[m]
    d[m] = -t[m] * rng.uniform(0.5, 2.0, size=(n, 1))[m]
    # camera-like rays from the eye
    m = kind == 4
    org[m] = cam12[:3]
    tgt = centres[rng.integers(0, len(radii), size=n)] + rng.normal(scale=0.3, size=(n, 3))
    d[m] = (tgt - cam12[:3])[m]
    return org, d


def main():
    skip_slow = "--skip-slow" in sys.argv
    assert ol.have_ref(), "build oracle/_ref/libref.so first (make -C oracle ref)"
    rng = np.random.default_rng(20261018)
    dc, dr = scenes.default_scene()
    bc, br = scenes.book_scene(11)
    dcam, aspect = ol.default_camera("ref")
    bcam = scenes.book_camera(300, 200).as12()
    meta = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libref.so (unmodified reference + shim)"}

    # 1. the reference's own main()
    t0 = time.time()
    ppm = ol.main_ppm("ref", MAIN_SEED)
    meta["main_ppm"] = {"seed": MAIN_SEED, "md5": hashlib.md5(ppm).hexdigest(), "bytes": len(ppm),
                        "head": ppm[:64].decode(), "seconds": round(time.time() - t0, 1)}
    print("main ppm", meta["main_ppm"]["md5"])

    # 2. small renders through the reference classes (bitwise targets for the C restatement, RAND15 streams)
    rgb, _, st = ol.render("ref", dc, dr, dcam, 64, 36, 8, 50, seed=1)
    np.savez_compressed(os.path.join(HERE, "ref_render_default_64x36.npz"), centres=dc, radii=dr, cam12=dcam, W=64, H=36,
                        spp=8, max_depth=50, seed=1, rgb=rgb, stats=np.array([st["samples"], st["casts"], st["black"]]))
    rgb, _, st = ol.render("ref", bc, br, bcam, 48, 32, 4, 50, seed=2)
    np.savez_compressed(os.path.join(HERE, "ref_render_book_48x32.npz"), centres=bc, radii=br, cam12=bcam, W=48, H=32, spp=4,
                        max_depth=50, seed=2, rgb=rgb, stats=np.array([st["samples"], st["casts"], st["black"]]))
    rgb, _, st = ol.render("ref", bc, br, bcam, 40, 28, 3, 5, seed=3)  # shallow depth limit
    np.savez_compressed(os.path.join(HERE, "ref_render_book_depth5.npz"), centres=bc, radii=br, cam12=bcam, W=40, H=28, spp=3,
                        max_depth=5, seed=3, rgb=rgb, stats=np.array([st["samples"], st["casts"], st["black"]]))

    # 3. primary hits (config 2 = default scene 400x225; plus the ~485-sphere scene)
    idx, t = ol.primary_hits("ref", dc, dr, dcam, 400, 225)
    np.savez_compressed(os.path.join(HERE, "ref_primary_c2_400x225.npz"), centres=dc, radii=dr, cam12=dcam, W=400, H=225,
                        idx=idx.astype(np.int16), t=t)
    idx, t = ol.primary_hits("ref", bc, br, bcam, 240, 160)
    np.savez_compressed(os.path.join(HERE, "ref_primary_book_240x160.npz"), centres=bc, radii=br, cam12=bcam, W=240, H=160,
                        idx=idx.astype(np.int16), t=t)

    # 4. hittable_list::hit on explicit rays, including the tmin=0 edge cases
    org, d = edge_rays(bc, br, bcam, rng, 3000)
    idx, rec = ol.hit_batch("ref", bc, br, org, d)
    np.savez_compressed(os.path.join(HERE, "ref_hit_book.npz"), centres=bc, radii=br, org=org, dir=d, tmin=0.0, tmax=np.inf,
                        idx=idx, rec=rec)
    idx2, rec2 = ol.hit_batch("ref", bc, br, org, d, 0.001, 50.0)
    np.savez_compressed(os.path.join(HERE, "ref_hit_book_tmin_tmax.npz"), centres=bc, radii=br, org=org, dir=d, tmin=0.001,
                        tmax=50.0, idx=idx2, rec=rec2)
    # second-bounce rays: origin = reference hit point p (exactly the doubles the reference produced)
    m = idx >= 0
    org_b = rec[m, 1:4]
    d_b = rec[m, 4:7] + rng.uniform(-1, 1, size=(m.sum(), 3)) * 0.9
    idx3, rec3 = ol.hit_batch("ref", bc, br, org_b, d_b)
    np.savez_compressed(os.path.join(HERE, "ref_hit_book_bounce.npz"), centres=bc, radii=br, org=org_b, dir=d_b, tmin=0.0,
                        tmax=np.inf, idx=idx3, rec=rec3)

    # 5. sphere::hit pairs
    k = rng.integers(0, len(br), size=2000)
    org, d = edge_rays(bc, br, bcam, rng, 2000)
    hit, rec = ol.sphere_hit_batch("ref", bc[k], br[k], org, d)
    np.savez_compressed(os.path.join(HERE, "ref_sphere_hit.npz"), centre=bc[k], radius=br[k], org=org, dir=d, tmin=0.0,
                        tmax=np.inf, hit=hit, rec=rec)

    # 6. ray_color with the shim stream
    org, d = edge_rays(bc, br, bcam, rng, 600)
    seeds = rng.integers(1, 2**63, size=600, dtype=np.uint64)
    rgbc, _ = ol.ray_color_batch("ref", bc, br, org, d, seeds, 50)
    np.savez_compressed(os.path.join(HERE, "ref_ray_color_book.npz"), centres=bc, radii=br, org=org, dir=d, seeds=seeds, depth=50,
                        rgb=rgbc)
    org, d = edge_rays(dc, dr, dcam, rng, 600)
    rgbc, _ = ol.ray_color_batch("ref", dc, dr, org, d, seeds, 50)
    np.savez_compressed(os.path.join(HERE, "ref_ray_color_default.npz"), centres=dc, radii=dr, org=org, dir=d, seeds=seeds,
                        depth=50, rgb=rgbc)

    # 7. get_ray / write_color / random_in_hemisphere / default camera
    uv = rng.uniform(-0.1, 1.1, size=(500, 2))
    np.savez_compressed(os.path.join(HERE, "ref_get_ray.npz"), cam12=bcam, uv=uv, out=ol.get_ray_batch("ref", bcam, uv),
                        default_cam12=dcam, default_aspect=aspect)
    sums = np.concatenate([rng.uniform(0, 120, size=(400, 3)), np.array([[0, 0, 0], [100, 100, 100], [99.8, 99.9, 100.1],
                                                                        [1e-9, 50, 200], [-1, 25, 1e6]])])
    np.savez_compressed(os.path.join(HERE, "ref_write_color.npz"), sums=sums, spp=100, out=ol.write_color_batch("ref", sums, 100))
    nrm = rng.normal(size=(500, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    seeds = rng.integers(1, 2**63, size=500, dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "ref_random_in_hemisphere.npz"), normals=nrm, seeds=seeds,
                        out=ol.random_in_hemisphere_batch("ref", nrm, seeds))

    # 8. high-spp renders of the reference itself: the PSNR >= 40 dB gate compares the CUDA image to these
    if not skip_slow:
        t0 = time.time()
        W, H = 200, 112
        rgb, _, st = ol.render("ref", dc, dr, dcam, W, H, 8192, 50, seed=11)
        np.savez_compressed(os.path.join(HERE, "ref_converged_default_200x112.npz"), centres=dc, radii=dr, cam12=dcam, W=W, H=H,
                            spp=8192, max_depth=50, seed=11, rgb=rgb,
                            stats=np.array([st["samples"], st["casts"], st["black"]]))
        print("converged default", round(time.time() - t0, 1), "s")
        t0 = time.time()
        W, H = 120, 80
        cam = scenes.book_camera(W, H).as12()
        rgb, _, st = ol.render("ref", bc, br, cam, W, H, 4096, 50, seed=12)
        np.savez_compressed(os.path.join(HERE, "ref_converged_book_120x80.npz"), centres=bc, radii=br, cam12=cam, W=W, H=H,
                            spp=4096, max_depth=50, seed=12, rgb=rgb,
                            stats=np.array([st["samples"], st["casts"], st["black"]]))
        print("converged book", round(time.time() - t0, 1), "s")

    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()
