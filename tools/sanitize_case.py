"""Small renders through every kernel variant for compute-sanitizer (memcheck / racecheck): exits non-zero on a frame mismatch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
c, r = scenes.book_scene(11)
W, H, spp = 64, 40, 2
cam = scenes.book_camera(W, H)
frames = []
with rt.Scene(c, r) as sc:
    for kw in (dict(scan_mode=0), dict(scan_mode=0, cull_smem=True), dict(scan_mode=0, paths_per_lane=1), dict(scan_mode=1), dict(scan_mode=2),
               dict(scan_mode=0, chunks=2)):
        img, _, st = rt.render(sc, cam, rt.make_params(W, H, spp, 50, seed=1, early_out=False, **kw))
        frames.append(img)
    acc, rgba, _ = rt.render_pass(sc, cam, rt.make_params(W, H, 1, 50, seed=1, early_out=False), 0)
    acc, rgba, _ = rt.render_pass(sc, cam, rt.make_params(W, H, 1, 50, seed=1, early_out=False), 1, acc)
    frames.append(rgba)
    idx, t = rt.primary_hits(sc, cam, W, H)
    rng = np.random.default_rng(0)
    rt.hit(sc, rng.normal(size=(500, 3)) * 5, rng.normal(size=(500, 3)))
    rt.ray_color(sc, np.tile([[13.0, 2.0, 3.0]], (200, 1)), rng.normal(size=(200, 3)) - [13, 2, 3], 50)
ok = all(np.array_equal(f, frames[0]) for f in frames)
print("frames equal across variants:", ok)
sys.exit(0 if ok else 1)
