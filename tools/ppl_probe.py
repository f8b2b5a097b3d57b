import sys, os, json
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
dc, dr = scenes.default_scene()
c3, r3 = scenes.book_scene(3)
for name, c, r, cam, W, H, spp in (("c1", dc, dr, rt.Camera.default(), 400, 225, 100), ("book3", c3, r3, scenes.book_camera(1200, 800), 1200, 800, 16)):
    for mode in (0, 2):
        for ppl in (1, 2, 4):
            if mode == 2 and ppl != 1: continue
            with rt.Scene(c, r) as sc:
                p = rt.make_params(W, H, spp, 50, seed=1, early_out=False, scan_mode=mode, paths_per_lane=ppl)
                rt.render(sc, cam, p); _, _, st = rt.render(sc, cam, p)
            print(name, "mode", mode, "ppl", ppl, round(st["samples"] / st["kernel_ms"] / 1e3, 1), "Msamples/s", round(st["kernel_ms"], 2), "ms", flush=True)
