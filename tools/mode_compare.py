"""Linear cull scan vs BVH traversal on the same scenes (kernel-time Msamples/s)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
def run(name, c, r, cam, W, H, spp, **kw):
    with rt.Scene(c, r) as sc:
        p = rt.make_params(W, H, spp, 50, seed=1, **kw)
        rt.render(sc, cam, p)
        _, _, st = rt.render(sc, cam, p)
    print(json.dumps(dict(name=name, n=len(r), kw=kw, ms=round(st["kernel_ms"], 2), msamples_s=round(st["samples"] / st["kernel_ms"] / 1e3, 1),
                          node_tests_per_cast=round(st["node_tests"] / max(st["casts"], 1), 1), exact_per_cast=round(st["exact_tests"] / max(st["casts"], 1), 2))), flush=True)
for g in (3, 6, 11, 22, 31):
    c, r = scenes.book_scene(g)
    cam = scenes.book_camera(1200, 800)
    for eo in (False, True):
        for mode in (0, 2):
            if mode == 0 and len(r) > 4080: continue
            run(f"book{g}", c, r, cam, 1200, 800, 16, early_out=eo, scan_mode=mode)
dc, dr = scenes.default_scene()
for mode in (0, 2):
    run("default", dc, dr, rt.Camera.default(), 400, 225, 100, early_out=False, scan_mode=mode)
