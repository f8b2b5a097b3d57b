"""BVH traversal kernel: kernel-time Msamples/s for paths-per-lane 1/2/4 on the 485-sphere and 100k-sphere scenes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes

def run(name, c, r, cam, W, H, spp, **kw):
    with rt.Scene(c, r) as sc:
        p = rt.make_params(W, H, spp, 50, seed=1, scan_mode=2, **kw)
        rt.render(sc, cam, p)
        _, _, st = rt.render(sc, cam, p)
    print(json.dumps(dict(name=name, n=len(r), kw=kw, ms=round(st["kernel_ms"], 2), msamples_s=round(st["samples"] / st["kernel_ms"] / 1e3, 1),
                          node_tests_per_cast=round(st["node_tests"] / max(st["casts"], 1), 1),
                          exact_per_cast=round(st["exact_tests"] / max(st["casts"], 1), 2), overflows=st["overflows"])), flush=True)

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c3, r3 = scenes.book_scene(11)
c4, r4 = scenes.book_scene(158)
for ppl in (1, 2, 4):
    for eo in (False, True):
        run("book11", c3, r3, scenes.book_camera(1200, 800), 1200, 800, spp, early_out=eo, paths_per_lane=ppl)
        run("book158", c4, r4, scenes.book_camera(1920, 1080), 1920, 1080, max(spp // 2, 4), early_out=eo, paths_per_lane=ppl)
