"""One render of the default scene (BASELINE config 1: 400x225, 100 spp, 2 spheres) for ncu; argv: mode ppl spp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ppl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 100
c, r = scenes.default_scene()
with rt.Scene(c, r) as sc:
    p = rt.make_params(400, 225, spp, 50, seed=1, early_out=False, scan_mode=mode, paths_per_lane=ppl)
    for _ in range(2):
        _, _, st = rt.render(sc, rt.Camera.default(), p)
print(st)
