"""One short render of the c3 scene for ncu (launch 0 = warm-up, launch 1 = the one to capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eo = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1200, 800)
ppl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
smem = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 0
grid = int(sys.argv[8]) if len(sys.argv) > 8 else 11   # 11 -> 485 spheres, 158 -> 99 856 (BASELINE config 4)
c, r = scenes.book_scene(grid)
cam = scenes.book_camera(W, H)
with rt.Scene(c, r) as sc:
    p = rt.make_params(W, H, spp, 50, seed=1, early_out=eo, paths_per_lane=ppl, cull_smem=smem, scan_mode=mode)
    for _ in range(2):
        _, _, st = rt.render(sc, cam, p)
print(st)
