"""Where does the host-buffer path (rt_upload_scene + rt_render) spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
c, r = scenes.book_scene(11); cam = scenes.book_camera(1200, 800)
p = rt.make_params(1200, 800, spp, 50, seed=0, early_out=False)
for it in range(4):
    t0 = time.perf_counter(); sc = rt.Scene(c, r); t1 = time.perf_counter()
    rgba, _, st = rt.render(sc, cam, p); t2 = time.perf_counter()
    sc.close(); t3 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.2f} ms, render call {1e3*(t2-t1):.2f} ms (kernel {st['kernel_ms']:.2f} ms), free {1e3*(t3-t2):.2f} ms")
sc = rt.Scene(c, r)
for it in range(3):
    t1 = time.perf_counter(); rgba, _, st = rt.render(sc, cam, p); t2 = time.perf_counter()
    print(f"reuse {it}: render call {1e3*(t2-t1):.2f} ms (kernel {st['kernel_ms']:.2f} ms)")
