"""Host cost of a scene (re-)upload: rt_upload_scene / rt_update_scene (flatten arrays + BVH build or refit + tie grid +
copies) for the 485-sphere and the 99 856-sphere book scenes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
print("host threads:", os.cpu_count())
for grid in (11, 158):
    c, r = scenes.book_scene(grid)
    ts = []
    for it in range(4):
        t0 = time.perf_counter(); sc = rt.Scene(c, r); t1 = time.perf_counter(); sc.close()
        ts.append(1e3 * (t1 - t0))
    sc = rt.Scene(c, r)
    tu, tf = [], []
    for it in range(4):
        t0 = time.perf_counter(); sc.update(c, r, refit=False); t1 = time.perf_counter(); tu.append(1e3 * (t1 - t0))
        t0 = time.perf_counter(); sc.update(c, r, refit=True); t1 = time.perf_counter(); tf.append(1e3 * (t1 - t0))
    sc.close()
    print(f"{len(r)} spheres: upload {min(ts):.1f} ms, update (rebuild) {min(tu):.1f} ms, update (refit) {min(tf):.1f} ms")
