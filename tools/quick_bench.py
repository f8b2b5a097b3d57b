"""Ad-hoc timing helper (not the contract bench): kernel-time Msamples/s for a few settings."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes

def run(name, c, r, cam, W, H, spp, **kw):
    with rt.Scene(c, r) as sc:
        kw.setdefault("scan_mode", 0)
        p = rt.make_params(W, H, spp, 50, seed=1, **kw)
        rt.render(sc, cam, p)  # warm-up
        _, _, st = rt.render(sc, cam, p)
    ms = st["kernel_ms"]
    n = len(r)
    out = dict(name=name, kw=kw, W=W, H=H, spp=spp, ms=round(ms, 3), msamples_s=round(st["samples"] / ms / 1e3, 2),
               casts_per_sample=round(st["casts"] / st["samples"], 3), exact_per_cast=round(st["exact_tests"] / st["casts"], 3),
               overflows=st["overflows"], gtests_s=round(st["casts"] * n / ms / 1e6, 2))
    print(json.dumps(out), flush=True)
    return out

if __name__ == "__main__":
    print(rt.device_info(0))
    peak, ms = rt.measure_fp32_peak(0)
    print(json.dumps(dict(ffma_per_s=peak, ms=ms, nominal=148 * 128 * 1.965e9)))
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    c, r = scenes.book_scene(11)
    cam = scenes.book_camera(1200, 800)
    for smem in (False, True):
        for ppl in (1, 2, 4):
            for eo in (False,):
                o = run("c3", c, r, cam, 1200, 800, spp, early_out=eo, paths_per_lane=ppl, cull_smem=smem)
                print("   frac_of_peak(11 slots/test) =", round(o["gtests_s"] * 1e9 * 11 / peak, 4))
    dc, dr = scenes.default_scene()
    run("c1", dc, dr, rt.Camera.default(), 400, 225, 100, early_out=False)
    run("c1", dc, dr, rt.Camera.default(), 400, 225, 100, early_out=True)
    if "--c4" in sys.argv:
        c4, r4 = scenes.book_scene(158)
        cam4 = scenes.book_camera(1920, 1080)
        for eo in (False, True):
            with rt.Scene(c4, r4) as sc:
                p = rt.make_params(1920, 1080, 8, 50, seed=1, early_out=eo)
                rt.render(sc, cam4, p)
                _, _, st = rt.render(sc, cam4, p)
            print(json.dumps(dict(name="c4_bvh", early_out=eo, n=len(r4), ms=round(st["kernel_ms"], 2),
                                  msamples_s=round(st["samples"] / st["kernel_ms"] / 1e3, 2),
                                  casts_per_sample=round(st["casts"] / st["samples"], 3),
                                  node_tests_per_cast=round(st["node_tests"] / st["casts"], 2),
                                  exact_per_cast=round(st["exact_tests"] / st["casts"], 3), overflows=st["overflows"])), flush=True)
