"""BVH mode A/B (kernel-time Msamples/s): wavefront kernel with / without the tie-grid fast path vs the round-1
one-path-per-lane traversal kernel, on the 485-sphere (C3) and 99 856-sphere (C4) scenes, reference semantics and
exact early-out, plus tmin = 0.001 (no self hits: every cast traverses)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes

VARIANTS = {"wave": 0, "wave_no_tie": 3, "r1_kernel": 2}


def run(name, c, r, cam, W, H, spp, variants, **kw):
    with rt.Scene(c, r) as sc:
        for vn in variants:
            p = rt.make_params(W, H, spp, 50, seed=1, scan_mode=2, variant=VARIANTS[vn], **kw)
            rt.render(sc, cam, p)
            best = None
            for _ in range(3):
                _, _, st = rt.render(sc, cam, p)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            st = best
            print(json.dumps(dict(name=name, variant=vn, n=len(r), kw=kw, spp=spp, ms=round(st["kernel_ms"], 2),
                                  msamples_s=round(st["samples"] / st["kernel_ms"] / 1e3, 1),
                                  casts_per_sample=round(st["casts"] / st["samples"], 2),
                                  self_resolved_frac=round(st["self_resolved"] / max(st["casts"], 1), 4),
                                  node_visits_per_cast=round(st["node_tests"] / 4 / max(st["casts"], 1), 2),
                                  exact_per_cast=round(st["exact_tests"] / max(st["casts"], 1), 2), overflows=st["overflows"])), flush=True)


if __name__ == "__main__":
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    c3, r3 = scenes.book_scene(11)
    c4, r4 = scenes.book_scene(158)
    for kw in (dict(early_out=False), dict(early_out=True), dict(early_out=False, tmin=0.001)):
        run("c3", c3, r3, scenes.book_camera(1200, 800), 1200, 800, spp, which, **kw)
        run("c4", c4, r4, scenes.book_camera(1920, 1080), 1920, 1080, max(spp // 2, 4), which, **kw)
    dc, dr = scenes.default_scene()
    run("c1", dc, dr, rt.Camera.default(), 400, 225, 100, which, early_out=False)
