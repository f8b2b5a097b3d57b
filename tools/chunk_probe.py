"""Work-unit grading probe (one GPU): kernel time of several frames under different chunk gradings (rt_params.reserved[1]:
0 = library default, -1 = ungraded, -(10 L + F) = L short levels holding F/4 long units of work per warp, n > 0 = n equal
chunks).  "shard" = rank 0 of an 8-way tile deal rendered alone, against 1/8 of the full-frame time: the launch's ramp-down
is what the 8-GPU scaling loses."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes

CODES = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,-1,-11,-12,-14,-21,-22,-24".split(","))]
book = scenes.book_scene(11)
cases = [("c3 shard 1/8 x500", 1200, 800, 500, book, scenes.book_camera, dict(shard_rank=0, shard_count=8), 2),
         ("c3 x500", 1200, 800, 500, book, scenes.book_camera, {}, 1),
         ("c3 x64", 1200, 800, 64, book, scenes.book_camera, {}, 3),
         ("c3 x8", 1200, 800, 8, book, scenes.book_camera, {}, 5),
         ("c1 x100", 400, 225, 100, scenes.default_scene(), None, {}, 5)]
for tag, W, H, spp, scene, cam_fn, kw, reps in cases:
    cam = cam_fn(W, H) if cam_fn else rt.Camera.default()
    buf = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
    with rt.Scene(*scene) as sc:
        for mode, name in ((0, "scan"), (3, "auto")):
            out = {}
            for code in CODES:
                if tag == "c3 x500" and code not in (0, -1):
                    continue
                p = rt.make_params(W, H, spp, 50, seed=0, early_out=False, scan_mode=mode, chunks=code, **kw)
                best = 1e30
                for _ in range(reps):
                    rt.render_device(sc, cam, p, buf.data_ptr())
                    best = min(best, rt.render_finish(sc)["kernel_ms"])
                out[str(code)] = round(best, 3)
            print(json.dumps(dict(frame=tag, mode=name, kernel_ms=out)), flush=True)
