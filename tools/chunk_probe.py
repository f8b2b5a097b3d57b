"""Per-rank kernel time of an 8-way tile deal, emulated on ONE GPU (rank r of 8 renders its shard alone), against 1/8 of the
full-frame kernel time, for several sample-chunk counts per tile (work-unit sizes): the launch's ramp-down is what the
8-GPU scaling loses."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes

W, H, spp = 1200, 800, 500
c, r = scenes.book_scene(11)
cam = scenes.book_camera(W, H)
buf = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
with rt.Scene(c, r) as sc:
    for mode, name in ((0, "scan"), (3, "auto")):
        def run(**kw):
            p = rt.make_params(W, H, spp, 50, seed=0, early_out=False, scan_mode=mode, **kw)
            best = 1e30
            for _ in range(2):
                rt.render_device(sc, cam, p, buf.data_ptr())
                best = min(best, rt.render_finish(sc)["kernel_ms"])
            return best
        full = run()
        print(json.dumps(dict(mode=name, full_frame_ms=round(full, 2), ideal_shard_ms=round(full / 8, 3))), flush=True)
        for chunks in (0, 32, 62, 125, 250):
            ms = [run(shard_rank=rk, shard_count=8, chunks=chunks) for rk in (0, 5)]
            print(json.dumps(dict(mode=name, chunks=chunks or "auto", shard_ms=[round(x, 3) for x in ms],
                                  efficiency=round(full / 8 / max(ms), 4))), flush=True)
