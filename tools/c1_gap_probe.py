"""Where does the gap between the torch-event step time and the kernel time of the 8 ms C1 frame come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes
c, r = scenes.default_scene()
cam = rt.Camera.default()
W, H, spp = 400, 225, 100
frame = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
with rt.Scene(c, r) as sc:
    for mode, name in ((0, "scan"), (3, "auto")):
        for do_flush in (True, False):
            p = rt.make_params(W, H, spp, 50, seed=0, early_out=False, scan_mode=mode)
            rt.render_device(sc, cam, p, frame.data_ptr(), 0, stream); rt.render_finish(sc)
            gaps = []
            for _ in range(6):
                if do_flush:
                    flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rt.render_device(sc, cam, p, frame.data_ptr(), 0, stream)
                e1.record(); e1.synchronize()
                st = rt.render_finish(sc)
                gaps.append((round(e0.elapsed_time(e1), 3), round(st["kernel_ms"], 3)))
            print(name, "flush" if do_flush else "no flush", gaps, flush=True)
