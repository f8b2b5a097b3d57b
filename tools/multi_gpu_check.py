"""torchrun --nproc-per-node N tools/multi_gpu_check.py : N-GPU sharded frame == 1-GPU frame, byte for byte."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import petershirleyraytracer_b200 as rt
from petershirleyraytracer_b200 import scenes, dist as rdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); torch.cuda.set_device(local)
c, r = scenes.book_scene(11); W, H, spp = 1200, 800, 8
cam = scenes.book_camera(W, H); p = rt.make_params(W, H, spp, 50, seed=3)
with rt.Scene(c, r, device=local) as sc:
    frame = rdist.render_sharded(sc, cam, p, rank, world)
    torch.cuda.synchronize()
    full, _, _ = rt.render(sc, cam, p)
same = np.array_equal(frame.cpu().numpy(), full)
t = torch.tensor([int(same)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print(f"world={world}: sharded frame == single-GPU frame on every rank: {bool(t.item())}")
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
