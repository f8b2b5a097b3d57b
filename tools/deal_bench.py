"""N-GPU deal comparison (run under torchrun): tiles dealt round-robin + all-gather vs samples split + integer all-reduce.
Device-timed (CUDA events, max over ranks), c3 / c5 workloads, linear scan and AUTO mode; prints one JSON line per leg
with the per-rank kernel times (slowest / mean = the load balance)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

def main():
    import petershirleyraytracer_b200 as rt
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c3"]
    for name in names:
        wl = bench.workload(name)
        for deal in (["tiles", "samples"] if world > 1 else ["tiles"]):
            b = bench.GpuBench(wl, rt, torch, dist, world, rank, local, deal=deal)
            for mode_name, mode in (("scan", rt.SCAN_FILTERED), ("auto", rt.SCAN_AUTO)):
                warm = b.params(mode, False, spp=max(world, wl["spp"] // 32))
                leg = b.timed(b.params(mode, False), 2 if name != "c5" else 1, 1, flush, warm_p=warm)
                k = torch.tensor([leg["kernel_ms_this_rank"] / leg["steps"]], dtype=torch.float64, device=dev)
                ks = [torch.zeros_like(k) for _ in range(world)]
                if world > 1:
                    dist.all_gather(ks, k)
                else:
                    ks = [k]
                ks = [x.item() for x in ks]
                if rank == 0:
                    print(json.dumps(dict(workload=name, gpus=world, deal=deal, mode=mode_name, msamples_s=round(leg["value"], 1),
                                          ms_per_step=round(leg["ms"] / leg["steps"], 3), kernel_ms_slowest=round(max(ks), 3),
                                          kernel_ms_mean=round(sum(ks) / len(ks), 3), imbalance=round(max(ks) / (sum(ks) / len(ks)) - 1, 5))), flush=True)
            b.close()
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
