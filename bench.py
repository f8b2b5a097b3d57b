#!/usr/bin/env python
"""bench.py -- Msamples/s of the path-tracing hot path on N B200s (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3] [--spp S]

A "step" is one full frame of the workload (BASELINE.json configs[2]: 1200x800 random-spheres scene, 485
spheres, 500 spp, max depth 50) traced by the CUDA path; for N > 1 the frame's 8x8 tiles are sharded
round-robin over the ranks, all-gathered with NCCL and de-interleaved (strong scaling).

One JSON line on stdout (rank 0):
  value        device-timed throughput, scene resident in HBM, frame left in HBM
  e2e          same metric through the host-buffer C ABI (rt_update_scene + rt_render): H2D of the scene (+ BVH build)
               and D2H of the RGBA frame into pinned memory inside the timed region
  roofline     FP32-FMA roofline of the render kernel: algorithmic work = sphere tests x 11 FP32-pipe
               instructions, peak = FFMA issue rate measured in this run
  cpu_baseline the reference's own CPU code (oracle/_ref) or its C restatement, timed on this host on a
               bounded sample of the same workload
`--impl reference` times that CPU implementation alone (rank 0), all host threads, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

SLOTS_PER_TEST = 11  # 3 FADD + 2 FMUL + 6 FFMA of programs/sphere.cc:6-14 with A hoisted, r^2 precomputed


def workload(name: str, spp_override: int | None):
    from petershirleyraytracer_b200 import scenes
    key = {"c1": "c1_default", "c3": "c3_book_1200x800", "c4": "c4_bvh_1920x1080", "c5": "c5_book_4k"}.get(name, name)
    scene_fn, cam_fn, W, H, spp, depth = scenes.CONFIGS[key]
    if spp_override:
        spp = spp_override
    c, r = scene_fn()
    return dict(name=key, centres=c, radii=r, cam=cam_fn(W, H), W=W, H=H, spp=spp, depth=depth)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (recipe: B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.rows, self.proc = device, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        load = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(wl, target_s: float, threads: int = 0):
    """Msamples/s of the reference CPU implementation on a bounded sample of the workload: full-width rows
    from the middle of the frame at the workload's depth; spp and row count sized for ~target_s seconds."""
    import oracle_lib as ol
    kind = "reference" if ol.have_ref() else "port"
    which = "ref" if kind == "reference" else "orc"
    # all the host cores this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1,
    # which would time the reference on one thread
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = avail if threads == 0 else threads
    cam12 = wl["cam"].as12()
    W, H = wl["W"], wl["H"]
    mid = H // 2

    def run(rows, spp):
        j0, j1 = max(0, mid - rows // 2), min(H, mid - rows // 2 + rows)
        t0 = time.perf_counter()
        _, _, st = ol.render(which, wl["centres"], wl["radii"], cam12, W, H, spp, wl["depth"], seed=7, j0=j0, j1=j1,
                             nthreads=cores)
        return st["samples"], time.perf_counter() - t0, (j0, j1)

    n, dt, _ = run(max(cores, 8), 1)                      # probe
    rate = n / max(dt, 1e-6)
    want = rate * target_s
    rows = int(min(H, max(cores, want / W)))
    spp = int(max(1, min(wl["spp"], want / (rows * W))))
    n, dt, (j0, j1) = run(rows, spp)
    if dt < 0.6 * target_s and spp < wl["spp"]:   # the probe under-estimated the rate (thread start-up): one longer sample
        spp = int(max(spp + 1, min(wl["spp"], spp * target_s / max(dt, 1e-3))))
        n, dt, (j0, j1) = run(rows, spp)
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": f"rows {j0}..{j1 - 1} of {H} x {W} px x {spp} spp, depth {wl['depth']} ({int(n)} samples, {dt:.1f} s)"}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    per_step = float(os.environ.get("RT_BENCH_REF_SECONDS", per_step))   # (tests shorten the sample)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(wl, per_step)
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    last = vals[-1]
    out = {"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(wl, args, extra={"note": "reference CPU renderer (unmodified classes, OpenMP over rows); "
                                                          "each step is a bounded sample of the workload"}),
           "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
           "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def config_dict(wl, args, extra=None):
    d = {"workload": f"{wl['name']}: {wl['W']}x{wl['H']} px, {len(wl['radii'])} spheres (book layout, seed 42), "
                     f"{wl['spp']} spp, max depth {wl['depth']}, tmin 0 (reference semantics)",
         "width": wl["W"], "height": wl["H"], "spp": wl["spp"], "max_depth": wl["depth"], "spheres": int(len(wl["radii"])),
         "parallelism": f"tiles8x8 round-robin over {args.gpus} GPU(s)" + (" + NCCL all-gather" if args.gpus > 1 else ""),
         "l2": "scene (8 KB cull array in the constant bank, 15 KB FP64 array) is cache resident by design; a 256 MB buffer is "
               "written between timed steps to flush L2"}
    if extra:
        d.update(extra)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--spp", type=int, default=None, help="override the workload's spp (quick checks; not the contract config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = workload(args.workload, args.spp)

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import petershirleyraytracer_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rt.lib()  # fail loudly if the CUDA library is missing

    W, H, spp, depth = wl["W"], wl["H"], wl["spp"], wl["depth"]
    cam = wl["cam"]
    scene = rt.Scene(wl["centres"], wl["radii"], device=local)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    frame = torch.empty(H * W * 4, dtype=torch.uint8, device=dev)

    # scenes beyond the linear scan's 4080 spheres (BASELINE config 4) are benchmarked through the BVH; their roofline
    # counts the executed box tests (14 FP32-pipe slots each: 6 FFMA + 8 min/max/compare) and sphere tests (11)
    big = len(wl["radii"]) > 4080
    SLOTS_PER_BOX = 14

    def make(early_out, scan_mode=None):
        if scan_mode is None:
            scan_mode = rt.SCAN_BVH if big else rt.SCAN_FILTERED
        # headline = the linear cull-scan kernel BASELINE.json's north_star specifies for this config (its metric,
        # % of the FP32-FMA roofline, is defined on the scan); the library's default AUTO mode (exact BVH
        # traversal) is timed too, device-side and end to end, and reported as "auto_mode"
        return rt.make_params(W, H, spp, depth, seed=0, early_out=early_out, scan_mode=scan_mode, shard_rank=rank,
                              shard_count=world)

    layout = rt.tile_layout(make(False))
    shard = torch.empty(layout.shard_bytes, dtype=torch.uint8, device=dev) if world > 1 else None
    gathered = torch.empty(world * layout.shard_bytes, dtype=torch.uint8, device=dev) if world > 1 else None

    def step(p):
        """One frame, device buffers only.  Returns the kernels launched."""
        if world == 1:
            rt.render_device(scene, cam, p, frame.data_ptr(), 0, stream)
            return 1
        rt.render_device(scene, cam, p, shard.data_ptr(), 0, stream)
        dist.all_gather_into_tensor(gathered, shard)
        rt.deinterleave(p, gathered.data_ptr(), frame.data_ptr(), local, stream)
        return 2

    def timed(p, nsteps, nwarm, sampler=None):
        for _ in range(nwarm):
            step(p); rt.render_finish(scene)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if sampler:
            sampler.start()
        tot_ms, kern_ms, launches, stats = 0.0, 0.0, 0, None
        for _ in range(nsteps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the timed region)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launches += step(p)
            e1.record()
            e1.synchronize()
            stats = rt.render_finish(scene)
            tot_ms += e0.elapsed_time(e1)
            kern_ms += stats["kernel_ms"]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([tot_ms, kern_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), launches, stats

    # ---- headline: reference semantics, every cast executed
    sampler = ClockSampler(local) if rank == 0 else None
    tot_ms, kern_ms, launches, stats = timed(make(False), args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    samples_per_step = W * H * spp
    value = samples_per_step * args.steps / (tot_ms * 1e-3) / 1e6

    # per-rank counters -> whole-job sums for the roofline
    slots = (stats["node_tests"] * SLOTS_PER_BOX + stats["exact_tests"] * SLOTS_PER_TEST) if big else stats["sphere_tests"] * SLOTS_PER_TEST
    cnt = torch.tensor([slots / SLOTS_PER_TEST, stats["casts"], stats["samples"], stats["exact_tests"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    tests_per_step, casts_per_step = cnt[0].item(), cnt[1].item()

    # ---- same frame with the exact early-out (bit-identical image, fewer casts)
    eo_ms, eo_kern_ms, _, eo_stats = timed(make(True), max(1, args.steps), 1)
    value_eo = samples_per_step * max(1, args.steps) / (eo_ms * 1e-3) / 1e6

    # ---- the library's AUTO mode on the same frame (exact BVH traversal for this scene size)
    auto_ms, _, _, auto_stats = timed(make(False, rt.SCAN_AUTO), max(1, args.steps), 1)
    value_auto = samples_per_step * max(1, args.steps) / (auto_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the public C ABI (upload scene, render, read the frame back)
    e2e = None
    h2d = wl["centres"].nbytes + wl["radii"].nbytes + 96 + 64
    d2h = H * W * 4
    host_frame = torch.empty(H * W * 4, dtype=torch.uint8).pin_memory()
    host_view = host_frame.numpy().reshape(H, W, 4)

    # One device scene for all e2e steps: every step re-uploads the flattened hittable_list into it (rt_update_scene
    # with a full rebuild: H2D of the sphere arrays + BVH build, no cudaMalloc / cudaFree, whose latency on shared
    # hosts is erratic -- 100-600 ms stalls were seen inside cudaFree) and reads the frame back into pinned memory.
    e2e_scene = rt.Scene(wl["centres"], wl["radii"], device=local)

    def e2e_step(p):
        t_a = time.perf_counter()
        sc = e2e_scene
        sc.update(wl["centres"], wl["radii"], refit=False)           # H2D of the flattened hittable_list + BVH rebuild
        t_b = time.perf_counter()
        if world == 1:
            rgba, _, st = rt.render(sc, cam, p, out=host_view)        # kernel + D2H into the pinned host frame
            if os.environ.get("RT_BENCH_DEBUG"):
                print(f"[e2e] upload {1e3 * (t_b - t_a):.1f} ms, render call {1e3 * (time.perf_counter() - t_b):.1f} ms "
                      f"(kernel {st['kernel_ms']:.1f} ms)", file=sys.stderr)
        else:
            rt.render_device(sc, cam, p, shard.data_ptr(), 0, stream)
            dist.all_gather_into_tensor(gathered, shard)
            rt.deinterleave(p, gathered.data_ptr(), frame.data_ptr(), local, stream)
            if rank == 0:
                host_frame.copy_(frame, non_blocking=False)
            rt.render_finish(sc)

    def e2e_rate(p):
        for _ in range(2):  # untimed: first-use allocator / module initialisation
            e2e_step(p)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        per_step = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            e2e_step(p)
            per_step.append(1e3 * (time.perf_counter() - ts))
            if os.environ.get("RT_BENCH_DEBUG"):
                print(f"[e2e] step {per_step[-1]:.1f} ms", file=sys.stderr)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        # value: all K steps over the whole wall time (max over ranks); the per-step list (rank 0) shows host hiccups
        return samples_per_step * args.steps / e2e_s.item() / 1e6, [round(x, 1) for x in per_step]

    e2e_v, e2e_steps = e2e_rate(make(False))
    e2e = {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step_rank0": e2e_steps}
    e2e_auto, e2e_auto_steps = e2e_rate(make(False, rt.SCAN_AUTO))

    if rank == 0:
        fma_per_s, _ = rt.measure_fp32_peak(local)
        kernel_s = kern_ms * 1e-3 / args.steps                       # average launch duration of the render kernel
        # tests are whole-job sums; with N ranks the kernels run concurrently, so per-GPU achieved = sum / N
        achieved = tests_per_step / world * SLOTS_PER_TEST * 2 / kernel_s / 1e12
        peak = fma_per_s * 2 / 1e12
        traffic = None
        tpath = os.path.join(REPO, "profiles", "render_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "note": "FP32-FMA roofline (no tensor cores, HBM traffic ~nil): achieved = sphere tests x 11 FP32-pipe "
                            "instructions x 2 flop / render-kernel time (CUDA events, avg over the timed launches, per GPU); "
                            "peak = FFMA rate measured in this run by rt_measure_fp32_peak (MEASURED_PEAKS.json has no FP32 "
                            f"figure; nominal 148 SMs x 128 lanes x 1.965 GHz x 2 = 74.4); the kernel itself issues "
                            "7 FMA-pipe + ~1.9 other instructions per test in the scan loop",
                    "sphere_tests_per_step": tests_per_step,   # (BVH workloads: executed box + sphere test slots / 11) "casts_per_sample": casts_per_step / samples_per_step,
                    "kernel_ms_per_step": kern_ms / args.steps,
                    "with_early_out": {"msamples_s": value_eo, "casts_per_sample": None if eo_stats is None else
                                       eo_stats["casts"] * world / samples_per_step}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not big:   # (the reference's O(N) scan of 1e5 spheres: ~1 ms per cast)
            cpu = cpu_reference_rate(wl, 12.0)
        out = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64 hit/shading + f32 cull", "data": "synthetic",
               "config": config_dict(wl, args, extra={"early_out": False, "paths_per_lane": 1 if big else 2,
                                                      "scan_mode": "4-wide BVH (RT_SCAN_BVH)" if big else "linear cull scan (RT_SCAN_FILTERED)",
                                                      "value_with_exact_early_out": value_eo}),
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
               # the same frame, same semantics, through the library's DEFAULT scan mode (RT_SCAN_AUTO -> exact BVH
               # traversal, SAH build at upload): what a caller of rt_render gets without asking for anything
               "auto_mode": {"scan_mode": "RT_SCAN_AUTO (flattened BVH, exact closest-hit semantics)", "value": value_auto,
                             "e2e": e2e_auto, "e2e_ms_per_step_rank0": e2e_auto_steps, "unit": "Msamples/s",
                             "box_tests_per_cast": None if not auto_stats else auto_stats["node_tests"] / max(1, auto_stats["casts"]),
                             "fp64_sphere_tests_per_cast": None if not auto_stats else auto_stats["exact_tests"] / max(1, auto_stats["casts"])}}
        print(json.dumps(out), flush=True)
    e2e_scene.close()
    scene.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
