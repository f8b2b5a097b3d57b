#!/usr/bin/env python
"""bench.py -- Msamples/s of the path-tracing hot path on N B200s (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3] [--spp S]

A "step" is one full frame of the workload (BASELINE.json configs[2]: 1200x800 random-spheres scene, 485
spheres, 500 spp, max depth 50) traced by the CUDA path; for N > 1 the frame's 8x8 tiles are dealt to the ranks,
all-gathered with NCCL and de-interleaved (strong scaling).

One JSON line on stdout (rank 0):
  value        device-timed throughput, scene resident in HBM, frame left in HBM.  Headline = the linear cull-scan
               kernel BASELINE.json's north_star specifies for this config (RT_SCAN_FILTERED), reference semantics
               (tmin 0, every cast the reference makes is executed)
  e2e          same metric through the host-buffer C ABI (rt_update_scene + rt_render): H2D of the scene (+ BVH build)
               and D2H of the RGBA frame into pinned memory inside the timed region
  roofline     FP32-FMA roofline of the render kernel: algorithmic work = sphere tests x 11 FP32-pipe
               instructions; peak = FFMA issue rate measured in this run, nominal peak beside it
  auto_mode    the same frame through the library default (RT_SCAN_AUTO: exact BVH traversal + tie grid)
  cpu_baseline the reference's own CPU code (oracle/_ref) or its C restatement, timed on this host on a
               bounded sample of the same workload: all cores, one thread, and (config 1) the reference's own
               unoptimised build configuration
  other_configs  short legs of BASELINE configs 1 and 4 (and 5 at 1 and 8 GPUs) with the same keys
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
SLOTS_PER_BOX = 14   # one BVH child box: 6 FFMA + 8 min/max/compare
NOMINAL_FMA_PER_S = 148 * 128 * 1.965e9   # SMs x FP32 lanes x max SM clock (MEASURED_PEAKS.json has no FP32 figure)
WORKLOAD_KEYS = {"c1": "c1_default", "c3": "c3_book_1200x800", "c4": "c4_bvh_1920x1080", "c5": "c5_book_4k"}


def workload(name: str, spp_override: int | None = None):
    from petershirleyraytracer_b200 import scenes
    key = WORKLOAD_KEYS.get(name, name)
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


def host_cores() -> int:
    # all the host cores this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(wl, target_s: float, threads: int = 0, which: str | None = None):
    """Msamples/s of the reference CPU implementation on a bounded sample of the workload: full-width rows
    from the middle of the frame at the workload's depth; spp and row count sized for ~target_s seconds.
    which: "ref" (oracle/_ref/libref.so, -O3), "refO0" (the reference's own unoptimised CMake configuration),
    "orc" (the C restatement); default: ref if built, else orc."""
    import oracle_lib as ol
    if which is None:
        which = "ref" if ol.have_ref() else "orc"
    kind = "port" if which == "orc" else "reference"
    cores = host_cores() if threads == 0 else threads
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
    build = {"ref": "-O3 -ffp-contract=off", "refO0": "-O0 (programs/CMakeLists.txt:1-6 sets no flags)", "orc": "-O3 -ffp-contract=off"}[which]
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind, "build": build, "seconds": dt,
            "sample": f"rows {j0}..{j1 - 1} of {H} x {W} px x {spp} spp, depth {wl['depth']} ({int(n)} samples, {dt:.1f} s)"}


def config_dict(wl, gpus: int):
    """Identical for the CUDA arm and the reference arm (the driver compares the two dicts)."""
    return {"workload": f"{wl['name']}: {wl['W']}x{wl['H']} px, {len(wl['radii'])} spheres (book layout, seed 42), "
                        f"{wl['spp']} spp, max depth {wl['depth']}, tmin 0 (reference semantics)",
            "width": wl["W"], "height": wl["H"], "spp": wl["spp"], "max_depth": wl["depth"], "spheres": int(len(wl["radii"])),
            "gpus": gpus}


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
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * statistics.mean(x["seconds"] for x in vals),    # measured duration of the bounded sample
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(wl, args.gpus),
           "notes": "reference CPU renderer (unmodified classes, OpenMP over rows); each step is a bounded sample of the workload "
                    f"sized for ~{per_step:.0f} s",
           "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                            "build": last["build"]},
           "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def load_traffic(wl_name: str, kernel: str):
    """DRAM bytes per launch of the render kernel from the ncu --set full capture of this workload at HEAD
    (profiles/r2_traffic.json, written from the committed ncu summaries); None if there is no capture for it."""
    path = os.path.join(REPO, "profiles", "r2_traffic.json")
    try:
        return json.load(open(path)).get(wl_name, {}).get(kernel)
    except Exception:
        return None


class GpuBench:
    """One workload on this rank's GPU: device-timed legs and the host-buffer e2e leg."""

    def __init__(self, wl, rt, torch, dist, world, rank, local, deal="tiles"):
        self.wl, self.rt, self.torch, self.dist = wl, rt, torch, dist
        self.world, self.rank, self.local = world, rank, local
        # how N ranks share one frame: "tiles" = 8x8 tiles dealt round-robin + all-gather of the 8-bit shards;
        # "samples" = every rank traces all tiles for 1/N of the samples + integer all-reduce of the radiance sums
        self.deal = deal if world > 1 else "tiles"
        self.dev = torch.device("cuda", local)
        self.W, self.H, self.spp, self.depth = wl["W"], wl["H"], wl["spp"], wl["depth"]
        self.cam = wl["cam"]
        self.scene = rt.Scene(wl["centres"], wl["radii"], device=local)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.frame = torch.empty(self.H * self.W * 4, dtype=torch.uint8, device=self.dev)
        layout = rt.tile_layout(self.params(rt.SCAN_AUTO, False))
        self.shard = torch.empty(layout.shard_bytes, dtype=torch.uint8, device=self.dev) if world > 1 else None
        self.gathered = torch.empty(world * layout.shard_bytes, dtype=torch.uint8, device=self.dev) if world > 1 else None
        self.samples_per_step = self.W * self.H * self.spp
        self.big = len(wl["radii"]) > 4080   # beyond the linear scan's constant bank: BVH only
        if self.deal == "samples":
            self.accum = torch.zeros(self.H * self.W * 3, dtype=torch.int64, device=self.dev)
            self.s_begin, self.s_end = rank * self.spp // world, (rank + 1) * self.spp // world

    def params(self, scan_mode, early_out, spp=None):
        if self.deal == "samples":
            return self.rt.make_params(self.W, self.H, spp or self.spp, self.depth, seed=0, early_out=early_out, scan_mode=scan_mode)
        return self.rt.make_params(self.W, self.H, spp or self.spp, self.depth, seed=0, early_out=early_out, scan_mode=scan_mode,
                                   shard_rank=self.rank, shard_count=self.world)

    def step(self, p, scene=None):
        """One frame, device buffers only.  Returns the kernels launched."""
        rt, sc = self.rt, scene or self.scene
        if self.deal == "samples":
            import copy
            total = p.spp
            b, e = self.rank * total // self.world, (self.rank + 1) * total // self.world
            q = copy.copy(p)
            q.spp = max(e - b, 1)
            self.accum.zero_()
            if e > b:
                rt.render_pass_device(sc, self.cam, q, b, self.accum.data_ptr(), 0, self.stream)
            self.dist.all_reduce(self.accum)
            rt.accum_to_frame(p, self.accum.data_ptr(), total, self.frame.data_ptr(), self.local, self.stream)
            return 2
        if self.world == 1:
            rt.render_device(sc, self.cam, p, self.frame.data_ptr(), 0, self.stream)
            return 1
        rt.render_device(sc, self.cam, p, self.shard.data_ptr(), 0, self.stream)
        self.dist.all_gather_into_tensor(self.gathered, self.shard)
        rt.deinterleave(p, self.gathered.data_ptr(), self.frame.data_ptr(), self.local, self.stream)
        return 2

    def timed(self, p, nsteps, nwarm, flush, sampler=None, warm_p=None):
        torch, dist, rt = self.torch, self.dist, self.rt
        for _ in range(nwarm):
            self.step(warm_p or p); rt.render_finish(self.scene)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        if sampler:
            sampler.start()
        tot_ms, kern_ms, launches, stats = 0.0, 0.0, 0, None
        for _ in range(nsteps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the timed region)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launches += self.step(p)
            e1.record()
            e1.synchronize()
            stats = rt.render_finish(self.scene)
            tot_ms += e0.elapsed_time(e1)
            kern_ms += stats["kernel_ms"]
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        local_kernel_ms = kern_ms
        t = torch.tensor([tot_ms, kern_ms], dtype=torch.float64, device=self.dev)
        cnt = torch.tensor([stats["sphere_tests"], stats["node_tests"], stats["exact_tests"], stats["casts"], stats["samples"],
                            stats["self_resolved"]], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)
        keys = ("sphere_tests", "node_tests", "exact_tests", "casts", "samples", "self_resolved")
        return {"ms": t[0].item(), "kernel_ms": t[1].item(), "kernel_ms_this_rank": local_kernel_ms, "launches": launches, "steps": nsteps,
                "value": self.samples_per_step * nsteps / (t[0].item() * 1e-3) / 1e6,
                "counts": dict(zip(keys, (x.item() for x in cnt)))}     # whole-job sums of the last step

    def e2e(self, p, nsteps, nwarm=2):
        """Host buffers through the public C ABI: every step re-uploads the flattened hittable_list into one reused device
        scene (rt_update_scene with a full rebuild: H2D of the sphere arrays + BVH / tie-grid build; no cudaMalloc /
        cudaFree, whose latency on shared hosts is erratic) and reads the frame back into pinned memory."""
        torch, dist, rt, wl = self.torch, self.dist, self.rt, self.wl
        if not hasattr(self, "e2e_scene"):
            self.e2e_scene = rt.Scene(wl["centres"], wl["radii"], device=self.local)
            self.host_frame = torch.empty(self.H * self.W * 4, dtype=torch.uint8).pin_memory()
            self.host_view = self.host_frame.numpy().reshape(self.H, self.W, 4)
        sc = self.e2e_scene

        def one():
            sc.update(wl["centres"], wl["radii"], refit=False)
            if self.world == 1:
                rt.render(sc, self.cam, p, out=self.host_view)
            else:
                self.step(p, scene=sc)
                if self.rank == 0:
                    self.host_frame.copy_(self.frame, non_blocking=False)
                rt.render_finish(sc)

        for _ in range(nwarm):
            one()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        per_step = []
        for _ in range(nsteps):
            ts = time.perf_counter()
            one()
            per_step.append(round(1e3 * (time.perf_counter() - ts), 1))
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.MAX)
        h2d = wl["centres"].nbytes + wl["radii"].nbytes + 96 + 128      # sphere arrays + camera + params
        return {"value": self.samples_per_step * nsteps / s.item() / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(self.H * self.W * 4), "ms_per_step_rank0": per_step}

    def roofline(self, leg, fma_per_s, mode_name):
        """FP32 roofline of one device-timed leg on EXECUTED work (so it can never exceed 1)."""
        c = leg["counts"]
        if mode_name == "scan":
            slots = c["sphere_tests"] * SLOTS_PER_TEST
        else:
            slots = c["node_tests"] * SLOTS_PER_BOX + c["exact_tests"] * SLOTS_PER_TEST
        kernel_s = leg["kernel_ms"] * 1e-3 / leg["steps"]              # average launch duration of the render kernel
        achieved = slots / self.world * 2 / kernel_s / 1e12             # counts are whole-job sums; ranks run concurrently
        peak, nominal = fma_per_s * 2 / 1e12, NOMINAL_FMA_PER_S * 2 / 1e12
        return {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_nominal": nominal, "frac_nominal": achieved / nominal,
                "kernel_ms_per_step": leg["kernel_ms"] / leg["steps"],
                "casts_per_sample": c["casts"] / max(c["samples"], 1),
                "sphere_tests_per_step": c["sphere_tests"], "box_tests_per_step": c["node_tests"],
                "fp64_sphere_tests_per_step": c["exact_tests"]}

    def close(self):
        if hasattr(self, "e2e_scene"):
            self.e2e_scene.close()
        self.scene.close()


def l2_note(wl, big):
    n = len(wl["radii"])
    if big:
        return (f"scene = {n * 32 / 1e6:.1f} MB FP64 sphere array + ~{n * 46 / 1e6:.1f} MB BVH / tie grid: L2 resident (126 MB), not L1; "
                "a 256 MB buffer is written between timed steps to flush L2")
    return (f"scene ({(n + 20) * 16 / 1e3:.0f} KB cull array in the constant bank, {n * 32 / 1e3:.0f} KB FP64 array) is cache resident by "
            "design; a 256 MB buffer is written between timed steps to flush L2")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--spp", type=int, default=None, help="override the workload's spp (quick checks; not the contract config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--deal", default=os.environ.get("RT_BENCH_DEAL", "tiles"), choices=["tiles", "samples"],
                    help="N > 1: tiles dealt to the ranks + all-gather (default), or samples split + integer all-reduce")
    args = ap.parse_args()
    wl = workload(args.workload, args.spp)

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fma_per_s, _ = rt.measure_fp32_peak(local)

    b = GpuBench(wl, rt, torch, dist, world, rank, local, deal=args.deal)
    head_mode = rt.SCAN_BVH if b.big else rt.SCAN_FILTERED
    head_name = "bvh" if b.big else "scan"

    # ---- headline: reference semantics, every cast executed
    sampler = ClockSampler(local) if rank == 0 else None
    head = b.timed(b.params(head_mode, False), args.steps, args.warmup, flush, sampler)
    clocks = sampler.stop() if sampler else None
    k = max(1, min(args.steps, 3))
    # ---- same frame with the exact early-out (bit-identical image, fewer casts)
    eo = b.timed(b.params(head_mode, True), k, 1, flush)
    # ---- the library's AUTO mode on the same frame (exact BVH traversal + tie grid for this scene size)
    auto = b.timed(b.params(rt.SCAN_AUTO, False), k, 1, flush)
    auto_eo = b.timed(b.params(rt.SCAN_AUTO, True), k, 1, flush)
    # ---- e2e: host buffers through the public C ABI
    e2e = b.e2e(b.params(head_mode, False), args.steps)
    e2e_auto = b.e2e(b.params(rt.SCAN_AUTO, False), k)

    # ---- short legs of the other BASELINE configs (device-timed value, e2e, executed-work roofline)
    other = {}
    if not args.no_other_configs and args.workload == "c3" and not args.spp:
        names = ["c1", "c4"] if world == 1 else []
        if world in (1, 8):
            names.append("c5")
        for name in names:
            owl = workload(name)
            ob = GpuBench(owl, rt, torch, dist, world, rank, local, deal=args.deal)
            mode = rt.SCAN_BVH if ob.big else rt.SCAN_FILTERED
            nst = {"c5": 1, "c1": 10}.get(name, 2)   # (c1 is an 8 ms frame: more steps, or launch jitter shows in the figure)
            # (a short frame warms caches / clocks; the timed steps are full frames.  c1's frame takes 8 ms: warm up with the
            #  frame itself, or the first timed step allocates the per-tile sums a 1 spp frame never needs)
            warm = None if name == "c1" else ob.params(mode, False, spp=max(1, owl["spp"] // 64))
            leg = ob.timed(ob.params(mode, False), nst, 1, flush, warm_p=warm)
            entry = {"config": config_dict(owl, world), "scan_mode": "bvh" if ob.big else "linear cull scan",
                     "value": leg["value"], "unit": "Msamples/s", "steps": nst, "ms_per_step": leg["ms"] / nst,
                     "roofline": ob.roofline(leg, fma_per_s, "bvh" if ob.big else "scan")}
            if name != "c5":
                entry["e2e"] = ob.e2e(ob.params(mode, False), nst, nwarm=1)
                if not ob.big:
                    aleg = ob.timed(ob.params(rt.SCAN_AUTO, False), nst, 1, flush)
                    entry["auto_mode"] = {"value": aleg["value"], "roofline": ob.roofline(aleg, fma_per_s, "bvh"),
                                          "self_resolved_per_cast": aleg["counts"]["self_resolved"] / max(aleg["counts"]["casts"], 1)}
            else:
                aleg = ob.timed(ob.params(rt.SCAN_AUTO, False), nst, 0, flush)
                entry["auto_mode"] = {"value": aleg["value"]}
            other[name] = entry
            ob.close()

    if rank == 0:
        roofline = b.roofline(head, fma_per_s, head_name)
        kern = "render_wave_kernel" if b.big else "render_kernel<2,1>"
        roofline["traffic"] = load_traffic(wl["name"], kern)
        roofline["note"] = ("FP32-FMA roofline (no tensor cores, HBM traffic ~nil): achieved = executed sphere tests x 11 FP32-pipe "
                            "instructions (+ BVH box tests x 14) x 2 flop / render-kernel time (CUDA events on the launching stream, "
                            "avg over the timed launches, per GPU); peak = FFMA rate measured in this run by rt_measure_fp32_peak, "
                            "peak_nominal = 148 SMs x 128 lanes x 1.965 GHz x 2 (MEASURED_PEAKS.json has no FP32 figure); the scan "
                            "loop itself issues 3.5 packed FFMA2 (= 7 FMAs) + ~2.3 other instructions per test")
        roofline["with_early_out"] = {"msamples_s": eo["value"], "casts_per_sample": eo["counts"]["casts"] / max(eo["counts"]["samples"], 1)}
        cpu = cpu1 = cpu_own = None
        if world == 1 and not args.no_cpu_baseline and not b.big:   # (the reference's O(N) scan of 1e5 spheres: ~1 ms per cast)
            cpu = cpu_reference_rate(wl, 12.0)
            cpu1 = cpu_reference_rate(wl, 8.0, threads=1)             # the reference itself is single-threaded (programs/main.cc:51-92)
            try:
                import oracle_lib as ol
                if ol.have_ref_O0():                                  # its own build configuration, on its own scene (config 1)
                    c1 = workload("c1")
                    cpu_own = {"O0_1thread": cpu_reference_rate(c1, 6.0, threads=1, which="refO0"),
                               "O3_1thread": cpu_reference_rate(c1, 4.0, threads=1, which="ref"),
                               "workload": config_dict(c1, 0)["workload"]}
            except Exception as e:  # noqa: BLE001
                cpu_own = {"error": repr(e)}
        ca = auto["counts"]
        out = {"metric": "Msamples/s", "value": head["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": head["ms"] / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64 hit/shading + f32 cull", "data": "synthetic",
               "config": config_dict(wl, world),
               "notes": {"scan_mode": "4-wide BVH + tie grid (RT_SCAN_BVH)" if b.big else "linear cull scan (RT_SCAN_FILTERED)",
                         "early_out": False, "paths_per_lane": 1 if b.big else 2,
                         "parallelism": (f"tiles8x8 dealt to {world} GPU(s)" + (" + NCCL all-gather" if world > 1 else "")) if b.deal == "tiles"
                         else f"samples split over {world} GPUs + NCCL all-reduce of the integer radiance sums",
                         "l2": l2_note(wl, b.big), "value_with_exact_early_out": eo["value"]},
               "clocks": clocks, "e2e": e2e, "gpu_launches": head["launches"], "roofline": roofline,
               "cpu_baseline": cpu, "cpu_baseline_1thread": cpu1, "cpu_baseline_reference_build": cpu_own,
               # the same frame, same semantics, through the library's DEFAULT scan mode (RT_SCAN_AUTO -> exact BVH
               # traversal, SAH build + tie grid at upload): what a caller of rt_render gets without asking for anything
               "auto_mode": {"scan_mode": "RT_SCAN_AUTO (flattened BVH + start-sphere test / tie grid, exact closest-hit semantics)",
                             "value": auto["value"], "value_with_exact_early_out": auto_eo["value"],
                             "e2e": e2e_auto["value"], "e2e_ms_per_step_rank0": e2e_auto["ms_per_step_rank0"], "unit": "Msamples/s",
                             "kernel_ms_per_step": auto["kernel_ms"] / auto["steps"],
                             "roofline_executed_work": b.roofline(auto, fma_per_s, "bvh"),
                             "box_tests_per_cast": ca["node_tests"] / max(ca["casts"], 1),
                             "fp64_sphere_tests_per_cast": ca["exact_tests"] / max(ca["casts"], 1),
                             "casts_decided_without_traversal": ca["self_resolved"] / max(ca["casts"], 1)},
               "other_configs": other}
        print(json.dumps(out), flush=True)
    b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
