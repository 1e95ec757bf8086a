#!/usr/bin/env python
"""bench.py — path samples/sec (Msamples/s) of the render hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm (libbrt, CUDA)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # N ranks, one per GPU, spp split + NCCL/P2P reduce
    python bench.py --impl reference                              # the CPU restatement of the reference on host cores

A "step" is ONE full render of the workload (default C3: synthetic random-spheres scene, 1920x1080, 256 spp, depth 10,
thin-lens aperture) through the path `RayTracer.render()` replaces (js/ray-tracer.js:166-281): zero the sums, trace all
samples, [reduce across GPUs], resolve (÷spp, tone map, gamma, RGBA8).  `value` keeps the scene resident in HBM;
`e2e` pushes the scene through the C ABI from host memory every step (flatten -> upload -> LBVH build -> render ->
D2H of the RGBA8 image into pinned host memory).  One path sample = one camera sample carried to termination.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "path samples/sec"
UNIT = "Msamples/s"

# per-test algorithmic flop counts from the reference's own arithmetic
FLOPS = dict(tests_sphere=24, tests_plane=18, tests_box=26, tests_tri_a=28, tests_tri_b=18, tests_tri_c=8, tests_aabb=23)

WORKLOADS = {
    # name: (scene factory kwargs, W, H, spp, depth, description)
    "c1": dict(fixture="sample_scene.json", W=600, H=400, spp=16, depth=10,
               desc="sample_scene.json 600x400 16spp depth10"),
    "c2": dict(fixture="sample_mesh.json", W=1280, H=720, spp=64, depth=10,
               desc="sample_mesh.json 1280x720 64spp depth10"),
    # BASELINE config 2 names "point-light shadow rays": the reference never calls its lights (lights.js has no call site),
    # so c2 renders as the reference does and c2d adds the direct-lighting EXTENSION (shadow rays to both lights)
    "c2d": dict(fixture="sample_mesh.json", W=1280, H=720, spp=64, depth=10, direct=True,
                desc="sample_mesh.json 1280x720 64spp depth10 + direct-lighting extension (point / directional shadow rays)"),
    "c3": dict(gen="c3", W=1920, H=1080, spp=256, depth=10,
               desc="synthetic random-spheres (486 objects, seed 42) 1920x1080 256spp depth10 thin-lens aperture 0.1"),
    "c4": dict(gen="c4", W=1920, H=1080, spp=1024, depth=16,
               desc="synthetic Cornell-style (planes, boxes, emissive quads, procedural sky) 1920x1080 1024spp depth16"),
    "c5": dict(gen="c5", W=3840, H=2160, spp=4096, depth=10,
               desc="synthetic 1,002,528-triangle terrain mesh 3840x2160 4096spp depth10"),
}


def load_workload(name: str):
    w = dict(WORKLOADS[name])
    if "fixture" in w:
        with open(os.path.join(ROOT, "tests", "golden", w["fixture"])) as f:
            scene = json.load(f)
    else:
        from tools import gen_scenes
        scene = gen_scenes.SCENES[w["gen"]]()
    w["scene"] = scene
    return w


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost", 0x100: "display_clock_setting"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_rate(w, threads: int, budget_s: float, spp: int = 1):
    """Times the float64 oracle (oracle/ — the CPU restatement of the reference; brute-force loops exactly as
    world.js:24-30 / geometry.js:253-259) on a bounded sample of the workload: full-width row bands at `spp`
    samples per pixel, spread evenly over the frame, until `budget_s` is spent.  -> (Msamples/s, sample description)."""
    from oracle.oracle import OracleRayTracer
    W, H = w["W"], w["H"]
    o = OracleRayTracer(W, H, seed=1, threads=threads)
    assert o.loadFromJSON(w["scene"])
    o.resizeCanvas(W, H)
    o.updateRenderSettings(dict(samples=spp, maxBounces=w["depth"]))
    o.directLighting = bool(w.get("direct"))
    band = max(1, min(H, 4 * max(1, threads)))
    # bands visited in a bit-reversed order so any prefix covers the frame evenly; when the frame is done and budget
    # remains, another pass renders the next sample index of every pixel
    nb = (H + band - 1) // band
    order = sorted(range(nb), key=lambda i: int(format(i, "016b")[::-1], 2))
    done, t_used, bands, passes = 0, 0.0, 0, 0
    while t_used < budget_s and passes < 64:
        o.sampleBegin = passes * spp
        for b in order:
            y0, y1 = b * band, min(H, (b + 1) * band)
            t0 = time.perf_counter()
            o.render(rect=(0, y0, W, y1), reuse=True)
            t_used += time.perf_counter() - t0
            done += (y1 - y0) * W * spp
            bands += 1
            if t_used >= budget_s:
                break
        passes += 1
    rate = done / t_used / 1e6
    return rate, (f"{bands} full-width {band}-row bands ({bands / nb:.2f} frames of {W}x{H} at {spp} spp, bit-reversed band order) "
                  f"= {done} path samples, {t_used:.1f} s on {threads} thread(s)")


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = load_workload(args.workload)
    threads = os.cpu_count() or 1
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        oracle_rate(w, threads, per_step * 0.25)
    rates, t0, sample = [], time.perf_counter(), ""
    for _ in range(args.steps):
        r, sample = oracle_rate(w, threads, per_step)
        rates.append(r)
    dt = time.perf_counter() - t0
    value = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "note": "CPU only; each step renders a bounded sample of the same frame (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle/ float64 C++ restatement of the reference JS (no JS engine in the image; cpp/ray-tracer-engine.cpp is a 0-byte file)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.append(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def algorithmic_flops(stats: dict) -> float:
    return float(sum(stats[k] * f for k, f in FLOPS.items()))


def run_ours(args, out):
    import numpy as np
    import torch
    import torch.distributed as dist
    import blenderraytracer_b200 as brt
    from blenderraytracer_b200.distributed import SppSplitRenderer, sample_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libbrt has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    w = load_workload(args.workload)
    W, H, spp, depth = w["W"], w["H"], args.spp or w["spp"], w["depth"]
    text = json.dumps(w["scene"]).encode()
    rt = brt.RayTracer(W, H, device=local, seed=args.seed)
    assert rt.loadFromJSON(text), getattr(rt, "lastError", "")
    rt.resizeCanvas(W, H)                                   # aspect = W/H as the UI path does (ray-tracer.js:505)
    rt.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    rt.sampler, rt.accel, rt.integrator = args.sampler, args.accel, args.integrator
    rt.directLighting = bool(w.get("direct"))
    rt.refillThreshold = args.refill
    rt.pathsInFlight = args.inflight
    info = rt.sceneInfo()
    sr = SppSplitRenderer(rt, reduce=args.reduce)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 0)):
        sr.step()
    barrier()

    # ---- timed: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        if world > 1:
            dist.barrier()
        e0, e1, e2 = ev[k]
        e0.record()
        rt.deviceMemset(sr.accum_ptr, 0, sr.nbytes)
        begin, count = sample_range(sr.spp(), sr.rank, sr.world)
        rt.renderAccumulate(sr.accum_ptr, begin, count)
        e1.record()                                          # e0..e1 = zero fill + the path-tracing megakernel
        _finish_step(sr)
        e2.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    kern_ms = [e[0].elapsed_time(e[1]) for e in ev]
    total_ms = float(sum(step_ms))
    tt = torch.tensor([total_ms, float(sum(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms = tt.tolist()
    samples_per_step = W * H * spp
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: host scene -> C ABI -> host pixels, every step
    desc_keep = None
    host_rgba = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    scene_flat = rt.sceneFlatDesc()                          # host descriptors as brt_scene_set_flat takes them
    h2d = rt.sceneInfo()["upload_bytes"]

    def e2e_step():
        rt.setSceneFlat(scene_flat)                          # marks the device scene dirty: re-upload + LBVH rebuild
        sr.step()
        if rank == 0:
            if sr.reduce == "p2p":
                rt.copyToHost(host_rgba.data_ptr(), sr._rgba_ptr, host_rgba.numel())
            else:
                host_rgba.copy_(sr.rgba, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / te.item() / 1e6

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (k_pathtrace): algorithmic flops from a counting build of the SAME traversal
        roof = None
        try:
            rt.countTests = True
            rt._push_params()
            begin, count = sample_range(sr.spp(), 0, world)
            rt.deviceMemset(sr.accum_ptr, 0, sr.nbytes)
            rt.renderAccumulate(sr.accum_ptr, begin, count)
            rt.synchronize()
            st = rt.stats()
            rt.countTests = False
            rt._push_params()
            flops_launch = algorithmic_flops(st)
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
            try:
                hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
            except Exception:
                pass
            peak = rt.measureFp32Peak()
            kern_ms_avg = kern_total_ms / args.steps
            achieved = flops_launch / (kern_ms_avg * 1e-3) / 1e12
            n_obj_flops = info["n_spheres"] * 24 + info["n_planes"] * 18 + info["n_boxes"] * 26 + info["n_triangles"] * 28
            traffic = None
            tp = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            roof = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "k_pathtrace", "kernel_ms": kern_ms_avg,
                    "peak_source": "FFMA micro-benchmark run live in this process (MEASURED_PEAKS.json has no fp32 entry); "
                                   "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4",
                    "flops_per_launch": flops_launch, "rays_per_launch": st["rays"],
                    "rays_per_sample": st["rays"] / max(1, W * H * count),
                    "tests": {k: st[k] for k in FLOPS},
                    "flops_bruteforce_per_launch": float(st["rays"]) * n_obj_flops,
                    # SURVEY 8(d) bytes model of the traversal: 64 B per node visit, 48 / 16 / 32 B per triangle / sphere / box test,
                    # 32 B per plane test — an UPPER bound on memory traffic (L1 / L2 serve nearly all of it; see `traffic`)
                    "bvh_bytes_model": {"bytes_per_launch": (bvh_bytes := st["tests_aabb"] // 2 * 64 + st["tests_tri_a"] * 48 + st["tests_sphere"] * 16
                                                             + st["tests_box"] * 32 + st["tests_plane"] * 32),
                                        "achieved_GBps": bvh_bytes / (kern_ms_avg * 1e-3) / 1e9, "hbm_peak_GBps": hbm_peak,
                                        "frac_of_hbm": bvh_bytes / (kern_ms_avg * 1e-3) / 1e9 / hbm_peak},
                    "note": "not a dense contraction: no tensor cores; scene + BVH are L1/L2 resident, HBM traffic is the accumulation buffer only",
                    # the same kernel against the HBM roofline, for completeness: algorithmic bytes = one read-modify-write of the
                    # W*H*16 B accumulation buffer per launch; it shows why "hbm" is not the bound of this path
                    "hbm": {"bound": "hbm", "achieved": (W * H * 32) / (kern_ms_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": (W * H * 32) / (kern_ms_avg * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                            "peak_source": hbm_src}}
        except Exception as ex:                               # keep the bench line even if the counting build fails
            roof = {"bound": "fp32", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None, "error": str(ex)}

        # ---- CPU baseline (rank 0, N = 1 only): the oracle on the box's host cores, bounded sample
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            v, sample = oracle_rate(w, threads, args.cpu_seconds)
            v1, sample1 = oracle_rate(w, 1, min(6.0, args.cpu_seconds / 2))
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                   "value_1thread": v1, "sample_1thread": sample1,
                   "note": "oracle/ float64 C++ restatement of the reference JS (brute-force loops as the reference); the reference "
                           "itself is single-threaded browser JavaScript and no JS engine exists in this image"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "spp_total": spp, "spp_per_gpu": sample_range(spp, 0, world)[1],
                       "sampler": args.sampler, "accel": "bvh" if info["n_bvh_nodes"] and args.accel != "brute" else "brute",
                       "integrator": "megakernel", "reduce": sr.reduce, "l2": "flushed between timed steps (256 MiB fill)",
                       "timing": "CUDA events per step on the launching stream, summed over steps, max over ranks",
                       "bvh_nodes": info["n_bvh_nodes"], "bvh_build_ms": info["bvh_build_ms"]},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 4),
                    "note": "per step: brt_scene_set_flat from host descriptors (upload + LBVH build) -> render -> RGBA8 to pinned host; wall clock, max over ranks"},
            "gpu_launches": int(sr.launches_per_step() * args.steps),
            "wall_s_timed_region": t_wall,
            "roofline": roof,
        }
        if cpu:
            line["cpu_baseline"] = cpu
    sr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        out.append(json.dumps(line))
    return 0


def _finish_step(sr):
    """The part of SppSplitRenderer.step() after the path-tracing launch: exchange + resolve."""
    rt = sr.rt
    if sr.reduce == "p2p":
        from blenderraytracer_b200.distributed import row_stripe
        sr._stream_barrier()
        r0, r1 = row_stripe(rt.height, sr.rank, sr.world)
        rt.reduceResolvePeers(sr._peers, r0, r1, sr._root_rgba)
        sr._stream_barrier()
    else:
        if sr.reduce == "nccl":
            from blenderraytracer_b200.distributed import reduce_sums
            reduce_sums(sr.accum, 0, sr.group)
        if sr.rank == 0:
            rt.resolveDevice(sr.accum_ptr, sr.rgba.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel")
    ap.add_argument("--sampler", default="fast", choices=["fast", "reference"])
    ap.add_argument("--accel", default="auto", choices=["auto", "brute", "bvh"])
    ap.add_argument("--integrator", default="auto", choices=["auto", "megakernel", "wavefront"])
    ap.add_argument("--reduce", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--refill", type=int, default=0, help="extend-phase refill threshold (idle lanes); 0 = library default")
    ap.add_argument("--inflight", type=int, default=0, help="samples of a pixel in flight per lane (1..4); 0 = library default")
    ap.add_argument("--cpu-seconds", type=float, default=14.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                       # timing rule: W >= 3
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "ours":
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version banner, make) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    line_holder = []
    try:
        rc = run_reference(args, line_holder) if args.impl == "reference" else run_ours(args, line_holder)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in line_holder:
        print(line, flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
