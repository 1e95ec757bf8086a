#!/usr/bin/env python
"""bench.py — path samples/sec (Msamples/s) of the render hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm (libbrt, CUDA)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # N ranks, one per GPU: spp split + fused peer exchange
    python bench.py --impl reference                              # the CPU restatement of the reference on host cores

A "step" is ONE full render of the workload (default C3: synthetic random-spheres scene, 1920x1080, 256 spp, depth 10,
thin-lens aperture) through the path `RayTracer.render()` replaces (js/ray-tracer.js:166-281): zero the sums, trace all
samples, [exchange across GPUs], resolve (÷spp, tone map, gamma, RGBA8).
  value  scene resident in HBM, CUDA events around every step on the launching stream, L2 flushed between steps, max over ranks.
  e2e    the plugin call itself, wall clock: at N = 1 `brt_scene_set_flat` (host descriptors) + `brt_render` into the caller's
         pinned HOST buffer — scene upload, LBVH build, render, resolve and the device-to-host copy all inside the call; at
         N > 1 the same per rank through the peer group (`brt_peer_render` + `brt_peer_fetch` into rank 0's host buffer).
  secondary  BASELINE config 5 (1 M-triangle terrain, 3840x2160) at 256 spp PER GPU (weak scaling: the config is 4096 spp,
         spp-split), same measurements, so the multi-GPU config is on the driver's record at every N.
  image_check (N > 1)  rank 0 renders the same sample set alone and compares RGBA8 with the N-rank image, for both exchanges.
One path sample = one camera sample carried to termination.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "path samples/sec"
UNIT = "Msamples/s"
L2_NOTE = "GPU arm: flushed between timed steps (256 MiB fill)"

# per-test algorithmic flop counts from the reference's own arithmetic
FLOPS = dict(tests_sphere=24, tests_plane=18, tests_box=26, tests_tri_a=28, tests_tri_b=18, tests_tri_c=8, tests_aabb=23)

WORKLOADS = {
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


def load_workload(name: str, binary: bool = False):
    """-> workload dict with `scene` (parsed JSON dict) and, with binary=True for generated meshes, `blob` (BRTSCN01 bytes:
    the same scene with mesh arrays in binary, ingested 10x faster than 38 MB of JSON text)."""
    w = dict(WORKLOADS[name])
    if "fixture" in w:
        with open(os.path.join(ROOT, "tests", "golden", w["fixture"])) as f:
            w["scene"] = json.load(f)
    else:
        from tools import gen_scenes
        if binary and w["gen"] == "c5":
            from tools import scene_binary
            w["scene"] = None
            w["blob"] = scene_binary.pack(gen_scenes.terrain(as_arrays=True))
        else:
            w["scene"] = gen_scenes.SCENES[w["gen"]]()
    return w


def workload_config(w, spp_total=None) -> dict:
    """The `config` object of the JSON line: identical in our arm and in the reference arm."""
    return {"workload": w["desc"], "width": w["W"], "height": w["H"], "spp_total": int(spp_total or w["spp"]), "max_depth": w["depth"],
            "l2": L2_NOTE}


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


def js_reference_rate(w, budget_s: float):
    """the reference's own JavaScript under Node (baseline/run_ref.mjs calls RayTracer.render() of
    /root/reference/js unchanged, single thread).  -> dict for cpu_baseline["js"]."""
    node = shutil.which("node") or shutil.which("nodejs")
    if not node:
        return "unavailable: no node / nodejs binary in this image (probed at bench time); baseline/run_ref.mjs is the harness to run where Node exists"
    ref = os.environ.get("BRT_REFERENCE_JS", os.path.join(ROOT, "baseline", "_ref", "js"))
    if not os.path.isdir(ref):
        return f"unavailable: node found at {node} but no copy of the reference's js/ under {ref} (see baseline/README.md)"
    try:
        scene_path = os.path.join(ROOT, "gpurun_out", "_bench_scene.json")
        os.makedirs(os.path.dirname(scene_path), exist_ok=True)
        json.dump(w["scene"], open(scene_path, "w"))
        rows = max(2, int(w["H"] * min(1.0, budget_s / 600.0)))
        cmd = [node, os.path.join(ROOT, "baseline", "run_ref.mjs"), "--ref", ref, "--scene", scene_path, "--width", str(w["W"]), "--height", str(w["H"]),
               "--samples", "1", "--bounces", str(w["depth"]), "--time-rows", str(rows)]
        r = json.loads(subprocess.check_output(cmd, timeout=max(60.0, 8 * budget_s), text=True).strip().splitlines()[-1])
        return {"value": r["msamples_per_s"], "unit": UNIT, "cores": 1, "kind": "reference", "sample": r["sample"]}
    except Exception as ex:                                   # never lose the bench line over the optional row
        return f"unavailable: node harness failed ({type(ex).__name__}: {ex})"


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
        "config": workload_config(w, args.spp or None),
        "run": {"note": "CPU only; each step renders a bounded sample of the same frame and reports a rate (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle/ float64 C++ restatement of the reference JS (no JS engine in the image; cpp/ray-tracer-engine.cpp is a 0-byte file)",
                         "js": js_reference_rate(w, 10.0)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.append(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def algorithmic_flops(stats: dict) -> float:
    return float(sum(stats[k] * f for k, f in FLOPS.items()))


class Env:
    """torch / distributed state of this rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — libbrt has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure(env: Env, args, name: str, spp_total: int, steps: int, warmup: int, cpu_seconds: float, check_image: bool, reduce: str):
    """All measurements of one workload on this rank set.  -> dict (rank 0) / None."""
    import numpy as np
    import blenderraytracer_b200 as brt
    from blenderraytracer_b200.distributed import SppSplitRenderer, sample_range
    torch, world, rank, local = env.torch, env.world, env.rank, env.local

    w = load_workload(name, binary=True)
    W, H, depth = w["W"], w["H"], w["depth"]
    rt = brt.RayTracer(W, H, device=local, seed=args.seed)
    t_ing = time.perf_counter()
    assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode()), getattr(rt, "lastError", "")
    ingest_s = time.perf_counter() - t_ing
    rt.resizeCanvas(W, H)                                   # aspect = W/H as the UI path does (ray-tracer.js:505)
    rt.updateRenderSettings(dict(samples=spp_total, maxBounces=depth))
    rt.sampler, rt.accel, rt.integrator = args.sampler, args.accel, args.integrator
    rt.directLighting = bool(w.get("direct"))
    rt.refillThreshold = args.refill
    rt.pathsInFlight = args.inflight
    rt.setStream(torch.cuda.current_stream().cuda_stream)
    rt._push_params()                                       # width / height / spp / depth reach the ctx (brt_set_render_params)
    info = rt.sceneInfo()
    my_begin, my_count = sample_range(spp_total, rank, world)
    samples_per_step = W * H * spp_total
    px_bytes = W * H * 16

    sr = SppSplitRenderer(rt, reduce=reduce) if world > 1 else None
    if world == 1:
        accum = torch.zeros((H, W, 4), dtype=torch.float32, device=env.dev)
        rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device=env.dev)

    def step_device():
        """the timed `value` step; returns the device time of the path-tracing launch of that step when known"""
        if world == 1:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            rt.deviceMemset(accum.data_ptr(), 0, px_bytes)
            rt.renderAccumulate(accum.data_ptr(), 0, spp_total)
            e1.record()                                      # e0..e1 = zero fill + the path-tracing megakernel
            rt.resolveDevice(accum.data_ptr(), rgba.data_ptr())
            e2.record()
            return e0, e1, e2
        e0, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sr.step()
        e2.record()
        return e0, None, e2

    for _ in range(max(warmup, 0)):
        step_device()
    env.barrier()

    # ---- timed: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
    clocks = ClockSampler(local)
    clocks.start()
    env.barrier()
    t_wall0 = time.perf_counter()
    step_ms, kern_ms = [], []
    for k in range(steps):
        env.flush.fill_(k & 0xFF)
        env.barrier()
        e0, e1, e2 = step_device()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e2))
        if e1 is not None:
            kern_ms.append(e0.elapsed_time(e1))
        else:
            rt.peerFetch(None)                               # no copy: refreshes the ctx's own event timings of this step
            kern_ms.append(rt.stats()["kernel_ms"])
    env.barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    total_ms, kern_total_ms = env.max_over_ranks([float(sum(step_ms)), float(sum(kern_ms))])
    value = samples_per_step * steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: host scene -> the plugin call -> host pixels, every step (wall clock, max over ranks)
    host_rgba = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    scene_flat = rt.sceneFlatDesc()                          # host descriptors as brt_scene_set_flat takes them
    h2d = rt.sceneInfo()["upload_bytes"]

    def e2e_step():
        rt.setSceneFlat(scene_flat)                          # marks the device scene dirty: re-upload + LBVH rebuild inside the call
        if world == 1:
            rt.renderInto(host_rgba.data_ptr())              # brt_render: the replaced seam (ray-tracer.js:166-281), host buffer out
        else:
            sr.step()
            sr.fetch_into(host_rgba.data_ptr())

    rt.setStream(None if world == 1 else torch.cuda.current_stream().cuda_stream)
    e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    env.barrier()
    (e2e_s,) = env.max_over_ranks([time.perf_counter() - t0])
    e2e_value = samples_per_step * steps / e2e_s / 1e6
    rt.setStream(torch.cuda.current_stream().cuda_stream)

    # ---- image check (N > 1): the N-rank image of both exchanges against rank 0 rendering the same sample set alone
    image_check = None
    if world > 1 and check_image:
        image_check = {"n_ranks": world}
        single = None
        for mode in ("fused", "nccl"):
            s2 = sr if mode == reduce else SppSplitRenderer(rt, reduce=mode)
            s2.step()
            img = s2.image()
            if s2 is not sr:
                s2.close()
            if rank == 0:
                if single is None:
                    rt.setStream(None)
                    single = rt.render(want_float=False).copy()      # brt_render on one GPU: all spp_total samples
                    rt.setStream(torch.cuda.current_stream().cuda_stream)
                d = np.abs(img.astype(np.int16) - single.astype(np.int16))
                image_check[mode] = {"max_lsb_diff": int(d.max()), "differing_bytes": int((d > 0).sum()), "bytes": int(d.size)}
            env.barrier()
        if rank == 0:
            image_check["max_lsb_diff"] = max(image_check[m]["max_lsb_diff"] for m in ("fused", "nccl"))
            image_check["differing_bytes"] = max(image_check[m]["differing_bytes"] for m in ("fused", "nccl"))
            image_check["note"] = "same Philox-keyed sample set; the ranks' fp32 partial sums are added in rank order, so bytes may differ by fp32 summation order only"

    res = None
    if rank == 0:
        # ---- roofline of the dominant kernel (k_pathtrace_mega): algorithmic flops from a counting build of the SAME traversal
        roof = None
        try:
            rt.countTests = True
            rt._push_params()
            cacc = accum if world == 1 else torch.zeros((H, W, 4), dtype=torch.float32, device=env.dev)
            rt.deviceMemset(cacc.data_ptr(), 0, px_bytes)
            rt.renderAccumulate(cacc.data_ptr(), my_begin, my_count)
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
            kern_ms_avg = kern_total_ms / steps
            achieved = flops_launch / (kern_ms_avg * 1e-3) / 1e12
            n_obj_flops = info["n_spheres"] * 24 + info["n_planes"] * 18 + info["n_boxes"] * 26 + info["n_triangles"] * 28
            traffic = None
            tp = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            rays = max(1, st["rays"])
            slots = 32 * max(1, st["trav_warp_iters"])
            bvh_bytes = st["tests_aabb"] // 2 * 64 + st["tests_tri_a"] * 48 + st["tests_sphere"] * 16 + st["tests_box"] * 32 + st["tests_plane"] * 32
            roof = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "k_pathtrace_mega", "kernel_ms": kern_ms_avg,
                    "peak_source": "FFMA micro-benchmark run live in this process (MEASURED_PEAKS.json has no fp32 entry); "
                                   "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4",
                    "flops_per_launch": flops_launch, "rays_per_launch": st["rays"],
                    "rays_per_sample": st["rays"] / max(1, W * H * my_count),
                    # quantities that only move the right way when the kernel or the hierarchy gets better (a worse tree inflates `achieved`)
                    "grays_per_s": st["rays"] / (kern_ms_avg * 1e-3) / 1e9,
                    "slab_tests_per_ray": st["tests_aabb"] / rays, "node_visits_per_ray": st["node_visits"] / rays,
                    "prim_tests_per_ray": (st["tests_sphere"] + st["tests_box"] + st["tests_tri_a"] + st["tests_plane"]) / rays,
                    "lane_slots": {"working": st["trav_lane_iters"] / slots, "waiting_for_slowest_ray": (st["trav_alive_lanes"] - st["trav_lane_iters"]) / slots,
                                   "drained": (slots - st["trav_alive_lanes"]) / slots,
                                   "note": "share of the 32 lane slots of every BVH-loop warp iteration (counting build, same rays)"},
                    "useful_lane_issue_frac": st["trav_lane_iters"] / slots,
                    "tests": {k: st[k] for k in FLOPS},
                    "flops_bruteforce_per_launch": float(st["rays"]) * n_obj_flops,
                    # SURVEY 8(d) bytes model of the traversal: 64 B per node visit, 48 / 16 / 32 B per triangle / sphere / box test,
                    # 32 B per plane test — an UPPER bound on memory traffic (L1 / L2 serve nearly all of it; see `traffic`)
                    "bvh_bytes_model": {"bytes_per_launch": bvh_bytes, "achieved_GBps": bvh_bytes / (kern_ms_avg * 1e-3) / 1e9, "hbm_peak_GBps": hbm_peak,
                                        "frac_of_hbm": bvh_bytes / (kern_ms_avg * 1e-3) / 1e9 / hbm_peak},
                    "note": "not a dense contraction: no tensor cores; scene + BVH are L1/L2 resident, HBM traffic is the accumulation buffer only",
                    # the same kernel against the HBM roofline, for completeness: algorithmic bytes = one read-modify-write of the
                    # W*H*16 B accumulation buffer per launch; it shows why "hbm" is not the bound of this path
                    "hbm": {"bound": "hbm", "achieved": (W * H * 32) / (kern_ms_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": (W * H * 32) / (kern_ms_avg * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "peak_source": hbm_src}}
        except Exception as ex:                               # keep the bench line even if the counting build fails
            roof = {"bound": "fp32", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None, "error": str(ex)}

        # ---- CPU baseline (rank 0, N = 1 only): the oracle on the box's host cores, bounded sample
        cpu = None
        if world == 1 and cpu_seconds > 0 and w.get("scene") is not None:
            threads = os.cpu_count() or 1
            v, sample = oracle_rate(w, threads, cpu_seconds)
            v1, sample1 = oracle_rate(w, 1, min(6.0, cpu_seconds / 2))
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                   "value_1thread": v1, "sample_1thread": sample1,
                   "js": js_reference_rate(w, cpu_seconds),
                   "note": "oracle/ float64 C++ restatement of the reference JS (brute-force loops as the reference); the reference "
                           "itself is single-threaded browser JavaScript"}

        integ = args.integrator if args.integrator != "auto" else "megakernel"
        res = {
            "value": value, "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
            "config": workload_config(w, spp_total),
            "run": {"spp_per_gpu": my_count, "sampler": args.sampler, "accel": "bvh" if info["n_bvh_nodes"] and args.accel != "brute" else "brute",
                    "integrator": integ, "exchange": (sr.reduce if sr else "none"),
                    "timing": "CUDA events per step on the launching stream, summed over steps, max over ranks",
                    "bvh_nodes": info["n_bvh_nodes"], "bvh_depth": info["bvh_depth"], "bvh_width": info["bvh_width"], "bvh_build_ms": info["bvh_build_ms"],
                    "scene_ingest_s": ingest_s, "scene_upload_ms": info["upload_ms"]},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 4),
                    "call": ("brt_scene_set_flat + brt_render(ctx, host_rgba8): scene upload, LBVH build, render, resolve and the D2H copy inside the plugin call"
                             if world == 1 else "per rank brt_scene_set_flat + brt_peer_render; rank 0 brt_peer_fetch(host_rgba8)"),
                    "note": "wall clock around the calls, pinned host output buffer, max over ranks"},
            "gpu_launches": int((sr.launches_per_step() if sr else 2) * steps),
            "wall_s_timed_region": t_wall,
            "roofline": roof,
        }
        if cpu:
            res["cpu_baseline"] = cpu
        if image_check:
            res["image_check"] = image_check
    if sr:
        sr.close()
    rt.close()
    return res


def run_ours(args, out):
    env = Env()
    try:
        spp = args.spp or WORKLOADS[args.workload]["spp"]
        prim = measure(env, args, args.workload, spp, args.steps, args.warmup, 0.0 if args.no_cpu else args.cpu_seconds,
                       check_image=True, reduce=args.reduce)
        sec = None
        if not args.no_secondary and args.workload != "c5":
            # BASELINE config 5: 1 M-triangle terrain at 4K, spp split; 256 spp per GPU (weak scaling), few steps
            sec = measure(env, args, "c5", args.secondary_spp * env.world, min(args.steps, 3), 3, 0.0, check_image=True, reduce=args.reduce)
        if env.rank == 0:
            line = {"metric": METRIC, "value": prim["value"], "unit": UNIT, "n_gpus": env.world, "steps": prim["steps"], "warmup": prim["warmup"],
                    "ms_per_step": prim["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic"}
            for k in ("config", "run", "clocks", "e2e", "gpu_launches", "wall_s_timed_region", "roofline", "cpu_baseline", "image_check"):
                if k in prim:
                    line[k] = prim[k]
            if sec:
                sec = dict(sec)
                sec.update({"metric": METRIC, "unit": UNIT, "scaling": "weak", "n_gpus": env.world,
                            "note": f"BASELINE config 5 at {args.secondary_spp} spp per GPU (the config's 4096 spp, spp-split, is a 13 s frame per GPU-eighth); "
                                    "efficiency at N = value(N) / (N x value(1)) of this block"})
                line["gpu_launches"] += sec.pop("gpu_launches", 0)
                line["secondary"] = sec
            out.append(json.dumps(line))
    finally:
        env.close()
    return 0


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
    ap.add_argument("--reduce", default="fused", choices=["fused", "nccl", "p2p"], help="N > 1 exchange: fused peer kernel (default) or NCCL reduce_scatter")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--refill", type=int, default=0, help="extend-phase refill threshold (idle lanes); 0 = library default")
    ap.add_argument("--inflight", type=int, default=0, help="samples of a pixel in flight per lane (1..4); 0 = library default")
    ap.add_argument("--cpu-seconds", type=float, default=14.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the BASELINE config 5 block")
    ap.add_argument("--secondary-spp", type=int, default=256, help="samples per pixel PER GPU of the config 5 block")
    args = ap.parse_args()
    if args.reduce == "p2p":
        args.reduce = "fused"
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
