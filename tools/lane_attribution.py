"""SIMD-lane attribution of k_pathtrace_mega (run on a B200): where the 32 lane slots of every warp iteration of the BVH
loop go — working, waiting for the slowest ray of the warp, or drained (the lane has no samples left) — and how often the
node / leaf blocks of an iteration are issued and with how many lanes.  Counted by the COUNT build of the same kernel
(same rays, same Philox streams), for the static pixel-per-lane binding and the balanced deal, next to the throughput of the
normal build.  Writes gpurun_out/lane_attribution.json and a markdown table (copied to profiles/ by hand).

    python tools/lane_attribution.py [c3:256 c5:64 c4:64 ...]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import blenderraytracer_b200 as brt
from bench import load_workload


def timed(rt, acc, spp, reps=3):
    best = 1e30
    for _ in range(reps + 1):
        rt.deviceMemset(acc.data_ptr(), 0, acc.numel() * 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rt.renderAccumulate(acc.data_ptr(), 0, spp)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


WIDTHS = [2]


def main():
    args = sys.argv[1:]
    if args and args[0].startswith("--widths="):
        WIDTHS[:] = [int(x) for x in args.pop(0).split("=")[1].split(",")]
    specs = args or ["c3:256", "c5:64", "c4:64", "c2:64"]
    out = []
    for spec in specs:
        name, spp = spec.split(":")[:2]
        spp = int(spp)
        w = load_workload(name, binary=True)
        if len(spec.split(":")) > 2:
            w["depth"] = int(spec.split(":")[2])           # c5:64:1 = primary rays only
        W, H = w["W"], w["H"]
        rt = brt.RayTracer(W, H, device=0, seed=1)
        assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode())
        rt.resizeCanvas(W, H)
        rt.updateRenderSettings(dict(samples=spp, maxBounces=w["depth"]))
        rt.directLighting = bool(w.get("direct"))
        rt.setStream(torch.cuda.current_stream().cuda_stream)
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        for sched in [("w%d" % x) for x in WIDTHS]:
            rt.bvhWidth = int(sched[1:])                        # 2 = the binary LBVH, 4 / 8 = its wide collapse
            rt.countTests = False
            rt._push_params()
            ms = timed(rt, acc, spp)
            rt.countTests = True
            rt._push_params()
            rt.deviceMemset(acc.data_ptr(), 0, acc.numel() * 4)
            rt.renderAccumulate(acc.data_ptr(), 0, spp)
            rt.synchronize()
            st = rt.stats()
            slots = 32 * st["trav_warp_iters"]
            row = dict(workload=name + (":d%d" % w["depth"]), spp=spp, schedule=sched, kernel_ms=ms, msamples_s=W * H * spp / ms / 1e3,
                       rays=st["rays"], rays_per_sample=st["rays"] / (W * H * spp),
                       node_visits_per_ray=st["node_visits"] / max(1, st["rays"]), box_tests_per_ray=st["tests_aabb"] / max(1, st["rays"]),
                       prim_tests_per_ray=(st["tests_sphere"] + st["tests_box"] + st["tests_tri_a"]) / max(1, st["rays"]),
                       trav_warp_iters=st["trav_warp_iters"],
                       frac_working=st["trav_lane_iters"] / max(1, slots),
                       frac_waiting=(st["trav_alive_lanes"] - st["trav_lane_iters"]) / max(1, slots),
                       frac_drained=(slots - st["trav_alive_lanes"]) / max(1, slots),
                       node_issue_frac=st["trav_node_issues"] / max(1, st["trav_warp_iters"]),
                       leaf_issue_frac=st["trav_leaf_issues"] / max(1, st["trav_warp_iters"]),
                       lanes_per_node_issue=(st["trav_lane_iters"] - st["trav_leaf_lanes"]) / max(1, st["trav_node_issues"]),
                       lanes_per_leaf_issue=st["trav_leaf_lanes"] / max(1, st["trav_leaf_issues"]),
                       path_lanes_per_iter=st["path_lane_iters"] / max(1, st["path_warp_iters"]),
                       trav_iters_per_path_iter=st["trav_warp_iters"] / max(1, st["path_warp_iters"]),
                       stats={k: int(v) if isinstance(v, int) else v for k, v in st.items()})
            out.append(row)
            print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items() if k != "stats"}), flush=True)
        rt.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lane_attribution.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "lane_attribution.md"), "w") as f:
        f.write("| workload | spp | hierarchy | Msamples/s | node visits / ray | box tests / ray | primitive tests / ray | working | waiting | drained | node-block issues / iter (lanes) | leaf-block issues / iter (lanes) | path-loop lanes |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in out:
            f.write(f"| {r['workload']} | {r['spp']} | {r['schedule']} | {r['msamples_s']:.0f} | {r['node_visits_per_ray']:.2f} | {r['box_tests_per_ray']:.2f} | {r['prim_tests_per_ray']:.2f} | {r['frac_working']:.3f} | {r['frac_waiting']:.3f} | {r['frac_drained']:.3f} | "
                    f"{r['node_issue_frac']:.2f} ({r['lanes_per_node_issue']:.1f}) | {r['leaf_issue_frac']:.2f} ({r['lanes_per_leaf_issue']:.1f}) | {r['path_lanes_per_iter']:.1f} |\n")


if __name__ == "__main__":
    main()
