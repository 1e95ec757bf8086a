"""Turns gpurun_out/ artefacts of tools/gpu_round.sh into the tracked summaries under profiles/ (round-tagged)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r01"

def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    out, tot = [], {}
    for r in rows:
        name, ns = r[4].split("(")[0].replace("void ", ""), float(r[-1])
        out.append((int(r[0]), name, r[7], r[8], ns))
        tot[name] = tot.get(name, 0) + ns
    ours = {k: v for k, v in tot.items() if k.startswith("brt::")}
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write("# command: python bench.py --steps 2 --warmup 3 --spp 16 --no-cpu   (C3 scene at 16 spp)\n")
        f.write("id,kernel,block,grid,duration_ns\n")
        for r in out:
            f.write("%d,%s,\"%s\",\"%s\",%.0f\n" % r)
        f.write("# --- share of libbrt kernel time by kernel ---\n")
        s = sum(ours.values())
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
            f.write(f"# {k}: {v/1e6:.3f} ms  {100*v/s:.2f}%\n")

def raw_metrics(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "lts__t_sectors.sum.per_second", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]

def sass_hist(rep, chunk=60):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    lines = [f"kernel: {rows[0][1]}", f"SASS instructions: {len(data)}; warp-level instructions executed: {tot}",
             "offset  share%  active-threads/inst  stall-samples  dominant opcodes"]
    for c in range(0, len(data), chunk):
        seg = data[c:c + chunk]
        ie = sum(int(r[ix["Instructions Executed"]]) for r in seg)
        if ie == 0:
            continue
        te = sum(int(r[ix["Thread Instructions Executed"]]) for r in seg)
        smp = sum(int(r[ix["# Samples"]]) for r in seg)
        ops = {}
        for r in seg:
            t = r[ix["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:6])
        lines.append(f"{c:6d}  {100*ie/tot:6.2f}  {te/ie:6.1f}  {smp:8d}  {top}")
    return "\n".join(lines)

def ncu_summary(rep, dst, title):
    m = raw_metrics(rep)
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 (one launch; report kept in gpurun_out/, not tracked)\n")
        for k in KEYS:
            if k in m:
                f.write(f"{k} [{m[k][0]}] = {m[k][1]}\n")
        f.write("\n# ---- SASS execution histogram (ncu --page source): where the warp-instructions go and how full the warps are ----\n")
        f.write(sass_hist(rep) + "\n")
    return m

def sass_loop_excerpt(dst):
    """A trimmed cuobjdump -sass listing of the hot kernel's node-visit block (LDG.E.128 node loads, FFMA slabs, FMNMX / FMNMX3)."""
    obj = os.path.join(ROOT, "blenderraytracer_b200", "csrc", "build", "pathtrace.o")
    fn = "_ZN3brt16k_pathtrace_megaILi0ELb1ELb0ELb0ELb0ELi1ELi0EEEvNS_8PTParamsE"     # <fast, bvh, !count, !direct, !hybrid, PRIMS_SPHERE, binary>
    txt = subprocess.check_output(["cuobjdump", "-sass", "-fun", fn, obj], text=True)
    lines = [l for l in txt.splitlines() if "/*" in l and not l.strip().startswith("/* 0x")]
    lines = [l.split("/* 0x")[0].rstrip() for l in lines]
    ix = [i for i, l in enumerate(lines) if "FMNMX3" in l]
    lo = max(0, ix[0] - 40)
    while lo < ix[0] and "LDG.E.128" not in lines[lo]:
        lo += 1
    hi = min(len(lines), ix[-1] + 32)
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass of k_pathtrace_mega<fast, bvh, PRIMS_SPHERE> (sm_100a): the node-visit block of the traversal loop\n"
                "# (three 128-bit + one 64-bit read-only loads of the 64-byte centre / half-extent node, nine FFMA2 = the 18 slab planes of both children,\n"
                "#  FMNMX3 / FMNMX reductions, one FMUL2 widening, near / far select, shared-memory push, sentinel pop)\n")
        f.write("\n".join(lines[max(0, lo - 3):hi]) + "\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    if os.path.exists(os.path.join(G, "launches_c3_spp16.csv")):
        launches(os.path.join(G, "launches_c3_spp16.csv"), os.path.join(P, f"{R}_launches_c3_spp16.csv"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from source_hist import hist
    for rep, name, title, tag in ((f"{R}_prof_c3_mega.ncu-rep", "ncu_k_pathtrace_mega_c3_spp16", "k_pathtrace_mega<fast, bvh, spheres> on C3 (1920x1080 random spheres) at 16 spp", "c3"),
                                  (f"{R}_prof_c5_mega.ncu-rep", "ncu_k_pathtrace_mega_c5_spp16", "k_pathtrace_mega<fast, bvh, triangles> on C5 (1 M-triangle terrain, 3840x2160) at 16 spp", "c5")):
        src = os.path.join(G, rep)
        if not os.path.exists(src):
            continue
        m = ncu_summary(src, os.path.join(P, f"{R}_{name}.txt"), title)
        with open(os.path.join(P, f"{R}_source_hist_{tag}.txt"), "w") as f:
            f.write(f"# {title}: warp-level instructions by CUDA source line (ncu --page source --print-source cuda,sass), hottest 60\n" + hist(src, 60) + "\n")
        def b(k):
            u, v = m[k]; v = float(v)
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        json.dump({"kernel": "k_pathtrace_mega", "dram_bytes_per_launch": tr, "dram_bytes_read": b("dram__bytes_read.sum"),
                   "dram_bytes_write": b("dram__bytes_write.sum"),
                   "note": f"ncu --set full, one launch of the {tag.upper()} bench at 16 spp; the traffic is the W*H*16 B accumulation buffer read + write "
                           "(the L2 was flushed before the launch) plus scene data, so it does not grow with spp"},
                  open(os.path.join(P, f"traffic_{tag}.json"), "w"), indent=1)
    try:
        sass_loop_excerpt(os.path.join(P, f"{R}_sass_node_visit.txt"))
    except Exception as ex:
        print("sass excerpt failed:", ex)
    for f in sorted(os.listdir(G)):
        if f.startswith("bench_") and f.endswith(".json"):
            txt = open(os.path.join(G, f)).read().strip().splitlines()
            if txt:
                open(os.path.join(P, f"{R}_{f}"), "w").write(txt[-1] + "\n")
    print(sorted(os.listdir(P)))
