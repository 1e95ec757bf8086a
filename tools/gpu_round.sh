#!/bin/bash
# One GPU-box session that produces the evidence copied into profiles/ (run under gpurun; 1 GPU).
set -u
O=gpurun_out
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
for w in c1 c2 c4; do python bench.py --workload $w --steps 3 --warmup 3 --cpu-seconds 6 > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; done
python bench.py --workload c5 --spp 64 --steps 2 --warmup 3 --no-cpu > $O/bench_c5_spp64.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"
python bench.py --integrator wavefront --steps 3 --warmup 3 --no-cpu > $O/bench_c3_wavefront.json 2>/dev/null; echo "bench wavefront rc=$?"
python bench.py --accel brute --steps 2 --warmup 3 --spp 16 --no-cpu > $O/bench_c3_brute_spp16.json 2>/dev/null; echo "bench brute rc=$?"
# ncu: launch list (cold-cache, serialised: shares only) and one full capture of the dominant kernel, same command line
CMD="python bench.py --steps 2 --warmup 3 --spp 16 --no-cpu"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c3_spp16.csv $CMD > $O/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/prof_c3_mega $CMD > $O/ncu2.log 2>&1
echo "ncu full rc=$?"
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(d.get("impl","ours"), d["config"]["workload"][:60], "| value", round(d["value"],2), d["unit"], "| ms/step", round(d["ms_per_step"],2), "| e2e", round(d["e2e"]["value"],2), "| frac", r.get("frac"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("unreadable", e)
PY
done
