set -u
O=gpurun_out; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
for w in c1 c2 c4; do python bench.py --workload $w --steps 3 --warmup 3 --cpu-seconds 6 --no-secondary > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; done
python bench.py --workload c5 --spp 64 --steps 3 --warmup 3 --no-cpu > $O/bench_c5_spp64.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"
for f in $O/bench_c1.json $O/bench_c2.json $O/bench_c3.json $O/bench_c4.json $O/bench_c5_spp64.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(d["config"]["workload"][:40], "| value", round(d["value"],1), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"],1), "| frac", round(r.get("frac") or 0,4), "| grays", round(r.get("grays_per_s") or 0,2), "| cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value_1thread"))
s=d.get("secondary")
if s: print("   secondary", round(s["value"],1), "e2e", round(s["e2e"]["value"],1), "frac", round(s["roofline"]["frac"],4))
PY
done
