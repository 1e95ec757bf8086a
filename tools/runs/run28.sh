# Round-2 evidence at the FINAL build (r02c) (1 GPU): tests, smoke, bench lines of every BASELINE config, launch list, ncu captures, lane attribution
set -u
O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gpu_tests.log
python bench.py --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
for w in c1 c2 c4; do python bench.py --workload $w --steps 3 --warmup 3 --cpu-seconds 6 --no-secondary > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; done
python bench.py --workload c5 --spp 64 --steps 3 --warmup 3 --no-cpu > $O/bench_c5_spp64.json 2> $O/bench_c5.err; echo "bench c5 rc=$?"
for f in $O/bench_c1.json $O/bench_c2.json $O/bench_c3.json $O/bench_c4.json $O/bench_c5_spp64.json $O/bench_ref_c3.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get("roofline") or {}
print(d.get("impl","ours"), d["config"]["workload"][:40], "| value", round(d["value"],2), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"],2), "| frac", round(r.get("frac") or 0,4), "| grays", round(r.get("grays_per_s") or 0,2), "| visits/ray", r.get("node_visits_per_ray"), "| cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value_1thread"))
s=d.get("secondary")
if s: print("   secondary", round(s["value"],1), "e2e", round(s["e2e"]["value"],1), "frac", round(s["roofline"]["frac"],4))
PY
done
CMD="python bench.py --steps 2 --warmup 3 --spp 16 --no-cpu --no-secondary"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c3_spp16.csv $CMD > $O/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02c_prof_c3_mega $CMD > $O/ncu2.log 2>&1
echo "ncu c3 rc=$?"
CMD5="python bench.py --workload c5 --steps 2 --warmup 3 --spp 16 --no-cpu"
$CMD5 > $O/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02c_prof_c5_mega $CMD5 > $O/ncu_c5.log 2>&1
echo "ncu c5 rc=$?"
python tools/lane_attribution.py --widths=2,4,8 c3:256 c5:64 c4:64 c2:64 > $O/lane_attr.log 2>&1; echo "lane rc=$?"
