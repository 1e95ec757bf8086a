set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gpu_tests.log
for w in c1 c2 c3; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-secondary > $O/bq_$w.json 2> $O/bq_$w.err; python - $O/bq_$w.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:30], "| value", round(d["value"],1), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"],1), "ratio", round(d["e2e"]["value"]/d["value"],3))
PY
done
