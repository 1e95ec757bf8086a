set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_wide_bvh.py -q -m gpu -x > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gpu_tests.log
for w in c3 c2 c4; do BRT_DEBUG=1 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-secondary > $O/bq_$w.json 2> $O/bq_$w.err; grep "candidate" $O/bq_$w.err | sort | uniq -c | head -4; python - $O/bq_$w.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:30], "| value", round(d["value"],1), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"],1), "| bvh_build_ms", round(d["run"]["bvh_build_ms"],3), "nodes", d["run"]["bvh_nodes"], "depth", d["run"]["bvh_depth"])
PY
done
