set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 4 > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench rc=$?"; tail -3 $O/bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
r=d['roofline']
print('C3 value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(r['frac'],4),'grays',round(r['grays_per_s'],2),'lanes',r['lane_slots'])
s=d.get('secondary')
if s: print('C5 value',round(s['value']),'e2e',round(s['e2e']['value']),'frac',round(s['roofline']['frac'],4),'nv/ray',round(s['roofline']['node_visits_per_ray'],2), s['run'])
print(d['cpu_baseline']['value'], d['cpu_baseline']['js'][:60])
PY
