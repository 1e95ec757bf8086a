#!/bin/bash
# final single-GPU verification at HEAD: smoke, the whole GPU suite, the sanitize script without the tool, default bench + reference arm
cd /root/repo
O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gpu_tests.log
timeout 600 python tools/sanitize_small.py > $O/variants_small.log 2>&1; echo "variants rc=$?"; tail -2 $O/variants_small.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
python bench.py > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref_c3.json", "gpurun_out/bench_c3.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(d.get("impl","ours"), "value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "frac", r.get("frac"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
    s=d.get("secondary")
    if s: print("   secondary", round(s["value"],1), "e2e", round(s["e2e"]["value"],1))
PY
