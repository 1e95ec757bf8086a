set -u
O=gpurun_out; mkdir -p $O
python tools/parity_vs_reference.py > $O/parity_vs_reference.json 2> $O/parity_vs_reference.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_vs_reference.json'))['cases']
for k,v in d.items(): print(f"{k:42s} median {v['median_abs_linear_err']:.2e} p99 {v['p99_abs_linear_err']:.2e} <=1LSB {v['within_1_lsb']:.4f} <=2LSB {v['within_2_lsb']:.4f} identical {v['identical_rgba8']:.4f}")
PY
