set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/gpu_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -30 $O/gpu_multi_tests.log
timeout 600 python tools/mgpu_check.py --inprocess 2 > $O/mgpu_inprocess_n2.log 2>&1; echo "inprocess rc=$?"; tail -12 $O/mgpu_inprocess_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/mgpu_check.py > $O/mgpu_torchrun_n2.log 2>&1; echo "torchrun rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*" $O/mgpu_torchrun_n2.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('C3 N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3), d.get('image_check'))
s=d.get('secondary')
if s: print('C5 N=2 value',round(s['value']),'e2e',round(s['e2e']['value']),'ms',round(s['ms_per_step'],2),'kernel',round(s['roofline']['kernel_ms'],2), s.get('image_check'))
PY
