set -u
O=gpurun_out; mkdir -p $O
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"; tail -3 $O/bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
print('C3 N=',n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3), d.get('image_check',{}).get('max_lsb_diff'), d['clocks'])
s=d.get('secondary')
if s: print('C5 N=',n,'value',round(s['value']),'e2e',round(s['e2e']['value']),'ms',round(s['ms_per_step'],2),'kernel',round(s['roofline']['kernel_ms'],2), s.get('image_check',{}).get('max_lsb_diff'))
PY
if [ "$N" = "8" ]; then
timeout 600 python tools/mgpu_check.py --inprocess 8 > $O/mgpu_inprocess_n8.log 2>&1; echo "inprocess8 rc=$?"; tail -3 $O/mgpu_inprocess_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 tools/mgpu_check.py > $O/mgpu_torchrun_n8.log 2>&1; echo "torchrun8 rc=$?"; grep "world=\|MGPU" $O/mgpu_torchrun_n8.log | tail -5
fi
