set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c4_crop" > $O/c4crop.log 2>&1; echo "c4 crop rc=$?"; tail -3 $O/c4crop.log
CMD="python bench.py --steps 2 --warmup 3 --spp 16 --no-cpu --no-secondary"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c3_spp16.csv $CMD > $O/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02_prof_c3_mega $CMD > $O/ncu2.log 2>&1
echo "ncu full c3 rc=$?"
CMD5="python bench.py --workload c5 --steps 2 --warmup 3 --spp 16 --no-cpu --no-secondary"
$CMD5 > $O/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02_prof_c5_mega $CMD5 > $O/ncu_c5.log 2>&1
echo "ncu full c5 rc=$?"
ls -la $O/*.ncu-rep
