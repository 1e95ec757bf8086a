set -u
O=gpurun_out; mkdir -p $O
export BRT_BVH_WIDTH=4
CMD5="python bench.py --workload c5 --steps 2 --warmup 3 --spp 16 --no-cpu"
$CMD5 > $O/plain_c5w4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02b_prof_c5_w4 $CMD5 > $O/ncu_c5w4.log 2>&1
echo "ncu c5 w4 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --spp 16 --no-cpu --no-secondary"
$CMD > $O/plain_c3w4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pathtrace -s 3 -c 1 -f -o $O/r02b_prof_c3_w4 $CMD > $O/ncu_c3w4.log 2>&1
echo "ncu c3 w4 rc=$?"
