set -u
O=gpurun_out; mkdir -p $O
L=$PWD/blenderraytracer_b200
BRT_LIBBRT=$L/libbrt_ch0.so timeout 900 python tools/bvh_exact_check.py c5 960 540 4 2 > $O/exact_ch0.log 2>&1; grep EXACT_CHECK $O/exact_ch0.log || tail -5 $O/exact_ch0.log
BRT_LIBBRT=$L/libbrt.so timeout 900 python tools/bvh_exact_check.py c5 960 540 4 2 4 8 > $O/exact_ch1.log 2>&1; grep EXACT_CHECK $O/exact_ch1.log || tail -5 $O/exact_ch1.log
