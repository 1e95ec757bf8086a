set -u
O=gpurun_out; mkdir -p $O
BRT_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_wide_bvh.py -x -q -m gpu > $O/wide_tests.log 2>&1; echo "wide tests rc=$?"; grep "wide-" $O/wide_tests.log | sort | uniq | head -12; tail -4 $O/wide_tests.log
L=blenderraytracer_b200/libbrt.so
timeout 900 python tools/ab.py w2=$L,BRT_BVH_WIDTH=2 w4=$L,BRT_BVH_WIDTH=4 w8=$L,BRT_BVH_WIDTH=8 -- c3:256 c5:64 c4:64 c2:64 > $O/ab_wide.log 2>&1; echo "ab rc=$?"; cat $O/ab_wide.log
