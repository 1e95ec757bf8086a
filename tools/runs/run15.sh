set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py prev=$L/libbrt_prev.so k1=$L/libbrt.so,BRT_LANE_PIXELS=1 k2=$L/libbrt.so,BRT_LANE_PIXELS=2 k4=$L/libbrt.so,BRT_LANE_PIXELS=4 k8=$L/libbrt.so,BRT_LANE_PIXELS=8 auto=$L/libbrt.so -- c3:256 c3:64 c3:32 c5:32 c4:32 > $O/ab_lanepx.log 2>&1; echo "ab rc=$?"; cat $O/ab_lanepx.log
timeout 900 python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_edge_cases.py tests/test_gpu_variants.py -x -q -m gpu > $O/stream_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/stream_tests.log
