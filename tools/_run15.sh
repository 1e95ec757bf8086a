set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py prev=$L/libbrt_prev.so stream=$L/libbrt.so -- c3:256 c3:32 c5:64 c4:64 c2:64 c1:16 > $O/ab_stream.log 2>&1; echo "ab rc=$?"; cat $O/ab_stream.log
timeout 900 python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_edge_cases.py tests/test_gpu_variants.py -x -q -m gpu > $O/stream_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/stream_tests.log
