set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py prev=$L/libbrt_prev.so sent=$L/libbrt.so p7=$L/libbrt_p7.so -- c5:64 c3:256 c4:64 c2:64 > $O/ab_sent.log 2>&1; echo "ab rc=$?"; cat $O/ab_sent.log
timeout 900 python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu -k "invisible or wide or tie or deep or direct" > $O/ch_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/ch_tests.log
