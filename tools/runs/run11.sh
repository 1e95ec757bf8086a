set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py ch0=$L/libbrt_ch0.so ch1=$L/libbrt_ch1.so f2=$L/libbrt.so -- c3:256 c5:64 c4:64 c2:64 c1:16 > $O/ab_ch.log 2>&1; echo "ab rc=$?"; cat $O/ab_ch.log
timeout 900 python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_parity.py -x -q -m gpu -k "invisible or wide or tie" > $O/ch_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/ch_tests.log
