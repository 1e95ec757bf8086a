set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200/libbrt.so
timeout 900 python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_variants.py -x -q -m gpu > $O/pair_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/pair_tests.log
timeout 900 python tools/ab.py off=$L,BRT_PAIR_LEAVES=0 on=$L,BRT_PAIR_LEAVES=1 -- c3:256 c4:64 c2:64 c1:16 > $O/ab_pair.log 2>&1; echo "ab rc=$?"; cat $O/ab_pair.log
