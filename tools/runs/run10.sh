set -u
O=gpurun_out; mkdir -p $O
BRT_LIBBRT=$PWD/blenderraytracer_b200/libbrt_dbg.so BRT_DEBUG=1 CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_wide_bvh.py -x -q -m gpu -k "ties" > $O/wide_tests.log 2>&1; echo "wide tests rc=$?"; grep -n "brt\]\|wide:" $O/wide_tests.log | head -30; tail -5 $O/wide_tests.log
