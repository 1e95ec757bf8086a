set -u
O=gpurun_out; mkdir -p $O
python tools/lane_attribution.py c3:256 c3:16 c5:64 c4:64 c2:64 > $O/lane_attr.log 2>&1; echo "attr rc=$?"; tail -12 $O/lane_attr.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/gpu_tests.log
