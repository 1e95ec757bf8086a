set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/gpu_tests.log
python tools/ab.py base=blenderraytracer_b200/libbrt.so p7=blenderraytracer_b200/libbrt_p7.so -- c3:256 c5:64 c4:64 c2:64 c1:16
python tools/lane_attribution.py c3:256 c5:64 c4:64 c2:64 c1:16 > $O/lane_attr.log 2>&1; cat $O/lane_attribution.md
