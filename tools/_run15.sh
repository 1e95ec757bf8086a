set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py base=$L/libbrt.so sah=$L/libbrt.so,BRT_SAH_MARGIN=2.0 p7=$L/libbrt_p7.so -- c3:256 c2:64 > $O/ab_sah.log 2>&1; echo "ab rc=$?"; cat $O/ab_sah.log
BRT_LIBBRT=$PWD/$L/libbrt_p7.so timeout 1500 python -m pytest tests -q -m gpu -x > $O/p7_tests.log 2>&1; echo "p7 tests rc=$?"; tail -5 $O/p7_tests.log
