set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py base=$L/libbrt.so b9=$L/libbrt_b9.so b10=$L/libbrt_b10.so -- c3:256 c5:64 c4:64 c2:64 > $O/ab_misc2.log 2>&1; echo "ab rc=$?"; cat $O/ab_misc2.log
