set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py base=$L/libbrt.so b64=$L/libbrt_b64.so b256=$L/libbrt_b256.so -- c3:256 c3:32 c5:64 c4:64 c2:64 > $O/ab_block.log 2>&1; echo "ab rc=$?"; cat $O/ab_block.log
