set -u
O=gpurun_out; mkdir -p $O
: > $O/exact_all.log
timeout 600 python tools/bvh_exact_check.py c3 1920 1080 16 2 4 8 2>&1 | grep EXACT_CHECK >> $O/exact_all.log
timeout 600 python tools/bvh_exact_check.py c4 1920 1080 8 2 4 8 2>&1 | grep EXACT_CHECK >> $O/exact_all.log
timeout 600 python tools/bvh_exact_check.py c2 1280 720 16 2 4 8 2>&1 | grep EXACT_CHECK >> $O/exact_all.log
timeout 900 python tools/bvh_exact_check.py c5 960 540 4 2 4 8 2>&1 | grep EXACT_CHECK >> $O/exact_all.log
cat $O/exact_all.log
