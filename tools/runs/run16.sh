set -u
O=gpurun_out; mkdir -p $O
python tools/lane_attribution.py --widths=2,4,8 c3:256 c5:64 c4:64 c2:64 > $O/lane_attr.log 2>&1; echo "lane rc=$?"; cat $O/lane_attribution.md
