set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 900 python tools/ab.py base=$L/libbrt.so hyb12=$L/libbrt.so,BRT_SMEM_STACK_MAX_DEPTH=12 hyb20=$L/libbrt.so,BRT_SMEM_STACK_MAX_DEPTH=20 -- c5:64 c3:64 > $O/ab_hyb.log 2>&1; echo "ab rc=$?"; cat $O/ab_hyb.log
timeout 1500 python tools/rare_event_check.py c5:64 ch0=$L/libbrt_ch0.so ch1=$L/libbrt.so > $O/rare_event.log 2>&1; echo rc=$?; grep RARE_EVENT $O/rare_event.log || tail -20 $O/rare_event.log
