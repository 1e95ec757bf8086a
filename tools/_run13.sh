set -u
O=gpurun_out; mkdir -p $O
L=blenderraytracer_b200
timeout 1500 python tools/rare_event_check.py c5:64 ch0=$L/libbrt_ch0.so ch1=$L/libbrt.so > $O/rare_event.log 2>&1; echo rc=$?; grep RARE_EVENT $O/rare_event.log || tail -20 $O/rare_event.log
