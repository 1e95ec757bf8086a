#!/bin/bash
# two-GPU verification at HEAD: the multi-GPU tests (in-process ctx, torchrun, N-API create([0,1])) and the N = 2 bench line
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_napi_mock.py tests/test_js_shim.py -q > $O/gpu2_tests.log 2>&1; echo "2-gpu tests rc=$?"; tail -3 $O/gpu2_tests.log
bash tools/runs/run22.sh 2
