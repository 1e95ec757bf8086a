#!/bin/bash
# JS shim end to end on the GPU, then the whole GPU suite
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_js_shim.py -x -q 2>&1 | tail -25 > gpurun_out/js_shim.log
cat gpurun_out/js_shim.log
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_suite.log
cat gpurun_out/gpu_suite.log
