#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
tail -5 gpurun_out/sanitize_memcheck.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "primary_visibility" -s 2>&1 | tail -5
