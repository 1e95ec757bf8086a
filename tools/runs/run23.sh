set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "reference" > $O/refvec_gpu.log 2>&1; echo "rc=$?"; tail -25 $O/refvec_gpu.log
