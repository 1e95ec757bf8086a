set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -25 $O/gpu_tests.log
timeout 900 python tools/parity_ids.py > $O/parity_ids.log 2>&1; echo "parity rc=$?"; tail -8 $O/parity_ids.log
