mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "cnv" -s > gpurun_out/r02_test16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test16.log
grep -E "worst|passed|failed|Error|assert" gpurun_out/r02_test16.log | cut -c1-250 | head -30
