mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02_test5a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test5a.log
timeout 1500 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_drivers_gpu.py tests/test_losscurve_gpu.py -m gpu -q -s > gpurun_out/r02_test5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test5.log
tail -5 gpurun_out/r02_test5a.log
grep -E "PARITY|LOSSCURVE|passed|failed|rc=|Error|error" gpurun_out/r02_test5.log | cut -c1-250 | head -60
