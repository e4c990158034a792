mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compact_eval_gpu.py -m gpu -q -s -x > gpurun_out/r02_test7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test7.log
tail -30 gpurun_out/r02_test7.log | cut -c1-250
