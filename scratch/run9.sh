mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bf16x3 or stencil or fused" > gpurun_out/r02_test9a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test9a.log
timeout 1500 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q -s -k "bf16x3" > gpurun_out/r02_test9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test9.log
tail -5 gpurun_out/r02_test9a.log | cut -c1-300
grep -E "PARITY|passed|failed|rc=|Error" gpurun_out/r02_test9.log | cut -c1-300 | head -40
export MPNN_BENCH_NO_CPU=1
timeout 300 python bench.py --precision bf16x3 --no-sweep --steps 30 --profile > gpurun_out/r02_bench9_x3.json 2> gpurun_out/r02_bench9_x3.txt; head -12 gpurun_out/r02_bench9_x3.txt; cut -c1-200 gpurun_out/r02_bench9_x3.json
timeout 300 python bench.py --precision bf16x3 --no-sweep --steps 50 --batch 128 > gpurun_out/r02_bench9_x3_b128.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench9_x3_b128.json
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"route|gather|scatter|compact|leaf_stats|softmax|talr|router_tail|pack_input|node_moments|fc_" --launch-skip 60 --launch-count 40 -o gpurun_out/r02_prof_route python scratch/mb_route.py > gpurun_out/r02_ncu_route.log 2>&1; tail -3 gpurun_out/r02_ncu_route.log
