mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_test1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test1.log
export MPNN_BENCH_NO_CPU=1
timeout 300 python bench.py --profile > gpurun_out/r02_b4096_a.json 2> gpurun_out/r02_b4096_a.txt
timeout 300 python bench.py --batch 128 > gpurun_out/r02_b128_a.json 2> gpurun_out/r02_b128_a.err
MPNN_FUSE_BNRED=0 timeout 300 python bench.py > gpurun_out/r02_b4096_nofuse.json 2> /dev/null
MPNN_TUNE_GENERIC=1 MPNN_FUSE_BNRED=0 timeout 300 python bench.py --profile > gpurun_out/r02_b4096_generic.json 2> gpurun_out/r02_b4096_generic.txt
B=4096 timeout 300 python scratch/mb_conv.py r2 > gpurun_out/r02_mb_conv.txt 2>&1
B=4096 timeout 300 python scratch/mb_bn.py > gpurun_out/r02_mb_bn.txt 2>&1
B=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stencil_gemm_umma -c 6 -o gpurun_out/r02_prof_conv python scratch/mb_conv.py r2 > gpurun_out/r02_ncu_conv.log 2>&1
tail -3 gpurun_out/r02_test1.log; cat gpurun_out/r02_b4096_a.json | cut -c1-400
