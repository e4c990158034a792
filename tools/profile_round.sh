#!/bin/bash
# One GPU call that regenerates the round's evidence under gpurun_out/ (copied into profiles/ afterwards):
#   tests, smoke, the default bench line + per-kernel table, the reference arm, B = 128 line, kernel timelines,
#   the ncu launch list of one step, ncu --set full captures of the dominant kernels, clocks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,driver_version --format=csv > gpurun_out/r02_nvidia_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py --profile > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_per_kernel.txt
MPNN_BENCH_NO_CPU=1 timeout 600 python bench.py --batch 128 --no-sweep --steps 300 --profile > gpurun_out/r02_bench_b128.json 2> gpurun_out/r02_bench_per_kernel_b128.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
B=128 timeout 300 python tools/timeline.py > gpurun_out/r02_timeline_b128.txt 2> /dev/null
B=4096 timeout 300 python tools/timeline.py > gpurun_out/r02_timeline_b4096.txt 2> /dev/null
B=128 timeout 300 python tools/critpath.py > gpurun_out/r02_critpath_b128.txt 2> /dev/null
B=4096 timeout 300 python tools/critpath.py > gpurun_out/r02_critpath_b4096.txt 2> /dev/null
MPNN_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-sweep > /dev/null 2> gpurun_out/r02_ncu_launch.err
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil_gemm_umma" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_conv_fwd python tools/mb_conv.py h32fwd > /dev/null 2>&1
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil_gemm_umma" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_conv_dgrad python tools/mb_conv.py h32dgrad > /dev/null 2>&1
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil_wgrad_umma" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_wgrad_h32 python tools/mb_conv.py h32wgrad > /dev/null 2>&1
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil_wgrad_umma" --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_wgrad_h8 python tools/mb_conv.py h8wgrad > /dev/null 2>&1
B=4096 timeout 900 ncu --set full --clock-control none -k regex:"bn_relu_pool_fwd|bn_relu_pool_bwd" --launch-skip 10 --launch-count 2 -o gpurun_out/r02_prof_bn python tools/mb_bn.py h32 > /dev/null 2>&1
# (what comes back is capped at 64 MiB: the 40-launch capture of the small kernels is exported to CSV here and dropped)
GRAPHS=0 B=4096 timeout 900 ncu --set full --clock-control none -k regex:"route_|router_tail|gather|scatter|compact|leaf_stats|softmax_ce|talr|node_moments" --launch-skip 40 --launch-count 40 -o /tmp/r02_prof_route python tools/mb_route.py > /dev/null 2>&1
ncu -i /tmp/r02_prof_route.ncu-rep --page raw --csv > gpurun_out/r02_ncu_routing_kernels_raw.csv 2> /dev/null
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/r02_prof_route.ncu-rep
du -sh gpurun_out
tail -3 gpurun_out/r02_gpu_tests.log; tail -2 gpurun_out/r02_smoke.log; head -c 300 gpurun_out/r02_bench_default.json; echo; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo; wc -l gpurun_out/r02_ncu_launch_list.csv; ls -la gpurun_out/*.ncu-rep | tail -8
