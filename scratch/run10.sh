mkdir -p gpurun_out
B=4096 timeout 300 python scratch/critpath.py > gpurun_out/r02_critpath_b4096.txt 2>&1
B=128 timeout 300 python scratch/critpath.py > gpurun_out/r02_critpath_b128.txt 2>&1
B=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"move_images|route_compact|leaf_stats|compact_paths" -c 30 -o gpurun_out/r02_prof_gather python scratch/mb_route.py > gpurun_out/r02_ncu_gather.log 2>&1
grep "critical path\|==" gpurun_out/r02_critpath_b4096.txt gpurun_out/r02_critpath_b128.txt
