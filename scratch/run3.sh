mkdir -p gpurun_out
B=128 timeout 300 python scratch/mb_chain.py > gpurun_out/r02_mb_chain.txt 2>&1
B=128 MPNN_PDL=0 timeout 300 python scratch/mb_chain.py > gpurun_out/r02_mb_chain_nopdl.txt 2>&1
B=4096 timeout 300 python scratch/mb_wgrad.py > gpurun_out/r02_mb_wgrad.txt 2>&1
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stencil_wgrad_umma --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_wgrad_a python scratch/mb_wgrad.py > gpurun_out/r02_ncu_wgrad.log 2>&1
B=128 timeout 300 python scratch/critpath.py > gpurun_out/r02_critpath_b128.txt 2>&1
cat gpurun_out/r02_mb_chain.txt gpurun_out/r02_mb_wgrad.txt
