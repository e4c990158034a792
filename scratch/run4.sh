mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r02_test4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test4.log
export MPNN_BENCH_NO_CPU=1
B=4096 timeout 300 python scratch/mb_wgrad.py > gpurun_out/r02_mb_wgrad4.txt 2>&1
B=4096 MPNN_TUNE_WGRAD_PER_SM=1 timeout 300 python scratch/mb_wgrad.py > gpurun_out/r02_mb_wgrad4_1persm.txt 2>&1
B=4096 timeout 300 python scratch/mb_conv.py r2 > gpurun_out/r02_mb_conv4.txt 2>&1
for B in 4096 128; do
  timeout 300 python bench.py --batch $B --profile > gpurun_out/r02_b${B}_c.json 2> gpurun_out/r02_b${B}_c.txt
done
tail -3 gpurun_out/r02_test4.log; cat gpurun_out/r02_mb_wgrad4.txt gpurun_out/r02_mb_wgrad4_1persm.txt gpurun_out/r02_mb_conv4.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_c.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
