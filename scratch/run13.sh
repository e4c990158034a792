mkdir -p gpurun_out
export MPNN_BENCH_NO_CPU=1
timeout 300 python bench.py --no-sweep --profile > gpurun_out/r02_b4096_f.json 2> gpurun_out/r02_b4096_f.txt
MPNN_TUNE_OCC=1 timeout 300 python bench.py --no-sweep > gpurun_out/r02_b4096_f_occ.json 2> /dev/null
timeout 300 python bench.py --no-sweep --batch 128 > gpurun_out/r02_b128_f.json 2> /dev/null
MPNN_TUNE_OCC=1 timeout 300 python bench.py --no-sweep --batch 128 > gpurun_out/r02_b128_f_occ.json 2> /dev/null
B=4096 timeout 300 python scratch/mb_one.py > gpurun_out/r02_mb_one2.txt 2>&1
B=4096 timeout 300 python scratch/mb_wgrad.py > gpurun_out/r02_mb_wgrad5.txt 2>&1
B=4096 timeout 300 python scratch/mb_conv.py r2 > gpurun_out/r02_mb_conv5.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_f*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/r02_mb_one2.txt gpurun_out/r02_mb_wgrad5.txt gpurun_out/r02_mb_conv5.txt; head -8 gpurun_out/r02_b4096_f.txt
