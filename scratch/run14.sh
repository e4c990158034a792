mkdir -p gpurun_out
export MPNN_BENCH_NO_CPU=1
B=4096 timeout 300 python scratch/mb_one.py > gpurun_out/r02_mb_one3.txt 2>&1
timeout 300 python bench.py --no-sweep --profile > gpurun_out/r02_b4096_g.json 2> gpurun_out/r02_b4096_g.txt
timeout 300 python bench.py --no-sweep --batch 128 > gpurun_out/r02_b128_g.json 2> /dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_g*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/r02_mb_one3.txt; head -4 gpurun_out/r02_b4096_g.txt
