mkdir -p gpurun_out
export MPNN_BENCH_NO_CPU=1
for B in 4096; do
  timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_e.json 2> /dev/null
  MPNN_LANE_PRIORITY=0 timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_e_noprio.json 2> /dev/null
  timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_e2.json 2> /dev/null
  MPNN_LANE_PRIORITY=0 timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_e2_noprio.json 2> /dev/null
done
B=4096 timeout 300 python scratch/mb_one.py > gpurun_out/r02_mb_one.txt 2>&1
B=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stencil_gemm_umma --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_k32n64 python scratch/mb_one.py > gpurun_out/r02_ncu_k32n64.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_e*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/r02_mb_one.txt
