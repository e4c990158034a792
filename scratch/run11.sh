mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/r02_test11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test11.log
tail -3 gpurun_out/r02_test11.log
export MPNN_BENCH_NO_CPU=1
for B in 4096 128; do
  timeout 300 python bench.py --batch $B --no-sweep --profile > gpurun_out/r02_b${B}_d.json 2> gpurun_out/r02_b${B}_d.txt
  MPNN_LANE_PRIORITY=0 timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_d_noprio.json 2> /dev/null
  MPNN_ROUTER_CLUSTER=8 timeout 300 python bench.py --batch $B --no-sweep > gpurun_out/r02_b${B}_d_cl8.json 2> /dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_d*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['ms_per_step'],4), d['families'].get('router_tail_bwd_batched'))
    except Exception as e: print(f, 'ERR', e)
PY
grep "conv_fwd    H8 K32+0 N64\|router_tail" gpurun_out/r02_b4096_d.txt
