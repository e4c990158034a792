mkdir -p gpurun_out
export MPNN_BENCH_NO_CPU=1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29521 bench.py --gpus 2 --steps 50 > gpurun_out/r02_dp2_b.json 2> gpurun_out/r02_dp2_b.err; echo "rc=$?" >> gpurun_out/r02_dp2_b.err
MPNN_DIST_OVERLAP=0 timeout 300 $R --master-port 29522 bench.py --gpus 2 --steps 50 > gpurun_out/r02_dp2_b_nooverlap.json 2> gpurun_out/r02_dp2_b_nooverlap.err; echo "rc=$?" >> gpurun_out/r02_dp2_b_nooverlap.err
timeout 300 python bench.py --steps 50 --no-sweep > gpurun_out/r02_dp1_b.json 2> /dev/null
timeout 300 python bench.py --steps 100 --batch 128 --no-sweep > gpurun_out/r02_dp1_b_b128.json 2> /dev/null
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/r02_test15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test15.log
tail -3 gpurun_out/r02_test15.log; tail -3 gpurun_out/r02_dp2_b.err | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_dp*_b*.json')):
    try:
        ls=[l for l in open(f) if l.startswith('{')]
        d=json.loads(ls[-1]); print(f, len(open(f).read().splitlines()), 'lines', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), [(c['batch_per_gpu'], round(c['value']), round(c['ms_per_step'],4)) for c in d['configs']])
    except Exception as e: print(f, 'ERR', e)
PY
