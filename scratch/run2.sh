mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r02_test2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test2.log
export MPNN_BENCH_NO_CPU=1
B=4096 timeout 300 python scratch/mb_conv.py r2 > gpurun_out/r02_mb_conv2.txt 2>&1
B=4096 timeout 300 python scratch/mb_bn.py > gpurun_out/r02_mb_bn2.txt 2>&1
B=4096 MPNN_BN_V2=0 timeout 300 python scratch/mb_bn.py > gpurun_out/r02_mb_bn2_v1.txt 2>&1
for B in 4096 128; do
  timeout 300 python bench.py --batch $B --profile > gpurun_out/r02_b${B}_b.json 2> gpurun_out/r02_b${B}_b.txt
  MPNN_FUSE_BNRED=0 timeout 300 python bench.py --batch $B > gpurun_out/r02_b${B}_b_nofuse.json 2> /dev/null
  MPNN_BN_V2=0 timeout 300 python bench.py --batch $B > gpurun_out/r02_b${B}_b_bnv1.json 2> /dev/null
done
tail -3 gpurun_out/r02_test2.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_b*_b*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],4))
    except Exception as e: print(f, 'ERR', e)
PY
