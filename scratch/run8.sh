mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_smi2.txt
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_compact_eval_gpu.py -m gpu -q -s > gpurun_out/r02_test8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test8.log
export MPNN_BENCH_NO_CPU=1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29511 bench.py --gpus 2 --steps 50 > gpurun_out/r02_dp2_a.json 2> gpurun_out/r02_dp2_a.err; echo "rc=$?" >> gpurun_out/r02_dp2_a.err
MPNN_DIST_OVERLAP=0 timeout 300 $R --master-port 29512 bench.py --gpus 2 --steps 50 > gpurun_out/r02_dp2_nooverlap.json 2> gpurun_out/r02_dp2_nooverlap.err; echo "rc=$?" >> gpurun_out/r02_dp2_nooverlap.err
MPNN_DIST_GRAPH=0 MPNN_DIST_OVERLAP=0 timeout 300 $R --master-port 29513 bench.py --gpus 2 --steps 50 > gpurun_out/r02_dp2_eager.json 2> gpurun_out/r02_dp2_eager.err; echo "rc=$?" >> gpurun_out/r02_dp2_eager.err
timeout 300 python bench.py --steps 50 --no-sweep > gpurun_out/r02_dp1_a.json 2> gpurun_out/r02_dp1_a.err
timeout 300 python bench.py --steps 100 --batch 128 --no-sweep > gpurun_out/r02_dp1_b128.json 2> /dev/null
mkdir -p /tmp/dp && cd /tmp/dp && timeout 300 $R --master-port 29514 $OLDPWD/multipath-nn_b200/train-nets cifar10-ac --synthetic --dp --n-iter 20 --t-log 10 --nets 0 --batch-size 256 > $OLDPWD/gpurun_out/r02_dp2_train.log 2>&1; echo "rc=$?" >> $OLDPWD/gpurun_out/r02_dp2_train.log; ls -R /tmp/dp/nets >> $OLDPWD/gpurun_out/r02_dp2_train.log; cd $OLDPWD
tail -8 gpurun_out/r02_test8.log | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_dp*.json')):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), [(c['batch_per_gpu'], round(c['value']), round(c['ms_per_step'],4)) for c in d['configs']])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:] if glob.glob(f.replace('.json','.err')) else '')
PY
tail -5 gpurun_out/r02_dp2_train.log
