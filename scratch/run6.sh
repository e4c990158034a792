mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_test6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_test6.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke6.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke6.log
timeout 900 python bench.py --profile > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.txt; echo "bench rc=$?" >> gpurun_out/r02_bench6.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench6_ref.json 2> gpurun_out/r02_bench6_ref.err
tail -4 gpurun_out/r02_test6.log; tail -4 gpurun_out/r02_smoke6.log; tail -3 gpurun_out/r02_bench6.txt; cut -c1-600 gpurun_out/r02_bench6_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench6.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])
for c in d['configs']: print(c['config'], c['batch_per_gpu'], c['dtype'], round(c['value']), round(c['ms_per_step'],4), round(c['tensor_frac_of_step'],4), round(c['hbm_frac_of_step'],4), c['launches_per_step'])
PY
