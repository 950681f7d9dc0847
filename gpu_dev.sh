cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python bench.py --dnms 4000 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_dev.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dev.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['roofline']['kernels'].items(): print(k, round(v['ms'],4), v.get('frac'))
PY
