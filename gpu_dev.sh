cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --dnms 4000 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_dev.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dev.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['roofline']['kernels'].items(): print(k, round(v['ms'],4), v.get('frac'))
PY
