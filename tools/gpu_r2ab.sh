cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-saturating --cpu-sample 10 2>gpurun_out/r2ab.err | tail -1 > gpurun_out/r2ab_c2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ab_c2.json'))
print(round(d['value']), d['ms_per_step'], d['parity'], d['spec_fallbacks'], round(d['e2e']['value']), d['e2e']['breakdown_ms'], d['clocks'])
PY
done
tail -3 gpurun_out/r2ab.err
