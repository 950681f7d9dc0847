cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 40 --no-saturating 2>gpurun_out/r2c_bench.err | tail -1 > gpurun_out/r2c_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['value'], d['ms_per_step'])
print(json.dumps(d['e2e']['breakdown_ms']), d['e2e']['value'])
print(d['roofline']['stages_ms'])
print(d.get('parity'), d.get('parity_detail'))
PY
tail -5 gpurun_out/r2c_bench.err
