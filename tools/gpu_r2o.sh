cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2o_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 40 --no-saturating 2>gpurun_out/r2o_bench.err | tail -1 > gpurun_out/r2o_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['value'], d['ms_per_step'], d.get('parity'))
print(json.dumps(d['e2e'])[:1800])
PY
tail -5 gpurun_out/r2o_bench.err
