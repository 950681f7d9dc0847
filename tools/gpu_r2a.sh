cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
print(d['roofline']['stages_ms'])
PY
tail -5 gpurun_out/r2a_bench.err
