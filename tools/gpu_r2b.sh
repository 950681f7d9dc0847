cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2b_bench.err | tail -1 > gpurun_out/r2b_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print(d['value'], d['ms_per_step'])
print(json.dumps(d['e2e'])[:1500])
print(d['roofline']['stages_ms'])
print(d.get('cpu_baseline'), d.get('parity'), d.get('parity_detail'))
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['read_allele_lookup'])
PY
tail -5 gpurun_out/r2b_bench.err
