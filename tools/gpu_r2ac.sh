cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_gpu_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-sample 30 2>gpurun_out/r2ac.err | tail -1 > gpurun_out/r2ac_c2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ac_c2.json'))
print(round(d['value']), d['ms_per_step'], d['parity'], d['spec_fallbacks'], round(d['e2e']['value']), d['e2e']['breakdown_ms'], d['e2e']['single_call']['breakdown_ms'], d['clocks'])
print(d['roofline']['classify_sites_saturating'])
PY
tail -3 gpurun_out/r2ac.err
