cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python bench.py --dnms 2000 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_dev.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dev.json'))
print(d['roofline']['classify_sites_saturating'])
PY
free -g | head -2; nproc
