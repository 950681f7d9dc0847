cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2n_pytest.log
(
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_lk16.so python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_lk32.so python tools/dbg_chain.py 10000
) 2>&1 | grep -v Warning | tee gpurun_out/r2n_lk.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 40 --no-saturating 2>gpurun_out/r2n_bench.err | tail -1 > gpurun_out/r2n_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print(d['value'], d['ms_per_step'], d.get('parity'))
print(json.dumps(d['e2e'])[:1800])
PY
tail -5 gpurun_out/r2n_bench.err
