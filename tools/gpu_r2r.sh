cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2r_pytest.log
(
python tools/dbg_chain.py 10000
for v in t64b32 t64b24 t128b12 t32b32 t256b8; do
UNFZ_LIB=$L/libunfazed_sm100_$v.so python tools/dbg_chain.py 10000
done
UNFZ_LIB=$L/libunfazed_sm100_t64b32.so python tools/dbg_chain.py 500 50000 60
python tools/dbg_chain.py 500 50000 60
) 2>&1 | grep -v Warning | tee gpurun_out/r2r_chain.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 40 --no-saturating 2>gpurun_out/r2r_bench.err | tail -1 > gpurun_out/r2r_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print(d['value'], d['ms_per_step'], d.get('parity'))
print(json.dumps(d['e2e'])[:1800])
PY
tail -5 gpurun_out/r2r_bench.err
