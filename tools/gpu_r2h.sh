cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
(
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_nopf.so python tools/dbg_chain.py 10000
python tools/dbg_chain.py 150 50000 60
UNFZ_LIB=$L/libunfazed_sm100_nopf.so python tools/dbg_chain.py 150 50000 60
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -3
) 2>&1 | grep -v Warning | tee gpurun_out/r2h_chain.log
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --no-saturating 2>gpurun_out/r2h_c5.err | tail -1 > gpurun_out/r2h_c5.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2h_c5.json'))
print('c5', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), json.dumps(d['e2e']['breakdown_ms']), 'parity', d.get('parity'), d.get('parity_detail'))
print('   ', d['roofline']['stages_ms'])
print('   ', d['secondary'], d.get('cpu_baseline',{}).get('value'))
PY
tail -3 gpurun_out/r2h_c5.err
