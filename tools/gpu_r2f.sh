cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_pytest.log
(
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_t64.so python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_t96.so python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_t64.so python tools/dbg_chain.py 150 50000 60
UNFZ_LIB=$L/libunfazed_sm100_t64.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
) 2>&1 | grep -v Warning | tee gpurun_out/r2f_chain.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 40 --no-saturating 2>gpurun_out/r2f_bench.err | tail -1 > gpurun_out/r2f_bench.json
UNFZ_NO_GRAPH=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-saturating 2>>gpurun_out/r2f_bench.err | tail -1 > gpurun_out/r2f_bench_nograph.json
python - <<'PY'
import json
for f in ('r2f_bench','r2f_bench_nograph'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(d['e2e']['breakdown_ms']), d.get('parity'))
    print(d['roofline']['stages_ms'])
PY
tail -5 gpurun_out/r2f_bench.err
