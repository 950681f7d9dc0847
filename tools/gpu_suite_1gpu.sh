# The 1-GPU measurement suite of a round (run on a B200 box through gpurun): GPU tests, every bench config with
# --full-parity where the oracle finishes in a minute, the reference arm.  Outputs under gpurun_out/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2 | tail -1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_gpu_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --full-parity 2>gpurun_out/r2y_c2.err | tail -1 > gpurun_out/r2_bench_c2_1gpu.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 2>gpurun_out/r2y_ref.err | tail -1 > gpurun_out/r2_bench_reference_arm_c2.json
timeout 900 python bench.py --config c2_many --steps 10 --warmup 3 --no-saturating 2>gpurun_out/r2y_c2_many.err | tail -1 > gpurun_out/r2_bench_c2_many_1gpu.json
for c in c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-saturating --full-parity 2>gpurun_out/r2y_$c.err | tail -1 > gpurun_out/r2_bench_${c}_1gpu.json
done
python - <<'PY'
import json
for c in ('c2','c2_many','c3','c4','c5'):
    try:
        d=json.load(open('gpurun_out/r2_bench_%s_1gpu.json'%c))
        r=d['roofline']
        print(c, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'single', round(d['e2e']['single_call']['value']), 'parity', d.get('parity'), d.get('parity_full'), 'cpu1', round(d.get('cpu_baseline',{}).get('value',0),1), 'roof', r['kernel'], round(r['frac'] or 0,3), 'lookup', round(r['read_allele_lookup']['frac'] or 0,3))
        print('    ', r['stages_ms'])
        print('    ', json.dumps(d['e2e']['breakdown_ms']))
    except Exception as e:
        print(c,'FAILED',e)
d=json.load(open('gpurun_out/r2_bench_reference_arm_c2.json'))
print('ref', d['value'], d['cpu_baseline']['cores'], d.get('threads_arm'))
PY
for f in gpurun_out/r2y_*.err; do echo $f; tail -n 2 $f; done
