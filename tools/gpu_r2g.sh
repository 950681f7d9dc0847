cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in c3 c4 c5 c2_many; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-saturating 2>gpurun_out/r2g_$c.err | tail -1 > gpurun_out/r2g_$c.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2g_$c.json'))
    print('$c', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), json.dumps(d['e2e']['breakdown_ms']), 'parity', d.get('parity'), d.get('parity_detail'))
    print('   ', d['roofline']['stages_ms'])
    print('   ', d['secondary'], d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('$c FAILED', e)
PY
  tail -3 gpurun_out/r2g_$c.err
done
