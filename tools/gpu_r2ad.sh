cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config c5 --steps 10 --warmup 3 --no-saturating 2>gpurun_out/r2ad_c5.err | tail -1 > gpurun_out/r2_bench_c5_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config c2 --dnms 4000 --steps 20 --warmup 5 --no-saturating 2>gpurun_out/r2ad_c2.err | tail -1 > gpurun_out/r2ad_c2_4000dnms_2gpu.json
python - <<'PY'
import json
for f in ('r2_bench_c5_2gpu','r2ad_c2_4000dnms_2gpu'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), json.dumps(d['e2e']['breakdown_ms']), d['spec_fallbacks'])
    except Exception as e:
        print(f,'FAILED',e)
PY
tail -2 gpurun_out/r2ad_c5.err gpurun_out/r2ad_c2.err
