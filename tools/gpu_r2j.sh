cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2j_multi_pytest.log
for c in c2 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 5 --warmup 3 2>gpurun_out/r2j_${c}_n2.err | tail -1 > gpurun_out/r2j_${c}_n2.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2j_${c}_n2.json'))
    print('$c n2', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), json.dumps(d['e2e']['breakdown_ms']), d['scaling'], d['n_gpus'])
except Exception as e:
    print('$c FAILED', e)
PY
  tail -3 gpurun_out/r2j_${c}_n2.err
done
