# gpu_suite_ngpu.sh N [configs...]: bench.py under an N-rank torchrun (default configs: c2 c5).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
shift
CFGS="${@:-c2 c5}"
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
for c in $CFGS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-saturating 2>gpurun_out/r2q_${c}_n$N.err | tail -1 > gpurun_out/r2_bench_${c}_${N}gpu.json
done
python - <<PY
import json
for c in "$CFGS".split():
    try:
        d=json.load(open('gpurun_out/r2_bench_%s_${N}gpu.json'%c))
        print(c, 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'single', round(d['e2e']['single_call']['value']), json.dumps(d['e2e']['breakdown_ms']))
    except Exception as e:
        print(c,'FAILED',e)
PY
for f in gpurun_out/r2q_*_n$N.err; do echo $f; tail -n 3 \$f 2>/dev/null; done
