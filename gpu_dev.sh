cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_full.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'], "gen", d['secondary']['dataset_gen_s'])
for k,v in d['roofline']['kernels'].items(): print(k, round(v['ms'],4), v.get('frac'))
print(d['roofline']['classify_sites_saturating'])
print(d['roofline']['read_allele_lookup_survey_bytes'])
print(d['cpu_baseline']); print(d['clocks'], d['gpu_launches'], d['roofline']['kernel'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --dnms 10000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"read_scan_pipe_kernel|classify_kernel|chain_kernel|chain_size_kernel|read_site_alleles_kernel|compact_kernel" -s 18 -c 9 -o gpurun_out/prof_r1c python bench.py --dnms 10000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json
cut -c1-300 gpurun_out/bench_ref.json
