cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export UNFZ_NO_GRAPH=1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturating"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"read_scan_warp|read_site_alleles|chain_setup|chain_bfs|chain_evidence|chain_size|classify_kernel" -s 14 -c 7 -o gpurun_out/r2_full_c2 $B > gpurun_out/r2_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chain_setup|chain_bfs|chain_evidence|read_scan_warp" -s 8 -c 4 -o gpurun_out/r2_full_c4 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-saturating > gpurun_out/r2_ncu_full_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2_ncu_full.log
