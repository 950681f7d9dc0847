# ncu launch list (default graph path), ncu --set full of the main kernels on c2 and c4 (per-call path), and
# compute-sanitizer memcheck / racecheck.  Summaries: tools/ncu_summary.py -> profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-saturating"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2f_launches_c2_graph.csv $B > gpurun_out/r2f_ncu_launch.log 2>&1
export UNFZ_NO_GRAPH=1
ncu --set full --clock-control none --import-source on -k regex:"read_scan_warp|read_site_alleles|chain_setup|chain_bfs|chain_evidence|chain_size|classify_kernel" -s 21 -c 7 -o gpurun_out/r2f_full_c2 $B > gpurun_out/r2f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chain_setup|chain_bfs|chain_evidence|read_scan_warp" -s 12 -c 4 -o gpurun_out/r2f_full_c4 python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-saturating > gpurun_out/r2f_ncu_full_c4.log 2>&1
unset UNFZ_NO_GRAPH
ls -la gpurun_out/r2f*.ncu-rep
(
echo "# compute-sanitizer on a B200 (final round-2 build)"
echo "## memcheck: python -m pytest tests/test_gpu_parity.py"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "Warning\|warn" | tail -6
echo "## memcheck: smoke() twice"
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); g.smoke(); print('smoke ok')" 2>&1 | tail -5
echo "## racecheck: smoke() twice"
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke(); g.smoke(); print('smoke ok')" 2>&1 | tail -5
) > gpurun_out/r2f_sanitizer.txt 2>&1
tail -30 gpurun_out/r2f_sanitizer.txt
