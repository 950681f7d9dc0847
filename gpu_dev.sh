cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"classify_kernel" -s 10 -c 1 -o gpurun_out/prof_cls python bench.py --dnms 1000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ncu -i gpurun_out/prof_cls.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); hdr=r[0]
for row in r[2:]:
    for k in ['gpu__time_duration.sum','launch__grid_size','dram__bytes_read.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__registers_per_thread']:
        if k in hdr: print(k, row[hdr.index(k)])
"
