cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -4 | tee gpurun_out/r2_multi_2gpu.log
