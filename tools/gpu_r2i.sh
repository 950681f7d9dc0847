cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
(
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_w64.so python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_w84.so python tools/dbg_chain.py 10000
) 2>&1 | grep -v Warning | tee gpurun_out/r2i_scan.log
