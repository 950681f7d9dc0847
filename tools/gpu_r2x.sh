cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2x_pytest.log
(
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_head.so python tools/dbg_chain.py 10000
python tools/dbg_chain.py 10000
) 2>&1 | grep -v Warning | tee gpurun_out/r2x_chain.log
