cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2t_pytest.log
(
python tools/dbg_chain.py 10000
UNFZ_LIB=$L/libunfazed_sm100_lkb6.so python tools/dbg_chain.py 10000
python tools/sat_classify.py
UNFZ_LIB=$L/libunfazed_sm100_clsb3.so python tools/sat_classify.py
UNFZ_LIB=$L/libunfazed_sm100_clsb5.so python tools/sat_classify.py
) 2>&1 | grep -v Warning | tee gpurun_out/r2t_chain.log
