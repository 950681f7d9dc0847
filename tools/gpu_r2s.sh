cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2s_pytest.log
(
python tools/dbg_chain.py 10000
for v in lkb8 lkb5; do
UNFZ_LIB=$L/libunfazed_sm100_$v.so python tools/dbg_chain.py 10000
done
) 2>&1 | grep -v Warning | tee gpurun_out/r2s_chain.log
