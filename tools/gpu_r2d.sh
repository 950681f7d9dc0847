cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=unfazed_b200
(
UNFZ_LIB=$PWD/$L/libunfazed_sm100_dbg.so python tools/dbg_chain.py 4000
CH_OLD=1 UNFZ_LIB=$PWD/$L/libunfazed_sm100_olddbg.so python tools/dbg_chain.py 4000
python tools/dbg_chain.py 4000
UNFZ_LIB=$PWD/$L/libunfazed_sm100_old.so python tools/dbg_chain.py 4000
UNFZ_LIB=$PWD/$L/libunfazed_sm100_dbg.so python tools/dbg_chain.py 150 50000 60
CH_OLD=1 UNFZ_LIB=$PWD/$L/libunfazed_sm100_olddbg.so python tools/dbg_chain.py 150 50000 60
python tools/dbg_chain.py 150 50000 60
UNFZ_LIB=$PWD/$L/libunfazed_sm100_old.so python tools/dbg_chain.py 150 50000 60
) 2>&1 | grep -v Warning | tee gpurun_out/r2d_chain_phases.log
