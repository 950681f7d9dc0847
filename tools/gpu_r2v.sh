cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/unfazed_b200
(
python tools/dbg_chain.py 500 50000 60
for v in t256b8 t256b4 t512b4 t512b2 t1024b2 t1024b1; do
UNFZ_LIB=$L/libunfazed_sm100_$v.so python tools/dbg_chain.py 500 50000 60
done
) 2>&1 | grep -v Warning | tee gpurun_out/r2v_chain.log
