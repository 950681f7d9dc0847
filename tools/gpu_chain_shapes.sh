# Chaining time for the two compiled CTA shapes (UNFZ_CHAIN_SHAPE=small|wide) over window sizes: where the switch sits.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2w_pytest.log
(
python tools/dbg_chain.py 10000
python tools/dbg_chain.py 500 50000 60
for s in small wide; do
echo "forced $s"
UNFZ_CHAIN_SHAPE=$s python tools/dbg_chain.py 2000 20000 30
UNFZ_CHAIN_SHAPE=$s python tools/dbg_chain.py 4000 10000 30
done
) 2>&1 | grep -v Warning | tee gpurun_out/r2w_chain.log
