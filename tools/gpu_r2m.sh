cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2m_pytest.log
python tools/dbg_chain.py 10000 2>&1 | grep -v Warning | tee gpurun_out/r2m_chain.log
python tools/dbg_chain.py 150 50000 60 2>&1 | grep -v Warning | tee -a gpurun_out/r2m_chain.log
