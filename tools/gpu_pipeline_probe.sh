# Per-step host issue / wait times of the pipelined resident loop (find and find_many plans): the probe that showed
# the resident step is stable at 1.4 ms when nothing allocates in the loop.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY'
import time, json, sys, os
sys.argv=['bench.py','--config','c2_many']
import bench, torch
args=bench.parse()
from unfazed_b200.engine import Engine, make_params
from unfazed_b200.phaser import BatchPhaser
ds=bench.make_ds(args,0,1)
eng=Engine(0)
for mpm in (1000, 10**9):
  with eng.on_stream():
    bp=BatchPhaser(eng, ds.sites, ds.reads, ds.pedigrees)
    params=make_params(readlen=151); cul=bp.cul(151,1000000,3)
    t0=time.perf_counter()
    plan,layout=bp.plan_batch(ds.dnms,[],threads=1,build="38",multiread_proc_min=mpm,search_dist=5000)
    print("mpm",mpm,"plan ms",(time.perf_counter()-t0)*1e3, "segs", plan.seg.shape, "mult max", int(plan.seg["mult"].max()), "nseg", int((plan.dnm["seg_hi"]-plan.dnm["seg_lo"]).sum()))
    eng.run(bp.dsites,bp.dreads,plan,params,blk_cul=cul,download=True,keep_device=False)
    def steps(k):
        pend=None; ts=[]
        for _ in range(k):
            t=time.perf_counter()
            h=eng.run(bp.dsites,bp.dreads,plan,params,blk_cul=cul,download=True,keep_device=False,defer=True)
            t1=time.perf_counter()
            if pend is not None: pend.finish()
            ts.append((round((t1-t)*1e3,2), round((time.perf_counter()-t1)*1e3,2)))
            pend=h
        pend.finish(); return ts
    steps(4)
    torch.cuda.synchronize(); t=time.perf_counter(); ts=steps(12); torch.cuda.synchronize(); print("ms/step",(time.perf_counter()-t)/12*1e3, "fallbacks", eng.spec_fallbacks); print(ts)
    # synchronous
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(12): eng.run(bp.dsites,bp.dreads,plan,params,blk_cul=cul,download=True,keep_device=False)
    torch.cuda.synchronize(); print("sync ms/step",(time.perf_counter()-t)/12*1e3)
    bp.release_device(); del bp
PY
