import time, torch, gc, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
def alloc():
    t0=time.perf_counter(); t=torch.empty(10_000_000, dtype=torch.float64, pin_memory=True); return t, time.perf_counter()-t0
torch.cuda.init()
t,dt=alloc(); print("first", dt)
a=t.numpy(); del t
t2,dt=alloc(); print("second while first alive", dt)
del a
t3,dt=alloc(); print("third after del numpy view", dt)
del t2, t3
t4,dt=alloc(); print("fourth after del tensors", dt)
# with a copy on side stream
s=torch.cuda.Stream(); d=torch.zeros(10_000_000, dtype=torch.float64, device="cuda")
with torch.cuda.stream(s):
    t4.copy_(d, non_blocking=True)
a=t4.numpy(); del t4
s.synchronize(); del a
t5,dt=alloc(); print("fifth after async copy + del", dt)
import nuspacesim_b200 as nss, numpy as np
cfg=nss.NssConfig(); n=2_000_000
rng=np.random.default_rng(1)
U=dict(u4=rng.random((4,n)),u_tau=rng.random(n),u_dec=rng.random(n))
for i in range(4):
    t0=time.perf_counter(); sim=nss.compute(cfg,uniforms=U); torch.cuda.synchronize(); print("compute",i,time.perf_counter()-t0, gc.get_count())
sim=None; gc.collect()
for i in range(3):
    t0=time.perf_counter(); sim=nss.compute(cfg,uniforms=U); torch.cuda.synchronize(); print("compute after gc",i,time.perf_counter()-t0)
    sim=None
