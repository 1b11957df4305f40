"""Development aid: time front_kernel (fused) and tau_kernel (column mode) per 2 Mi events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nuspacesim_b200 as nss
cfg = nss.NssConfig()
n = 1 << 21
t = nss.Taus(cfg)
rng = np.random.default_rng(0)
b = torch.from_numpy(np.radians(rng.uniform(0.1, 41, n))).cuda()
le = torch.full((n,), 8.0, dtype=torch.float64, device="cuda")
u = torch.from_numpy(rng.random(n)).cuda()
for _ in range(3):
    t(b, le, u=u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    t(b, le, u=u)
e1.record(); torch.cuda.synchronize()
tau_us = e0.elapsed_time(e1) / 20 * 1e3
g = nss.RegionGeom(cfg, seed=1)
for _ in range(3):
    g.throw(n)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    g.throw(n)
e1.record(); torch.cuda.synchronize()
print(os.path.basename(os.environ.get("NSS_B200_LIB", "repo")), "tau() %.1f us" % tau_us,
      " throw(philox)+scan %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
