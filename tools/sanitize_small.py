"""Small run of every kernel variant, for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nuspacesim_b200 as nss
from nuspacesim_b200.config import MonoCloud, PowerSpectrum, PressureMapCloud

n = 3001
for prof in ("Greisen", "Gaisser-Hillas Parameterized", "Gaisser-Hillas Fluctuated"):
    for cloud in (None, MonoCloud(altitude=4.0), PressureMapCloud(month=3)):
        cfg = nss.NssConfig()
        cfg.simulation.thrown_events = n
        cfg.simulation.eas_long_profile = prof
        if cloud is not None:
            cfg.simulation.cloud_model = cloud
        if prof != "Greisen":
            cfg.simulation.spectrum = PowerSpectrum(index=2.0, lower_bound=7.0, upper_bound=10.0)
        sim = nss.compute(cfg, seed=5)
        acc = nss.run_acceptance(cfg, n, seed=5, batch=1024)
        assert sim.meta["ONEVPASS"][0] == acc[2], (prof, cloud, sim.meta["ONEVPASS"], acc)
        print(prof, getattr(cloud, "id", None), len(sim), acc)
rng = np.random.default_rng(1)
u = dict(u4=rng.random((4, 777)), u_tau=rng.random(777), u_dec=rng.random(777))
print(len(nss.compute(nss.NssConfig(), uniforms=u)))
print("sanitize_small ok")
