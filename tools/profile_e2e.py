"""cProfile of compute(config, uniforms=<pinned host arrays>) (development aid)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nuspacesim_b200 as nss

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cfg = nss.NssConfig()
rng = np.random.Generator(np.random.PCG64(1))
pinned = {}
for name, shape in (("u4", (4, n)), ("u_tau", (n,)), ("u_dec", (n,))):
    t = torch.empty(shape, dtype=torch.float64).pin_memory()
    t.numpy()[...] = rng.random(shape)
    pinned[name] = t
for i in range(2):
    t0 = time.perf_counter(); sim = nss.compute(cfg, uniforms=pinned); torch.cuda.synchronize()
    print("warm", i, time.perf_counter() - t0)
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
sim = nss.compute(cfg, uniforms=pinned)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("compute s", dt, "events/s", n / dt)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
