"""Development aid: time repeated compute(config, uniforms=<pinned host arrays>) calls."""
import os, sys, time
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
ts = []
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = nss.compute(cfg, uniforms=pinned)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(os.environ.get("NSS_HOST_COLUMNS", "pool"), "pipeline", os.environ.get("NSS_PIPELINE_THROW", "1"),
      " ".join(f"{t:.4f}" for t in ts), "steady", n / np.mean(ts[3:]))
