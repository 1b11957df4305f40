"""CUDA timeline of compute(config, uniforms=<pinned host arrays>) with torch.profiler
(development aid): per-kernel / per-memcpy device time and the idle gaps of one call."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import nuspacesim_b200 as nss

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cfg = nss.NssConfig()
rng = np.random.Generator(np.random.PCG64(1))
pinned = {}
for name, shape in (("u4", (4, n)), ("u_tau", (n,)), ("u_dec", (n,))):
    t = torch.empty(shape, dtype=torch.float64).pin_memory()
    t.numpy()[...] = rng.random(shape)
    pinned[name] = t
n_warm = int(os.environ.get("NSS_TL_WARM", "3"))   # un-profiled calls before the profiled one
n_after = int(os.environ.get("NSS_TL_AFTER", "0"))  # ... and after it (keeps the load up when
# several GPUs are profiled side by side: CUDA_VISIBLE_DEVICES=i python tools/e2e_timeline.py &)
for i in range(n_warm):
    t0 = time.perf_counter(); sim = nss.compute(cfg, uniforms=pinned); torch.cuda.synchronize()
    print("warm", i, time.perf_counter() - t0, flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    sim = nss.compute(cfg, uniforms=pinned)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("compute s", dt, "events/s", n / dt)
for i in range(n_after):
    t0 = time.perf_counter(); sim2 = nss.compute(cfg, uniforms=pinned); torch.cuda.synchronize()
    print("after", i, time.perf_counter() - t0, flush=True)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t_first = ev[0].time_range.start
rows = []
for e in ev:
    rows.append((e.time_range.start - t_first, e.time_range.end - t_first, e.name[:60]))
# merge same-name consecutive events
out = []
for s, e, nme in rows:
    if out and out[-1][2] == nme and s - out[-1][1] < 200:
        out[-1] = (out[-1][0], e, nme, out[-1][3] + 1)
    else:
        out.append((s, e, nme, 1))
for s, e, nme, c in out:
    print("%9.3f ms -> %9.3f ms  (%8.3f)  x%-3d %s" % (s / 1e3, e / 1e3, (e - s) / 1e3, c, nme))
