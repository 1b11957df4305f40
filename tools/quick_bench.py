"""Quick GPU timing of the acceptance pipeline (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nuspacesim_b200 as nss

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 20
cfg = nss.NssConfig()
if len(sys.argv) > 2 and sys.argv[2]:  # longitudinal profile, e.g. "Gaisser-Hillas Fluctuated"
    cfg.simulation.eas_long_profile = sys.argv[2]
stats = os.environ.get("NSS_QB_STATS", "1") != "0"  # 0: the path without work counters
r = nss.AcceptanceRunner(cfg, batch=1 << 20, collect_stats=stats)
r.enqueue(1234, 0, 1 << 16)
torch.cuda.synchronize()
r.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r.enqueue(1234, 0, n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
s = r.sums(False)
print("n", n, "ms", ms, "events/s", n / ms * 1e3)
print("sums", s, "result", r.finalize(s))
if stats:
    st = r.stats.cpu().numpy()
    print("stats per EAS event", st[1:] / max(st[0], 1), "n_eas", st[0])
