"""Development aid: host <-> device copy rates with all ranks of a box copying at once.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mb/pcie_n.py"""
import os, time
import torch, torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 60_000_000  # 480 MB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.zero_()
d = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(label, fn, reps=5, solo=False):
    world = dist.get_world_size()
    out = []
    for who in ([0, 4] if solo else [None]):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if who is None or dist.get_rank() == who:
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        if dist.get_rank() == 0:
            rates = [reps * n * 8 / float(x.item()) / 1e9 for x in g]
            if who is None:
                print(f"{label:28s} all ranks at once: " + " ".join(f"{r:5.1f}" for r in rates)
                      + f"  sum {sum(rates):6.1f} GB/s", flush=True)
            else:
                print(f"{label:28s} rank {who} alone: {rates[who]:5.1f} GB/s", flush=True)


def h2d(): d.copy_(h, non_blocking=True)
def d2h(): h.copy_(d, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)

h2 = torch.empty(n, dtype=torch.float64).pin_memory(); h2.zero_()
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
run("H2D 480 MB pinned", h2d, solo=True)
run("H2D 480 MB pinned", h2d)
run("D2H 480 MB pinned", d2h, solo=True)
run("D2H 480 MB pinned", d2h)
run("H2D + D2H concurrently", both)
dist.barrier(); dist.destroy_process_group()
