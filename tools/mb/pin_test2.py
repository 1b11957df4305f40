"""Development aid: how fast can 1.2 GB of result columns reach host memory in a COLD process?
Measures (1) page-locking cost: sequential vs threaded torch pinned allocation, cudaHostRegister
of faulted / unfaulted numpy memory; (2) device -> pageable copy rate; (3) device -> small pinned
ring -> pageable memcpy rate (1 and 4 helper threads)."""
import ctypes
import threading
import time

import numpy as np
import torch

torch.cuda.init()
torch.zeros(1, device="cuda")
N, COLS = 10_000_000, 15
rt = ctypes.CDLL("libcudart.so.12")


def timed(label, fn):
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:60s} {dt:7.3f} s   {COLS * N * 8 / dt / 1e9:6.2f} GB/s", flush=True)
    return r


def seq_pinned():
    return [torch.empty(N, dtype=torch.float64, pin_memory=True) for _ in range(COLS)]


def par_pinned(k):
    out = [None] * COLS

    def work(i0):
        for i in range(i0, COLS, k):
            out[i] = torch.empty(N, dtype=torch.float64, pin_memory=True)

    th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    [t.start() for t in th]
    [t.join() for t in th]
    return out


def host_register(touch):
    arrs = [np.empty(N) for _ in range(COLS)]
    if touch:
        for a in arrs:
            a[::512] = 0.0
    t0 = time.perf_counter()
    for a in arrs:
        rc = rt.cudaHostRegister(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 0)
        assert rc == 0, rc
    dt = time.perf_counter() - t0
    for a in arrs:
        rt.cudaHostUnregister(ctypes.c_void_p(a.ctypes.data))
    return dt


a = timed("torch pinned, sequential (15 x 80 MB), cold", seq_pinned)
del a
torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
b = timed("torch pinned, 4 threads", lambda: par_pinned(4))
del b
torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
print("cudaHostRegister of untouched numpy memory: %.3f s" % host_register(False))
print("cudaHostRegister of touched numpy memory:   %.3f s" % host_register(True))

dev = [torch.rand(N, dtype=torch.float64, device="cuda") for _ in range(COLS)]
torch.cuda.synchronize()


def to_pageable():
    out = [torch.empty(N, dtype=torch.float64) for _ in range(COLS)]
    for o, d in zip(out, dev):
        o.copy_(d, non_blocking=True)
    return out


p = timed("device -> fresh pageable tensors (copy_)", to_pageable)
p = timed("device -> the same pageable tensors again", lambda: [o.copy_(d) for o, d in zip(p, dev)])


def ring_path(nthreads, chunk=1 << 20):
    ring = [torch.empty(chunk, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    out = [np.empty(N) for _ in range(COLS)]
    jobs = []
    sem = threading.Semaphore(0)
    done = threading.Event()

    def worker():
        while True:
            sem.acquire()
            if not jobs:
                return
            ev, src, dst = jobs.pop(0)
            ev.synchronize()
            np.copyto(dst, src.numpy()[: dst.size])

    # simple synchronous ring for the measurement: copy chunk -> pinned, memcpy -> pageable
    s = torch.cuda.Stream()
    t0 = time.perf_counter()
    k = 0
    pend = []
    for c in range(COLS):
        for lo in range(0, N, chunk):
            hi = min(lo + chunk, N)
            r = ring[k % 4]
            if len(pend) >= 4:
                ev, src, dst = pend.pop(0)
                ev.synchronize()
                np.copyto(dst, src.numpy()[: dst.size])
            with torch.cuda.stream(s):
                r[: hi - lo].copy_(dev[c][lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            pend.append((ev, r, out[c][lo:hi]))
            k += 1
    for ev, src, dst in pend:
        ev.synchronize()
        np.copyto(dst, src.numpy()[: dst.size])
    return time.perf_counter() - t0


dt = ring_path(1)
print(f"device -> 4 x 8 MB pinned ring -> fresh numpy (1 thread memcpy) {dt:7.3f} s  {COLS * N * 8 / dt / 1e9:6.2f} GB/s")
dt = ring_path(1)
print(f"same again                                                      {dt:7.3f} s  {COLS * N * 8 / dt / 1e9:6.2f} GB/s")
