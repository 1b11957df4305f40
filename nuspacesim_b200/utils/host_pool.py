"""Page-locked host memory for the result columns of compute().

The table compute() returns is 120 bytes per thrown event (15 float64 columns), all of it
copied device -> host.  Page-locking fresh memory through the CUDA allocator costs ~0.8 s per
GB, almost all of it first-touch page faults taken one 4 KB page at a time by one thread
(measured on the B200 box, tools/mb/pin_test2.py: cudaHostAlloc 1.5 GB/s; cudaHostRegister of
memory that is already faulted in 9 GB/s).  Here the memory comes from one anonymous mapping
per table that is advised to use huge pages, faulted in by a few threads in parallel (memset
through ctypes releases the GIL) and then registered with ``cudaHostRegister`` -- on a helper
thread, while the GPU already computes the first slices.  Blocks return to a small pool when
the table that used them is garbage-collected, so a process that calls compute() repeatedly
page-locks nothing after its first calls (what torch's caching host allocator did before).

Only compute() uses this; nothing here touches the data path of the kernels.
"""
from __future__ import annotations

import atexit
import ctypes
import mmap
import threading
import weakref

import numpy as np

from .. import _lib

#: free blocks kept for reuse (largest first is not needed: tables of one run are equal-sized)
MAX_POOLED_BLOCKS = 3
TOUCH_THREADS = 4

_lock = threading.Lock()
_free = []      # (nbytes, mm, addr, registered)
_live = {}      # addr -> (mm, nbytes, registered) of blocks currently lent out
_libc = ctypes.CDLL(None)
_libc.memset.restype = ctypes.c_void_p
_libc.memset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]


def _cudart():
    return _lib.torch_cuda().cuda.cudart()


def _new_block(nbytes):
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        mm.madvise(mmap.MADV_HUGEPAGE)
    except (AttributeError, OSError, ValueError):
        pass
    addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    # fault the pages in, in parallel
    step = -(-nbytes // TOUCH_THREADS)
    step = (step + 4095) & ~4095
    th = [threading.Thread(target=_libc.memset, args=(addr + lo, 0, min(step, nbytes - lo)))
          for lo in range(0, nbytes, step)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    rc = int(_cudart().cudaHostRegister(addr, nbytes, 0))
    return mm, addr, rc == 0


def _release(addr):
    """Finalizer of a lent block's base array: back to the pool, or unregister and unmap."""
    with _lock:
        ent = _live.pop(addr, None)
        if ent is None:
            return
        mm, nbytes, registered = ent
        if len(_free) < MAX_POOLED_BLOCKS:
            _free.append((nbytes, mm, addr, registered))
            return
    _drop(mm, addr, registered)


def _drop(mm, addr, registered):
    try:
        if registered:
            _cudart().cudaHostUnregister(addr)
    except Exception:
        pass
    try:
        mm.close()
    except (BufferError, ValueError):
        pass  # an exported view is still alive somewhere: the mapping goes with it


@atexit.register
def _cleanup():
    with _lock:
        blocks, _free[:] = list(_free), []
    for _, mm, addr, registered in blocks:
        _drop(mm, addr, registered)


class PendingColumns:
    """Host columns that may still be being page-locked on a helper thread.

    ``ready()`` tells whether ``columns`` can be used without waiting; ``wait()`` joins the
    helper.  ``columns`` maps each name to a float64 torch CPU tensor of length n; they are
    views of one block whose lifetime follows the NumPy arrays made from them."""

    def __init__(self, names, n):
        self.names, self.n = list(names), int(n)
        self.columns = None
        self.pinned = False
        nbytes = max(len(self.names) * self.n * 8, 8)
        with _lock:
            hit = next((i for i, b in enumerate(_free) if b[0] == nbytes), None)
            blk = _free.pop(hit) if hit is not None else None
        self._thread = None
        if blk is not None:
            self._adopt(blk[1], blk[2], blk[3], nbytes)
        else:
            torch = _lib.torch_cuda()
            dev = torch.cuda.current_device()

            def work():
                try:
                    torch.cuda.set_device(dev)
                    mm, addr, registered = _new_block(nbytes)
                    self._adopt(mm, addr, registered, nbytes)
                except Exception:  # no mapping (address space, memlock limit ...): plain memory
                    self.pinned = False
                    self.columns = {nm: torch.empty(self.n, dtype=torch.float64)
                                    for nm in self.names}

            self._thread = threading.Thread(target=work, name="nss-host-columns", daemon=True)
            self._thread.start()

    def _adopt(self, mm, addr, registered, nbytes):
        torch = _lib.torch_cuda()
        base = np.frombuffer(mm, dtype=np.float64, count=len(self.names) * self.n)
        with _lock:
            _live[addr] = (mm, nbytes, registered)
        weakref.finalize(base, _release, addr)
        self.pinned = registered
        self.columns = {nm: torch.from_numpy(base[i * self.n:(i + 1) * self.n])
                        for i, nm in enumerate(self.names)}

    def ready(self) -> bool:
        return self._thread is None or not self._thread.is_alive()

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        return self.columns
