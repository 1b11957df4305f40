"""Development aid: run compute() (column mode, one 2 Mi-event slice, uniforms given on the
host) so that ncu can capture the HBM-bound stage kernels (K1/K2/K4 in column
mode).  See tools/column_mode_report.py for the capture command and the summary."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import nuspacesim_b200 as nss

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 21
cfg = nss.NssConfig()
rng = np.random.Generator(np.random.PCG64(5))
U = dict(u4=rng.random((4, n)), u_tau=rng.random(n), u_dec=rng.random(n))
for i in range(2):  # the second call is the one to read (tables cached, allocator warm)
    sim = nss.compute(cfg, uniforms=U, slice_events=0)
    torch.cuda.synchronize()
print("n_thrown", n, "n_valid", len(sim), "mcint", sim.meta["OMCINT"][0])
