#!/usr/bin/env python
"""Development tool: compile nss_eas_core.cuh for the host (tests/host_emul, W = 1) -- optionally
with extra -D flags for an experimental variant -- and report its agreement with the oracle
(exact sum), the unmodified reference (31 k events, tests/golden/eas_large.npz) and a
previously saved variant, plus the work counters.  The oracle columns are cached in
/tmp/nss_emul_eval_cache.npz.

    python tools/emul_eval.py [-D NAME[=V] ...] [--save tag] [--against tag] [--sets c1,e10]
"""
from __future__ import annotations

import argparse
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/nss_emul_eval_cache.npz"
TAGS = ("c1", "e7", "e9", "e10", "pl10")


def rel(a, b):
    with np.errstate(all="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    r[a == b] = 0
    return r


def oracle_cache():
    if os.path.exists(CACHE):
        return dict(np.load(CACHE))
    from oracle import nss_oracle as O

    G = np.load(os.path.join(ROOT, "tests", "golden", "eas_large.npz"))
    out = {}
    nt = os.cpu_count() or 1
    for tag in TAGS:
        b, a, e = (G[f"{tag}_{k}"] for k in ("beta_rad", "altDec", "showerEnergy"))
        d0, c0, _ = O.cphotang(b, a, e, 525.0, None, nthreads=nt)
        dr, _, _ = O.cphotang(b, a, e, 525.0, None, nthreads=nt, accum="reference")
        out[f"{tag}_od"], out[f"{tag}_oc"], out[f"{tag}_odr"] = d0, c0, dr
        out[f"{tag}_sens"] = O.conditioning(b, a, e, 525.0, None, nthreads=nt)
    np.savez(CACHE, **out)
    return out


def build(defs, tag, csrc=None):
    src = os.path.join(ROOT, "tests", "host_emul", "eas_host_emul.cpp")
    lib = f"/tmp/libeas_emul_{tag}.so"
    cmd = ["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-fopenmp", "-I",
           csrc or os.path.join(ROOT, "nuspacesim_b200", "csrc"), src, "-o", lib]
    cmd += [f"-D{d}" for d in defs]
    subprocess.check_call(cmd)
    return ctypes.CDLL(lib)


def run(L, beta, alt, es):
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    beta, alt, es = (np.ascontiguousarray(x, float) for x in (beta, alt, es))
    n = beta.size
    npe, cc, d, c = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    st = (ctypes.c_ulonglong * 6)()
    L.nss_emul_eas_profile(ctypes.c_longlong(n), dp(beta), dp(alt), dp(es), ctypes.c_double(525.0),
                           ctypes.c_double(2.5), ctypes.c_double(0.2), ctypes.c_double(10.0),
                           ctypes.c_int(0), ctypes.c_float(0.0), dp(npe), dp(cc), dp(d), dp(c), st,
                           ctypes.c_int(0), None, None)
    return npe, d, c, list(st)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-D", action="append", default=[])
    ap.add_argument("--save")
    ap.add_argument("--against")
    ap.add_argument("--sets", default=",".join(TAGS))
    ap.add_argument("--csrc", help="directory holding an alternative nss_eas_core.cuh")
    args = ap.parse_args()
    oc = oracle_cache()
    G = np.load(os.path.join(ROOT, "tests", "golden", "eas_large.npz"))
    L = build(args.D, args.save or "tmp", args.csrc)
    prev = dict(np.load(f"/tmp/nss_emul_{args.against}.npz")) if args.against else None
    saved = {}
    print(f"{'set':5s} {'n':>6s} | vs oracle: med p99 max | vs ref: med p99 p99.9 max nbad | "
          f"deficit-corr med | cang max | flips | vs {args.against}: med p99 max | N_val J nexp")
    for tag in args.sets.split(","):
        b, a, e = (G[f"{tag}_{k}"] for k in ("beta_rad", "altDec", "showerEnergy"))
        t0 = time.time()
        npe, d, c, st = run(L, b, a, e)
        dt = time.time() - t0
        ref_d, ref_c = G[f"{tag}_dphots"], G[f"{tag}_cang"]
        od, odr, sens = oc[f"{tag}_od"], oc[f"{tag}_odr"], oc[f"{tag}_sens"]
        live = ref_d != 0
        ro = rel(d, od)[live]
        rr = rel(d, ref_d)[live]
        with np.errstate(all="ignore"):
            deficit = np.where(od > 0, (od - odr) / od, 0.0)
        rc = rel(d * (1 - deficit), ref_d)[live]
        nbad = int((rel(d, ref_d) > np.maximum(2.5e-3, 5 * sens)).sum())
        flips = int(((d * 0.5 >= 10) != (ref_d * 0.5 >= 10)).sum())
        s = (f"{tag:5s} {b.size:6d} | {np.median(ro):.2e} {np.percentile(ro, 99):.2e} {ro.max():.2e} | "
             f"{np.median(rr):.2e} {np.percentile(rr, 99):.2e} {np.percentile(rr, 99.9):.2e} "
             f"{rr.max():.2e} {nbad} | {np.median(rc):.2e} | {rel(c, ref_c).max():.2e} | {flips} |")
        if prev is not None:
            rp = rel(d, prev[f"{tag}_d"])[live]
            s += f" {np.median(rp):.2e} {np.percentile(rp, 99):.2e} {rp.max():.2e} |"
        ne = max(st[0], 1)
        s += f" {st[2] / ne:.1f} {st[3] / ne:.0f} {st[5] / ne:.0f}  ({dt:.1f}s)"
        print(s, flush=True)
        saved[f"{tag}_d"], saved[f"{tag}_c"] = d, c
    if args.save:
        np.savez(f"/tmp/nss_emul_{args.save}.npz", **saved)


if __name__ == "__main__":
    main()
