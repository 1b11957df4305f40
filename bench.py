#!/usr/bin/env python
"""bench.py -- thrown neutrino events per second through the whole optical path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): simulated neutrino events/sec, end-to-end optical acceptance
(throw -> spectrum -> taus -> decay altitude -> EAS Cherenkov -> acceptance sums).

Workload (SURVEY.md 8(d) config 2 = BASELINE.json configs[1]): sample_input_file.toml
physics (525 km detector, monospectrum log10 E_nu = 8.0, theta_ch_max 3 deg, Greisen,
nupyprop table v3, NoCloud), 1e7 thrown events per GPU per step.  Weak scaling: rank r of
G owns thrown-event indices [step*G*n + r*n, step*G*n + (r+1)*n); the only exchange is
one all-reduce of 8 float64 acceptance sums per step.

  value  in-kernel Philox RNG, nothing but 8 doubles leaves the GPU (nss_run_acceptance).
  e2e    the reference-facing call: nuspacesim_b200.compute(config, uniforms=...) with the
         uniforms in pinned HOST memory (the reference draws them on the host too) and the
         full 15-column results table + acceptance returned to the HOST, every step.
  roofline  the EAS kernel (compute bound: FP32 pipe + MUFU); achieved =
         algorithmic FLOPs (8(d) formula evaluated on the kernel's own work counters) /
         CUDA-event duration of the launches inside the timed region; peak = FP32 FMA rate
         measured on this device by nss_microbench_peaks ("of measured").
  cpu_baseline  the UNMODIFIED reference (kind "reference": its own source files, staged into
         oracle/_ref/ by build() and run through oracle/ref_bench.py with the EAS stage fanned
         out over all host cores like its dask bag) on a bounded sample of the same workload;
         the oracle's C/OpenMP port beside it (cpu_baseline.port), and one core alone.
  workloads  events/s (device-resident, 1e7 events, this GPU) of BASELINE.json configs 3/4 and
         the non-default shower profiles / cloud model, each with its own work per event.

--impl reference times the reference's CPU implementation alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "simulated neutrino events/sec (end-to-end optical acceptance)"
UNIT = "events/s"
WORKLOAD = ("sample_input_file.toml physics, monospectrum log_nu_energy=8.0, "
            "1e7 thrown events per GPU per step, optical Cherenkov")
NW = 28  # wavelength bins the reference really has (29 edges, cphotang.py:84-117)
SEED = 1234


# ------------------------------------------------------------------ helpers


def oracle_tables():
    from nuspacesim_b200.utils.grid import NssGrid

    d = {}
    base = os.path.join(ROOT, "nuspacesim_b200", "data", "nupyprop_tables")
    for kind in ("pexit", "cdf"):
        g = NssGrid.read(os.path.join(base, f"nu2tau_{kind}.3.h5"))
        d[kind] = g.data
        for n, a in zip(g.axis_names, g.axes):
            d[f"{kind}_{n}"] = a
    return d


def time_cpu_path(n: int, cores: int, seed: int = 20240601):
    """The oracle's restatement of compute()'s optical branch on n thrown events."""
    from oracle import nss_oracle as O

    tabs = oracle_tables()
    rng = np.random.Generator(np.random.PCG64(seed))
    u4, ut, ud = rng.random((4, n)), rng.random(n), rng.random(n)
    O.build_c_oracle()
    t0 = time.perf_counter()
    mc, _, _, st = O.run_path(O.OracleParams(), tabs, u4, None, ut, ud, nthreads=cores)
    dt = time.perf_counter() - t0
    time_cpu_path.last_eas_seconds = st.get("eas_seconds")
    return dt, mc


def cpu_baseline(target_s: float, cores: int):
    """Calibrate on 2000 events, then time a sample sized for ~target_s seconds."""
    dt, _ = time_cpu_path(2000, cores)
    n = int(max(2000, min(2_000_000, 2000 * target_s / max(dt, 1e-3))))
    dt, mc = time_cpu_path(n, cores)
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} thrown events of the same workload in {dt:.1f} s "
                       f"(oracle/: C engine with OpenMP over events + NumPy stages)",
                mcint=mc[0], mcint_unc=mc[3],
                eas_stage_share=(time_cpu_path.last_eas_seconds or 0.0) / dt,
                eas_stage_value=n / time_cpu_path.last_eas_seconds
                if time_cpu_path.last_eas_seconds else None)


def time_reference(n: int, procs: int, seed: int = 1):
    """The unmodified reference on n thrown events (oracle/ref_bench.py in its own process:
    it forks workers and installs stub modules).  -> dict or None if it cannot run here."""
    try:
        r = subprocess.run([sys.executable, "-m", "oracle.ref_bench", "--events", str(int(n)),
                            "--procs", str(int(procs)), "--seed", str(int(seed))],
                           cwd=ROOT, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def reference_baseline(target_s: float, cores: int):
    """Calibrate on 2000 events, then time a sample sized for ~target_s seconds; plus one core
    on 1e3 events (SURVEY.md 8(d) CPU baseline plan)."""
    cal = time_reference(2000, cores)
    if cal is None:
        return None
    n = int(max(2000, min(2_000_000, 2000 * target_s / max(cal["seconds"], 1e-3))))
    r = time_reference(n, cores, seed=2)
    if r is None:
        return None
    one = time_reference(1000, 1, seed=3)
    return dict(value=r["events_per_s"], unit=UNIT, cores=cores, kind="reference",
                sample=f"{n} thrown events of the same workload in {r['seconds']:.1f} s: the "
                       f"reference's own source files ({r['source']} copy) -- NumPy stages "
                       f"single-threaded, CphotAng.run over {cores} worker processes "
                       "(multiprocessing in place of its dask bag, cphotang.py:719-723)",
                mcint=r["mcint"], mcint_unc=r["mcint_unc"], eas_stage_share=r["eas_seconds"] / r["seconds"],
                eas_stage_value=r["eas_events_per_s"],
                single_core=None if one is None else dict(
                    value=one["events_per_s"], events=one["events"], seconds=one["seconds"]))


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                    power_w_max=float(max(pw)), samples=len(sm), reasons=sorted(reasons))


def algorithmic_work(st):
    """FLOPs and transcendentals from the kernel's work counters
    st = {events, N_raw, N_val, J, J*N_e, n_exp, N_val*N_e, -} (sums over events)."""
    ev, n_raw, n_val, J, JNe, n_exp, n_val_ne = (float(x) for x in st[:7])
    flops = 45 * n_raw + (60 + 8 * NW + 25) * n_val + 25 * n_val_ne + 10 * J + 10 * JNe
    transc = 6 * n_raw + (12 + 3 * NW + 2) * n_val + 2 * n_val_ne + 4 * J + 2 * JNe
    return flops, transc


# ------------------------------------------------------------------ reference arm


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = []
    cal = time_reference(2000, cores)
    kind = "reference" if cal is not None else "port"
    if cal is not None:
        n = int(max(2000, min(500_000, 2000 * 8.0 / max(cal["seconds"], 1e-3))))  # ~8 s per step
        src = cal["source"]
        for i in range(args.warmup + args.steps):
            r = time_reference(n, cores, seed=100 + i)
            if r is None:
                kind = "port"
                break
            if i >= args.warmup:
                per_step.append(r["seconds"])
    if kind == "port":  # no staged reference on this machine: the oracle's restatement
        per_step = []
        dt, _ = time_cpu_path(2000, cores)
        n = int(max(2000, min(500_000, 2000 * 8.0 / max(dt, 1e-3))))
        for i in range(args.warmup + args.steps):
            dt, mc = time_cpu_path(n, cores, seed=100 + i)
            if i >= args.warmup:
                per_step.append(dt)
    tot = sum(per_step)
    v = n * args.steps / tot
    what = (f"the UNMODIFIED reference ({src} copy of its source files, oracle/ref_bench.py): "
            "NumPy stages single-threaded, CphotAng.run over all host cores"
            if kind == "reference" else
            "the oracle port (C engine + NumPy stages): no staged reference on this machine")
    line = dict(
        # n_gpus echoes the launch (the driver pairs the two arms by it); no GPU is used
        metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=1e3 * tot / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32 engine / f64 stages", data="synthetic", impl="reference",
        config=dict(workload=WORKLOAD, events_per_step=n, implementation=what,
                    note="same physics and metric as the B200 arm; each step is a BOUNDED "
                         f"sample of {n} thrown events of that workload (the full 1e7 would "
                         "take hours on the host), so the rate, not the step size, is compared"),
        cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind=kind,
                          sample=f"{n} thrown events per step"),
        e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line))


# ------------------------------------------------------------------ secondary workloads


def workload_configs(nss):
    """BASELINE.json configs[2] (energy sweep) and [3] (power law, wide angle), the two
    Gaisser-Hillas profiles and the pressure-map cloud model -> {name: (config, description)}."""
    import math

    from nuspacesim_b200.config import PowerSpectrum, PressureMapCloud

    out = {}
    for tag, le in (("e7", 7.0), ("e7.5", 7.5), ("e8.5", 8.5), ("e9", 9.0), ("e9.5", 9.5),
                    ("e10", 10.0)):
        c = nss.NssConfig()
        c.simulation.spectrum.log_nu_energy = le
        out[tag] = (c, f"monospectrum log_nu_energy={le}")
    c = nss.NssConfig()
    c.simulation.spectrum = PowerSpectrum(index=2.0, lower_bound=6.0, upper_bound=12.0)
    c.simulation.max_cherenkov_angle = math.radians(10.0)
    out["powerlaw_w10"] = (c, "power-law spectrum (index 2, 1e6-1e12 GeV), max_cherenkov_angle 10 deg")
    for tag, prof in (("ghp", "Gaisser-Hillas Parameterized"), ("ghf", "Gaisser-Hillas Fluctuated")):
        c = nss.NssConfig()
        c.simulation.eas_long_profile = prof
        out[tag] = (c, f"monospectrum 8.0, {prof} longitudinal profile")
    c = nss.NssConfig()
    c.simulation.cloud_model = PressureMapCloud(month=7)
    out["cloudmap"] = (c, "monospectrum 8.0, PressureMapCloud month 7 (per-event cloud tops)")
    return out


def bench_workloads(nss, torch, n, batch, first):
    """Timed like the headline: the acceptance runner as compute() launches it (work counters
    compiled out); the counters come from a second, untimed pass over the same events."""
    res = {}
    for name, (cfg, desc) in workload_configs(nss).items():
        r = nss.AcceptanceRunner(cfg, batch=batch)
        r.enqueue(SEED, first, min(n, batch))  # warm-up (tables, CONEX / cloud map upload)
        torch.cuda.synchronize()
        r.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.enqueue(SEED, first, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        mc = r.finalize(r.sums(False))
        del r
        rs = nss.AcceptanceRunner(cfg, batch=batch, collect_stats=True)
        rs.enqueue(SEED, first, n)
        torch.cuda.synchronize()
        st = rs.stats.cpu().numpy().astype(float)
        del rs
        ne = max(st[0], 1.0)
        flops, transc = algorithmic_work(st)
        res[name] = dict(
            value=n / ms * 1e3, unit=UNIT, ms=ms, events=n, config=desc,
            acceptance=dict(mcint=mc[0], geo=mc[1], n_pass=mc[2], unc=mc[3]),
            eas_events_fraction=st[0] / n,
            work_per_eas_event=dict(N_raw=st[1] / ne, N_val=st[2] / ne, J=st[3] / ne,
                                    JNe=st[4] / ne, exp_evaluated=st[5] / ne),
            algorithmic_tflops=flops / (ms * 1e-3) * 1e-12)
    return res


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one eas_walk_kernel launch from the
    newest committed `ncu --set full` summary (profiles/r*_eas_v*_ncu_key_metrics.json)."""
    import glob
    import re

    best = None
    for f in glob.glob(os.path.join(ROOT, "profiles", "r*_eas_v*_ncu_key_metrics.json")):
        m = re.search(r"r(\d+)_eas_v(\d+)_", os.path.basename(f))
        if m and (best is None or (int(m.group(1)), int(m.group(2))) > best[0]):
            best = ((int(m.group(1)), int(m.group(2))), f)
    if best is None:
        return None, None, None
    try:
        with open(best[1]) as fh:
            d = json.load(fh)
        def nbytes(v):
            if isinstance(v, dict):  # {"value": "306.5", "unit": "Mbyte"}
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v["unit"]]
                return float(v["value"]) * scale
            return float(v)

        rd, wr = nbytes(d["dram__bytes_read.sum"]), nbytes(d["dram__bytes_write.sum"])
        ncu_traffic.hardware = {k: d.get(src) for k, src in (
            ("issue_slots_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            ("warps_per_sm", "sm__warps_active.avg.per_cycle_active"),
            ("registers", "launch__registers_per_thread"),
            ("warp_instructions_per_event", "warp_instructions_per_event"))}
        return rd + wr, d.get("events_in_launch"), os.path.relpath(best[1], ROOT)
    except Exception:
        return None, None, os.path.relpath(best[1], ROOT)


ncu_traffic.hardware = None


# ------------------------------------------------------------------ the B200 arm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=float, default=1e7, help="thrown events per GPU per step")
    ap.add_argument("--batch", type=int, default=5_000_000,
                    help="thrown events per launch of the stage kernels (measured on one B200: "
                         "2 Mi 1.197e8, 5e6 1.206e8, 1e7 1.207e8 events/s)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)  # honoured as given; the contract asks for >= 3

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    numa = None
    if world > 1 and os.environ.get("NSS_BENCH_NUMA", "1") != "0":
        from nuspacesim_b200.utils.affinity import bind_to_device_node

        numa = bind_to_device_node(local)  # host buffers of this rank stay next to its GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import nuspacesim_b200 as nss
    from nuspacesim_b200 import _lib

    L = _lib.lib()
    cfg = nss.NssConfig()
    n = int(args.events)
    runner = nss.AcceptanceRunner(cfg, batch=args.batch)

    def first_event(step):
        return (step * world + rank) * n

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident acceptance mode -> `value`
    results = []

    def accept_step(step):
        runner.reset()
        runner.enqueue(SEED, first_event(step), n)
        results.append(runner.sums())  # all-reduce (N > 1) + 64-byte read-back

    for s in range(args.warmup):
        accept_step(s)
    sync_all()
    clocks = ClockSampler(local)
    L.nss_eas_profile_enable(1)
    launches0 = runner.n_kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        accept_step(args.warmup + s)
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    L.nss_eas_profile_enable(0)
    eas_ms, eas_launches = C.c_double(0), C.c_int64(0)
    _lib.check(L.nss_eas_profile_collect(C.byref(eas_ms), C.byref(eas_launches)))
    clk = clocks.stop()
    gpu_launches = runner.n_kernel_launches - launches0
    value = n * world * args.steps / (ms_total * 1e-3)
    mc = runner.finalize(results[-1])

    # ---- work counters of one step's events (untimed) -> algorithmic FLOPs per launch
    line_extra = {}
    if rank == 0:
        st_runner = nss.AcceptanceRunner(cfg, batch=args.batch, collect_stats=True)
        st_runner.enqueue(SEED, first_event(args.warmup), n)
        torch.cuda.synchronize()
        st = st_runner.stats.cpu().numpy()
        flops_step, transc_step = algorithmic_work(st)
        launches_per_step = eas_launches.value / args.steps
        eas_s_per_step = eas_ms.value * 1e-3 / args.steps
        fp32_peak, mufu_peak = C.c_double(0), C.c_double(0)
        _lib.check(L.nss_microbench_peaks(C.byref(fp32_peak), C.byref(mufu_peak),
                                          _lib.stream_ptr()))
        ach = flops_step / eas_s_per_step * 1e-12
        traffic, traffic_events, traffic_src = ncu_traffic()
        ev_per_launch = float(st[0]) / launches_per_step
        if traffic is not None and traffic_events:
            # the capture's launch may hold a different number of events than this run's
            traffic = traffic * ev_per_launch / float(traffic_events)
        line_extra["roofline"] = dict(
            bound="fp32", kernel="eas_walk_kernel", achieved=ach, peak=fp32_peak.value,
            unit="TFLOP/s", frac=ach / fp32_peak.value if fp32_peak.value else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE eas_walk_kernel launch, read from
            # the newest committed ncu --set full summary under profiles/ (scaled to this run's
            # events per launch).  The walk must read 96 B of each event's EasPrep record (the
            # Gaisser-Hillas third is not touched by the Greisen instantiation) and write the
            # 48 B EasAcc record: 144 B per event; the kernel is compute bound (bound = "fp32").
            traffic=traffic, traffic_source=traffic_src,
            traffic_unit="bytes per launch (ncu, cold cache)",
            algorithmic_bytes=144.0 * ev_per_launch,
            peak_source="nss_microbench_peaks on this device (FFMA chains; MEASURED_PEAKS.json "
                        "has no FP32 entry), of measured",
            sfu=dict(achieved=transc_step / eas_s_per_step * 1e-12, peak=mufu_peak.value,
                     unit="1e12 transcendentals/s (reference count) vs MUFU.EX2 issue rate",
                     frac=(transc_step / eas_s_per_step * 1e-12) / mufu_peak.value
                     if mufu_peak.value else None),
            per_launch=dict(flops=flops_step / launches_per_step,
                            transcendentals=transc_step / launches_per_step,
                            ms=eas_ms.value / max(eas_launches.value, 1),
                            events=float(st[0]) / launches_per_step),
            share_of_step=eas_ms.value / ms_total,
            work_per_eas_event=dict(N_raw=st[1] / st[0], N_val=st[2] / st[0], J=st[3] / st[0],
                                    JNe=st[4] / st[0], exp_evaluated=st[5] / st[0]),
            # `achieved` counts the REFERENCE's operations; the kernel reaches the
            # same numbers with about a third of them (closed-form stepping, column totals,
            # closed-form radial sums, Gauss rule over wavelengths), so frac can exceed 1.  What
            # the hardware does is in the ncu summary the traffic comes from: issue-slot bound.
            hardware_from_ncu=ncu_traffic.hardware,
            # live: warp-instructions issued per second by the walk (instructions per event from
            # that ncu summary x the events this run's launches processed / their CUDA-event
            # time) against the SM's issue rate (148 SMs x 4 schedulers x the measured clock)
            issue_slots=(lambda ipe, mhz: None if not ipe or not mhz else dict(
                achieved=ipe * float(st[0]) / eas_s_per_step, peak=148 * 4 * mhz * 1e6,
                frac=ipe * float(st[0]) / eas_s_per_step / (148 * 4 * mhz * 1e6),
                unit="warp-instructions/s"))(
                (ncu_traffic.hardware or {}).get("warp_instructions_per_event"),
                clk.get("sm_mhz")),
            hbm_note="acceptance mode has no algorithmic DRAM traffic (0 B/event in, 64 B/step "
                     "out); HBM peak 6535 GB/s (MEASURED_PEAKS.json) bounds only the column-mode "
                     "stages timed in e2e",
        )
        del st_runner

    # ---- end to end through compute() with HOST buffers -> `e2e`
    e2e = None
    if not args.no_e2e:
        rng = np.random.Generator(np.random.PCG64(20240601 + rank))
        pinned = {}
        for name, shape in (("u4", (4, n)), ("u_tau", (n,)), ("u_dec", (n,))):
            t = torch.empty(shape, dtype=torch.float64).pin_memory()
            t.numpy()[...] = rng.random(shape)
            pinned[name] = t
        h2d = sum(t.numel() * 8 for t in pinned.values())
        d2h = 0
        sim = None
        sync_all()
        k_e2e = max(1, min(args.steps, 3))
        # warm-ups: the first two calls pin ~2.4 GB of host memory for two generations of the
        # results table (cudaHostAlloc, ~1 s each); from the third on the pinned-host cache
        # recycles them
        n_warm = int(os.environ.get("NSS_BENCH_E2E_WARMUP", "3"))
        step_s = []
        for i in range(n_warm + k_e2e):
            if i == n_warm:
                sync_all()
                t0 = time.perf_counter()
            t1 = time.perf_counter()
            sim = nss.compute(cfg, uniforms=pinned)
            d2h = sum(np.asarray(sim[c]).nbytes for c in sim.colnames) + 4 * 8
            step_s.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        dt = torch.tensor([dt_local], dtype=torch.float64, device="cuda")
        per_rank = [dt_local]
        if world > 1:
            gathered = [torch.zeros_like(dt) for _ in range(world)]
            dist.all_gather(gathered, dt)
            per_rank = [float(g.item()) for g in gathered]
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = dict(value=n * world * k_e2e / float(dt.item()), unit=UNIT,
                   h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), steps=k_e2e,
                   api="nuspacesim_b200.compute(config, uniforms=<pinned host arrays>) -> "
                       "15-column table + OMCINT meta on the host",
                   mcint=float(sim.meta["OMCINT"][0]), n_valid=len(sim),
                   per_rank_s=per_rank, rank0_call_s=step_s, rank0_numa_binding=numa,
                   # first compute() call of this process (library and CUDA context already
                   # up from the device-resident leg; no page-locked host memory cached yet):
                   # what a one-shot `nuspacesim run` pays on top of start-up
                   cold_call_s=step_s[0], cold_value=n / step_s[0])
        # the same call with in-kernel Philox instead of host uniforms (`compute(config, seed=s)`,
        # what `nuspacesim run` does): no host -> device bytes, the same table back
        cfg_seed = nss.NssConfig()
        cfg_seed.simulation.thrown_events = n
        sync_all()
        ts = []
        for i in range(1 + k_e2e):
            if i == 1:
                sync_all()
                t0 = time.perf_counter()
            t1 = time.perf_counter()
            sim = nss.compute(cfg_seed, seed=SEED + 7 * rank + i)
            ts.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        dts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dts, op=dist.ReduceOp.MAX)
        e2e["seed_variant"] = dict(
            value=n * world * k_e2e / float(dts.item()), unit=UNIT, h2d_bytes_per_step=0,
            d2h_bytes_per_step=int(d2h), steps=k_e2e, rank0_call_s=ts,
            api="nuspacesim_b200.compute(config, seed=s) -> 15-column table on the host")
        del sim

    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms_total / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="f32 engine / f64 stages", data="synthetic",
            config=dict(workload=WORKLOAD, events_per_gpu_per_step=n, batch=args.batch,
                        seed=SEED, rng="in-kernel Philox4x32-10, counter = thrown-event index",
                        parallelism=f"event-sharded x{world}, one 8-double all-reduce per step",
                        l2="no input arrays; per-batch stage buffers (%.0f MB) exceed the 126 MB L2"
                           % (args.batch * 65 / 1e6),
                        cpu_arm="cpu_baseline and `--impl reference` time the same workload's "
                                "RATE on a bounded sample of it (1e4-1e5 thrown events per "
                                "step: 1e7 would take hours on the host)"),
            acceptance=dict(mcint=mc[0], geo=mc[1], n_pass=mc[2], unc=mc[3]),
            clocks=clk, gpu_launches=int(gpu_launches), e2e=e2e,
        )
        line.update(line_extra)
        if not args.no_workloads and world == 1:
            line["workloads"] = bench_workloads(nss, torch, n, args.batch,
                                                first_event(args.warmup + args.steps))
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            port = cpu_baseline(min(args.cpu_seconds, 8.0), cores)
            ref = reference_baseline(args.cpu_seconds, cores)
            if ref is not None:
                ref["port"] = port
                line["cpu_baseline"] = ref
            else:
                port["note"] = "no staged reference on this machine (oracle/_ref/): port only"
                line["cpu_baseline"] = port
        print(json.dumps(line, default=float))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
