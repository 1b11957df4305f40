"""compute(): one simulation of the optical hot path, stage by stage.

Mirror of ``nuspacesim.compute`` (src/nuspacesim/compute.py:73-298) for the optical
branch: same signature, the same 15 result columns in the same order, the same meta keys
(``OMCINT``, ``OMCINTGO``, ``ONEVPASS``, ``OMCINTUN``, ``simTime``, ``HIERARCH Config``).
All stage arrays stay on the GPU between stages; columns are copied to the host once
each for the returned table.  The radio branch (compute.py:261-294) is out of scope and
is skipped.

``run_acceptance`` is the column-free mode used for large N: the whole chain runs from
in-kernel Philox (``nss_run_acceptance``) and only 8 doubles leave the GPU; with
torch.distributed initialised the thrown-event range is split across ranks and the sums
are all-reduced (the only exchange step on the path).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Iterable

import numpy as np

from . import _lib, results_table
from .params import params_from_config
from .simulation.atmosphere.clouds import CloudTopHeight
from .simulation.eas_optical.eas import EAS
from .simulation.geometry.region_geometry import RegionGeom
from .simulation.spectra.spectra import Spectra
from .simulation.taus.taus import Taus

__all__ = ["compute", "run_acceptance"]


#: thrown events per slice of compute()'s pipelined path (the acceptance mode's batch size)
SLICE_EVENTS = int(os.environ.get("NSS_SLICE_EVENTS", 1 << 21))
#: slices whose uniforms are uploaded ahead of the one being computed
PREFETCH_SLICES = 2
#: enqueue the next slice's throw ahead of the current slice's EAS kernels
PIPELINE_THROW = os.environ.get("NSS_PIPELINE_THROW", "1") != "0"
#: result columns larger than this in total are not page-locked (pinning is ~1 GB/s and
#: page-locked memory is a scarce resource: 120 B per thrown event)
PINNED_MAX_BYTES = int(os.environ.get("NSS_PINNED_MAX_BYTES", 16 << 30))

COLUMN_GROUPS = (
    ("beta_rad", "theta_rad", "path_len"), ("init_lat", "init_lon"), ("log_e_nu",),
    ("tauBeta", "tauLorentz", "tauEnergy", "showerEnergy", "tauExitProb"),
    ("altDec", "lenDec"), ("numPEs", "costhetaChEff"),
)


def compute(config, verbose: bool = False, output_file: str = None, to_plot: list = [],
            write_stages: bool = False, seed: int = 0, uniforms: dict = None,
            slice_events: int = None, shard: bool = False):
    """Simulate an upward going shower (compute.py:73-298, optical branch).

    Extensions: ``seed`` keys the Philox streams (the reference uses the global unseeded
    NumPy state); ``uniforms`` may hold pre-sampled arrays ``u4`` (4, N), ``u_spec``,
    ``u_tau``, ``u_dec``, ``u_profile`` (N,) indexed by thrown event, for shared-input
    validation.  Large runs are pipelined in slices of ``slice_events`` thrown events
    (default ``SLICE_EVENTS``; 0 disables): the stage sequence runs slice by slice so that the
    upload of the next slice's uniforms and the download of the previous slice's columns hide
    behind the EAS kernel.  Per-event results do not depend on the slicing; staged writes and
    plots use the unsliced path (they need whole columns after every stage).

    ``shard=True`` with ``torch.distributed`` initialised (one process per GPU): rank r of G
    simulates the contiguous block ``shard_range(N, r, G)`` of the thrown events and returns
    ITS rows of the table (concatenating the ranks' tables in rank order gives the one-process
    table; nothing per event is exchanged); the acceptance meta keys are global -- the ranks'
    five sums are gathered once and merged exactly.
    """
    n_thrown = config.simulation.thrown_events
    if uniforms and "u4" in uniforms:
        n_thrown = int(uniforms["u4"].shape[1]) if getattr(uniforms["u4"], "ndim", 0) == 2 else 0
    step = SLICE_EVENTS if slice_events is None else int(slice_events)
    sliceable = (not write_stages and not to_plot and config.simulation.mode != "Target"
                 and config.detector.optical.enable and n_thrown > 0)
    sliced = step > 0 and n_thrown >= 2 * step and sliceable
    lo, hi = 0, int(n_thrown)
    if shard:
        import torch.distributed as dist

        if not sliceable:
            raise ValueError("shard=True needs the optical branch without staged writes/plots")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            lo, cnt = shard_range(int(n_thrown), dist.get_rank(), dist.get_world_size())
            hi = lo + cnt
        sliced = True
        step = step if step > 0 else SLICE_EVENTS
    try:
        with _lib.deferred_host_copies() as sink:
            if sliced:
                return _compute_sliced(config, verbose, output_file, seed, uniforms, sink,
                                       int(n_thrown), step, lo, hi, shard)
            return _compute(config, verbose, output_file, to_plot, write_stages, seed, uniforms,
                            sink)
    finally:
        _lib.h2d_synchronize()  # the caller's pinned uniforms are free to be reused


class StagedWriter:
    """Optionally write intermediate values to file (compute.py:166-183).

    Module-level on purpose: a class created inside compute() would keep the table (and its
    pinned host columns) alive in a reference cycle until the next gc pass."""

    def __init__(self, sim, sink, output_file, write_stages):
        self.sim, self.sink = sim, sink
        self.output_file, self.write_stages = output_file, write_stages

    def __call__(self, col_names: Iterable[str], columns: Iterable, *args, **kwargs):
        self.sim.add_columns(columns, names=col_names, *args, **kwargs)
        if self.write_stages:
            self.sink.synchronize()
            self.sim.write(self.output_file, format="fits", overwrite=True)

    def add_meta(self, name: str, value: Any, comment: str):
        self.sim.meta[name] = (value, comment)
        if self.write_stages:
            self.sink.synchronize()
            self.sim.write(self.output_file, format="fits", overwrite=True)


def _compute(config, verbose, output_file, to_plot, write_stages, seed, uniforms, sink):
    """Body of compute(); column copies to the host are asynchronous on `sink`'s stream and
    complete when compute() returns (or before any staged write)."""

    def logv(*args):
        if verbose:
            print(*args)

    sim = results_table.init(config)
    geom = RegionGeom(config, seed=seed)
    cloud = CloudTopHeight(config)
    spec = Spectra(config, seed=seed)
    tau = Taus(config, seed=seed)
    eas = EAS(config, seed=seed)
    if config.simulation.mode == "Target":
        raise NotImplementedError("Target (ToO) mode is outside the accelerated path")

    sw = StagedWriter(sim, sink, output_file, write_stages)
    U = uniforms or {}
    # Pinned host uniforms: u4 goes first (geometry needs it), the per-stage ones follow on a
    # side stream and overlap the geometry kernels.
    Ud = {}
    for name in ("u4", "u_spec", "u_tau", "u_dec", "u_profile"):
        if name in U:
            Ud[name] = _lib.to_device_prefetch(U[name])

    logv(f"Running NuSpaceSim with Energy Spectrum ({config.simulation.spectrum})")
    logv("Computing Geometries.")
    thrown = config.simulation.thrown_events
    if "u4" in U:
        thrown, ev = Ud["u4"]
        if ev is not None:
            _lib.torch_cuda().cuda.current_stream().wait_event(ev)
    beta_tr, thetaArr, pathLenArr, *_ = geom(thrown, store=sw, plot=to_plot, device=True)
    n_valid = int(beta_tr.numel())
    logv(f"\tThrew {config.simulation.thrown_events} neutrinos. {n_valid} were valid.")
    if n_valid == 0:
        print("\tWARNING: No valid events thrown! Exiting early! Check geometry!")
        return sim

    torch = _lib.torch_cuda()

    WHICH = {"u_spec": 0, "u_tau": 1, "u_dec": 2, "u_profile": 3}

    def valid_u(name):
        """Per-valid-event uniforms: the caller's array, else (when the geometry itself came
        from Philox) the event's own counter-based stream, so that compute(seed=s) and
        run_acceptance(seed=s) simulate identical events."""
        if name in Ud:
            full, ev = Ud[name]
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
            return geom.compact(full)
        if "u4" not in U:
            return geom.event_uniforms(WHICH[name])
        return None

    init_lat, init_long = geom.find_lat_long_along_traj(torch.zeros_like(beta_tr), device=True)
    sw(("init_lat", "init_lon"), (_lib.to_host(init_lat), _lib.to_host(init_long)))

    logv("Computing Energy Spectra.")
    log_e_nu, mc_spec_norm, spec_weights_sum = spec(
        n_valid, store=sw, plot=to_plot, device=True, u=valid_u("u_spec"))

    logv("Computing Taus.")
    # the caller's u_tau is used as it is (shared-input parity with the reference's `u`
    # argument); Philox and host-drawn uniforms are plain draws the kernel maps onto
    # (cdf[0], 1) like the reference's own np.random.uniform(cdfs[:, 0], 1.0)
    tauBeta, tauLorentz, tauEnergy, showerEnergy, tauExitProb = tau(
        beta_tr, log_e_nu, store=sw, plot=to_plot, u=valid_u("u_tau"), u_raw="u_tau" not in Ud)

    logv("Computing Decay Altitudes.")
    altDec, lenDec = eas.altDec(beta_tr, tauBeta, tauLorentz, valid_u("u_dec"), store=sw)

    if config.detector.optical.enable:
        logv("Computing EAS Optical Cherenkov light.")
        u_prof = valid_u("u_profile") if eas.params.eas_profile == 2 else None
        numPEs, costhetaChEff = eas(beta_tr, altDec, showerEnergy, init_lat, init_long,
                                    cloudf=cloud, u_profile=u_prof, store=sw, plot=to_plot)
        logv("Computing Optical Monte Carlo Integral.")
        mcint, mcintgeo, passEV, mcunc = geom.mcintegral(
            numPEs, costhetaChEff, tauExitProb,
            config.detector.optical.photo_electron_threshold, mc_spec_norm, spec_weights_sum,
            lenDec=lenDec, method="Optical", store=sw)
        sw.add_meta("OMCINT", mcint, "Optical MonteCarlo Integral")
        sw.add_meta("OMCINTGO", mcintgeo, "Optical MonteCarlo Integral, GEO Only")
        sw.add_meta("ONEVPASS", passEV, "Optical Number of Passing Events")
        sw.add_meta("OMCINTUN", mcunc, "Stat unc of MonteCarlo Integral")
        logv(f"\tMonte Carlo Integral [Optical]: {mcint}")
        logv(f"\tMonte Carlo Integral, GEO Only [Optical]: {mcintgeo}")
        logv(f"\tNumber of Passing Events [Optical]: {passEV}")
        logv(f"\tStat uncert of MC Integral [Optical]: {mcunc}")

    # like the reference, compute() itself writes only the staged files (write_stages); the
    # final table is written by the caller (apps/run.py:209-210)
    return sim


def _merge_sums(S):
    """Rows of {sum f_geo, sum f, sum (f - mean)^2, n_pass, n_valid} -> one row (exact merge of
    the squared deviations, Chan et al.)."""
    S = np.asarray(S, dtype=np.float64).reshape(-1, 5)
    S = S[S[:, 4] > 0]
    if len(S) == 0:
        return np.zeros(5)
    n_valid = S[:, 4].sum()
    mean = S[:, 1].sum() / n_valid
    m2 = float((S[:, 2] + S[:, 4] * (S[:, 1] / S[:, 4] - mean) ** 2).sum())
    return np.array([S[:, 0].sum(), S[:, 1].sum(), m2, S[:, 3].sum(), n_valid])


def _gather_rank_sums(row):
    """All ranks' merged sums (one 5-vector each) -> (world, 5) on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return row[None, :]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(row, dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def _slice_bounds(first, last, step):
    """Slices of about `step` thrown events; with four or more of them the first and the last
    are quarter-sized: nothing overlaps the first slice's upload or the last slice's download,
    so those two are kept short."""
    n = last - first
    if n <= 0:
        return []
    if n < 4 * step or step < 4 or os.environ.get("NSS_SLICE_EDGES", "1") == "0":
        return [(lo, min(lo + step, last)) for lo in range(first, last, step)]
    edge = step // int(os.environ.get("NSS_SLICE_EDGE_DIV", 4))
    inner = n - 2 * edge
    k = -(-inner // step)
    cuts = [first, first + edge] + [first + edge + (inner * i) // k for i in range(1, k + 1)]
    cuts.append(last)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


class _HostColumns:
    """The full-length host arrays the slices' columns are copied into (utils/host_pool.py:
    page-locked on a helper thread, recycled between calls).  Until the block is ready the
    copies are only queued -- the device tensors stay alive -- so that a cold first call
    computes while its result memory is being page-locked instead of waiting for it."""

    def __init__(self, names, n, sink):
        from .utils import host_pool

        torch = _lib.torch_cuda()
        self.sink, self.queue = sink, []
        self.pend = None
        mode = os.environ.get("NSS_HOST_COLUMNS", "pool")
        if len(names) * n * 8 > PINNED_MAX_BYTES or mode == "pageable":
            # too large to page-lock (or asked not to): plain pageable memory
            self.cols = {nm: torch.empty(n, dtype=torch.float64) for nm in names}
        elif mode == "torch":  # torch's caching host allocator (cudaHostAlloc)
            self.cols = {nm: torch.empty(n, dtype=torch.float64, pin_memory=True) for nm in names}
        else:
            self.pend = host_pool.PendingColumns(names, n)

    def _ready(self):
        if self.pend is not None and self.pend.ready():
            self.cols = self.pend.wait()
            self.pend = None
        return self.pend is None

    def put(self, name, off, nv, t):
        self.queue.append((name, off, nv, t))
        if self._ready():
            self.flush()

    def flush(self):
        for name, off, nv, t in self.queue:
            self.sink.copy_into(self.cols[name][off:off + nv], t)
        self.queue = []

    def finish(self):
        if self.pend is not None:
            self.cols = self.pend.wait()
            self.pend = None
        self.flush()
        return self.cols


def _compute_sliced(config, verbose, output_file, seed, uniforms, sink, n_thrown, step,
                    first=0, last=None, shard=False):
    """compute() for large N: the same stage objects, called per slice of thrown events.

    Slice k owns thrown events [lo, hi); valid events keep their thrown order, so the result
    columns are the concatenation of the slices' columns.  A slice's uniforms are uploaded
    PREFETCH_SLICES ahead on the upload stream; its throw is enqueued ahead of the previous
    slice's EAS kernels (the valid-event count is on the host before they finish); its columns
    go to their rows of full-length page-locked host arrays on the copy stream while the next
    slice computes; the five acceptance sums per slice stay on the device until the end and
    are merged exactly (Chan's update for the sum of squared deviations)."""

    def logv(*args):
        if verbose:
            print(*args)

    torch = _lib.torch_cuda()
    sim = results_table.init(config)
    cloud = CloudTopHeight(config)
    spec = Spectra(config, seed=seed)
    tau = Taus(config, seed=seed)
    eas = EAS(config, seed=seed)
    U = uniforms or {}
    last = n_thrown if last is None else last
    bounds = _slice_bounds(first, last, step)
    logv(f"Running NuSpaceSim with Energy Spectrum ({config.simulation.spectrum})")
    logv(f"Computing thrown events [{first}, {last}) of {n_thrown} in {len(bounds)} pipelined "
         "slices.")

    # uploads run PREFETCH_SLICES ahead of the slice being computed, in consumption order, on
    # the side stream: the device never holds more than that many slices' uniforms
    def upload(k):
        lo, hi = bounds[k]
        d = {}
        if "u4" in U:
            d["u4"] = _lib.rows_to_device_prefetch(U["u4"], lo, hi)
        for name in ("u_spec", "u_tau", "u_dec", "u_profile"):
            if name in U:
                d[name] = _lib.to_device_prefetch(U[name][lo:hi])
        return d

    pre = {k: upload(k) for k in range(min(PREFETCH_SLICES, len(bounds)))}

    names = [n for g in COLUMN_GROUPS for n in g]
    host = _HostColumns(names, max(last - first, 1), sink)
    WHICH = {"u_spec": 0, "u_tau": 1, "u_dec": 2, "u_profile": 3}
    sums, off = [], 0
    thr = config.detector.optical.photo_electron_threshold
    mc_spec_norm = spec_weights_sum = 1.0
    def start_throw(k):
        """Enqueue slice k's throw + mask scan; its valid-event count travels to pinned host
        memory right behind them, so reading it later never waits for kernels enqueued after
        this point (the previous slice's EAS in particular)."""
        lo, hi = bounds[k]
        d = pre.pop(k)
        if k + PREFETCH_SLICES < len(bounds):
            pre[k + PREFETCH_SLICES] = upload(k + PREFETCH_SLICES)
        g = RegionGeom(config, seed=seed)
        g.first_event = lo
        thrown = hi - lo
        if "u4" in d:
            thrown, ev = d["u4"]
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        g.throw(thrown, defer_count=True)
        return g, d

    nxt = start_throw(0) if bounds else None
    for k, (lo, hi) in enumerate(bounds):
        geom, d = nxt
        nxt = None
        nv = geom._n_valid  # waits for THIS slice's throw + scan only
        if nv == 0:
            if k + 1 < len(bounds):
                nxt = start_throw(k + 1)
            continue
        beta_tr, theta, path_len = geom.beta_rad(True), geom.thetas(True), geom.pathLens(True)

        def valid_u(name):
            if name in d:
                full, ev = d[name]
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                return geom.compact(full)
            return geom.event_uniforms(WHICH[name]) if "u4" not in U else None

        cols = dict(beta_rad=beta_tr, theta_rad=theta, path_len=path_len)
        cols["init_lat"], cols["init_lon"] = geom.find_lat_long_along_traj(
            torch.zeros_like(beta_tr), device=True)
        cols["log_e_nu"], mc_spec_norm, spec_weights_sum = spec(
            nv, device=True, u=valid_u("u_spec"))
        (cols["tauBeta"], cols["tauLorentz"], cols["tauEnergy"], cols["showerEnergy"],
         cols["tauExitProb"]) = tau(beta_tr, cols["log_e_nu"], u=valid_u("u_tau"),
                                    u_raw="u_tau" not in U)
        cols["altDec"], cols["lenDec"] = eas.altDec(beta_tr, cols["tauBeta"], cols["tauLorentz"],
                                                    valid_u("u_dec"))
        front = [n for n in names if n not in ("numPEs", "costhetaChEff")]
        for n in front:  # these travel while the EAS kernel runs
            host.put(n, off, nv, cols[n])
        u_prof = valid_u("u_profile") if eas.params.eas_profile == 2 else None
        # the next slice's throw goes in front of this slice's EAS kernels: its valid-event
        # count is on the host long before they finish, and the host never idles the GPU
        if k + 1 < len(bounds) and PIPELINE_THROW:
            nxt = start_throw(k + 1)
        cols["numPEs"], cols["costhetaChEff"] = eas(
            beta_tr, cols["altDec"], cols["showerEnergy"], cols["init_lat"], cols["init_lon"],
            cloudf=cloud, u_profile=u_prof)
        for n in ("numPEs", "costhetaChEff"):
            host.put(n, off, nv, cols[n])
        sums.append(geom.mcintegral(cols["numPEs"], cols["costhetaChEff"], cols["tauExitProb"],
                                    thr, mc_spec_norm, spec_weights_sum, device_sums=True))
        off += nv
        if nxt is None and k + 1 < len(bounds):
            nxt = start_throw(k + 1)

    logv(f"\tThrew {last - first} neutrinos. {off} were valid.")
    if off == 0 and not shard:
        print("\tWARNING: No valid events thrown! Exiting early! Check geometry!")
        host.finish()
        return sim
    # (slices, 8) on the device: sum geo, sum f, M2, n_pass, n_valid, ... -> one host read
    S = torch.stack(sums).cpu().numpy()[:, :5] if sums else np.zeros((0, 5))
    hcols = host.finish()
    sink.synchronize()
    for group in COLUMN_GROUPS:
        sim.add_columns([hcols[n].numpy()[:off] for n in group], names=group)
    mine = _merge_sums(S)
    tot = _merge_sums(_gather_rank_sums(mine)) if shard else mine
    mcnorm = float(RegionGeom(config, seed=seed).mcnorm)
    with np.errstate(invalid="ignore", divide="ignore"):
        mcunc = float(np.sqrt(tot[2] / (tot[4] - 1.0) / n_thrown) * mcnorm)
    mcint = float(tot[1] * mcnorm / n_thrown)
    mcintgeo = float(tot[0] * mcnorm / n_thrown)
    passEV = int(tot[3])
    sim.meta["OMCINT"] = (mcint, "Optical MonteCarlo Integral")
    sim.meta["OMCINTGO"] = (mcintgeo, "Optical MonteCarlo Integral, GEO Only")
    sim.meta["ONEVPASS"] = (passEV, "Optical Number of Passing Events")
    sim.meta["OMCINTUN"] = (mcunc, "Stat unc of MonteCarlo Integral")
    logv(f"\tMonte Carlo Integral [Optical]: {mcint}")
    logv(f"\tMonte Carlo Integral, GEO Only [Optical]: {mcintgeo}")
    logv(f"\tNumber of Passing Events [Optical]: {passEV}")
    logv(f"\tStat uncert of MC Integral [Optical]: {mcunc}")
    return sim


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of thrown-event indices owned by ``rank``."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


class AcceptanceRunner:
    """Column-free optical acceptance from in-kernel RNG (``nss_run_acceptance``)."""

    def __init__(self, config, batch: int = 1 << 21, collect_stats: bool = False):
        self.config = config
        self.params = params_from_config(config)
        self.taus = Taus(config)
        if self.params.eas_profile == 2:
            from .simulation.eas_optical import cphotang

            cphotang._ensure_conex_on_device()
        self.cloud = CloudTopHeight(config) if self.params.has_cloud == 2 else None
        self.batch = int(batch)
        L = _lib.lib()
        self.ws = _lib.empty(int(L.nss_acceptance_workspace_bytes(self.batch)), "uint8")
        self.out = _lib.zeros(8)
        self.flags = _lib.zeros(1, "int32")
        self.stats = _lib.zeros(8, "int64") if collect_stats else None
        self.n_kernel_launches = 0

    def enqueue(self, seed: int, first_event: int, n: int):
        """Accumulate events [first_event, first_event + n) into the device sums (async)."""
        if self.cloud is not None:
            self.cloud.register()  # the library holds one map per process
        _lib.check(
            _lib.lib().nss_run_acceptance(
                C.byref(self.params), self.taus.tables, int(seed), int(first_event), int(n),
                self.batch, _lib.ptr(self.out), _lib.ptr(self.stats), _lib.ptr(self.flags),
                _lib.ptr(self.ws), _lib.stream_ptr()),
            "nss_run_acceptance",
        )
        # per batch: front, eas_prep, eas_walk, eas_finish, accept, accept_fold
        self.n_kernel_launches += 6 * ((int(n) + self.batch - 1) // self.batch)

    def reset(self):
        self.out.zero_()
        self.flags.zero_()
        if self.stats is not None:
            self.stats.zero_()

    def sums(self, all_reduce: bool = True):
        """Device sums -> host; all-reduced over ranks when torch.distributed is up."""
        import torch.distributed as dist

        t = self.out.clone()
        if all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        Taus._raise_flags(self.flags)
        return _lib.to_host(t)

    def finalize(self, sums):
        s = np.ascontiguousarray(sums, dtype=np.float64)
        r = np.zeros(4)
        _lib.check(
            _lib.lib().nss_acceptance_finalize(
                C.byref(self.params), s.ctypes.data_as(_lib.c_f64p), r.ctypes.data_as(_lib.c_f64p)),
            "nss_acceptance_finalize",
        )
        return float(r[0]), float(r[1]), int(r[2]), float(r[3])


def run_acceptance(config, n_events: int = None, seed: int = 1234, batch: int = 1 << 21):
    """(mcintegral, mcintegralgeoonly, numEvPass, mcintegraluncert) for ``n_events`` thrown
    events, sharded over the ranks of torch.distributed if it is initialised."""
    import torch.distributed as dist

    n_total = int(n_events if n_events is not None else config.simulation.thrown_events)
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    first, n = shard_range(n_total, rank, world)
    r = AcceptanceRunner(config, batch=batch)
    r.enqueue(seed, first, n)
    return r.finalize(r.sums())
