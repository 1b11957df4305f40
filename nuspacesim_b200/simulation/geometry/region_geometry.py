"""RegionGeom: trajectory sampling, line-of-sight geometry and the MC acceptance integral.

Host-side mirror of the reference class
``nuspacesim.simulation.geometry.region_geometry.RegionGeom``
(src/nuspacesim/simulation/geometry/region_geometry.py:46-409): same constructor,
methods, argument meaning and exceptions.  All arithmetic runs in the sm_100a kernels
of libnss_b200.so (``nss_geom_throw_*``, ``nss_mask_scan`` / ``nss_compact_f64``,
``nss_geom_lat_long``, ``nss_mcintegral``); the object's state lives in device memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib
from ...params import params_from_config
from ...utils import decorators

__all__ = ["RegionGeom"]


class RegionGeom:
    """Region Geometry of trajectories."""

    def __init__(self, config, seed: int = 0):
        self.config = config
        self.params = params_from_config(config)
        p = self.params
        # the per-run constants of region_geometry.py:49-92, computed by nss_params_init
        self.earth_radius = p.earth_radius
        self.earth_rad_2 = p.earth_radius**2
        self.core_alt = p.core_alt
        self.detection_mode = config.simulation.mode
        self.detLat = p.det_lat
        self.detLong = p.det_long
        self.minLOSpathLen = p.min_los
        self.maxLOSpathLen = p.max_los
        self.sinOfMaxThetaTrSubV = p.sin_max_theta
        self.maxPhiS = p.max_phi_s
        self.minPhiS = p.min_phi_s
        self.mcnorm = p.mcnorm
        #: Philox key for ``throw(int)`` (extension: the reference uses the global,
        #: unseeded np.random state, region_geometry.py:99)
        self.seed = int(seed)
        self.first_event = 0
        self._dev = {}
        self._n = 0
        self._valid_cache = {}
        self._pos = None
        self._nv, self._nv_pending = 0, None

    # ------------------------------------------------------------------ throw

    def throw(self, u=None, defer_count: bool = False):
        """Throw N events with 4 * u random numbers (region_geometry.py:94-239).

        ``defer_count`` (extension, used by compute()'s pipelined path): do not wait for the
        number of valid events; it is copied to pinned host memory behind the scan and read --
        waiting for THAT copy only -- when first needed, so kernels enqueued after this call do
        not delay it."""
        L = _lib.lib()
        if u is None:
            raise RuntimeError(
                "Provide a number of trajectories, or a 2D set of uniform random "
                "numbers in [0, 1]"
            )
        soa = _lib.NssGeomSoA()
        if isinstance(u, (int, np.integer)):
            n = int(u)
            self._alloc(n, soa)
            _lib.check(
                L.nss_geom_throw_philox(C.byref(self.params), self.seed, self.first_event, n,
                                        C.byref(soa), _lib.stream_ptr()),
                "nss_geom_throw_philox",
            )
        else:
            if u.shape[0] != 4:
                raise RuntimeError("u random numbers must be of shape (4, N)")
            ud = _lib.to_device(u)
            n = int(ud.shape[1])
            self._alloc(n, soa)
            _lib.check(
                L.nss_geom_throw_u(C.byref(self.params), _lib.ptr(ud), n, C.byref(soa),
                                   _lib.stream_ptr()),
                "nss_geom_throw_u",
            )
        self._scan(defer_count)

    def _alloc(self, n, soa):
        self._n = n
        self._dev = {name: _lib.empty(n) for name in _lib.GEOM_FIELDS}
        self._dev["event_mask"] = _lib.empty(n, "uint8")
        for name in _lib.GEOM_FIELDS:
            setattr(soa, name, self._dev[name].data_ptr())
        soa.event_mask = self._dev["event_mask"].data_ptr()
        self._valid_cache = {}

    @property
    def _n_valid(self):
        if self._nv is None:
            pinned, ev = self._nv_pending
            ev.synchronize()
            self._nv = int(pinned.item())
            self._nv_pending = None
        return self._nv

    def _scan(self, defer_count=False):
        L = _lib.lib()
        n = self._n
        self._pos = _lib.empty(max(n, 1), "int64")
        nset = _lib.zeros(1, "int64")
        ws = _lib.empty(max(int(L.nss_scan_workspace_bytes(n)), 8), "uint8")
        _lib.check(
            L.nss_mask_scan(_lib.ptr(self._dev["event_mask"]), n, _lib.ptr(self._pos),
                            _lib.ptr(nset), _lib.ptr(ws), _lib.stream_ptr()),
            "nss_mask_scan",
        )
        if defer_count and n > 0:
            # on a side stream: behind the scan, but not in front of whatever the caller
            # enqueues next on its own stream (an 8-byte copy queues behind megabytes of
            # column copies in the copy engine; the compute stream must not wait for that)
            torch = _lib.torch_cuda()
            pinned = torch.empty(1, dtype=torch.int64, pin_memory=True)
            done = torch.cuda.Event()
            done.record()
            side = _lib.side_stream()
            side.wait_event(done)
            with torch.cuda.stream(side):
                pinned.copy_(nset, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            nset.record_stream(side)
            self._nv, self._nv_pending = None, (pinned, ev)
        else:
            self._nv = int(nset.item())

    def compact(self, full, device: bool = True):
        """Entries of a length-N (thrown-order) array that belong to valid events, in thrown
        order -- ``full[event_mask]`` without the host synchronisation of boolean indexing
        (the positions are known from the scan throw() already did)."""
        f = _lib.to_device(full)
        if f.numel() != self._n:
            raise RuntimeError(f"expected one entry per thrown event ({self._n}), got {f.numel()}")
        out = _lib.empty(self._n_valid)
        if self._n_valid == 0:  # nothing to move (and an empty tensor has no address to pass)
            return out if device else _lib.to_host(out)
        _lib.check(
            _lib.lib().nss_compact_f64(
                _lib.ptr(self._dev["event_mask"]), _lib.ptr(self._pos), _lib.ptr(f),
                _lib.ptr(out), self._n, 0, _lib.stream_ptr()),
            "nss_compact_f64",
        )
        return out if device else _lib.to_host(out)

    def event_uniforms(self, which: int, device: bool = True):
        """Uniform `which` (0 spectrum, 1 tau energy, 2 decay length) of every VALID event's
        Philox stream, in thrown order: the numbers the column-free mode draws in-kernel for
        the same seed (extension; the reference draws from the global NumPy state)."""
        out = _lib.empty(max(self._n_valid, 1))[: self._n_valid]
        if self._n_valid:
            _lib.check(
                _lib.lib().nss_event_uniforms(
                    self.seed, self.first_event, self._n, _lib.ptr(self._dev["event_mask"]),
                    _lib.ptr(self._pos), int(which), _lib.ptr(out), _lib.stream_ptr()),
                "nss_event_uniforms",
            )
        return out if device else _lib.to_host(out)

    # full-length attributes of the reference object, materialised on demand
    def __getattr__(self, name):
        dev = self.__dict__.get("_dev", {})
        if name in dev:
            a = _lib.to_host(dev[name])
            return a.astype(bool) if name == "event_mask" else a
        raise AttributeError(name)

    # ------------------------------------------------------------------ accessors (:241-287)

    #: the accessors compute() uses on every throw: compacted together, in ONE launch, the
    #: first time any of them is asked for (mask and positions are read once, not 13 times)
    _COMMON = (("betaTrSubN", True), ("thetaTrSubV", False), ("losPathLen", False),
               ("phiTrSubV", False), ("elevAngVSubN", False), ("aziAngVSubN", False),
               ("longS", False), ("latS", False), ("costhetaTrSubN", False),
               ("costhetaNSubV", False), ("costhetaTrSubV", False))

    def _compact_many(self, keys):
        keys = [k for k in keys if k not in self._valid_cache]
        outs = [_lib.empty(self._n_valid) for _ in keys]
        if self._n_valid and keys:
            # a slice or rank shard with no valid event yields empty columns (compute.py:200-204)
            k = len(keys)
            ins = (C.c_void_p * k)(*[self._dev[name].data_ptr() for name, _ in keys])
            dst = (C.c_void_p * k)(*[o.data_ptr() for o in outs])
            rad = (C.c_int * k)(*[int(r) for _, r in keys])
            _lib.check(
                _lib.lib().nss_compact_columns_f64(
                    _lib.ptr(self._dev["event_mask"]), _lib.ptr(self._pos), ins, dst, rad, k,
                    self._n, _lib.stream_ptr()),
                "nss_compact_columns_f64",
            )
        for key, o in zip(keys, outs):
            self._valid_cache[key] = o

    def _valid(self, name, to_radians=False, device=False):
        key = (name, to_radians)
        if key not in self._valid_cache:
            self._compact_many(self._COMMON if key in self._COMMON else (key,))
        t = self._valid_cache[key]
        return t if device else _lib.to_host(t)

    def betas(self, device=False):
        """Earth-emergence angles (degrees) for valid events."""
        return self._valid("betaTrSubN", device=device)

    def beta_rad(self, device=False):
        """Radian Earth-emergence angles for valid events."""
        return self._valid("betaTrSubN", True, device=device)

    def thetas(self, device=False):
        return self._valid("thetaTrSubV", device=device)

    def phis(self, device=False):
        return self._valid("phiTrSubV", device=device)

    def pathLens(self, device=False):
        return self._valid("losPathLen", device=device)

    def valid_costhetaTrSubN(self, device=False):
        return self._valid("costhetaTrSubN", device=device)

    def valid_costhetaNSubV(self, device=False):
        return self._valid("costhetaNSubV", device=device)

    def valid_costhetaTrSubV(self, device=False):
        return self._valid("costhetaTrSubV", device=device)

    def valid_longS(self, device=False):
        return self._valid("longS", device=device)

    def valid_latS(self, device=False):
        return self._valid("latS", device=device)

    def valid_longS_rad(self, device=False):
        return self._valid("longS", True, device=device)

    def valid_latS_rad(self, device=False):
        return self._valid("latS", True, device=device)

    def valid_elevAngVSubN(self, device=False):
        return self._valid("elevAngVSubN", device=device)

    def valid_aziAngVSubN(self, device=False):
        return self._valid("aziAngVSubN", device=device)

    def find_lat_long_along_traj(self, dist_along_traj, device=False):
        """Latitude/longitude (rad) at a distance along each valid trajectory (:289-355)."""
        nv = self._n_valid
        dev_in = _lib.is_device_array(dist_along_traj)
        d = _lib.to_device(dist_along_traj)
        if d.numel() != nv:
            raise ValueError(f"operands could not be broadcast together with shapes ({d.numel()},) ({nv},)")
        lat, lon = _lib.empty(nv), _lib.empty(nv)
        _lib.check(
            _lib.lib().nss_geom_lat_long(
                C.byref(self.params), _lib.ptr(self.thetas(True)), _lib.ptr(self.phis(True)),
                _lib.ptr(self.valid_elevAngVSubN(True)), _lib.ptr(self.valid_aziAngVSubN(True)),
                _lib.ptr(self.valid_longS(True)), _lib.ptr(self.valid_latS(True)), _lib.ptr(d),
                nv, _lib.ptr(lat), _lib.ptr(lon), _lib.stream_ptr()),
            "nss_geom_lat_long",
        )
        if device or dev_in:
            return lat, lon
        return _lib.to_host(lat), _lib.to_host(lon)

    @decorators.nss_result_plot()
    @decorators.nss_result_store("beta_rad", "theta_rad", "path_len")
    def __call__(self, numtrajs, *args, device=False, **kwargs):
        """Throw numtrajs events and return valid betas (:357-362)."""
        self.throw(numtrajs)
        return self.beta_rad(device), self.thetas(device), self.pathLens(device)

    # ------------------------------------------------------------------ mcintegral (:364-409)

    def mcintegral(self, triggers, costheta, tauexitprob, threshold, spec_norm,
                   spec_weights_sum, **kwargs):
        """Monte Carlo integral -> (mcintegral, mcintegralgeoonly, numEvPass, mcintegraluncert)."""
        L = _lib.lib()
        nv = self._n_valid
        trig = _lib.to_device(triggers)
        pex = _lib.to_device(tauexitprob)
        if np.ndim(costheta) == 0 and not hasattr(costheta, "is_cuda"):
            cth = _lib.to_device(np.full(nv, float(costheta)))
        else:
            cth = _lib.to_device(costheta)
        for a in (trig, pex, cth):
            if a.numel() != nv:
                raise ValueError(
                    f"operands could not be broadcast together with shapes ({nv},) ({a.numel()},)")
        out = _lib.zeros(8)
        ws = _lib.empty(int(L.nss_mc_workspace_bytes()), "uint8")
        _lib.check(
            L.nss_mcintegral(
                C.byref(self.params), _lib.ptr(self.valid_costhetaTrSubN(True)),
                _lib.ptr(self.valid_costhetaNSubV(True)),
                _lib.ptr(self.valid_costhetaTrSubV(True)), _lib.ptr(trig), _lib.ptr(cth),
                _lib.ptr(pex), None, nv, max(self._n, 1), float(threshold), float(spec_norm),
                float(spec_weights_sum), _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr()),
            "nss_mcintegral",
        )
        if kwargs.get("device_sums", False):
            # extension (compute()'s sliced pipeline): the raw device sums {sum f_geo, sum f,
            # sum (f - mean)^2, n_pass, n_valid, ...} without a host synchronisation
            return out
        o = _lib.to_host(out)
        return float(o[5]), float(o[6]), int(o[3]), float(o[7])
