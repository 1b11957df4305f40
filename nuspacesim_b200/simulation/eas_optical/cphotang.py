"""CphotAng: Cherenkov photon density and angle per event.

Mirror of ``nuspacesim.simulation.eas_optical.cphotang.CphotAng``
(src/nuspacesim/simulation/eas_optical/cphotang.py:71-724): same constructor and
``__call__`` signature; the per-event engine (``run`` and everything below it, including
the pybind11 ``zsteps``) is the fused warp-per-event kernel ``nss_cphotang``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ... import _lib

__all__ = ["CphotAng", "PROFILES", "conex_table"]

#: simulation.eas_long_profile -> NssParams.eas_profile (cphotang.py:256-276)
PROFILES = {"Greisen": 0, "Gaisser-Hillas Parameterized": 1, "Gaisser-Hillas Fluctuated": 2}

_CONEX = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                      "data", "CONEX_table", "gh_conex_pi_E17_95deg_0km_eposlhc.npy")
_conex_uploaded = None


def conex_table():
    """(1000, 6) float64 rows Nm, Xm, X0, p1, p2, p3: columns 4..9 of the CONEX fit table the
    reference loads with np.loadtxt(usecols=(4..9)) (cphotang.py:266-276)."""
    return np.load(_CONEX)


def _ensure_conex_on_device():
    """Upload the CONEX table once per process (and again after a device change)."""
    global _conex_uploaded
    dev = _lib.torch_cuda().cuda.current_device()
    if _conex_uploaded != dev:
        t = np.ascontiguousarray(conex_table(), dtype=np.float64)
        _lib.check(_lib.lib().nss_eas_set_conex_table(t.ctypes.data_as(_lib.c_f64p), t.shape[0]),
                   "nss_eas_set_conex_table")
        _conex_uploaded = dev


def cloud_args(cloudf, lat, long, n):
    """-> (has_cloud, cloud_top, per-event device array or None) for ``cloudf(lat, long)``
    (cphotang.py:644).  NoCloud/MonoCloud are constants (clouds.py:54-82); a
    ``CloudTopHeight`` built from a PressureMapCloud is looked up on the GPU; any other
    callable is evaluated on the host, vectorised over the events if it accepts arrays."""
    if cloudf is None:
        return 0, 0.0, None
    if getattr(cloudf, "per_event", False):
        return 2, 0.0, cloudf.device_tops(lat, long)
    if hasattr(cloudf, "per_event"):  # our own constant models
        tops = cloudf(0.0, 0.0)
    else:
        la = _lib.to_host(lat) if _lib.is_device_array(lat) else np.asarray(lat, dtype=float)
        lo = _lib.to_host(long) if _lib.is_device_array(long) else np.asarray(long, dtype=float)
        try:
            tops = np.asarray(cloudf(la, lo), dtype=np.float64)
        except Exception:
            tops = np.zeros(0)
        if tops.ndim != 0 and tops.shape != (n,):
            tops = np.array([float(cloudf(a, b)) for a, b in zip(la, lo)])
    tops = np.asarray(tops, dtype=np.float64)
    if tops.ndim == 0:
        top = float(tops)
        return (0, 0.0, None) if np.isneginf(top) else (1, top, None)
    return 2, 0.0, _lib.to_device(tops)


class CphotAng:
    def __init__(self, detector_altitude, longitudinal_profile_func="Greisen", seed: int = 0):
        if longitudinal_profile_func not in PROFILES:
            # the reference leaves particle_count undefined here and fails on first use
            raise ValueError(f"unknown longitudinal profile {longitudinal_profile_func!r}")
        self.longitudinal_profile_func = longitudinal_profile_func
        self._rng = np.random.Generator(np.random.Philox(key=int(seed) + 3))
        self.detector_altitude = detector_altitude
        self.dtype = np.float32
        self.params = _lib.NssParams()
        self.params.detector_altitude = float(detector_altitude)
        self.params.telescope_effective_area = 1.0
        self.params.quantum_efficiency = 1.0
        self.params.photo_electron_threshold = 1.0
        self.params.eas_profile = PROFILES[longitudinal_profile_func]

    def profile_uniforms(self, n, u_profile=None):
        """Device array of the per-event uniforms that pick the CONEX row (row = floor(u *
        nrows); the reference calls np.random.randint per event, shower_properties.py:165),
        or None for the profiles that draw nothing."""
        if self.params.eas_profile != 2:
            return None
        _ensure_conex_on_device()
        if u_profile is None:
            u_profile = self._rng.random(n)
        up = _lib.to_device(u_profile)
        if up.numel() != n:
            raise RuntimeError("u_profile must hold one uniform per event")
        return up

    def __call__(self, betaE, alt, Eshow100PeV, init_lat, init_long, cloudf=None, stats=None,
                 u_profile=None):
        """-> (Dphots [photons/m^2], Cang [deg]) per event (cphotang.py:703-724)."""
        if (len(betaE) < 1 or len(alt) < 1 or len(Eshow100PeV) < 1 or len(init_lat) < 1
                or len(init_long) < 1):
            return np.empty([]), np.empty([])
        dev = any(_lib.is_device_array(x) for x in (betaE, alt, Eshow100PeV))
        b, a, e = _lib.to_device(betaE), _lib.to_device(alt), _lib.to_device(Eshow100PeV)
        n = b.numel()
        self.params.has_cloud, self.params.cloud_top, ctops = cloud_args(
            cloudf, init_lat, init_long, n)
        d, c = _lib.empty(n), _lib.empty(n)
        up = self.profile_uniforms(n, u_profile)
        ws = _lib.eas_workspace(n)
        _lib.check(
            _lib.lib().nss_cphotang(C.byref(self.params), _lib.ptr(b), _lib.ptr(a), _lib.ptr(e),
                                    _lib.ptr(up), _lib.ptr(ctops), n, _lib.ptr(d), _lib.ptr(c),
                                    _lib.ptr(stats), _lib.ptr(ws), _lib.stream_ptr()),
            "nss_cphotang",
        )
        return (d, c) if dev else (_lib.to_host(d), _lib.to_host(c))

    def run(self, betaE, alt, Eshow100PeV, lat, long, cloudf=None, u_profile=None):
        """One event (cphotang.py:623-701)."""
        d, c = self(np.array([betaE], float), np.array([alt], float),
                    np.array([Eshow100PeV], float), [lat], [long], cloudf,
                    u_profile=None if u_profile is None else np.array([u_profile], float))
        return np.float32(d[0]), np.float32(c[0])
