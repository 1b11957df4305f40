"""CphotAng: Cherenkov photon density and angle per event.

Mirror of ``nuspacesim.simulation.eas_optical.cphotang.CphotAng``
(src/nuspacesim/simulation/eas_optical/cphotang.py:71-724): same constructor and
``__call__`` signature; the per-event engine (``run`` and everything below it, including
the pybind11 ``zsteps``) is the fused warp-per-event kernel ``nss_cphotang``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib

__all__ = ["CphotAng"]


def cloud_scalar(cloudf):
    """cloudf(lat, long) of the NoCloud/MonoCloud models is a constant (clouds.py:54-82)."""
    if cloudf is None:
        return 0, 0.0
    top = float(cloudf(0.0, 0.0))
    if np.isneginf(top):
        return 0, 0.0
    return 1, top


class CphotAng:
    def __init__(self, detector_altitude, longitudinal_profile_func="Greisen"):
        if longitudinal_profile_func != "Greisen":
            raise NotImplementedError(
                "only the Greisen longitudinal profile is on the accelerated path "
                "(Gaisser-Hillas variants: SURVEY.md 8(f) rank 4)"
            )
        self.detector_altitude = detector_altitude
        self.dtype = np.float32
        self.params = _lib.NssParams()
        self.params.detector_altitude = float(detector_altitude)
        self.params.telescope_effective_area = 1.0
        self.params.quantum_efficiency = 1.0
        self.params.photo_electron_threshold = 1.0

    def __call__(self, betaE, alt, Eshow100PeV, init_lat, init_long, cloudf=None, stats=None):
        """-> (Dphots [photons/m^2], Cang [deg]) per event (cphotang.py:703-724)."""
        if (len(betaE) < 1 or len(alt) < 1 or len(Eshow100PeV) < 1 or len(init_lat) < 1
                or len(init_long) < 1):
            return np.empty([]), np.empty([])
        dev = any(_lib.is_device_array(x) for x in (betaE, alt, Eshow100PeV))
        b, a, e = _lib.to_device(betaE), _lib.to_device(alt), _lib.to_device(Eshow100PeV)
        n = b.numel()
        self.params.has_cloud, self.params.cloud_top = cloud_scalar(cloudf)
        d, c = _lib.empty(n), _lib.empty(n)
        _lib.check(
            _lib.lib().nss_cphotang(C.byref(self.params), _lib.ptr(b), _lib.ptr(a), _lib.ptr(e), n,
                                    _lib.ptr(d), _lib.ptr(c), _lib.ptr(stats), _lib.stream_ptr()),
            "nss_cphotang",
        )
        return (d, c) if dev else (_lib.to_host(d), _lib.to_host(c))

    def run(self, betaE, alt, Eshow100PeV, lat, long, cloudf=None):
        """One event (cphotang.py:623-701)."""
        d, c = self(np.array([betaE], float), np.array([alt], float),
                    np.array([Eshow100PeV], float), [lat], [long], cloudf)
        return np.float32(d[0]), np.float32(c[0])
