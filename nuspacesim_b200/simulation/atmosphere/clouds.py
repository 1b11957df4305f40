"""CloudTopHeight: cloud-top altitude callable handed to the EAS stage.

Mirror of ``nuspacesim.simulation.atmosphere.clouds.CloudTopHeight``
(src/nuspacesim/simulation/atmosphere/clouds.py:54-107).  NoCloud and MonoCloud are constants
and travel to the kernel as one scalar.  PressureMapCloud looks the cloud-top pressure up in
a monthly (361, 576) map and converts it with the 1976 US standard atmosphere; on the GPU the
map is registered once as altitudes (``nss_cloud_set_altitude_map``) and every event gets its
own cloud top (``nss_cloud_top_from_map``).

The lookup reproduces the reference AS WRITTEN, which has two quirks (clouds.py:85-96,
compute.py:214-216): both map indices are taken from the latitude
(``j = searchsorted(longitudes, lat)``), and compute() hands in latitudes in radians while
the axes are laid out in degrees -- so only a handful of map cells around the equator are
ever read.  Fixing either would change the reference's numbers; they are kept and documented
(DESIGN.md 7).
"""
from __future__ import annotations

import os

import numpy as np

from ... import _lib

__all__ = ["CloudTopHeight", "us_std_atm_altitude_from_pressure",
           "us_std_atm_pressure_from_altitude", "pressure_map"]

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "data",
                     "cloud_maps")

# 1976 US Standard Atmosphere layer table (src/nuspacesim/constants.py:55-75)
_H_B = np.array([0, 11, 20, 32, 47, 51, 71, 84.852, np.inf])
_LM_B = np.array([-6.5, 0.0, 1.0, 2.8, 0.0, -2.8, -2.0, 0.0, 0.0])
_T_B = np.array([288.15, 216.65, 216.65, 228.65, 270.65, 270.65, 214.65, 186.946, 0.0])
_P_B = 1.01325e5 * np.array([1.0, 2.233611e-1, 5.403295e-2, 8.5666784e-3, 1.0945601e-3,
                             6.6063531e-4, 3.9046834e-5, 3.68501e-6, 0.0])
_GMR = 34.163195
_EARTH_RADIUS = 6371.0  # km, constants.py:48 (not the 6378.1 of the geometry)


def us_std_atm_altitude_from_pressure(P):
    """Geometric altitude [km] at pressure P [Pa]
    (eas_optical/atmospheric_models.py:105-123)."""
    P = np.asarray(P, dtype=np.float64)
    layer = np.zeros(P.shape, dtype=int)
    for j in range(1, len(_P_B)):
        layer[_P_B[j] >= P] = j
    Hb, Lb, Tb, Pb = _H_B[layer], _LM_B[layer], _T_B[layer], _P_B[layer]
    positive = P > 0
    isothermal = Lb == 0
    with np.errstate(all="ignore"):
        dh_iso = Tb * (1.0 / _GMR) * np.log(Pb / P)
        dh_grad = (Tb / Lb) * ((Pb / P) ** ((1.0 / _GMR) * Lb) - 1)
        H = Hb + np.where(positive, np.where(isothermal, dh_iso, dh_grad), 0.0)
        z = _EARTH_RADIUS * H / (_EARTH_RADIUS - H)
    return np.where(positive, z, np.inf)


def us_std_atm_pressure_from_altitude(z):
    """Pressure [Pa] at geometric altitude z [km]: the inverse of the function above
    (eas_optical/atmospheric_models.py:126-145)."""
    z = np.asarray(z, dtype=np.float64)
    finite = z < np.inf
    with np.errstate(all="ignore"):
        h = np.where(finite, z * _EARTH_RADIUS / (z + _EARTH_RADIUS), np.inf)
    layer = np.zeros(h.shape, dtype=int)
    for j in range(1, len(_H_B)):
        layer[_H_B[j] <= h] = j
    Hb, Lb, Tb, Pb = _H_B[layer], _LM_B[layer], _T_B[layer], _P_B[layer]
    isothermal = Lb == 0
    with np.errstate(all="ignore"):
        f_iso = np.exp((-_GMR / Tb) * (h - Hb))
        f_grad = (Tb / (Tb + Lb * (h - Hb))) ** (_GMR / Lb)
    return np.where(finite, Pb * np.where(isothermal, f_iso, f_grad), Pb)


def pressure_map(month: int, version=0) -> np.ndarray:
    """The month's cloud-top pressure map [Pa], (361, 576) float32 (clouds.py:99-107)."""
    fn = os.path.join(_DATA, f"nss_map_CloudTopPressure.v{version}.npz")
    if not os.path.exists(fn):
        raise FileNotFoundError(f"no cloud maps of version {version!r} ({fn})")
    with np.load(fn) as z:
        key = f"m{int(month):02d}"
        if key not in z.files:
            raise ValueError(f"cloud map month must be 1..12, got {month!r}")
        return z[key]


_registered = None  # (device, month, version) of the map currently held by the library


class CloudTopHeight:
    def __init__(self, config):
        self.cloud_model = config.simulation.cloud_model
        cid = getattr(self.cloud_model, "id", "no_cloud") if self.cloud_model else "no_cloud"
        self.map = None
        if cid == "no_cloud":
            self._alt = 0.0
        elif cid == "monocloud":
            self._alt = float(self.cloud_model.altitude)
        elif cid == "pressure_map":
            self.map = pressure_map(self.cloud_model.month, self.cloud_model.version)
            self.latitudes = np.linspace(-90, 90, self.map.shape[0])
            self.longitudes = np.linspace(-180, 180, self.map.shape[1])
            self._alt_map = np.ascontiguousarray(
                us_std_atm_altitude_from_pressure(self.map), dtype=np.float64)
        else:
            raise RuntimeError(f"Unrecognized Cloud Model Type: {cid!r}!")

    @property
    def per_event(self) -> bool:
        """True when the cloud top depends on the event's position (PressureMapCloud)."""
        return self.map is not None

    def __call__(self, lat=0.0, long=0.0, *args, **kwargs):
        if self.map is None:
            return np.single(self._alt)
        # host evaluation of the same lookup, for scalar or array (lat, long)
        lat = np.asarray(lat, dtype=np.float64)
        i = np.minimum(np.searchsorted(self.latitudes, lat), self.map.shape[0] - 1)
        j = np.minimum(np.searchsorted(self.longitudes, lat), self.map.shape[1] - 1)
        return self._alt_map[i, j]

    def register(self):
        """Make this map the one the library looks events up in (once per process/device)."""
        global _registered
        key = (_lib.torch_cuda().cuda.current_device(), int(self.cloud_model.month),
               str(self.cloud_model.version))
        if _registered != key:
            _lib.check(
                _lib.lib().nss_cloud_set_altitude_map(
                    self._alt_map.ctypes.data_as(_lib.c_f64p), self.map.shape[0],
                    self.map.shape[1], self.latitudes.ctypes.data_as(_lib.c_f64p),
                    self.longitudes.ctypes.data_as(_lib.c_f64p)),
                "nss_cloud_set_altitude_map",
            )
            _registered = key

    def device_tops(self, lat, long=None):
        """Cloud top [km] of every event as a device array (``nss_cloud_top_from_map``)."""
        self.register()
        la = _lib.to_device(lat)
        out = _lib.empty(la.numel())
        _lib.check(
            _lib.lib().nss_cloud_top_from_map(_lib.ptr(la), None, la.numel(), _lib.ptr(out),
                                              _lib.stream_ptr()),
            "nss_cloud_top_from_map",
        )
        return out
