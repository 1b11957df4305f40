"""Flatten an ``NssConfig``-shaped object into the POD ``NssParams`` the kernels take.

The object is duck-typed: the reference's own pydantic ``NssConfig``
(src/nuspacesim/config.py:370-382) works as well as ``nuspacesim_b200.config.NssConfig``.
"""
from __future__ import annotations

import math

from . import _lib


def params_from_config(config) -> _lib.NssParams:
    p = _lib.NssParams()
    det, sim = config.detector, config.simulation
    p.detector_altitude = float(det.initial_position.altitude)
    p.det_lat = float(det.initial_position.latitude)
    p.det_long = float(det.initial_position.longitude)
    p.max_cherenkov_angle = float(sim.max_cherenkov_angle)
    p.max_azimuth_angle = float(sim.max_azimuth_angle)
    p.angle_from_limb = float(sim.angle_from_limb)
    opt = det.optical
    p.telescope_effective_area = float(opt.telescope_effective_area) if opt else 2.5
    p.quantum_efficiency = float(opt.quantum_efficiency) if opt else 0.2
    p.photo_electron_threshold = float(opt.photo_electron_threshold) if opt else 10.0
    p.etau_frac = float(sim.tau_shower.etau_frac)
    spec = sim.spectrum
    if getattr(spec, "id", "monospectrum") == "powerspectrum":
        p.spectrum_kind = 1
        p.spec_index = float(spec.index)
        p.spec_lower = float(spec.lower_bound)
        p.spec_upper = float(spec.upper_bound)
        p.log_nu_energy = 0.0
    else:
        p.spectrum_kind = 0
        p.log_nu_energy = float(spec.log_nu_energy)
        p.spec_index, p.spec_lower, p.spec_upper = 2.0, 6.0, 12.0
    prof = getattr(sim, "eas_long_profile", "Greisen")
    p.eas_profile = {"Default": 0, "Greisen": 0, "Gaisser-Hillas Parameterized": 1,
                     "Gaisser-Hillas Fluctuated": 2}[prof]
    cloud = getattr(sim, "cloud_model", None)
    cid = getattr(cloud, "id", "no_cloud")
    if cid == "monocloud":
        # CloudTopHeight -> mono_altitude(altitude) (atmosphere/clouds.py:54-82)
        p.has_cloud = 1
        p.cloud_top = float(cloud.altitude)
        if math.isinf(p.cloud_top) and p.cloud_top < 0:
            p.has_cloud = 0
            p.cloud_top = 0.0
    elif cid == "pressure_map":
        # one cloud top per event, looked up in the month's map (clouds.py:67-69, 85-107);
        # CloudTopHeight.register() hands the map to the library
        p.has_cloud = 2
        p.cloud_top = 0.0
    else:
        # NoCloud: cloudf returns 0.0 km, below every step -> identical to no cloud
        p.has_cloud = 0
        p.cloud_top = 0.0
    _lib.check(_lib.lib().nss_params_init(p), "nss_params_init")
    return p
