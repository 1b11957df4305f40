"""EAS: decay altitude and the optical Cherenkov air-shower stage.

Mirror of ``nuspacesim.simulation.eas_optical.eas.EAS``
(src/nuspacesim/simulation/eas_optical/eas.py:48-138): same constructor, ``altDec`` and
``__call__`` signatures.  ``altDec`` runs in ``nss_alt_dec``; ``__call__`` (altDec mask,
CphotAng, numPEs, angle widening, cosine) is the single fused kernel ``nss_eas_optical``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib
from ...params import params_from_config
from ...utils import decorators
from .cphotang import CphotAng, cloud_args

__all__ = ["EAS"]


class EAS:
    """Electromagnetic Air Shower wrapper class."""

    def __init__(self, config, seed: int = 0):
        self.config = config
        self.params = params_from_config(config)
        self.CphotAng = CphotAng(
            config.detector.initial_position.altitude, config.simulation.eas_long_profile,
            seed=seed,
        )
        self._rng = np.random.Generator(np.random.Philox(key=int(seed) + 2))
        #: optional device uint64[6] work counters filled by the kernel
        self.stats = None

    @decorators.nss_result_store("altDec", "lenDec")
    def altDec(self, beta, tauBeta, tauLorentz, u=None, *args, **kwargs):
        """Decay altitude (eas.py:61-81) -> (altDec [km], lenDec [km])."""
        dev = any(_lib.is_device_array(x) for x in (beta, tauBeta, tauLorentz))
        b, tb, tl = _lib.to_device(beta), _lib.to_device(tauBeta), _lib.to_device(tauLorentz)
        n = b.numel()
        if u is None:
            u = 1.0 - self._rng.random(n)  # (0, 1]: log(u) must be finite
        ud = _lib.to_device(u)
        alt, length = _lib.empty(n), _lib.empty(n)
        _lib.check(
            _lib.lib().nss_alt_dec(_lib.ptr(b), _lib.ptr(tb), _lib.ptr(tl), _lib.ptr(ud), n,
                                   _lib.ptr(alt), _lib.ptr(length), _lib.stream_ptr()),
            "nss_alt_dec",
        )
        return (alt, length) if dev else (_lib.to_host(alt), _lib.to_host(length))

    @decorators.nss_result_plot()
    @decorators.nss_result_store("numPEs", "costhetaChEff")
    def __call__(self, beta, altDec, showerEnergy, init_lat, init_long, *args, cloudf=None,
                 u_profile=None, **kwargs):
        """Electromagnetic Air Shower operation (eas.py:83-138) -> (numPEs, costhetaChEff).

        ``u_profile`` (extension, "Gaisser-Hillas Fluctuated" only): one uniform per event
        selecting the CONEX row; drawn from the seeded generator when omitted."""
        dev = any(_lib.is_device_array(x) for x in (beta, altDec, showerEnergy))
        b, a, e = _lib.to_device(beta), _lib.to_device(altDec), _lib.to_device(showerEnergy)
        n = b.numel()
        p = self.params
        p.has_cloud, p.cloud_top, ctops = cloud_args(cloudf, init_lat, init_long, n)
        npe, cth = _lib.empty(n), _lib.empty(n)
        up = self.CphotAng.profile_uniforms(n, u_profile)
        ws = _lib.eas_workspace(n)
        _lib.check(
            _lib.lib().nss_eas_optical(C.byref(p), _lib.ptr(b), _lib.ptr(a), _lib.ptr(e), None,
                                       _lib.ptr(up), _lib.ptr(ctops), n, _lib.ptr(npe),
                                       _lib.ptr(cth), None, None,
                                       _lib.ptr(self.stats), _lib.ptr(ws), _lib.stream_ptr()),
            "nss_eas_optical",
        )
        return (npe, cth) if dev else (_lib.to_host(npe), _lib.to_host(cth))
