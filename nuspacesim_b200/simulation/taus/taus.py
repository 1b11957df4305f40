"""Taus: tau exit probability and tau-energy sampling from the nu2tau tables.

Host-side mirror of ``nuspacesim.simulation.taus.taus.Taus``
(src/nuspacesim/simulation/taus/taus.py:57-183): same constructor, methods, return
values and exceptions (``ValueError`` when ``log_e_nu`` leaves the table, as scipy's
``bounds_error=True`` does at taus.py:112-118).  The arithmetic runs in
``nss_tau_exit_prob`` / ``nss_tau_energy`` / ``nss_tau_sample`` (sm_100a, float64).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ... import _lib
from ...params import params_from_config
from ...utils import decorators
from ...utils.grid import NssGrid

__all__ = ["Taus", "massTau", "mean_Tau_life"]

massTau = 1.77686  # GeV/c^2
mean_Tau_life = 2.903e-13  # seconds

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "data")


def _f64p(a):
    return a.ctypes.data_as(_lib.c_f64p)


# compute() builds a Taus per call: the parsed table files and their device copies are kept
# per process.  Cached grids are shared between instances (treat them as read-only; assign
# a new NssGrid to ``pexit_grid`` / ``tau_cdf_grid`` to use other tables -- that instance
# then uploads and owns its own copy).
_GRID_CACHE = {}
_HANDLE_CACHE = {}


def _read_tables(tid, ver):
    key = (str(tid), str(ver))
    if key not in _GRID_CACHE:
        d = os.path.join(_DATA, f"{tid}_tables")
        _GRID_CACHE[key] = (
            NssGrid.read(os.path.join(d, f"nu2tau_pexit.{ver}.h5"), path="/", format="hdf5"),
            NssGrid.read(os.path.join(d, f"nu2tau_cdf.{ver}.h5"), format="hdf5"))
    return _GRID_CACHE[key]


class Taus:
    def __init__(self, config, seed: int = 0):
        self.config = config
        self.params = params_from_config(config)
        tid = config.simulation.tau_shower.id
        ver = config.simulation.tau_shower.table_version
        self.pexit_grid, self.tau_cdf_grid = _read_tables(tid, ver)
        self._cached = (self.pexit_grid, self.tau_cdf_grid)
        self._rng = np.random.Generator(np.random.Philox(key=int(seed)))
        self._handle = None
        self._owns_handle = False

    # device copy of the two grids
    @property
    def tables(self):
        if self._handle is not None and not self._owns_handle and (
                self.pexit_grid is not self._cached[0] or self.tau_cdf_grid is not self._cached[1]):
            self._handle = None  # the grids were replaced after the shared copy was taken
        if self._handle is None:
            torch = _lib.torch_cuda()
            shared = (self.pexit_grid is self._cached[0] and self.tau_cdf_grid is self._cached[1])
            ckey = (id(self.pexit_grid), id(self.tau_cdf_grid), torch.cuda.current_device())
            if shared and ckey in _HANDLE_CACHE:
                self._handle = _HANDLE_CACHE[ckey]
                return self._handle
            pg, cg = self.pexit_grid, self.tau_cdf_grid
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (
                pg.data, pg["log_e_nu"], pg["beta_rad"], cg.data, cg["log_e_nu"],
                cg["beta_rad"], cg["e_tau_frac"])]
            h = C.c_void_p()
            _lib.check(
                _lib.lib().nss_tables_create(
                    _f64p(arrs[0]), _f64p(arrs[1]), len(arrs[1]), _f64p(arrs[2]), len(arrs[2]),
                    _f64p(arrs[3]), _f64p(arrs[4]), len(arrs[4]), _f64p(arrs[5]), len(arrs[5]),
                    _f64p(arrs[6]), len(arrs[6]), C.byref(h)),
                "nss_tables_create",
            )
            self._handle = h
            self._owns_handle = not shared
            if shared:
                _HANDLE_CACHE[ckey] = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None and self._owns_handle:
                _lib.lib().nss_tables_free(self._handle)
                self._handle = None
        except Exception:
            pass

    @staticmethod
    def _check_flags(flags):
        """Raise now, or -- inside compute()'s deferred region -- when its copies are drained
        (a flag read-back would otherwise queue behind the column copies)."""
        sink = _lib.host_sink()
        if sink.deferred:
            sink.defer_check(flags, Taus._raise_flags)
        else:
            Taus._raise_flags(flags)

    @staticmethod
    def _raise_flags(flags):
        f = int(flags.item())
        if f & _lib.FLAG_LOGE_RANGE:
            raise ValueError("One of the requested xi is out of bounds in dimension 0")
        if f & _lib.FLAG_U_RANGE:
            raise ValueError("u must lie inside (cdf[0], cdf[-1]] of the interpolated CDF")

    def _inputs(self, betas, log_e_nu):
        dev = _lib.is_device_array(betas) or _lib.is_device_array(log_e_nu)
        b, le = _lib.to_device(betas), _lib.to_device(log_e_nu)
        if b.shape != le.shape:
            raise RuntimeError(f"Input Array Shape Mismatch: {tuple(le.shape)}, {tuple(b.shape)}")
        return b, le, dev

    def _uniform(self, n, u, u_raw):
        """-> (device uniforms, raw flag).  Uniforms drawn here are plain (0, 1) draws that the
        kernel maps onto (cdf[0], 1) of each event's blended CDF row, like the reference's
        np.random.uniform(cdfs[:, 0], 1.0) (utils/cdf.py:419-423); a caller's ``u`` is used as
        it is unless ``u_raw`` says it is such a plain draw too."""
        if u is None:
            # open interval (0, 1): u == cdf[0] breaks the crossing search
            u = (self._rng.integers(0, 2**53, size=n).astype(np.float64) + 0.5) / 2.0**53
            u_raw = True
        return _lib.to_device(u), int(bool(u_raw))

    def tau_exit_prob(self, betas, log_e_nu):
        """Tau Exit Probability (taus.py:99-122)."""
        b, le, dev = self._inputs(betas, log_e_nu)
        out, flags = _lib.empty(b.numel()), _lib.zeros(1, "int32")
        _lib.check(
            _lib.lib().nss_tau_exit_prob(self.tables, _lib.ptr(b), _lib.ptr(le), b.numel(),
                                         _lib.ptr(out), _lib.ptr(flags), _lib.stream_ptr()),
            "nss_tau_exit_prob",
        )
        self._check_flags(flags)
        return out if dev else _lib.to_host(out)

    def tau_energy(self, betas, log_e_nu, u=None, u_raw=False):
        """Tau energies sampled from the interpolated CDF (taus.py:124-146).

        ``u`` holds one uniform per event (event i consumes u[i]); ``u_raw`` (extension): the
        given ``u`` are plain (0, 1) draws to be mapped onto (cdf[0], 1) per event."""
        b, le, dev = self._inputs(betas, log_e_nu)
        ud, raw = self._uniform(b.numel(), u, u_raw)
        out, flags = _lib.empty(b.numel()), _lib.zeros(1, "int32")
        _lib.check(
            _lib.lib().nss_tau_energy(self.tables, _lib.ptr(b), _lib.ptr(le), _lib.ptr(ud), raw,
                                      b.numel(), _lib.ptr(out), _lib.ptr(flags),
                                      _lib.stream_ptr()),
            "nss_tau_energy",
        )
        self._check_flags(flags)
        return out if dev else _lib.to_host(out)

    @decorators.nss_result_plot()
    @decorators.nss_result_store(
        "tauBeta", "tauLorentz", "tauEnergy", "showerEnergy", "tauExitProb"
    )
    def __call__(self, betas, log_e_nu, *args, u=None, u_raw=False, **kwargs):
        """taus.py:154-183 -> (tauBeta, tauLorentz, tauEnergy, showerEnergy, tauExitProb)."""
        b, le, dev = self._inputs(betas, log_e_nu)
        n = b.numel()
        ud, raw = self._uniform(n, u, u_raw)
        outs = [_lib.empty(n) for _ in range(5)]
        flags = _lib.zeros(1, "int32")
        _lib.check(
            _lib.lib().nss_tau_sample(C.byref(self.params), self.tables, _lib.ptr(b),
                                      _lib.ptr(le), _lib.ptr(ud), raw, n,
                                      *[_lib.ptr(o) for o in outs],
                                      _lib.ptr(flags), _lib.stream_ptr()),
            "nss_tau_sample",
        )
        self._check_flags(flags)
        return tuple(outs) if dev else tuple(_lib.to_host(o) for o in outs)
