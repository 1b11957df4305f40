"""ctypes binding of libnss_b200.so (C ABI declared in include/nss_b200.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, the
operators raise.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# NSS_B200_LIB: development override to A/B-test builds of the same library on one box
LIB_PATH = os.environ.get("NSS_B200_LIB") or os.path.join(HERE, "libnss_b200.so")

c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_ulonglong)

FLAG_LOGE_RANGE = 1
FLAG_U_RANGE = 2


class NssParams(C.Structure):
    """Mirror of ``struct NssParams`` (include/nss_b200.h)."""

    _fields_ = [
        ("detector_altitude", C.c_double),
        ("det_lat", C.c_double),
        ("det_long", C.c_double),
        ("max_cherenkov_angle", C.c_double),
        ("max_azimuth_angle", C.c_double),
        ("angle_from_limb", C.c_double),
        ("telescope_effective_area", C.c_double),
        ("quantum_efficiency", C.c_double),
        ("photo_electron_threshold", C.c_double),
        ("etau_frac", C.c_double),
        ("spectrum_kind", C.c_int32),
        ("has_cloud", C.c_int32),
        ("eas_profile", C.c_int32),
        ("reserved0", C.c_int32),
        ("log_nu_energy", C.c_double),
        ("spec_index", C.c_double),
        ("spec_lower", C.c_double),
        ("spec_upper", C.c_double),
        ("cloud_top", C.c_double),
        ("earth_radius", C.c_double),
        ("core_alt", C.c_double),
        ("min_los", C.c_double),
        ("max_los", C.c_double),
        ("sin_max_theta", C.c_double),
        ("max_phi_s", C.c_double),
        ("min_phi_s", C.c_double),
        ("mcnorm", C.c_double),
    ]


GEOM_FIELDS = (
    "thetaTrSubV", "costhetaTrSubV", "phiTrSubV", "phiS", "losPathLen", "thetaS",
    "costhetaNSubV", "costhetaTrSubN", "betaTrSubN", "latS", "longS", "elevAngVSubN",
    "aziAngVSubN",
)


class NssGeomSoA(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GEOM_FIELDS] + [("event_mask", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/nss_b200.h declares
SIGNATURES = {
    "nss_params_init": (C.c_int, [C.POINTER(NssParams)]),
    "nss_tables_create": (C.c_int, [c_f64p, c_f64p, C.c_int, c_f64p, C.c_int, c_f64p, c_f64p,
                                    C.c_int, c_f64p, C.c_int, c_f64p, C.c_int,
                                    C.POINTER(C.c_void_p)]),
    "nss_tables_free": (None, [C.c_void_p]),
    "nss_geom_throw_u": (C.c_int, [C.POINTER(NssParams), C.c_void_p, C.c_int64,
                                   C.POINTER(NssGeomSoA), C.c_void_p]),
    "nss_geom_throw_philox": (C.c_int, [C.POINTER(NssParams), C.c_uint64, C.c_uint64, C.c_int64,
                                        C.POINTER(NssGeomSoA), C.c_void_p]),
    "nss_scan_workspace_bytes": (C.c_int64, [C.c_int64]),
    "nss_mask_scan": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "nss_compact_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_int, C.c_void_p]),
    "nss_compact_columns_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int,
                                          C.c_int64, C.c_void_p]),
    "nss_event_uniforms": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p]),
    "nss_geom_lat_long": (C.c_int, [C.POINTER(NssParams)] + [C.c_void_p] * 7 + [C.c_int64] +
                          [C.c_void_p] * 3),
    "nss_spectrum_powerlaw": (C.c_int, [C.POINTER(NssParams), C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "nss_tau_exit_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "nss_tau_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nss_tau_sample": (C.c_int, [C.POINTER(NssParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 7),
    "nss_alt_dec": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 3),
    "nss_eas_workspace_bytes": (C.c_int64, [C.c_int64]),
    "nss_eas_optical": (C.c_int, [C.POINTER(NssParams)] + [C.c_void_p] * 6 + [C.c_int64] +
                        [C.c_void_p] * 7),
    "nss_cphotang": (C.c_int, [C.POINTER(NssParams)] + [C.c_void_p] * 5 + [C.c_int64] +
                     [C.c_void_p] * 5),
    "nss_cloud_set_altitude_map": (C.c_int, [c_f64p, C.c_int, C.c_int, c_f64p, c_f64p]),
    "nss_cloud_top_from_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p]),
    "nss_eas_set_conex_table": (C.c_int, [c_f64p, C.c_int]),
    "nss_eas_profile_enable": (C.c_int, [C.c_int]),
    "nss_eas_profile_collect": (C.c_int, [c_f64p, c_i64p]),
    "nss_microbench_peaks": (C.c_int, [c_f64p, c_f64p, C.c_void_p]),
    "nss_mc_workspace_bytes": (C.c_int64, []),
    "nss_mcintegral": (C.c_int, [C.POINTER(NssParams)] + [C.c_void_p] * 7 +
                       [C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double] +
                       [C.c_void_p] * 3),
    "nss_acceptance_workspace_bytes": (C.c_int64, [C.c_int64]),
    "nss_run_acceptance": (C.c_int, [C.POINTER(NssParams), C.c_void_p, C.c_uint64, C.c_uint64,
                                     C.c_int64, C.c_int64] + [C.c_void_p] * 5),
    "nss_acceptance_finalize": (C.c_int, [C.POINTER(NssParams), c_f64p, c_f64p]),
    "nss_philox4x32": (None, [C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]),
    "nss_strerror": (C.c_char_p, [C.c_int]),
    "nss_version": (C.c_int, []),
    "nss_sizeof_params": (C.c_int, []),
}

_LIB = None


class NssLibraryError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (build it first with ``python -m nuspacesim_b200.build``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise NssLibraryError(
                f"{LIB_PATH} not found: build it with `python -m nuspacesim_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback for the hot path."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if L.nss_sizeof_params() != C.sizeof(NssParams):
            raise NssLibraryError("NssParams mirror does not match the compiled library")
        _LIB = L
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().nss_strerror(rc).decode()
        raise NssLibraryError(f"libnss_b200 {what} failed: {msg} (code {rc})")


# ------------------------------------------------------------------ device plumbing (torch)


def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        raise NssLibraryError(
            "no CUDA device visible: the nuspacesim_b200 operators run only on a GPU "
            "(there is no CPU fallback)"
        )
    return torch


def stream_ptr() -> C.c_void_p:
    torch = torch_cuda()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def is_device_array(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def to_device(x, dtype="float64"):
    """NumPy array / scalar sequence / torch tensor -> contiguous CUDA tensor."""
    torch = torch_cuda()
    td = getattr(torch, dtype)
    if hasattr(x, "is_cuda"):
        return x.to(device="cuda", dtype=td).contiguous()
    a = np.ascontiguousarray(x, dtype=dtype)
    return torch.from_numpy(a).to("cuda", non_blocking=False)


_H2D_STREAM = None
_SIDE_STREAM = None


def side_stream():
    """A stream for small read-backs that must not sit in front of the caller's kernels."""
    global _SIDE_STREAM
    if _SIDE_STREAM is None:
        _SIDE_STREAM = torch_cuda().cuda.Stream()
    return _SIDE_STREAM


def to_device_prefetch(x, dtype="float64"):
    """Start the host -> device copy of `x` on a side stream and return (tensor, event): the
    copy of a PINNED host tensor overlaps whatever runs on the current stream; the consumer
    calls ``torch.cuda.current_stream().wait_event(event)`` before its first use.  Anything
    that is not a pinned torch tensor takes the plain synchronous path (event = None)."""
    global _H2D_STREAM
    torch = torch_cuda()
    if not (hasattr(x, "is_pinned") and not x.is_cuda and x.is_pinned()
            and x.dtype == getattr(torch, dtype) and x.is_contiguous()):
        return to_device(x, dtype), None
    if _H2D_STREAM is None:
        _H2D_STREAM = torch.cuda.Stream()
    with torch.cuda.stream(_H2D_STREAM):
        t = x.to("cuda", non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(_H2D_STREAM)
    t.record_stream(torch.cuda.current_stream())
    return t, ev


def rows_to_device_prefetch(x, lo, hi, dtype="float64"):
    """Columns [lo, hi) of a 2-D host array as a contiguous CUDA tensor (rows copied one by one
    so a pinned source needs no host-side repacking) -> (tensor, event or None)."""
    global _H2D_STREAM
    torch = torch_cuda()
    td = getattr(torch, dtype)
    pinned = (hasattr(x, "is_pinned") and not x.is_cuda and x.is_pinned() and x.dtype == td)
    if not pinned:
        if hasattr(x, "is_cuda"):
            return x[:, lo:hi].to(device="cuda", dtype=td).contiguous(), None
        return to_device(np.ascontiguousarray(np.asarray(x)[:, lo:hi], dtype=dtype)), None
    if _H2D_STREAM is None:
        _H2D_STREAM = torch.cuda.Stream()
    with torch.cuda.stream(_H2D_STREAM):
        t = torch.empty((x.shape[0], hi - lo), dtype=td, device="cuda")
        for r in range(x.shape[0]):
            t[r].copy_(x[r, lo:hi], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(_H2D_STREAM)
    t.record_stream(torch.cuda.current_stream())
    return t, ev


def h2d_synchronize():
    """Wait for every prefetch started by `to_device_prefetch` (the caller may reuse or free
    its pinned host buffers afterwards)."""
    if _H2D_STREAM is not None:
        _H2D_STREAM.synchronize()


def empty(n, dtype="float64"):
    torch = torch_cuda()
    return torch.empty(int(n), dtype=getattr(torch, dtype), device="cuda")


def zeros(n, dtype="float64"):
    torch = torch_cuda()
    return torch.zeros(int(n), dtype=getattr(torch, dtype), device="cuda")


def eas_workspace(n):
    """Device scratch for one nss_eas_optical / nss_cphotang call of n events (torch's caching
    allocator hands the same block back to the next call on the same stream)."""
    torch = torch_cuda()
    return torch.empty(int(lib().nss_eas_workspace_bytes(int(n))), dtype=torch.uint8,
                       device="cuda")


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class _HostSink:
    """Device -> pinned host copies on a side stream (async while `deferred` is set)."""

    def __init__(self):
        torch = torch_cuda()
        self.stream = torch.cuda.Stream()
        self.deferred = 0
        self.pending_checks = []

    def copy(self, t) -> np.ndarray:
        torch = torch_cuda()
        t = t.detach()
        if t.numel() < 4096:  # scalars / tiny arrays: the plain path is cheaper
            return t.cpu().numpy()
        t = t.contiguous()
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            out.copy_(t, non_blocking=True)
        t.record_stream(self.stream)  # the caching allocator must not recycle t before the copy
        if not self.deferred:
            self.stream.synchronize()
        return out.numpy()  # the array keeps the pinned tensor alive

    def copy_into(self, dst_pinned, t):
        """Asynchronous device -> host copy of `t` into a slice of a pinned host tensor, on the
        copy stream, after everything enqueued so far on the current stream."""
        torch = torch_cuda()
        t = t.detach().contiguous()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            dst_pinned.copy_(t, non_blocking=True)
        t.record_stream(self.stream)
        if not self.deferred:
            self.stream.synchronize()

    def defer_check(self, flags, raiser):
        """Error flags of a kernel are read back with everything else at the end of the
        deferred region instead of stalling the pipeline behind queued column copies."""
        self.pending_checks.append((flags, raiser))

    def synchronize(self):
        self.stream.synchronize()
        checks, self.pending_checks = self.pending_checks, []
        for flags, raiser in checks:
            raiser(flags)


_SINK = None


def host_sink() -> _HostSink:
    global _SINK
    if _SINK is None:
        _SINK = _HostSink()
    return _SINK


class deferred_host_copies:
    """Within this context `to_host` returns arrays that are filled asynchronously on a copy
    stream (overlapping the next stage's kernels); they are complete when the context
    exits or after `host_sink().synchronize()`."""

    def __enter__(self):
        host_sink().deferred += 1
        return host_sink()

    def __exit__(self, *exc):
        s = host_sink()
        s.deferred -= 1
        s.synchronize()
        return False


def to_host(t) -> np.ndarray:
    """CUDA tensor -> NumPy array through pinned memory (fresh array owned by the caller)."""
    return host_sink().copy(t)
