"""``store=`` / ``plot=`` hooks with the calling convention of the reference.

Reference: src/nuspacesim/utils/decorators.py:53-95 (nss_result_store) and :156-219
(nss_result_plot).  ``store`` is any callable ``(names, columns)``; ``plot`` may be a
callable, or an iterable of callables, invoked as ``plot(args, values)``.  The
reference's named matplotlib plots are not part of this package: plot names (strings)
are accepted and ignored.
"""
from __future__ import annotations

from functools import wraps
from typing import Callable, Iterable


def _host_columns(values):
    """Columns handed to ``store`` are NumPy arrays, as in the reference."""
    from .. import _lib

    return [_lib.to_host(v) if _lib.is_device_array(v) else v for v in values]


def nss_result_store(*names):
    def decorator_store(func):
        @wraps(func)
        def store_f(*args, store=None, **kwargs):
            values = func(*args, **kwargs)
            if store is not None:
                if isinstance(values, tuple):
                    assert len(names) == len(values)
                    store(names, _host_columns(values))
                else:
                    assert len(names) == 1
                    store(names, _host_columns([values]))
            return values

        return store_f

    return decorator_store


def nss_result_plot(*plot_fs):
    def decorator_plot(func):
        @wraps(func)
        def wrapper_f(*args, plot=None, **kwargs):
            values = func(*args, **kwargs)
            if callable(plot):
                plot(args, values)
            elif isinstance(plot, Iterable) and not isinstance(plot, str):
                plot = list(plot)
                if plot and all(callable(p) for p in plot):
                    for p in plot:
                        p(args, values)
            return values

        return wrapper_f

    return decorator_plot
