"""NssGrid: N-D gridded data with named axes, in the reference's table format.

Reference: src/nuspacesim/utils/grid.py:118-267 (class, an astropy NDDataArray subclass)
and :286-367 (HDF5 / FITS readers and writers registered with astropy's I/O registry).
Here the class is a plain NumPy container and the HDF5 layout
(root attrs ``AXIS0..n``, dataset ``__nss_grid_data__``, group ``__nss_grid_axes__/<name>``)
is read and written by this package's own ``hdf5_min`` (h5py / astropy are not needed).
"""
from __future__ import annotations

from typing import Any, List, Sequence, Union

import numpy as np

from . import hdf5_min

__all__ = ["NssGrid"]


class NssGrid:
    def __init__(self, data, axes: Sequence[np.ndarray], axis_names: Sequence[str]):
        data = np.asarray(data)
        axes = [np.asarray(a) for a in axes]
        axis_names = list(axis_names)
        if len(axes) != len(axis_names):
            raise ValueError("Must give same number of names as axes.")
        if len(axes) != data.ndim:
            raise ValueError(
                f"Must give same number of axes({len(axes)}) as grid dimensions ({data.ndim})."
            )
        for a in axes:
            if a.ndim != 1:
                raise ValueError(f"Each axis must be 1D. Got {a.ndim} from {a.shape}.")
        for i in range(data.ndim):
            if axes[i].shape[0] != data.shape[i]:
                raise ValueError("Axes lengths must correspond to grid dimensions.")
        self._data = data
        self._axes = axes
        self._names = axis_names
        self.meta = {f"AXIS{i}": n for i, n in enumerate(axis_names)}

    # ---- NDDataArray-like surface
    @property
    def data(self) -> np.ndarray:
        return self._data

    @property
    def ndim(self) -> int:
        return self._data.ndim

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def axes(self) -> List[np.ndarray]:
        return self._axes

    @property
    def axis_names(self) -> List[str]:
        return self._names

    def __repr__(self) -> str:
        rep = "NssGrid {\n" + f"meta : {self.meta!r}\n"
        for axis, name in zip(self.axes, self.axis_names):
            rep += f"{name} ({len(axis)}): {axis}\n"
        rep += f"data: {self._data!r}\ndims: {self._data.shape}\n}}"
        return rep

    __str__ = __repr__

    def __eq__(self, other) -> bool:
        return (
            np.array_equal(self.data, other.data)
            and self.axis_names == other.axis_names
            and all(np.array_equal(s, o) for s, o in zip(self.axes, other.axes))
        )

    def __getitem__(self, item) -> Union[np.ndarray, Any]:
        if self._data.shape == ():
            raise TypeError("scalars cannot be sliced.")
        if isinstance(item, str):
            if item not in self._names:
                raise KeyError(f"Axis name ({item}) not found.")
            return self._axes[self._names.index(item)]
        if isinstance(item, (int, np.integer)):
            if item < len(self._axes):
                return self._axes[item]
            raise IndexError(f"Array index ({item}) out of bounds.")
        # slicing: keep the axes that survive (slices / masks selecting > 1 element)
        if not isinstance(item, tuple):
            item = (item,)
        keep_axes, keep_names = [], []
        for i, s in enumerate(item):
            if isinstance(s, slice) or np.count_nonzero(s) > 1:
                keep_axes.append(self._axes[i][s])
                keep_names.append(self._names[i])
        for i in range(len(item), self.ndim):
            keep_axes.append(self._axes[i])
            keep_names.append(self._names[i])
        return NssGrid(np.squeeze(self._data[item]), keep_axes, keep_names)

    # ---- I/O (grid.py:314-359)
    @classmethod
    def read(cls, filename, path: str = "/", format: str = None, **_) -> "NssGrid":
        if format is None:
            format = "fits" if str(filename).lower().endswith((".fits", ".fit")) else "hdf5"
        if format == "fits":
            return cls._read_fits(filename)
        if format != "hdf5":
            raise NotImplementedError(f"unknown NssGrid format {format!r}")
        f = hdf5_min.H5File(str(filename))
        g = f[path] if path not in ("/", "", None) else f
        data = np.asarray(g["__nss_grid_data__"][()])
        names = [g.attrs[f"AXIS{i}"] for i in range(data.ndim)]
        names = [n.decode() if isinstance(n, bytes) else str(n) for n in names]
        axes = [np.asarray(g["__nss_grid_axes__"][n][()]) for n in names]
        return cls(data, axes, names)

    def write(self, filename, format: str = None, overwrite: bool = True, **_) -> None:
        import os

        if format is None:
            format = "fits" if str(filename).lower().endswith((".fits", ".fit")) else "hdf5"
        if os.path.exists(filename) and not overwrite:
            raise OSError(f"{filename} exists")
        if format == "fits":
            return self._write_fits(filename)
        if format != "hdf5":
            raise NotImplementedError(f"unknown NssGrid format {format!r}")
        hdf5_min.write_hdf5(
            str(filename),
            {
                "__nss_grid_data__": self._data,
                "__nss_grid_axes__": {
                    "__kind__": "group",
                    "members": {n: a for n, a in zip(self._names, self._axes)},
                    "attrs": {},
                },
            },
            dict(self.meta),
        )

    # ---- FITS layout (grid.py:286-311): primary image + one BinTable HDU per axis
    @classmethod
    def _read_fits(cls, filename) -> "NssGrid":
        from . import fits_min

        hdus = fits_min.read_hdus(filename)
        hdr, data = hdus[0]
        naxis = hdr["NAXIS"]
        names = [str(hdr[f"AXIS{i}"]) for i in range(naxis)]
        axes = [next(iter(hdus[i + 1][1].values())) for i in range(naxis)]
        return cls(data, axes, names)

    def _write_fits(self, filename) -> None:
        from . import fits_min

        with open(filename, "wb") as f:
            f.write(fits_min.image_hdu(self._data, dict(self.meta), primary=True))
            for axis, name in zip(self._axes, self._names):
                f.write(fits_min.bintable_hdu({name: np.asarray(axis)},
                                              {"AXIS": name, "EXTNAME": name}))
