"""ResultsTable: the table ``compute()`` returns, and its FITS BinTable writer.

Reference: src/nuspacesim/results_table.py:54-78 (an astropy ``Table`` whose meta holds
``simTime`` and the flattened ``HIERARCH Config ...`` keys) and
src/nuspacesim/apps/run.py:209-210 (``simulation.write(output, format="fits")``).
astropy is not a dependency here: this is a small column container with the same
surface (``add_columns``, ``colnames``, ``meta``, ``__getitem__``, ``write``) and a
from-scratch FITS writer (2880-byte blocks, big-endian ``D`` columns, HIERARCH cards).
"""
from __future__ import annotations

import datetime
from collections.abc import MutableMapping
from typing import Any, Dict, Iterable, List

import numpy as np

__all__ = ["ResultsTable", "init", "output_filename", "flatten_dict", "read_fits_table"]


def flatten_dict(d: MutableMapping, parent_key: str = "", sep: str = ".") -> dict:
    """src/nuspacesim/utils/misc.py:14-23."""
    out = {}
    for k, v in d.items():
        key = parent_key + sep + k if parent_key else k
        if isinstance(v, MutableMapping):
            out.update(flatten_dict(v, key, sep=sep))
        else:
            out[key] = v
    return out


class ResultsTable:
    def __init__(self, meta: Dict[str, Any] | None = None):
        self.meta: Dict[str, Any] = dict(meta or {})
        self._cols: Dict[str, np.ndarray] = {}

    # -- astropy.table.Table surface used by compute() / StagedWriter
    def add_columns(self, cols: Iterable, names: Iterable[str], *args, **kwargs) -> None:
        cols, names = list(cols), list(names)
        if len(cols) != len(names):
            raise ValueError("number of columns and names differ")
        for n, c in zip(names, cols):
            c = np.asarray(c)
            if self._cols and len(c) != len(self):
                raise ValueError(f"Inconsistent data column lengths: {len(c)} vs {len(self)}")
            if n in self._cols:
                raise ValueError(f"Duplicate column names: {n}")
            self._cols[n] = c

    def add_column(self, col, name: str) -> None:
        self.add_columns([col], [name])

    @property
    def colnames(self) -> List[str]:
        return list(self._cols)

    def __len__(self) -> int:
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __getitem__(self, name: str) -> np.ndarray:
        return self._cols[name]

    def __contains__(self, name: str) -> bool:
        return name in self._cols

    def __repr__(self) -> str:
        return f"<ResultsTable length={len(self)} columns={self.colnames}>"

    # -- FITS
    def write(self, filename, format: str = "fits", overwrite: bool = False) -> None:
        if format != "fits":
            raise NotImplementedError("only format='fits' is implemented")
        import os

        if os.path.exists(filename) and not overwrite:
            raise OSError(f"File {filename} already exists.")
        with open(filename, "wb") as f:
            f.write(_primary_hdu())
            f.write(_bintable_hdu(self._cols, self.meta))


def init(config=None) -> ResultsTable:
    """Initialize a simulation results table with metadata (results_table.py:54-71)."""
    if config is None:
        from .config import NssConfig

        config = NssConfig()
    if isinstance(config, ResultsTable):
        t = ResultsTable(config.meta)
        t.add_columns([config[n] for n in config.colnames], config.colnames)
        return t
    if hasattr(config, "model_dump"):
        now = f"{datetime.datetime.now():%Y%m%d%H%M%S}"
        return ResultsTable(
            meta={
                "simTime": (now, "Start time of Simulation"),
                **flatten_dict(config.model_dump(), "HIERARCH Config", sep=" "),
            }
        )
    return ResultsTable()


def output_filename(filename=None, now=None) -> str:
    if filename is not None:
        return filename
    now = now if now else f"{datetime.datetime.now():%Y%m%d%H%M%S}"
    return f"nuspacesim_run_{now}.fits"


# ------------------------------------------------------------------ minimal FITS


def _fmt_value(v, room: int = 68) -> str:
    """FITS value field; a string is cut so that it still closes its quote within `room`."""
    if isinstance(v, (bool, np.bool_)):
        return "T" if v else "F"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)) and np.isfinite(v):
        s = repr(float(v)).upper()
        return s if ("." in s or "E" in s) else s + ".0"
    s = "".join(ch if " " <= ch <= "~" else "?" for ch in str(v)).replace("'", "''")
    s = s[:max(room - 2, 0)]
    if s.endswith("'") and (len(s) - len(s.rstrip("'"))) % 2:  # do not split an escaped quote
        s = s[:-1]
    return "'" + f"{s:<8}" + "'"


def _card(key: str, value=None, comment: str = "") -> bytes:
    """One 80-character header card (FITS 4.0 sec. 4.1-4.2).  Keywords of up to 8 characters
    are written upper-case, like astropy writes the reference's ``simTime`` meta key as
    ``SIMTIME`` (results_table.py:60-66); longer or blank-containing keys use the ESO
    HIERARCH convention astropy uses for ``HIERARCH Config ...``."""
    if value is None:
        text = key
    elif key.startswith("HIERARCH ") or len(key) > 8 or " " in key:
        k = key if key.startswith("HIERARCH ") else "HIERARCH " + key
        head = f"{k} = "
        text = head + _fmt_value(value, 80 - len(head))
        if comment and len(text) + 3 + len(comment) <= 80:
            text += " / " + comment
    else:
        val = _fmt_value(value)
        val = f"{val:<20}" if val.startswith("'") else f"{val:>20}"
        text = f"{key.upper():<8}= {val}"
        if comment and len(text) + 3 < 80:
            text += " / " + "".join(ch if " " <= ch <= "~" else "?" for ch in comment)
    return text[:80].ljust(80).encode("ascii", "replace")


def _pad(b: bytes, fill: bytes = b" ") -> bytes:
    r = (-len(b)) % 2880
    return b + fill * r


def _primary_hdu() -> bytes:
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", 8),
             _card("NAXIS", 0), _card("EXTEND", True), _card("END")]
    return _pad(b"".join(cards))


_TFORM = {"f8": ("D", ">f8", 8), "f4": ("E", ">f4", 4), "i8": ("K", ">i8", 8),
          "i4": ("J", ">i4", 4), "i2": ("I", ">i2", 2), "u1": ("B", "u1", 1), "b1": ("L", "S1", 1)}


def _bintable_hdu(cols: Dict[str, np.ndarray], meta: Dict[str, Any]) -> bytes:
    names = list(cols)
    nrows = len(cols[names[0]]) if names else 0
    specs = []
    for n in names:
        a = np.asarray(cols[n])
        code = a.dtype.str[1:] if a.dtype.kind != "b" else "b1"
        if code not in _TFORM:
            a = a.astype(np.float64)
            code = "f8"
        specs.append((n, a, *_TFORM[code]))
    rowbytes = sum(s[4] for s in specs)
    cards = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8),
             _card("NAXIS", 2), _card("NAXIS1", rowbytes), _card("NAXIS2", nrows),
             _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(specs))]
    for i, (n, a, tform, _, _) in enumerate(specs, 1):
        cards.append(_card(f"TTYPE{i}", n))
        cards.append(_card(f"TFORM{i}", tform))
    for k, v in meta.items():
        if isinstance(v, tuple) and len(v) == 2:
            cards.append(_card(k, v[0], str(v[1])))
        else:
            cards.append(_card(k, v))
    cards.append(_card("END"))
    header = _pad(b"".join(cards))
    if specs and nrows:
        rec = np.empty(nrows, dtype=[(f"c{i}", s[3]) for i, s in enumerate(specs)])
        for i, s in enumerate(specs):
            rec[f"c{i}"] = s[1] if s[2] != "L" else np.where(s[1], b"T", b"F")
        data = rec.tobytes()
    else:
        data = b""
    return header + _pad(data, b"\0")


def read_fits_table(filename):
    """Read back a file written by ``ResultsTable.write`` (round-trip tests)."""
    with open(filename, "rb") as f:
        raw = f.read()

    def parse_header(off):
        hdr = {}
        while True:
            block = raw[off:off + 2880]
            off += 2880
            done = False
            for i in range(0, 2880, 80):
                c = block[i:i + 80].decode("ascii")
                if c.startswith("END"):
                    done = True
                    break
                if "=" not in c:
                    continue
                k, rest = c.split("=", 1)
                k = k.strip()
                val = rest.strip()
                if val.startswith("'"):
                    end = val.find("'", 1)
                    while end + 1 < len(val) and val[end + 1] == "'":
                        end = val.find("'", end + 2)
                    v = val[1:end].replace("''", "'").rstrip()
                else:
                    tok = val.split("/")[0].strip()
                    v = True if tok == "T" else False if tok == "F" else (
                        int(tok) if tok.lstrip("+-").isdigit() else float(tok))
                hdr[k] = v
            if done:
                return hdr, off

    _, off = parse_header(0)
    hdr, off = parse_header(off)
    n, nf = hdr["NAXIS2"], hdr["TFIELDS"]
    rev = {v[0]: v[1] for v in _TFORM.values()}
    dt = [(hdr[f"TTYPE{i}"], rev[hdr[f"TFORM{i}"]]) for i in range(1, nf + 1)]
    rec = np.frombuffer(raw, dtype=dt, count=n, offset=off) if n else np.empty(0, dtype=dt)
    t = ResultsTable({k: v for k, v in hdr.items()
                      if not k.startswith(("TTYPE", "TFORM", "NAXIS", "XTENSION", "BITPIX",
                                           "PCOUNT", "GCOUNT", "TFIELDS"))})
    if nf:
        t.add_columns([rec[name].astype(rec[name].dtype.newbyteorder("=")) for name, _ in dt],
                      [name for name, _ in dt])
    return t
