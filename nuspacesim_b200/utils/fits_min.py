"""Minimal FITS reader/writer for the two layouts the hot path touches.

astropy is not a dependency.  Only what the reference writes is covered:
  * the NssGrid FITS layout (src/nuspacesim/utils/grid.py:286-311): primary HDU = the N-D
    grid (float64 image) with ``AXISi`` name cards, followed by one single-column BinTable HDU
    per axis;
  * the results table (one BinTable HDU; see ``results_table.py``).
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

from ..results_table import _TFORM, _bintable_hdu, _card, _pad

_BITPIX = {"f8": -64, "f4": -32, "i8": 64, "i4": 32, "i2": 16, "u1": 8}
_BITPIX_REV = {-64: ">f8", -32: ">f4", 64: ">i8", 32: ">i4", 16: ">i2", 8: "u1"}


def image_hdu(data: np.ndarray, meta: Dict[str, Any], primary: bool = True) -> bytes:
    a = np.asarray(data)
    code = a.dtype.str[1:]
    if code not in _BITPIX:
        a = a.astype(np.float64)
        code = "f8"
    cards = [_card("SIMPLE", True, "conforms to FITS standard") if primary
             else _card("XTENSION", "IMAGE", "image extension"),
             _card("BITPIX", _BITPIX[code]), _card("NAXIS", a.ndim)]
    # FITS counts the fastest-varying axis first: NAXIS1 is the LAST NumPy dimension
    for i, n in enumerate(reversed(a.shape), 1):
        cards.append(_card(f"NAXIS{i}", int(n)))
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0), _card("GCOUNT", 1)]
    for k, v in meta.items():
        cards.append(_card(k, v[0], str(v[1])) if isinstance(v, tuple) and len(v) == 2
                     else _card(k, v))
    cards.append(_card("END"))
    be = a.astype(a.dtype.newbyteorder(">")) if a.dtype.itemsize > 1 else a
    return _pad(b"".join(cards)) + _pad(np.ascontiguousarray(be).tobytes(), b"\0")


def _parse_header(raw: bytes, off: int) -> Tuple[Dict[str, Any], int]:
    hdr: Dict[str, Any] = {}
    while True:
        block = raw[off:off + 2880]
        if len(block) < 2880:
            raise ValueError("truncated FITS header")
        off += 2880
        for i in range(0, 2880, 80):
            c = block[i:i + 80].decode("ascii")
            if c.startswith("END"):
                return hdr, off
            if c.startswith("HIERARCH "):
                k, rest = c[9:].split("=", 1)
            elif c[8:10] == "= ":
                k, rest = c[:8], c[10:]
            else:
                continue
            k, val = k.strip(), rest.strip()
            if val.startswith("'"):
                end = val.find("'", 1)
                while end + 1 < len(val) and val[end + 1] == "'":
                    end = val.find("'", end + 2)
                v: Any = val[1:end].replace("''", "'").rstrip()
            else:
                tok = val.split("/")[0].strip()
                v = True if tok == "T" else False if tok == "F" else (
                    int(tok) if tok.lstrip("+-").isdigit() else float(tok))
            hdr[k] = v


def read_hdus(filename) -> List[Tuple[Dict[str, Any], Any]]:
    """[(header, data)] for every HDU: images as ndarrays, BinTables as {name: column}."""
    with open(filename, "rb") as f:
        raw = f.read()
    out, off = [], 0
    rev = {v[0]: v[1] for v in _TFORM.values()}
    while off < len(raw):
        hdr, off = _parse_header(raw, off)
        naxis = hdr.get("NAXIS", 0)
        if hdr.get("XTENSION", "IMAGE").strip() == "BINTABLE":
            n, nf = hdr["NAXIS2"], hdr["TFIELDS"]
            dt = [(hdr[f"TTYPE{i}"], rev[str(hdr[f"TFORM{i}"]).lstrip("1")])
                  for i in range(1, nf + 1)]
            rec = (np.frombuffer(raw, dtype=dt, count=n, offset=off) if n
                   else np.empty(0, dtype=dt))
            data: Any = {name: rec[name].astype(rec[name].dtype.newbyteorder("="))
                         for name, _ in dt}
            nbytes = hdr["NAXIS1"] * n
        elif naxis:
            shape = tuple(hdr[f"NAXIS{i}"] for i in range(naxis, 0, -1))
            dt = np.dtype(_BITPIX_REV[hdr["BITPIX"]])
            cnt = int(np.prod(shape))
            a = np.frombuffer(raw, dtype=dt, count=cnt, offset=off).reshape(shape)
            data = a.astype(dt.newbyteorder("="))
            nbytes = cnt * dt.itemsize
        else:
            data, nbytes = None, 0
        off += nbytes + ((-nbytes) % 2880)
        out.append((hdr, data))
    return out


def bintable_hdu(cols: Dict[str, np.ndarray], meta: Dict[str, Any]) -> bytes:
    return _bintable_hdu(cols, meta)
