"""Independent structural checker for the FITS files the results table writes (tests only).

Shares no code with nuspacesim_b200/results_table.py or utils/fits_min.py: it re-derives the
layout from the FITS 4.0 standard -- 2880-byte logical records, 80-character ASCII cards,
keyword character set, fixed-format value fields, mandatory-keyword order of the primary HDU
and of a BINTABLE extension (sec. 4.4.1.1, 7.3.1), TFORMn codes, NAXIS1 = row width, big-endian
IEEE data, zero fill behind the data, blank fill behind END -- and decodes the table itself.
It stands in for astropy.io.fits (absent here), which reads the reference's result files
(src/nuspacesim/results_table.py:54-71, apps/run.py:209-210).
"""
import re
import struct

BLOCK = 2880
CARD = 80
KEYWORD_RE = re.compile(r"^[A-Z0-9_-]{1,8} *$")
TFORM_RE = re.compile(r"^(\d*)([LXBIJKAEDCMPQ])(.*)$")
WIDTH = {"L": 1, "B": 1, "I": 2, "J": 4, "K": 8, "A": 1, "E": 4, "D": 8, "C": 8, "M": 16}
STRUCT = {"B": "B", "I": "h", "J": "i", "K": "q", "E": "f", "D": "d"}


class FitsError(AssertionError):
    pass


def _need(cond, msg):
    if not cond:
        raise FitsError(msg)


def _parse_value(field, where):
    """Value field (columns 11-80 of a standard card, or the text right of '=' in a HIERARCH
    card) -> python value."""
    s = field.lstrip()
    if s.startswith("'"):
        i, out = 1, []
        while True:
            _need(i < len(s), f"{where}: string value does not close its quote")
            if s[i] == "'":
                if i + 1 < len(s) and s[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(s[i])
            i += 1
        rest = s[i + 1:].strip()
        _need(rest == "" or rest.startswith("/"), f"{where}: text after the string value")
        return "".join(out).rstrip()
    tok = s.split("/", 1)[0].strip()
    _need(tok != "", f"{where}: empty value field")
    if tok in ("T", "F"):
        return tok == "T"
    if re.fullmatch(r"[+-]?\d+", tok):
        return int(tok)
    _need(re.fullmatch(r"[+-]?(\d+\.?\d*|\.\d+)([ED][+-]?\d+)?", tok) is not None,
          f"{where}: value {tok!r} is neither logical, integer, real nor string")
    return float(tok.replace("D", "E"))


def _read_header(raw, off, what):
    """-> (ordered list of (keyword, value), offset of the first byte behind the header)."""
    cards, end = [], False
    while not end:
        _need(off + BLOCK <= len(raw), f"{what}: header runs past the end of the file")
        block = raw[off:off + BLOCK]
        off += BLOCK
        for i in range(0, BLOCK, CARD):
            c = block[i:i + CARD]
            _need(all(0x20 <= b <= 0x7E for b in c), f"{what}: non-printable byte in a card")
            c = c.decode("ascii")
            if end:
                _need(c.strip() == "", f"{what}: text behind END inside the last header block")
                continue
            if c.startswith("END") and c[3:].strip() == "":
                end = True
                continue
            where = f"{what}: card {c[:30]!r}"
            if c.startswith("HIERARCH "):
                _need("=" in c, f"{where}: HIERARCH card without '='")
                key, val = c.split("=", 1)
                toks = key.split()
                _need(len(toks) >= 2 and all(re.fullmatch(r"[A-Za-z0-9_.\-]+", t) for t in toks),
                      f"{where}: bad HIERARCH keyword tokens")
                cards.append((" ".join(toks), _parse_value(val, where)))
                continue
            kw = c[:8]
            if kw.strip() in ("", "COMMENT", "HISTORY"):
                continue
            _need(KEYWORD_RE.match(kw) is not None,
                  f"{where}: keyword must be upper-case A-Z 0-9 _ - and left-justified")
            _need(c[8:10] == "= ", f"{where}: value indicator '= ' expected in columns 9-10")
            v = _parse_value(c[10:], where)
            if not isinstance(v, str):  # fixed format: numbers/logicals right-justified to col 30
                fld = c[10:30]
                _need(fld.strip() == c[10:].split("/", 1)[0].strip() and not fld.endswith(" "),
                      f"{where}: numeric value is not right-justified in columns 11-30")
            else:
                _need(c[10] == "'", f"{where}: string value must start in column 11")
            cards.append((kw.strip(), v))
    return cards, off


def check_results_file(path):
    """Validate the file; -> (meta dict of the table HDU, {column: list of values})."""
    with open(path, "rb") as f:
        raw = f.read()
    _need(len(raw) % BLOCK == 0, "file size is not a multiple of 2880 bytes")
    # ---- primary HDU
    cards, off = _read_header(raw, 0, "primary")
    keys = [k for k, _ in cards]
    _need(keys[:3] == ["SIMPLE", "BITPIX", "NAXIS"], f"primary: mandatory order, got {keys[:3]}")
    d = dict(cards)
    _need(d["SIMPLE"] is True and d["BITPIX"] in (8, 16, 32, 64, -32, -64), "primary: values")
    _need(d["NAXIS"] == 0, "primary: a table file carries no primary data (NAXIS = 0)")
    _need(d.get("EXTEND") is True, "primary: EXTEND = T expected when extensions follow")
    # ---- binary table extension
    cards, off = _read_header(raw, off, "bintable")
    keys = [k for k, _ in cards]
    want = ["XTENSION", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "PCOUNT", "GCOUNT", "TFIELDS"]
    _need(keys[:8] == want, f"bintable: mandatory keyword order, got {keys[:8]}")
    _need(len(set(keys)) == len(keys), "bintable: duplicated keyword")
    d = dict(cards)
    _need(d["XTENSION"] == "BINTABLE" and d["BITPIX"] == 8 and d["NAXIS"] == 2, "bintable: values")
    _need(d["PCOUNT"] == 0 and d["GCOUNT"] == 1, "bintable: PCOUNT/GCOUNT")
    nf, nrows, rowbytes = d["TFIELDS"], d["NAXIS2"], d["NAXIS1"]
    _need(0 <= nf <= 999 and nrows >= 0, "bintable: TFIELDS/NAXIS2 range")
    cols, width = [], 0
    for i in range(1, nf + 1):
        _need(f"TFORM{i}" in d, f"bintable: TFORM{i} missing")
        _need(f"TTYPE{i}" in d and isinstance(d[f"TTYPE{i}"], str), f"bintable: TTYPE{i} missing")
        m = TFORM_RE.match(d[f"TFORM{i}"])
        _need(m is not None, f"bintable: TFORM{i} = {d[f'TFORM{i}']!r} is not a valid code")
        rep = int(m.group(1) or 1)
        code = m.group(2)
        _need(code in WIDTH and rep == 1, f"bintable: unsupported TFORM{i}")
        cols.append((d[f"TTYPE{i}"], code, width))
        width += WIDTH[code] * rep
    _need(len({c[0] for c in cols}) == len(cols), "bintable: duplicated column name")
    _need(width == rowbytes, f"bintable: NAXIS1 = {rowbytes} but the TFORMs add up to {width}")
    ndata = rowbytes * nrows
    padded = -(-ndata // BLOCK) * BLOCK
    _need(off + padded == len(raw), "bintable: data size + fill does not end the file")
    _need(all(b == 0 for b in raw[off + ndata:off + padded]), "bintable: fill behind the data must be 0")
    table = {}
    for name, code, start in cols:
        if code in STRUCT:
            fmt = ">" + STRUCT[code]
            table[name] = [struct.unpack_from(fmt, raw, off + r * rowbytes + start)[0]
                           for r in range(nrows)]
        else:
            table[name] = [raw[off + r * rowbytes + start:off + r * rowbytes + start + 1]
                           for r in range(nrows)]
    meta = {k: v for k, v in cards if k not in want and not re.match(r"T(TYPE|FORM)\d+$", k)}
    return meta, table
