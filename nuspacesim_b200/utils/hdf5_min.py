"""Minimal pure-Python HDF5 reader/writer for NssGrid table files.

The reference reads/writes its ``nu2tau_*.h5`` tables through h5py
(reference: src/nuspacesim/utils/grid.py:314-359).  Neither h5py nor astropy is
available to this build, so this module implements the subset of the HDF5 file
format those files use:

* superblock version 0/1, 8-byte offsets and lengths,
* "old style" groups (symbol-table message -> v1 B-tree -> SNOD -> local heap),
* version-1 object headers with continuation blocks,
* contiguous (or compact) little/big-endian IEEE float and integer datasets,
* attributes holding fixed-length or variable-length (global heap) strings and
  numeric scalars/arrays.

The writer emits the same layout, so files written here can be opened by stock
h5py / the reference's ``hdf5_nssgrid_reader`` and vice versa.
"""

from __future__ import annotations

import struct
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

__all__ = ["H5File", "H5Group", "H5Dataset", "read_hdf5", "write_hdf5", "is_hdf5"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


def is_hdf5(filename) -> bool:
    try:
        with open(filename, "rb") as f:
            return f.read(8) == _SIG
    except OSError:
        return False


# --------------------------------------------------------------------------- reader


class H5Dataset:
    def __init__(self, data: np.ndarray, attrs: Dict[str, Any]):
        self.data = data
        self.attrs = attrs

    def __getitem__(self, item):
        if item == ():
            return self.data
        return self.data[item]

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim


class H5Group:
    def __init__(self, members: Dict[str, Any], attrs: Dict[str, Any]):
        self.members = members
        self.attrs = attrs

    def __getitem__(self, name):
        if isinstance(name, bytes):
            name = name.decode()
        if name in ("/", ""):
            return self
        node = self
        for part in name.strip("/").split("/"):
            node = node.members[part]
        return node

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def keys(self):
        return self.members.keys()


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file (bad signature)")
        ver = buf[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {ver} not supported")
        self.O = buf[13]
        self.L = buf[14]
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("only 8-byte offsets/lengths supported")
        p = 24 if ver == 0 else 28
        self.base, _fs, self.eof, _drv = struct.unpack_from("<4Q", buf, p)
        p += 32
        # root group symbol table entry
        _lno, self.root_ohdr, ctype = struct.unpack_from("<QQI", buf, p)

    # -- object headers
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.b
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver} not supported")
        blocks = [(addr + 16, hsize)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8 : p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((coff, clen))
                out.append((mtype, body))
        return out

    # -- datatype
    def _dtype(self, body: bytes):
        cv = body[0]
        cls = cv & 0x0F
        bits0 = body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        if cls == 0:  # fixed point
            order = ">" if (bits0 & 1) else "<"
            signed = (bits0 >> 3) & 1
            return ("num", np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size)
        if cls == 1:  # float
            order = ">" if (bits0 & 1) else "<"
            return ("num", np.dtype(f"{order}f{size}"), size)
        if cls == 3:  # fixed string
            return ("str", None, size)
        if cls == 9:  # variable length
            vtype = bits0 & 0x0F
            return ("vlenstr" if vtype == 1 else "vlen", None, size)
        raise NotImplementedError(f"HDF5 datatype class {cls} not supported")

    def _dataspace(self, body: bytes) -> Tuple[int, ...]:
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", body, p)) if rank else ()

    def _gheap_object(self, coll_addr: int, index: int) -> bytes:
        b = self.b
        if b[coll_addr : coll_addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        csize = struct.unpack_from("<Q", b, coll_addr + 8)[0]
        p = coll_addr + 16
        end = coll_addr + csize
        while p + 16 <= end:
            idx, _ref, _r, osize = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16 : p + 16 + osize]
            p += 16 + ((osize + 7) & ~7)
        raise KeyError(f"global heap object {index} not found")

    def _decode_values(self, kind, dt, esize, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if kind == "num":
            arr = np.frombuffer(raw, dtype=dt, count=n).astype(dt.newbyteorder("="))
            return arr.reshape(shape) if shape else arr.reshape(())[()]
        if kind == "str":
            vals = [
                raw[i * esize : (i + 1) * esize].split(b"\x00")[0].decode("utf-8")
                for i in range(n)
            ]
        elif kind == "vlenstr":
            vals = []
            for i in range(n):
                ln, caddr, idx = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._gheap_object(caddr, idx)[:ln].decode("utf-8"))
        else:
            raise NotImplementedError(kind)
        if not shape:
            return vals[0]
        return np.array(vals, dtype=object).reshape(shape)

    def _attribute(self, body: bytes) -> Tuple[str, Any]:
        ver = body[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8
            pad = lambda x: (x + 7) & ~7  # noqa: E731
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8 if ver == 2 else 9
            pad = lambda x: x  # noqa: E731
        else:
            raise NotImplementedError(f"attribute version {ver}")
        name = body[p : p + nsz].split(b"\x00")[0].decode("utf-8")
        p += pad(nsz)
        kind, dt, esize = self._dtype(body[p : p + dsz])
        p += pad(dsz)
        shape = self._dataspace(body[p : p + ssz])
        p += pad(ssz)
        return name, self._decode_values(kind, dt, esize, shape, body[p:])

    # -- groups
    def _local_heap_data(self, addr: int) -> int:
        if self.b[addr : addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        return struct.unpack_from("<Q", self.b, addr + 24)[0]

    def _btree_entries(self, addr: int, heap_data: int, out: Dict[str, int]):
        b = self.b
        if b[addr : addr + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise NotImplementedError("only group B-trees supported")
        p = addr + 24
        for _ in range(used):
            p += 8  # key
            child = struct.unpack_from("<Q", b, p)[0]
            p += 8
            if level > 0:
                self._btree_entries(child, heap_data, out)
            else:
                if b[child : child + 4] != b"SNOD":
                    raise ValueError("bad symbol table node")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                q = child + 8
                for _ in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, q)
                    s = heap_data + noff
                    e = b.index(b"\x00", s)
                    out[b[s:e].decode("utf-8")] = ohdr
                    q += 40

    def read_object(self, addr: int):
        msgs = self._messages(addr)
        attrs: Dict[str, Any] = {}
        shape = None
        dtype = None
        layout = None
        symtab = None
        for mtype, body in msgs:
            if mtype == 0x0C:
                k, v = self._attribute(body)
                attrs[k] = v
            elif mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x11:
                symtab = struct.unpack_from("<QQ", body, 0)
        if symtab is not None:
            entries: Dict[str, int] = {}
            self._btree_entries(symtab[0], self._local_heap_data(symtab[1]), entries)
            return H5Group({k: self.read_object(a) for k, a in entries.items()}, attrs)
        if dtype is None or shape is None or layout is None:
            raise ValueError("object is neither a group nor a simple dataset")
        kind, dt, esize = dtype
        n = int(np.prod(shape)) if shape else 1
        lver = layout[0]
        if lver == 3:
            lclass = layout[1]
            if lclass == 1:
                daddr, dsize = struct.unpack_from("<QQ", layout, 2)
                raw = b"" if daddr == _UNDEF else self.b[daddr : daddr + dsize]
            elif lclass == 0:
                csz = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4 : 4 + csz]
            else:
                raise NotImplementedError("chunked datasets not supported")
        elif lver in (1, 2):
            rank, lclass = layout[1], layout[2]
            if lclass != 1:
                raise NotImplementedError("only contiguous v1/v2 layouts supported")
            daddr = struct.unpack_from("<Q", layout, 8)[0]
            raw = self.b[daddr : daddr + n * esize]
        else:
            raise NotImplementedError(f"layout version {lver}")
        if len(raw) < n * esize:
            raw = raw + b"\x00" * (n * esize - len(raw))
        return H5Dataset(np.asarray(self._decode_values(kind, dt, esize, shape, raw)), attrs)


class H5File(H5Group):
    """Read-only view of an HDF5 file (h5py-like ``f[path]``, ``.attrs``)."""

    def __init__(self, filename, mode: str = "r"):
        if mode != "r":
            raise ValueError("H5File is read-only; use write_hdf5 to write")
        with open(filename, "rb") as f:
            buf = f.read()
        rd = _Reader(buf)
        root = rd.read_object(rd.root_ohdr)
        if not isinstance(root, H5Group):
            raise ValueError("root object is not a group")
        super().__init__(root.members, root.attrs)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def read_hdf5(filename) -> H5File:
    return H5File(filename)


# --------------------------------------------------------------------------- writer


class _Writer:
    """Serialises a nested dict tree into an HDF5 (superblock v0) byte string."""

    def __init__(self):
        self.buf = bytearray()
        self.gheap_objs: List[bytes] = []

    def _align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def _alloc(self, data: bytes) -> int:
        self._align()
        addr = len(self.buf)
        self.buf += data
        return addr

    # datatype / dataspace message bodies
    @staticmethod
    def _dt_body(dt: np.dtype) -> bytes:
        dt = np.dtype(dt)
        size = dt.itemsize
        if dt.kind == "f":
            if size == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                bits = (0x20, 0x3F, 0x00)
            elif size == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                bits = (0x20, 0x1F, 0x00)
            else:
                raise NotImplementedError("float size")
            return struct.pack("<BBBBI", 0x11, *bits, size) + props
        if dt.kind in "iu":
            bits0 = 0x08 if dt.kind == "i" else 0x00
            return struct.pack("<BBBBI", 0x10, bits0, 0, 0, size) + struct.pack(
                "<HH", 0, size * 8
            )
        raise NotImplementedError(f"dtype {dt}")

    @staticmethod
    def _vlenstr_dt_body() -> bytes:
        # class 9 (vlen), type=string(1), padding null-terminate(0), charset UTF-8(1)
        base = struct.pack("<BBBBI", 0x10 | 0x03, 0x00, 0x00, 0x00, 1)  # 1-byte string base
        return struct.pack("<BBBBI", 0x10 | 0x09, 0x01, 0x01, 0x00, 16) + base

    @staticmethod
    def _space_body(shape: Tuple[int, ...]) -> bytes:
        rank = len(shape)
        return struct.pack("<BBBB4x", 1, rank, 0, 0) + struct.pack(f"<{rank}Q", *shape)

    @staticmethod
    def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
        pad = (-len(body)) % 8
        body = body + b"\x00" * pad
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _attr_msg(self, name: str, value) -> bytes:
        nb = name.encode("utf-8") + b"\x00"
        pad8 = lambda x: x + b"\x00" * ((-len(x)) % 8)  # noqa: E731
        if isinstance(value, str):
            dtb = self._vlenstr_dt_body()
            spb = self._space_body(())
            vb = value.encode("utf-8")
            idx = self.gheap_index[vb]
            data = struct.pack("<IQI", len(vb), self.gheap_addr, idx)
        else:
            arr = np.asarray(value)
            if arr.dtype.kind == "f":
                arr = arr.astype("<f8")
            elif arr.dtype.kind in "iub":
                arr = arr.astype("<i8")
            else:
                raise NotImplementedError(f"attribute type {arr.dtype}")
            dtb = self._dt_body(arr.dtype)
            spb = self._space_body(arr.shape)
            data = arr.tobytes()
        body = (
            struct.pack("<BxHHH", 1, len(nb), len(dtb), len(spb))
            + pad8(nb)
            + pad8(dtb)
            + pad8(spb)
            + data
        )
        return self._msg(0x0C, body)

    def _object_header(self, msgs: List[bytes]) -> int:
        payload = b"".join(msgs)
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(payload))
        return self._alloc(hdr + payload)

    def _write_dataset(self, arr: np.ndarray, attrs: Dict[str, Any]) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == "f":
            arr = arr.astype(f"<f{arr.dtype.itemsize}")
        elif arr.dtype.kind in "iu":
            arr = arr.astype(f"<{arr.dtype.kind}{arr.dtype.itemsize}")
        daddr = self._alloc(arr.tobytes())
        msgs = [
            self._msg(0x01, self._space_body(arr.shape)),
            self._msg(0x03, self._dt_body(arr.dtype), flags=1),
            self._msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)),
        ]
        msgs += self._attr_msgs(attrs)
        return self._object_header(msgs)

    def _attr_msgs(self, attrs: Dict[str, Any]) -> List[bytes]:
        return [self._attr_msg(k, v) for k, v in (attrs or {}).items()]

    def _write_group(self, members: Dict[str, Any], attrs: Dict[str, Any]) -> int:
        # children first
        child_addrs = []
        for name in sorted(members.keys()):
            node = members[name]
            if isinstance(node, dict) and node.get("__kind__") == "group":
                a = self._write_group(node["members"], node.get("attrs", {}))
            else:
                if isinstance(node, dict):
                    a = self._write_dataset(np.asarray(node["data"]), node.get("attrs", {}))
                else:
                    a = self._write_dataset(np.asarray(node), {})
            child_addrs.append((name, a))
        # local heap: offset 0 is the empty string
        heap = bytearray(b"\x00" * 8)
        offs = []
        for name, _ in child_addrs:
            offs.append(len(heap))
            nb = name.encode("utf-8") + b"\x00"
            heap += nb + b"\x00" * ((-len(nb)) % 8)
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)  # free block: next=1 (none), size
        heap_data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(
            b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data_addr)
        )
        # one SNOD leaf (capacity 2K = 8 entries by default group leaf K=4);
        # split into several leaves when there are more members.
        leaf_cap = 8
        leaves = [child_addrs[i : i + leaf_cap] for i in range(0, max(len(child_addrs), 1), leaf_cap)]
        leaf_offs = [offs[i : i + leaf_cap] for i in range(0, max(len(offs), 1), leaf_cap)]
        snod_addrs = []
        for leaf, lo in zip(leaves, leaf_offs):
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf))
            for (name, a), o in zip(leaf, lo):
                body += struct.pack("<QQII16x", o, a, 0, 0)
            body += b"\x00" * (40 * (leaf_cap - len(leaf)))
            snod_addrs.append(self._alloc(body))
        # B-tree root (level 0) with keys = heap offsets
        k_internal = 16
        if len(snod_addrs) > 2 * k_internal:
            raise NotImplementedError("too many group members")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snod_addrs), _UNDEF, _UNDEF)
        tree += struct.pack("<Q", 0)
        for lo, sa in zip(leaf_offs, snod_addrs):
            tree += struct.pack("<QQ", sa, lo[-1] if lo else 0)
        tree += b"\x00" * (16 * (2 * k_internal - len(snod_addrs)))
        tree_addr = self._alloc(tree)
        msgs = [self._msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))]
        msgs += self._attr_msgs(attrs)
        ohdr = self._object_header(msgs)
        self._group_cache[ohdr] = (tree_addr, heap_addr)
        return ohdr

    def _collect_strings(self, members, attrs):
        for v in (attrs or {}).values():
            if isinstance(v, str):
                vb = v.encode("utf-8")
                if vb not in self.gheap_index:
                    self.gheap_objs.append(vb)
                    self.gheap_index[vb] = len(self.gheap_objs)
        for node in (members or {}).values():
            if isinstance(node, dict):
                self._collect_strings(node.get("members"), node.get("attrs"))

    def serialise(self, members: Dict[str, Any], attrs: Dict[str, Any]) -> bytes:
        self._group_cache: Dict[int, Tuple[int, int]] = {}
        self.gheap_index: Dict[bytes, int] = {}
        self.gheap_addr = 0
        # superblock placeholder (56 bytes header + 40 byte root entry = 96)
        self.buf += b"\x00" * 96
        # global heap collection for variable-length strings goes first so that
        # its address is known when the attribute messages are emitted
        self._collect_strings(members, attrs)
        if self.gheap_objs:
            body = bytearray()
            for i, obj in enumerate(self.gheap_objs, start=1):
                body += struct.pack("<HHIQ", i, 1, 0, len(obj)) + obj
                body += b"\x00" * ((-len(obj)) % 8)
            total = max(16 + len(body) + 16, 4096)
            free = total - 16 - len(body)
            body += struct.pack("<HHIQ", 0, 0, 0, free)
            body += b"\x00" * (free - 16)
            self.gheap_addr = self._alloc(
                b"GCOL" + struct.pack("<B3xQ", 1, total) + bytes(body)
            )
        root = self._write_group(members, attrs)
        self._align()
        eof = len(self.buf)
        tree_addr, heap_addr = self._group_cache[root]
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_hdf5(filename, members: Dict[str, Any], attrs: Optional[Dict[str, Any]] = None):
    """Write a tree to ``filename``.

    ``members`` maps names to ndarray (dataset) or
    ``{"__kind__": "group", "members": {...}, "attrs": {...}}``.
    """
    data = _Writer().serialise(members, attrs or {})
    with open(filename, "wb") as f:
        f.write(data)
