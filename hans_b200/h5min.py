"""Minimal HDF5 writer / reader for the restart file (no h5py, no libhdf5: neither is in this image).

Writes the subset of the HDF5 file format (format specification 1.1 / what HDF5 1.8 calls "the
earliest" layout) that `restart.hdf5` needs -- the schema of /root/reference/NesSa/NSio.py:69-111:
root attributes (every parameter + prev_iters + step sizes) and one group per walker holding two
contiguous float64 datasets.  Structures used:

    superblock version 0 (8-byte offsets and lengths, group leaf K = 4, internal K = 16)
    old-style groups: version 1 object header with a Symbol Table message (0x0011) -> version 1
        B-tree (node type 0, as many levels as the number of links needs) -> symbol table nodes
        (SNOD, <= 8 entries, sorted by name) + a local heap with the link names
    datasets: version 1 object header with Dataspace (0x0001, v1), Datatype (0x0003, v1),
        Fill Value (0x0005, v2, default value) and Data Layout (0x0008, v3, contiguous) messages
    attributes: Attribute messages (0x000C, v1) in the object header; int64, float64 and
        fixed-length ASCII / UTF-8 strings, scalar or one-dimensional

`read()` parses exactly that subset (it is the resume path, mpihans.py:77-94,166-181).  Files
written by the reference itself (h5py, variable-length strings, possibly newer object headers) are
read through h5py when it is importable (NSio.open_restart); this module does not try to.

There is no libhdf5 in this image to open the result with.  What pins the format instead: the same
reader parses the one libhdf5-written file on the system (scipy's MATLAB v7.3 test file: superblock
0, local heap, B-tree, symbol table node, version 1 object header, float64 datatype, dataspace and
attribute messages are byte-for-byte the structures written here; tests/test_host_cpu.py), and the
layout keeps to libhdf5's reader-side rules: B-tree and symbol table nodes are allocated at their
full size, group B-tree keys bound the names of child i as (key[i], key[i+1]], the local heap's
free-list head is 1 (H5HL_FREE_NULL) when the data segment is full, object header messages are
8-byte aligned and fill the header exactly, the end-of-file address equals the file size.  Where
h5py exists, tests/test_host_cpu.py::test_h5min_file_opens_with_h5py opens a written file with it.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16
SNOD_CAP, NODE_CAP = 2 * LEAF_K, 2 * INTERNAL_K
SNOD_SIZE = 8 + 40 * SNOD_CAP
NODE_SIZE = 24 + 8 * NODE_CAP + 8 * (NODE_CAP + 1)

MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT, MSG_ATTRIBUTE, MSG_STAB = 0x0001, 0x0003, 0x0005, 0x0008, 0x000C, 0x0011


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace encodings ---------------------------------------------------------------
def _dt_f64() -> bytes:
    # class 1 (floating point), version 1; little endian, implied mantissa msb, sign bit 63
    return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_i64() -> bytes:
    # class 0 (fixed point), version 1; little endian, two's complement
    return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)


def _dt_str(n: int, utf8: bool) -> bytes:
    # class 3 (string), version 1; null padded; character set in bits 4-7
    return struct.pack("<BBBBI", 0x13, 0x01 | (0x10 if utf8 else 0x00), 0x00, 0x00, n)


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _encode_value(v):
    """-> (datatype bytes, shape, raw data bytes) for an attribute value."""
    if isinstance(v, (bytes, str, np.str_, np.bytes_)):
        raw = v if isinstance(v, (bytes, np.bytes_)) else str(v).encode("utf-8")
        utf8 = any(c > 127 for c in raw)
        n = max(len(raw), 1)
        return _dt_str(n, utf8), (), raw.ljust(n, b"\0")
    a = np.asarray(v)
    if a.dtype == np.bool_ or np.issubdtype(a.dtype, np.integer):
        return _dt_i64(), tuple(a.shape), np.ascontiguousarray(a, dtype="<i8").tobytes()
    if np.issubdtype(a.dtype, np.floating):
        return _dt_f64(), tuple(a.shape), np.ascontiguousarray(a, dtype="<f8").tobytes()
    if a.dtype.kind in "US":  # array of strings -> one string, like str(list) would be useless; join
        return _encode_value(",".join(str(x) for x in a.ravel()))
    raise TypeError(f"h5min: cannot store attribute of type {a.dtype}")


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attribute_message(name: str, value) -> bytes:
    dt, shape, raw = _encode_value(value)
    ds = _dataspace(shape)
    nm = name.encode("utf-8") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw
    if len(body) + 8 > 0xFFF8:
        raise ValueError(f"h5min: attribute {name} too large for an object header message")
    return _message(MSG_ATTRIBUTE, body)


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # the superblock is filled in at the end

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, array: np.ndarray) -> int:
        a = np.ascontiguousarray(array, dtype="<f8")
        raw = a.tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [
            _message(MSG_DATASPACE, _dataspace(a.shape), 1),
            _message(MSG_DATATYPE, _dt_f64(), 1),
            _message(MSG_FILL, struct.pack("<BBBBI", 2, 2, 2, 1, 0), 1),  # v2: late allocation, fill if set, default value
            _message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),
        ]
        return self.alloc(_object_header(msgs))

    def group(self, links: Dict[str, Tuple[int, int, int]], attr_messages=()) -> Tuple[int, int, int]:
        """links: name -> (object header address, B-tree address or 0, heap address or 0) [the latter
        two for sub-groups: cached in the symbol table entry].  Returns (header, btree, heap)."""
        names = sorted(links, key=lambda s: s.encode("utf-8"))
        # local heap: "" at offset 0, then the names
        seg = bytearray(8)
        off = {}
        for n in names:
            off[n] = len(seg)
            seg += _pad8(n.encode("utf-8") + b"\0")
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, seg_addr))
        # symbol table nodes
        level = []  # (address, largest name) per child of the level being built
        for i in range(0, len(names), SNOD_CAP):
            chunk = names[i:i + SNOD_CAP]
            ent = b""
            for n in chunk:
                hdr, bt, hp = links[n]
                if bt:
                    ent += struct.pack("<QQII", off[n], hdr, 1, 0) + struct.pack("<QQ", bt, hp)
                else:
                    ent += struct.pack("<QQII16x", off[n], hdr, 0, 0)
            node = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + ent
            level.append((self.alloc(node.ljust(SNOD_SIZE, b"\0")), chunk[-1]))
        if not level:  # an empty group still has a (childless) B-tree root
            level = []
        # B-tree levels, bottom up
        depth = 0
        while True:
            nodes = [level[i:i + NODE_CAP] for i in range(0, max(len(level), 1), NODE_CAP)]
            addrs = []
            base = len(self.buf) + (-len(self.buf) % 8)
            for k in range(len(nodes)):
                addrs.append(base + k * NODE_SIZE)  # NODE_SIZE is a multiple of 8
            nxt = []
            for k, children in enumerate(nodes):
                left = addrs[k - 1] if k > 0 else UNDEF
                right = addrs[k + 1] if k + 1 < len(nodes) else UNDEF
                first_key = off[nodes[k - 1][-1][1]] if k > 0 else 0  # names in child i lie in (key[i], key[i+1]]
                body = struct.pack("<Q", first_key)
                for addr, largest in children:
                    body += struct.pack("<QQ", addr, off[largest])
                node = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(children), left, right) + body
                a = self.alloc(node.ljust(NODE_SIZE, b"\0"))
                assert a == addrs[k]
                nxt.append((a, children[-1][1] if children else ""))
            if len(nxt) == 1:
                btree_addr = nxt[0][0]
                break
            level, depth = nxt, depth + 1
        hdr = self.alloc(_object_header([_message(MSG_STAB, struct.pack("<QQ", btree_addr, heap_addr))] + list(attr_messages)))
        return hdr, btree_addr, heap_addr

    def finish(self, root) -> bytes:
        hdr, bt, hp = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write(filename: str, attrs: dict, groups: Dict[str, Dict[str, np.ndarray]]) -> None:
    """Root attributes `attrs` and `groups[name][dataset] = float64 array` -> an HDF5 file."""
    w = _Writer()
    links = {}
    for g, datasets in groups.items():
        links[g] = w.group({d: (w.dataset(a), 0, 0) for d, a in datasets.items()})
    root = w.group(links, [_attribute_message(k, v) for k, v in attrs.items()])
    data = w.finish(root)
    with open(filename, "wb") as f:
        f.write(data)


# ---- reader (of the subset above) -----------------------------------------------------------------
class _Reader:
    def __init__(self, data: bytes):
        # the superblock may sit behind a user block of 512, 1024, ... bytes; addresses count from it
        base = 0
        while data[base:base + 8] != SIG:
            base = 512 if base == 0 else 2 * base
            if base >= len(data):
                raise ValueError("not an HDF5 file")
        data = data[base:]
        self.d = data
        ver, so, sl = data[8], data[13], data[14]
        if ver != 0 or so != 8 or sl != 8:
            raise ValueError("h5min reads superblock version 0 with 8-byte offsets only (use h5py for other files)")
        self.eof = struct.unpack_from("<Q", data, 40)[0]
        self.root_header = struct.unpack_from("<Q", data, 56 + 8)[0]

    def messages(self, addr: int):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.d, addr)
        if ver != 1:
            raise ValueError("h5min reads version 1 object headers only (use h5py for other files)")
        p, end, out = addr + 16, addr + 16 + size, []
        while p < end and len(out) < nmsg:
            mtype, msize, _flags = struct.unpack_from("<HHB", self.d, p)
            out.append((mtype, self.d[p + 8:p + 8 + msize]))
            p += 8 + msize
        return out

    def heap_name(self, heap_addr: int, off: int) -> str:
        if self.d[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        seg = struct.unpack_from("<Q", self.d, heap_addr + 24)[0]
        end = self.d.index(b"\0", seg + off)
        return self.d[seg + off:end].decode("utf-8")

    def links(self, header_addr: int) -> Dict[str, int]:
        stab = [m for t, m in self.messages(header_addr) if t == MSG_STAB]
        if not stab:
            raise ValueError("not an old-style group")
        btree, heap = struct.unpack_from("<QQ", stab[0], 0)
        out = {}

        def walk(addr):
            if self.d[addr:addr + 4] == b"SNOD":
                n = struct.unpack_from("<H", self.d, addr + 6)[0]
                for i in range(n):
                    off, hdr = struct.unpack_from("<QQ", self.d, addr + 8 + 40 * i)
                    out[self.heap_name(heap, off)] = hdr
                return
            if self.d[addr:addr + 4] != b"TREE":
                raise ValueError("bad B-tree node")
            ntype, _lvl, used = struct.unpack_from("<BBH", self.d, addr + 4)
            if ntype != 0:
                raise ValueError("not a group B-tree")
            for i in range(used):
                walk(struct.unpack_from("<Q", self.d, addr + 24 + 8 + 16 * i)[0])

        walk(btree)
        return out

    @staticmethod
    def _decode(dt: bytes, ds: bytes, raw: bytes):
        cls, bits0, size = dt[0] & 0x0F, dt[1], struct.unpack_from("<I", dt, 4)[0]
        rank = ds[1]
        shape = struct.unpack_from("<%dQ" % rank, ds, 8) if rank else ()
        count = int(np.prod(shape)) if rank else 1
        if cls == 3:
            s = raw[:size].rstrip(b"\0").decode("utf-8")
            return s
        dtype = {0: "<i8", 1: "<f8"}.get(cls)
        if dtype is None or size != 8:
            raise ValueError("h5min reads int64 / float64 / string values only")
        a = np.frombuffer(raw[:8 * count], dtype=dtype).reshape(shape).copy()
        return a.item() if rank == 0 else a

    def attributes(self, header_addr: int) -> dict:
        out = {}
        for t, m in self.messages(header_addr):
            if t != MSG_ATTRIBUTE:
                continue
            _ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", m, 0)
            p = 8
            name = m[p:p + nsz - 1].decode("utf-8"); p += nsz + (-nsz % 8)
            dt = m[p:p + dsz]; p += dsz + (-dsz % 8)
            ds = m[p:p + ssz]; p += ssz + (-ssz % 8)
            out[name] = self._decode(dt, ds, m[p:])
        return out

    def dataset(self, header_addr: int) -> np.ndarray:
        ms = dict(self.messages(header_addr))
        ds, dt, lay = ms[MSG_DATASPACE], ms[MSG_DATATYPE], ms[MSG_LAYOUT]
        rank = ds[1]
        shape = struct.unpack_from("<%dQ" % rank, ds, 8)
        if lay[0] == 3 and lay[1] == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
        elif lay[0] in (1, 2) and lay[2] == 1:  # the older message: rank + 1, class, address, 4-byte dimensions
            addr = struct.unpack_from("<Q", lay, 8)[0]
            size = 8 * int(np.prod(shape))
        else:
            raise ValueError("h5min reads contiguous layouts only")
        if (dt[0] & 0x0F) != 1 or struct.unpack_from("<I", dt, 4)[0] != 8:
            raise ValueError("h5min reads float64 datasets only")
        raw = self.d[addr:addr + size] if size else b""
        return np.frombuffer(raw, dtype="<f8").reshape(shape).copy()


def read(filename: str):
    """-> (root attributes, {group: {dataset: array}}) of a file written by write()."""
    with open(filename, "rb") as f:
        r = _Reader(f.read())
    attrs = r.attributes(r.root_header)
    groups = {}
    for g, hdr in r.links(r.root_header).items():
        groups[g] = {d: r.dataset(a) for d, a in r.links(hdr).items()}
    return attrs, groups


def is_hdf5(filename: str) -> bool:
    try:
        with open(filename, "rb") as f:
            return f.read(8) == SIG
    except OSError:
        return False
