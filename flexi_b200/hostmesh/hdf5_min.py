"""Minimal HDF5 parser (and, at the end of the file, writer) for the files HOPR and FLEXI write (io_hdf5/hdf5_input.f90 reads them through libhdf5; this image
has neither libhdf5 nor h5py).  Supported subset -- exactly what those writers produce with the library defaults: superblock
version 0 (optionally behind a FLEXI userblock), version-1 object headers with continuation blocks, old-style groups (symbol table:
B-tree v1 + local heap), contiguous or compact dataset layout, fixed-point / IEEE float / fixed-length string datatypes, attribute
messages v1-v3.  Anything else (chunked or filtered data, new-style groups, compound types) raises NotImplementedError.

Test and bench harness code (the stand-in for the Fortran host's mesh reader), not part of the product library."""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Dataset:
    def __init__(self, f, dtype, shape, layout, attrs):
        self._f, self.dtype, self.shape, self._layout, self.attrs = f, dtype, shape, layout, attrs

    def read(self):
        """Array in C order of the stored dims (the reverse of the Fortran dims the writer declared)."""
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        kind, a, b = self._layout
        if kind == "compact":
            raw = a
        else:
            if a == _UNDEF:
                return np.zeros(self.shape, dtype=self.dtype)
            raw = self._f.buf[self._f.base + a: self._f.base + a + n * self.dtype.itemsize]
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()


class File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        off = 0
        while self.buf[off:off + 8] != _SIG:  # the superblock sits at 0, 512, 1024, ... (behind a userblock)
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise ValueError(f"{path}: no HDF5 signature")
        b = self.buf
        if b[off + 8] != 0:
            raise NotImplementedError(f"superblock version {b[off + 8]}")
        if b[off + 13] != 8 or b[off + 14] != 8:
            raise NotImplementedError("offsets / lengths other than 8 bytes")
        base_stored = struct.unpack_from("<Q", b, off + 24)[0]
        self.base = off + base_stored if base_stored == 0 else base_stored  # addresses are relative to the superblock of a userblock file
        self.userblock = bytes(b[:off])
        root_hdr = struct.unpack_from("<Q", b, off + 56 + 8)[0]
        self.attrs, msgs = self._object(root_hdr)
        self.items = {}
        for t, d in msgs:
            if t == 0x11:
                btree, heap = struct.unpack_from("<QQ", d, 0)
                self._walk_group(btree, heap)

    # ---- groups -----------------------------------------------------------------------------------------------------------
    def _heap_data(self, heap):
        p = self.base + heap
        assert self.buf[p:p + 4] == b"HEAP"
        return self.base + struct.unpack_from("<Q", self.buf, p + 24)[0]

    def _walk_group(self, btree, heap):
        hd = self._heap_data(heap)
        p = self.base + btree
        assert self.buf[p:p + 4] == b"TREE" and self.buf[p + 4] == 0
        level, used = self.buf[p + 5], struct.unpack_from("<H", self.buf, p + 6)[0]
        q = p + 24
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, q + 8 + 16 * i)[0]
            if level > 0:
                self._walk_group(child, heap)
                continue
            s = self.base + child
            assert self.buf[s:s + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", self.buf, s + 6)[0]
            for k in range(nsym):
                e = s + 8 + 40 * k
                name_off, hdr = struct.unpack_from("<QQ", self.buf, e)
                end = self.buf.index(b"\0", hd + name_off)
                self.items[self.buf[hd + name_off:end].decode()] = hdr

    # ---- object headers -----------------------------------------------------------------------------------------------------
    def _messages(self, addr):
        p = self.base + addr
        if self.buf[p] != 1:
            raise NotImplementedError(f"object header version {self.buf[p]}")
        nmsg, _, size = struct.unpack_from("<HII", self.buf, p + 2)
        blocks = [(p + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                t, sz, _flags = struct.unpack_from("<HHB", self.buf, q)
                d = self.buf[q + 8:q + 8 + sz]
                if t == 0x10:  # continuation
                    o, ln = struct.unpack_from("<QQ", d, 0)
                    blocks.append((self.base + o, ln))
                out.append((t, d))
                q += 8 + sz
                left -= 8 + sz
        return out

    @staticmethod
    def _dtype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from("<I", d, 4)[0]
        if d[1] & 1 and cls in (0, 1):
            raise NotImplementedError("big-endian data")
        if cls == 0:
            return np.dtype(("<i" if d[1] & 8 else "<u") + str(size)), 8 + 4
        if cls == 1:
            return np.dtype("<f" + str(size)), 8 + 12
        if cls == 3:
            return np.dtype("S" + str(size)), 8
        raise NotImplementedError(f"datatype class {cls} (version {ver})")

    @staticmethod
    def _dspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            o = 8
        elif ver == 2:
            o = 4
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        dims = struct.unpack_from("<" + "Q" * rank, d, o) if rank else ()
        return tuple(int(x) for x in dims), o + 8 * rank * (2 if flags & 1 else 1)

    def _attr(self, d):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        o = 8 if ver < 3 else 9
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = d[o:o + nsz].split(b"\0")[0].decode()
        o += pad(nsz)
        dt, _ = self._dtype(d[o:o + tsz])
        o += pad(tsz)
        shape, _ = self._dspace(d[o:o + ssz])
        o += pad(ssz)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        val = np.frombuffer(d[o:o + n * dt.itemsize], dtype=dt, count=n).reshape(shape).copy()
        return name, val

    def _object(self, addr):
        msgs = self._messages(addr)
        attrs = {}
        for t, d in msgs:
            if t == 0x0C:
                k, v = self._attr(d)
                attrs[k] = v
        return attrs, msgs

    # ---- datasets -----------------------------------------------------------------------------------------------------------
    def __contains__(self, name):
        return name in self.items

    def keys(self):
        return list(self.items)

    def __getitem__(self, name):
        attrs, msgs = self._object(self.items[name])
        dt = shape = layout = None
        for t, d in msgs:
            if t == 0x01:
                shape, _ = self._dspace(d)
            elif t == 0x03:
                dt, _ = self._dtype(d)
            elif t == 0x0B:
                raise NotImplementedError(f"{name}: filtered (compressed) dataset")
            elif t == 0x08:
                ver = d[0]
                if ver != 3:
                    raise NotImplementedError(f"{name}: data layout message version {ver}")
                cls = d[1]
                if cls == 1:
                    a, ln = struct.unpack_from("<QQ", d, 2)
                    layout = ("contiguous", a, ln)
                elif cls == 0:
                    ln = struct.unpack_from("<H", d, 2)[0]
                    layout = ("compact", d[4:4 + ln], ln)
                else:
                    raise NotImplementedError(f"{name}: chunked dataset")
        if dt is None or shape is None or layout is None:
            raise KeyError(f"{name} is not a dataset")
        return Dataset(self, dt, shape, layout, attrs)


# ---- writer: the same subset, laid out the way libhdf5 1.8+ lays out FLEXI's files (io_hdf5/hdf5_output.f90:785-839 WriteState) ----
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 8:
        return bytes.fromhex("11203f000800000000004000340b0034ff030000")
    if dt.kind == "i" and dt.itemsize == 4:
        return bytes.fromhex("100800000400000000002000")
    if dt.kind == "S":
        return bytes.fromhex("13020000") + struct.pack("<I", dt.itemsize)  # Fortran strings: space padded
    raise NotImplementedError(f"datatype {dt}")


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dspace_msg(shape):
    shape = tuple(int(x) for x in shape) or (1,)
    dims = struct.pack("<" + "Q" * len(shape), *shape)
    return bytes([1, len(shape), 1, 0, 0, 0, 0, 0]) + dims + dims  # version 1, dims + maximum dims


def _msg(mtype, data):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _attr_msg(name, value):
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = value.astype("S")
    if value.dtype.kind == "i" and value.dtype.itemsize != 4:
        value = value.astype("<i4")
    nm = name.encode() + b"\0"
    dt, ds = _dtype_msg(value.dtype), _dspace_msg(value.shape or (1,))
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + value.tobytes()
    return _msg(0x0C, body)


def _ohdr(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def write_file(path, datasets, attrs=None, userblock=b""):
    """Write root-level datasets {name: array} (float64 / int32 / fixed strings, C order as given) and root attributes
    {name: array-like} as an HDF5 file of the subset this module reads: superblock 0, old-style root group (B-tree v1, local heap,
    one symbol-table node: at most 8 datasets), version-1 object headers, contiguous layout.  userblock: bytes stored in front
    of the superblock (FLEXI keeps its parameter file there), padded to 512 * 2^k."""
    attrs = attrs or {}
    names = sorted(datasets)
    if not 0 < len(names) <= 8:
        raise NotImplementedError("1..8 root datasets")
    ub = 0
    if userblock:
        ub = 512
        while ub < len(userblock):
            ub *= 2
    # local heap data: the empty name at offset 0, then the dataset names
    heap, name_off = bytearray(8), {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b"\0")
    root_msgs_fixed = [_attr_msg(k, v) for k, v in attrs.items()]
    a_root = 96
    root_len = 16 + 24 + sum(len(m) for m in root_msgs_fixed)
    a_tree = a_root + root_len
    a_heap = a_tree + 544
    a_heap_data = a_heap + 32
    a_snod = a_heap_data + len(heap)
    pos = a_snod + 328
    hdrs, blobs = {}, {}
    for n in names:
        arr = np.ascontiguousarray(datasets[n])
        if arr.dtype.kind == "i" and arr.dtype.itemsize != 4:
            arr = arr.astype("<i4")
        ds, dt = _dspace_msg(arr.shape), _dtype_msg(arr.dtype)
        fixed = 16 + (8 + len(_pad8(ds))) + (8 + len(_pad8(dt))) + (8 + 8) + (8 + 24)
        a_data = pos + fixed
        layout = bytes([3, 1]) + struct.pack("<QQ", a_data, arr.nbytes)
        hdrs[n] = (pos, _ohdr([_msg(0x01, ds), _msg(0x03, dt), _msg(0x05, bytes.fromhex("0202020100000000")), _msg(0x08, layout)]))
        assert len(hdrs[n][1]) == fixed
        blobs[n] = (a_data, arr.tobytes())
        pos = a_data + len(_pad8(blobs[n][1]))
    eof = pos
    out = bytearray(ub + eof)
    out[:len(userblock)] = userblock
    sb = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0) + struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_tree, a_heap)
    assert len(sb) == 96
    root = _ohdr([_msg(0x11, struct.pack("<QQ", a_tree, a_heap))] + root_msgs_fixed)
    assert len(root) == root_len
    tree = b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", 1, _UNDEF, _UNDEF) + struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])
    hp = b"HEAP" + bytes(4) + struct.pack("<QQQ", len(heap), 1, a_heap_data)  # free-list head 1 = no free block
    snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], hdrs[n][0], 0, 0)
    for a, b in [(0, sb), (a_root, root), (a_tree, tree), (a_heap, hp), (a_heap_data, bytes(heap)), (a_snod, snod)]:
        out[ub + a:ub + a + len(b)] = b
    for n in names:
        out[ub + hdrs[n][0]:ub + hdrs[n][0] + len(hdrs[n][1])] = hdrs[n][1]
        out[ub + blobs[n][0]:ub + blobs[n][0] + len(blobs[n][1])] = blobs[n][1]
    with open(path, "wb") as fh:
        fh.write(out)
