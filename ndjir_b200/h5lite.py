"""A pure-Python reader / writer for the subset of HDF5 that nnabla's `.h5` parameter files use.

The reference stores and restores its parameters with `nn.save_parameters("...h5")` / `nn.load_parameters`
(python/train.py:101, python/render_image.py:43, python/extract_by_mc.py:300), i.e. through h5py: one float dataset per
parameter at the path given by its scope name (`geometric-network/affine-00/affine/W`), nested groups for the scopes,
and two scalar attributes per dataset (`need_grad`, a bool stored as the int8 enum {FALSE, TRUE}; `index`, the
registration order).  h5py is not part of this image, so this module implements the file format itself, following
the published "HDF5 File Format Specification Version 3.0":

  written   superblock version 0, version-1 object headers, old-style groups (symbol-table message, version-1 group
            B-tree, SNOD symbol nodes, local heap), contiguous little-endian datasets, version-1 dataspace / datatype
            / attribute messages, fill-value message version 2, layout message version 3 - what libhdf5 writes with
            its default ("earliest") format bounds, so h5py / nnabla read the files
  read      the same, plus what h5py adds on its side: superblock version 1, object-header continuation chunks, NIL
            messages, attribute messages version 2 / 3 (h5py names attributes in UTF-8, which selects version 3),
            dataspace version 2, compact datasets, layout versions 1 / 2, big-endian data, fixed-length strings.
            Chunked / filtered datasets, new-style groups (superblock 2 / 3, `OHDR` headers) and shared messages are
            NOT read: nnabla never writes them; the reader raises NotImplementedError naming the construct.

Status: the format is pinned on the specification, not on libhdf5 (absent from this image): tests/test_h5lite.py
checks the writer's bytes against the layout the specification prescribes and the round trip through the reader.
"""
import struct
from collections import OrderedDict

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4            # a symbol node holds up to 2 * LEAF_K entries
INTERNAL_K = 16       # a group B-tree node holds up to 2 * INTERNAL_K children
SNOD_SIZE = 8 + 2 * LEAF_K * 40
TREE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
HEAP_FREE_NULL = 1    # libhdf5's H5HL_FREE_NULL: "no free block" in a local heap

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0xC, 0x10, 0x11


class Dataset:
    """A dataset read from / to be written to a file: `data` (numpy array) and `attrs` (name -> numpy scalar / array)."""

    def __init__(self, data, attrs=None):
        self.data = np.asarray(data)
        self.attrs = OrderedDict(attrs or {})

    def __repr__(self):
        return f"Dataset(shape={self.data.shape}, dtype={self.data.dtype}, attrs={dict(self.attrs)})"


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------------------------
# datatype / dataspace messages
def _encode_datatype(dt):
    dt = np.dtype(dt)
    if dt == np.bool_:
        # h5py's mapping of numpy bool: enumeration over int8 with members FALSE = 0, TRUE = 1
        base = _encode_datatype(np.int8)
        names = _pad8(b"FALSE\0") + _pad8(b"TRUE\0")
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    order = 1 if dt.byteorder == ">" else 0
    if dt.kind == "f":
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        bits = dt.itemsize * 8
        head = struct.pack("<BBBBI", 0x11, 0x20 | order, bits - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, bits, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, (0x08 if dt.kind == "i" else 0) | order, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated, ASCII
    raise TypeError(f"h5lite cannot store dtype {dt}")


def _decode_datatype(buf, pos):
    """-> (numpy dtype, is_bool_enum, bytes consumed)"""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, pos)
    cls, ver = cv & 0x0F, cv >> 4
    order = ">" if b0 & 1 else "<"
    if cls == 0:
        return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}"), False, 8 + 4
    if cls == 1:
        return np.dtype(f"{order}f{size}"), False, 8 + 12
    if cls == 3:
        return np.dtype(f"S{size}"), False, 8
    if cls == 8:
        n = b0 | (b1 << 8)
        base, _, used = _decode_datatype(buf, pos + 8)
        p = pos + 8 + used
        names = []
        for _ in range(n):
            e = buf.index(b"\0", p)
            names.append(bytes(buf[p:e]))
            ln = e - p + 1
            p += (ln + 7) // 8 * 8 if ver < 3 else ln
        vals = np.frombuffer(bytes(buf[p:p + n * base.itemsize]), dtype=base)
        p += n * base.itemsize
        is_bool = sorted(zip(vals.tolist(), names)) == [(0, b"FALSE"), (1, b"TRUE")]
        return base, is_bool, p - pos
    raise NotImplementedError(f"h5lite: datatype class {cls} (only fixed-point, floating-point, string, enum are read)")


def _encode_dataspace(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _decode_dataspace(buf, pos):
    ver, rank, flags = struct.unpack_from("<BBB", buf, pos)
    if ver == 1:
        p = pos + 8
    elif ver == 2:
        if buf[pos + 3] == 2:          # null dataspace
            return None
        p = pos + 4
    else:
        raise NotImplementedError(f"h5lite: dataspace message version {ver}")
    return tuple(struct.unpack_from("<Q", buf, p + 8 * i)[0] for i in range(rank))


# ------------------------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)          # the superblock is written last, in place

    def alloc(self, data):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def message(mtype, data, flags=0):
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def object_header(self, messages):
        body = b"".join(messages)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def attribute(self, name, value):
        v = np.asarray(value)
        nm = name.encode("utf-8") + b"\0"
        dt, ds = _encode_datatype(v.dtype), _encode_dataspace(v.shape)
        raw = v.astype(np.uint8).tobytes() if v.dtype == np.bool_ else np.ascontiguousarray(v).tobytes()
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw
        return self.message(MSG_ATTRIBUTE, body)

    def dataset(self, d):
        a = np.ascontiguousarray(d.data)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        raw = a.astype(np.uint8).tobytes() if a.dtype == np.bool_ else a.tobytes()
        addr = self.alloc(raw) if raw else UNDEF
        msgs = [self.message(MSG_DATASPACE, _encode_dataspace(d.data.shape)),
                self.message(MSG_DATATYPE, _encode_datatype(a.dtype), flags=1),
                self.message(MSG_FILL, struct.pack("<BBBBi", 2, 2, 2, 1, 0)),
                self.message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, len(raw)))]
        msgs += [self.attribute(k, v) for k, v in d.attrs.items()]
        return self.object_header(msgs)

    def group(self, tree):
        """-> (object header address, B-tree address, local heap address)"""
        entries = []                                   # (name bytes, header address, cache type, scratch)
        for name, child in tree.items():
            nb = name.encode("utf-8")
            if not nb or b"/" in nb or b"\0" in nb:
                raise ValueError(f"invalid link name {name!r}")
            if isinstance(child, dict):
                hdr, bt, hp = self.group(child)
                entries.append((nb, hdr, 1, struct.pack("<QQ", bt, hp)))
            else:
                d = child if isinstance(child, Dataset) else Dataset(child)
                entries.append((nb, self.dataset(d), 0, b"\0" * 16))
        entries.sort(key=lambda e: e[0])               # libhdf5 orders the links of a group with strcmp
        heap = bytearray(8)                            # offset 0: the empty string, key 0 of the B-tree
        offs = []
        for nb, *_ in entries:
            offs.append(len(heap))
            heap += _pad8(nb + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), HEAP_FREE_NULL, heap_data))
        # symbol nodes, then B-tree levels bottom-up; a node is (address, first key, last key)
        nodes = []
        for i in range(0, len(entries), 2 * LEAF_K):
            part = entries[i:i + 2 * LEAF_K]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for j, (nb, hdr, cache, scratch) in enumerate(part):
                body += struct.pack("<QQI4x", offs[i + j], hdr, cache) + scratch
            addr = self.alloc(body.ljust(SNOD_SIZE, b"\0"))
            nodes.append((addr, offs[i - 1] if i else 0, offs[i + len(part) - 1]))
        level = 0
        while True:
            groups = [nodes[i:i + 2 * INTERNAL_K] for i in range(0, len(nodes), 2 * INTERNAL_K)] or [[]]
            base = len(self.buf) + (-len(self.buf) % 8)
            addrs = [base + k * TREE_SIZE for k in range(len(groups))]
            parents = []
            for k, ch in enumerate(groups):
                left = addrs[k - 1] if k else UNDEF
                right = addrs[k + 1] if k + 1 < len(groups) else UNDEF
                body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(ch), left, right)
                body += struct.pack("<Q", ch[0][1] if ch else 0)
                for a, _, last in ch:
                    body += struct.pack("<QQ", a, last)
                got = self.alloc(body.ljust(TREE_SIZE, b"\0"))
                assert got == addrs[k]
                parents.append((got, ch[0][1] if ch else 0, ch[-1][2] if ch else 0))
            if len(parents) == 1:
                bt = parents[0][0]
                break
            nodes, level = parents, level + 1
        hdr = self.object_header([self.message(MSG_SYMBOL_TABLE, struct.pack("<QQ", bt, heap_addr))])
        return hdr, bt, heap_addr

    def finish(self, tree):
        hdr, bt, hp = self.group(tree)
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, hdr, 1, bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def dumps(tree):
    """File image of a nested dict {name: dict | Dataset | array}."""
    return _Writer().finish(tree)


def write(path, tree):
    with open(path, "wb") as f:
        f.write(dumps(tree))


# ------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf):
        self.buf = buf
        start = 0
        while buf[start:start + 8] != SIGNATURE:       # a user block may precede the superblock: 512, 1024, ...
            start = 512 if start == 0 else start * 2
            if start + 8 > len(buf):
                raise ValueError("not an HDF5 file (signature not found)")
        ver = buf[start + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f"h5lite: superblock version {ver} (files written with libver='latest'; nnabla "
                                      f"writes version 0)")
        so, sl = buf[start + 13], buf[start + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError(f"h5lite: offset / length sizes {so} / {sl} (only 8 / 8)")
        p = start + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", buf, p)
        if self.base == UNDEF:
            self.base = 0
        self.root_header = struct.unpack_from("<Q", buf, p + 32 + 8)[0]

    def at(self, addr):
        return self.base + addr

    def messages(self, addr):
        """[(type, flags, offset of the message data, size)] of the version-1 object header at addr, all chunks."""
        p = self.at(addr)
        if self.buf[p:p + 4] == b"OHDR":
            raise NotImplementedError("h5lite: version-2 object header (file written with libver='latest')")
        ver, _, _, _, size = struct.unpack_from("<BBHII", self.buf, p)
        if ver != 1:
            raise ValueError(f"object header version {ver} at {addr:#x}")
        chunks, out = [(p + 16, size)], []
        while chunks:
            q, left = chunks.pop(0)
            end = q + left
            while q + 8 <= end:
                mtype, msize, flags = struct.unpack_from("<HHB", self.buf, q)
                if flags & 0x02 and mtype not in (MSG_NIL,):
                    raise NotImplementedError("h5lite: shared object-header message")
                out.append((mtype, flags, q + 8, msize))
                if mtype == MSG_CONTINUATION:
                    off, ln = struct.unpack_from("<QQ", self.buf, q + 8)
                    chunks.append((self.at(off), ln))
                q += 8 + msize
        return out

    def heap_name(self, heap_data, off):
        s = heap_data + off
        return bytes(self.buf[s:self.buf.index(b"\0", s)]).decode("utf-8")

    def group_entries(self, btree, heap):
        hp = self.at(heap)
        if self.buf[hp:hp + 4] != b"HEAP":
            raise ValueError(f"local heap signature missing at {heap:#x}")
        heap_data = self.at(struct.unpack_from("<Q", self.buf, hp + 24)[0])
        out = []

        def walk(addr):
            p = self.at(addr)
            sig = bytes(self.buf[p:p + 4])
            if sig == b"TREE":
                ntype, _, used = struct.unpack_from("<BBH", self.buf, p + 4)
                if ntype != 0:
                    raise ValueError("group B-tree expected")
                for i in range(used):
                    walk(struct.unpack_from("<Q", self.buf, p + 24 + 8 + 16 * i)[0])
            elif sig == b"SNOD":
                n = struct.unpack_from("<H", self.buf, p + 6)[0]
                for i in range(n):
                    name_off, hdr = struct.unpack_from("<QQ", self.buf, p + 8 + 40 * i)
                    out.append((self.heap_name(heap_data, name_off), hdr))
            else:
                raise ValueError(f"unexpected node signature {sig!r} at {addr:#x}")

        walk(btree)
        return out

    def attribute(self, p, size):
        ver = self.buf[p]
        if ver == 1:
            nlen, dlen, slen = struct.unpack_from("<HHH", self.buf, p + 2)
            q = p + 8
            al = lambda n: (n + 7) // 8 * 8  # noqa: E731
        elif ver in (2, 3):
            if self.buf[p + 1] & 0x03:
                raise NotImplementedError("h5lite: attribute with shared datatype / dataspace")
            nlen, dlen, slen = struct.unpack_from("<HHH", self.buf, p + 2)
            q = p + 8 + (1 if ver == 3 else 0)
            al = lambda n: n  # noqa: E731
        else:
            raise NotImplementedError(f"h5lite: attribute message version {ver}")
        name = bytes(self.buf[q:q + nlen]).split(b"\0")[0].decode("utf-8")
        q += al(nlen)
        dt, is_bool, _ = _decode_datatype(self.buf, q)
        q += al(dlen)
        shape = _decode_dataspace(self.buf, q)
        q += al(slen)
        if shape is None:
            return name, None
        n = int(np.prod(shape, dtype=np.int64))
        v = np.frombuffer(bytes(self.buf[q:q + n * dt.itemsize]), dtype=dt).reshape(shape)
        v = v.astype(np.bool_) if is_bool else v.astype(dt.newbyteorder("="))
        return name, (v[()] if shape == () else v)

    def dataset(self, msgs):
        shape = dt = None
        is_bool = False
        raw = None
        attrs = OrderedDict()
        for mtype, _, p, size in msgs:
            if mtype == MSG_DATASPACE:
                shape = _decode_dataspace(self.buf, p)
            elif mtype == MSG_DATATYPE:
                dt, is_bool, _ = _decode_datatype(self.buf, p)
            elif mtype == MSG_ATTRIBUTE:
                k, v = self.attribute(p, size)
                attrs[k] = v
        if shape is None or dt is None:
            raise ValueError("dataset without dataspace / datatype")
        n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        for mtype, _, p, size in msgs:
            if mtype != MSG_LAYOUT:
                continue
            ver = self.buf[p]
            if ver == 3:
                cls = self.buf[p + 1]
                if cls == 0:
                    ln = struct.unpack_from("<H", self.buf, p + 2)[0]
                    raw = bytes(self.buf[p + 4:p + 4 + ln])
                elif cls == 1:
                    addr, ln = struct.unpack_from("<QQ", self.buf, p + 2)
                    raw = b"" if addr == UNDEF else bytes(self.buf[self.at(addr):self.at(addr) + ln])
                else:
                    raise NotImplementedError("h5lite: chunked dataset (nnabla writes contiguous datasets)")
            elif ver in (1, 2):
                rank, cls = self.buf[p + 1], self.buf[p + 2]
                if cls == 1:
                    addr = struct.unpack_from("<Q", self.buf, p + 8)[0]
                    raw = b"" if addr == UNDEF else bytes(self.buf[self.at(addr):self.at(addr) + n])
                elif cls == 0:
                    q = p + 8 + 4 * rank
                    ln = struct.unpack_from("<I", self.buf, q)[0]
                    raw = bytes(self.buf[q + 4:q + 4 + ln])
                else:
                    raise NotImplementedError("h5lite: chunked dataset (nnabla writes contiguous datasets)")
            else:
                raise NotImplementedError(f"h5lite: data layout message version {ver}")
        if raw is None:
            raise ValueError("dataset without a layout message")
        raw = raw.ljust(n, b"\0")[:n]                  # storage not allocated yet reads as the (zero) fill value
        a = np.frombuffer(raw, dtype=dt).reshape(shape)
        a = a.astype(np.bool_) if is_bool else a.astype(dt.newbyteorder("="))
        return Dataset(a, attrs)

    def object(self, hdr):
        msgs = self.messages(hdr)
        for mtype, _, p, _ in msgs:
            if mtype == MSG_SYMBOL_TABLE:
                bt, hp = struct.unpack_from("<QQ", self.buf, p)
                return OrderedDict((name, self.object(h)) for name, h in self.group_entries(bt, hp))
        if any(m[0] == MSG_LAYOUT for m in msgs):
            return self.dataset(msgs)
        if any(m[0] in (0x2, 0x6) for m in msgs):
            raise NotImplementedError("h5lite: new-style group (link messages; file written with libver='latest')")
        raise ValueError(f"object at {hdr:#x} is neither a group nor a dataset")


def loads(buf):
    """Nested OrderedDict {name: OrderedDict | Dataset} of a file image."""
    r = _Reader(bytes(buf))
    return r.object(r.root_header)


def read(path):
    with open(path, "rb") as f:
        return loads(f.read())


# ------------------------------------------------------------------------------------------------------------------
# nnabla's parameter-file conventions on top of the container format
def save_parameters(path, params, need_grad=None):
    """`nn.save_parameters(path)` for `.h5`: one dataset per parameter at its scope path, attributes `need_grad`
    (bool) and `index` (registration order) - nnabla/utils/... `_h5_parameter_file_saver` as used by the reference's
    python/train.py:101.  params: ordered {name: array}; need_grad: {name: bool} (default True)."""
    tree = OrderedDict()
    for i, (name, v) in enumerate(params.items()):
        node = tree
        parts = name.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, OrderedDict())
            if not isinstance(node, dict):
                raise ValueError(f"{name}: {part} is both a parameter and a scope")
        ng = True if need_grad is None else bool(need_grad.get(name, True))
        node[parts[-1]] = Dataset(np.asarray(v), OrderedDict(need_grad=np.bool_(ng), index=np.int64(i)))
    write(path, tree)


def load_parameters(path):
    """`nn.load_parameters(path)` for `.h5`: ordered {name: array} in the order of the `index` attributes (files
    without them: visiting order), and {name: need_grad}."""
    found = []

    def visit(prefix, node):
        for k, v in node.items():
            name = f"{prefix}/{k}" if prefix else k
            if isinstance(v, dict):
                visit(name, v)
            else:
                idx = v.attrs.get("index")
                found.append((len(found) if idx is None else int(idx), name, v))

    visit("", read(path))
    found.sort(key=lambda t: t[0])
    params = OrderedDict((name, d.data) for _, name, d in found)
    need_grad = {name: bool(d.attrs.get("need_grad", True)) for _, name, d in found}
    return params, need_grad
