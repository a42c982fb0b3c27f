"""CPU: nnabla `.h5` parameter files (reference python/train.py:101, render_image.py:43).

h5py / libhdf5 are absent from this image, so the container format is checked against the HDF5 File Format
Specification directly: the writer's bytes field by field, the reader on a hand-assembled file that uses the
constructs h5py adds (attribute message version 3, continuation chunks, NIL messages, superblock version 1, compact
layout, dataspace version 2), and the round trip of a whole parameter set under the reference's scope names."""
import struct
from collections import OrderedDict

import numpy as np
import pytest

from ndjir_b200 import h5lite, nnabla_names, scene
from ndjir_b200.config import make_conf
from ndjir_b200.engine import ParamStore

U = h5lite.UNDEF


def test_round_trip_nested_tree():
    rng = np.random.RandomState(0)
    tree = OrderedDict()
    tree["a"] = OrderedDict(W=h5lite.Dataset(rng.randn(5, 7).astype(np.float32),
                                             OrderedDict(need_grad=np.bool_(True), index=np.int64(3))),
                            b=rng.randn(7))
    tree["z-last"] = np.arange(12, dtype=np.int32).reshape(3, 2, 2)
    tree["empty"] = np.zeros((0, 4), np.float32)
    tree["scalar"] = h5lite.Dataset(np.float32(2.5), {"note": np.bytes_(b"hello"), "vec": np.arange(3, dtype=np.uint8)})
    tree["flag"] = np.asarray([True, False, True])
    tree["deep"] = {"er": {"est": np.float64(1.0)}, "void": {}}
    back = h5lite.loads(h5lite.dumps(tree))
    assert list(back) == sorted(tree)                       # links come back in the group's (strcmp) order
    W = back["a"]["W"]
    assert W.data.dtype == np.float32 and np.array_equal(W.data, tree["a"]["W"].data)
    assert W.attrs["need_grad"] == np.bool_(True) and W.attrs["need_grad"].dtype == np.bool_
    assert W.attrs["index"] == 3 and W.attrs["index"].dtype == np.int64
    assert back["a"]["b"].data.dtype == np.float64 and np.array_equal(back["a"]["b"].data, tree["a"]["b"])
    assert np.array_equal(back["z-last"].data, tree["z-last"]) and back["z-last"].data.dtype == np.int32
    assert back["empty"].data.shape == (0, 4)
    assert back["scalar"].data.shape == () and back["scalar"].data == np.float32(2.5)
    assert back["scalar"].attrs["note"] == b"hello" and np.array_equal(back["scalar"].attrs["vec"], [0, 1, 2])
    assert back["flag"].data.dtype == np.bool_ and back["flag"].data.tolist() == [True, False, True]
    assert back["deep"]["er"]["est"].data == 1.0 and back["deep"]["void"] == {}


@pytest.mark.parametrize("n", [1, 8, 9, 256, 257, 700])
def test_group_btree_levels(n):
    """8 links per symbol node, 32 children per B-tree node: 257 links need a second B-tree level."""
    tree = OrderedDict((f"p{(i * 7919) % n:04d}", np.full((2,), i, np.float32)) for i in range(n))
    img = h5lite.dumps(tree)
    back = h5lite.loads(img)
    assert list(back) == sorted(tree)
    for k, v in tree.items():
        assert np.array_equal(back[k].data, v)
    # root B-tree node: level = ceil(log32(ceil(n / 8))) - 1 ... checked through the node header
    bt = struct.unpack_from("<Q", img, 24 + 32 + 24)[0]
    assert img[bt:bt + 4] == b"TREE"
    ntype, level, used = struct.unpack_from("<BBH", img, bt + 4)
    snods = -(-n // 8)
    assert ntype == 0 and level == (0 if snods <= 32 else 1) and used == (snods if snods <= 32 else -(-snods // 32))
    # keys of a node ascend in name order (key = heap offset of the largest name of the child to its left)
    heap = struct.unpack_from("<Q", img, 24 + 32 + 32)[0]
    heap_data = struct.unpack_from("<Q", img, heap + 24)[0]
    keys = [struct.unpack_from("<Q", img, bt + 24 + 16 * i)[0] for i in range(used + 1)]
    names = [img[heap_data + k:img.index(b"\0", heap_data + k)] for k in keys]
    assert names[0] == b"" and names == sorted(names)


def test_writer_bytes_follow_the_specification():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    img = h5lite.dumps({"g": {"x": h5lite.Dataset(a, OrderedDict(need_grad=np.bool_(False), index=np.int64(7)))}})
    # superblock version 0 (spec III.A): signature, versions, 8-byte offsets and lengths, K values, addresses
    assert img[:8] == b"\x89HDF\r\n\x1a\n"
    assert tuple(img[8:16]) == (0, 0, 0, 0, 0, 8, 8, 0)
    assert struct.unpack_from("<HHI", img, 16) == (4, 16, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", img, 24)
    assert (base, free, eof, drv) == (0, U, len(img), U) and len(img) % 8 == 0
    name_off, root, cache, _, bt, heap = struct.unpack_from("<QQIIQQ", img, 56)
    assert name_off == 0 and cache == 1
    # root object header (version 1, spec IV.A.1.a): one symbol-table message pointing at the same B-tree / heap
    ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", img, root)
    assert (ver, nmsg, refc, size) == (1, 1, 1, 24)
    assert struct.unpack_from("<HHB", img, root + 16) == (0x11, 16, 0)
    assert struct.unpack_from("<QQ", img, root + 24) == (bt, heap)
    # local heap (III.D): signature, version 0, free list "none", data segment holding "" and "g"
    assert img[heap:heap + 4] == b"HEAP" and img[heap + 4] == 0
    dsize, free_off, daddr = struct.unpack_from("<QQQ", img, heap + 8)
    assert free_off == 1 and dsize == 16 and img[daddr:daddr + 16] == b"\0" * 8 + b"g" + b"\0" * 7
    # B-tree node (III.A.1) -> symbol node (III.C) -> entry for "g" with cached B-tree / heap addresses
    assert img[bt:bt + 4] == b"TREE" and struct.unpack_from("<BBHQQ", img, bt + 4) == (0, 0, 1, U, U)
    k0, snod, k1 = struct.unpack_from("<QQQ", img, bt + 24)
    assert (k0, k1) == (0, 8) and img[snod:snod + 4] == b"SNOD" and struct.unpack_from("<BBH", img, snod + 4) == (1, 0, 1)
    g_name, g_hdr, g_cache = struct.unpack_from("<QQI", img, snod + 8)
    assert g_name == 8 and g_cache == 1
    g_bt, g_heap = struct.unpack_from("<QQ", img, snod + 8 + 24)
    assert struct.unpack_from("<QQ", img, g_hdr + 24) == (g_bt, g_heap)
    g_snod = struct.unpack_from("<Q", img, g_bt + 32)[0]
    x_hdr = struct.unpack_from("<Q", img, g_snod + 8 + 8)[0]
    # dataset header: dataspace, datatype, fill value, layout, two attributes; every message 8-byte aligned
    ver, _, nmsg, _, size = struct.unpack_from("<BBHII", img, x_hdr)
    assert (ver, nmsg) == (1, 6)
    p, seen = x_hdr + 16, []
    while p < x_hdr + 16 + size:
        t, s, fl = struct.unpack_from("<HHB", img, p)
        assert s % 8 == 0
        seen.append((t, p + 8))
        p += 8 + s
    assert [t for t, _ in seen] == [0x1, 0x3, 0x5, 0x8, 0xC, 0xC]
    q = seen[0][1]      # dataspace version 1, rank 2, dims 2 x 3
    assert tuple(img[q:q + 8]) == (1, 2, 0, 0, 0, 0, 0, 0) and struct.unpack_from("<QQ", img, q + 8) == (2, 3)
    q = seen[1][1]      # IEEE little-endian float32: class 1 version 1, sign bit 31, exponent 23 / 8, mantissa 0 / 23, bias 127
    assert tuple(img[q:q + 4]) == (0x11, 0x20, 31, 0) and struct.unpack_from("<I", img, q + 4)[0] == 4
    assert struct.unpack_from("<HHBBBBI", img, q + 8) == (0, 32, 23, 8, 0, 23, 127)
    q = seen[3][1]      # layout version 3, contiguous, raw data where it says
    lv, lc, addr, ln = struct.unpack_from("<BBQQ", img, q)
    assert (lv, lc, ln) == (3, 1, 24) and img[addr:addr + ln] == a.tobytes() and addr % 8 == 0
    q = seen[4][1]      # attribute version 1: "need_grad", enum over int8 {FALSE = 0, TRUE = 1}, scalar, value 0
    av, _, nlen, dlen, slen = struct.unpack_from("<BBHHH", img, q)
    assert (av, nlen, dlen, slen) == (1, 10, 38, 8) and img[q + 8:q + 18] == b"need_grad\0"
    d = q + 8 + 16
    assert tuple(img[d:d + 4]) == (0x18, 2, 0, 0) and struct.unpack_from("<I", img, d + 4)[0] == 1
    assert tuple(img[d + 8:d + 12]) == (0x10, 0x08, 0, 0) and struct.unpack_from("<IHH", img, d + 12) == (1, 0, 8)
    assert img[d + 20:d + 36] == b"FALSE\0\0\0TRUE\0\0\0\0" and img[d + 36:d + 38] == b"\0\1"
    assert tuple(img[d + 40:d + 48]) == (1, 0, 0, 0, 0, 0, 0, 0) and img[d + 48] == 0
    q = seen[5][1]      # "index": int64 scalar 7
    assert img[q + 8:q + 14] == b"index\0" and struct.unpack_from("<q", img, q + 8 + 8 + 16 + 8)[0] == 7


def _foreign_file():
    """A file assembled here, independently of the writer, the way h5py lays a small parameter file out: superblock
    version 1, a dataset header whose attributes live in a continuation chunk after a NIL message, attribute message
    version 3 (UTF-8 names), dataspace version 2, a compact dataset, big-endian data."""
    buf = bytearray(100)

    def put(b):
        buf.extend(b"\0" * (-len(buf) % 8))
        a = len(buf)
        buf.extend(b)
        return a

    def msg(t, data, flags=0):
        data = data + b"\0" * (-len(data) % 8)
        return struct.pack("<HHB3x", t, len(data), flags) + data

    def header(msgs, nmsg=None):
        body = b"".join(msgs)
        return put(struct.pack("<BBHII4x", 1, 0, nmsg or len(msgs), 1, len(body)) + body)

    f32 = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    f64be = struct.pack("<BBBBI", 0x11, 0x21, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    i64 = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)
    enum3 = struct.pack("<BBBBI", 0x38, 2, 0, 0, 1) + struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 1) + struct.pack("<HH", 0, 8) \
        + b"FALSE\0TRUE\0" + b"\0\1"                                             # enum version 3: names not padded
    scalar2 = struct.pack("<BBBB", 2, 0, 0, 0)                                   # dataspace version 2, scalar

    def attr3(name, dt, raw):
        nm = name.encode() + b"\0"
        return msg(0xC, struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(scalar2), 1) + nm + dt + scalar2 + raw)

    W = np.arange(8, dtype=np.float32).reshape(2, 4)
    w_raw = put(W.tobytes())
    cont = put(msg(0x0, b"\0" * 16) + attr3("need_grad", enum3, b"\1") + attr3("index", i64, struct.pack("<q", 1)))
    cont_len = len(buf) - cont
    w_hdr = header([msg(0x1, struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 2, 4)), msg(0x3, f32, 1),
                    msg(0x5, struct.pack("<BBBBi", 2, 2, 2, 1, 0)), msg(0x8, struct.pack("<BBQQ", 3, 1, w_raw, 32)),
                    msg(0x10, struct.pack("<QQ", cont, cont_len))], nmsg=8)
    b = np.asarray([1.5, -2.0], dtype=">f8")
    b_hdr = header([msg(0x1, struct.pack("<BBBBI", 1, 1, 0, 0, 0) + struct.pack("<Q", 2)), msg(0x3, f64be, 1),
                    msg(0x8, struct.pack("<BBH", 3, 0, 16) + b.tobytes()),
                    attr3("index", i64, struct.pack("<q", 0))])
    heap_data = put(b"\0" * 8 + b"W\0\0\0\0\0\0\0" + b"b\0\0\0\0\0\0\0" + b"\0" * 8)
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, 32, 24, heap_data))           # a free block at offset 24
    buf[heap_data + 24:heap_data + 32] = struct.pack("<Q", 1)                    # ... whose "next" is FREE_NULL
    snod = put((b"SNOD" + struct.pack("<BBH", 1, 0, 2) + struct.pack("<QQI4x16x", 8, w_hdr, 0)
                + struct.pack("<QQI4x16x", 16, b_hdr, 0)).ljust(328, b"\0"))
    bt = put((b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, U, U) + struct.pack("<QQQ", 0, snod, 16)).ljust(544, b"\0"))
    root = header([msg(0x11, struct.pack("<QQ", bt, heap))])
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 1, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", 0, U, len(buf), U) + struct.pack("<QQI4xQQ", 0, root, 1, bt, heap)
    assert len(sb) == 100
    buf[:100] = sb
    return bytes(buf), W, b


def test_reader_handles_what_h5py_adds(tmp_path):
    img, W, b = _foreign_file()
    t = h5lite.loads(img)
    assert list(t) == ["W", "b"]
    assert np.array_equal(t["W"].data, W) and t["W"].attrs["need_grad"] == np.bool_(True) and t["W"].attrs["index"] == 1
    assert t["b"].data.dtype == np.float64 and np.array_equal(t["b"].data, [1.5, -2.0]) and t["b"].attrs["index"] == 0
    p = tmp_path / "foreign.h5"
    p.write_bytes(img)
    params, ng = h5lite.load_parameters(str(p))
    assert list(params) == ["b", "W"] and ng == {"b": True, "W": True}      # ordered by the `index` attribute


def test_reader_names_what_it_does_not_read():
    img = bytearray(h5lite.dumps({"x": np.zeros(3, np.float32)}))
    with pytest.raises(ValueError):
        h5lite.loads(b"not an hdf5 file" * 8)
    v2 = bytearray(img)
    v2[8] = 2
    with pytest.raises(NotImplementedError, match="superblock version 2"):
        h5lite.loads(bytes(v2))
    chunked = bytearray(img)
    # flip the layout class of the only dataset to "chunked"
    hdr = struct.unpack_from("<Q", img, img.index(b"SNOD") + 8 + 8)[0]
    p = hdr + 16
    while struct.unpack_from("<H", img, p)[0] != 0x8:
        p += 8 + struct.unpack_from("<H", img, p + 2)[0]
    chunked[p + 8 + 1] = 2
    with pytest.raises(NotImplementedError, match="chunked"):
        h5lite.loads(bytes(chunked))


@pytest.mark.parametrize("kind", ["default", "triplaneline", "no_voxel"])
def test_parameter_file_round_trip_under_reference_names(kind, tmp_path):
    conf = make_conf(kind, geometric_network={"feature_size": 128, "voxel": {"grid_size": 8}})
    P = scene.init_params(conf, seed=5)
    P["pl_gain"] = np.asarray([3.25], np.float32)
    ps = ParamStore(conf, "cpu")
    ps.load_reference(P)
    path = str(tmp_path / "param_00010.h5")
    ps.save_parameters(path)
    params, need_grad = h5lite.load_parameters(path)
    want = nnabla_names.to_nnabla(conf, P)
    assert list(params) == list(want)                       # registration order kept through the `index` attributes
    for k, v in want.items():
        assert params[k].dtype == np.float32 and np.array_equal(params[k], v), k
    assert "roughness-network/affine--1/affine/W" in params                  # the reference's off-by-one scope (q16)
    assert need_grad["geometric-network/gain"] and not need_grad["photogrammetric-light-network/gain"]
    assert params["geometric-network/affine-00/affine/W"].shape == P["geo"][0][0].shape      # (in, out) like PF.affine
    ps2 = ParamStore(conf, "cpu")
    ps2.load_parameters(path)
    assert np.array_equal(ps2.data.numpy(), ps.data.numpy()) and ps2.pl_gain == 3.25
    for k in ps.grid:
        assert np.array_equal(ps2.grid[k].numpy(), ps.grid[k].numpy())


@pytest.mark.parametrize("off,scope", [("implicit_illumination_network", "implicit-illumination-network"),
                                       ("photogrammetric_light_network", "photogrammetric-light-network")])
def test_parameter_file_of_a_config_that_switches_a_network_off(off, scope, tmp_path):
    """config/no_implicit_illumination.yaml / no_lightp.yaml: the reference never creates the network, so its parameter
    file has no entry under that scope (network.py:308-309, renderer.py:161).  Ours writes the same set of names and
    reads such a file back: the active networks bit for bit, the unused block of the flat store zero."""
    conf = make_conf("default", geometric_network={"feature_size": 128, "voxel": {"grid_size": 8}}, **{off: {"use_me": False}})
    full = make_conf("default", geometric_network={"feature_size": 128, "voxel": {"grid_size": 8}})
    P = scene.init_params(conf, seed=5)
    ps = ParamStore(conf, "cpu")
    ps.load_reference(P)
    path = str(tmp_path / "param.h5")
    ps.save_parameters(path)
    params, _ = h5lite.load_parameters(path)
    names_full = [n for n, _ in nnabla_names.parameter_names(full)]
    assert not [n for n in params if n.startswith(scope + "/")]
    assert list(params) == [n for n in names_full if not n.startswith(scope + "/")]
    ps2 = ParamStore(conf, "cpu")
    ps2.load_parameters(path)
    net = {"implicit_illumination_network": "ii", "photogrammetric_light_network": "pl"}[off]
    a, b = ps.reference_dict(), ps2.reference_dict()
    for name in scene.NET_ORDER:
        for (W1, b1), (W2, b2) in zip(a[name], b[name]):
            if name == net:
                assert not W2.any() and not b2.any()
            else:
                assert np.array_equal(W1, W2) and np.array_equal(b1, b2), name
