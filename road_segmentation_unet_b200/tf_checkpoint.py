"""TensorFlow-1.x checkpoint bundles (what `tf.train.Saver.save` writes and `.restore` reads:
tf_aerial_images.py:171, :343-379, run.py:135,164) without TensorFlow.

A V2 checkpoint `prefix` is
    prefix.index                  an SSTable (LevelDB table format: prefix-compressed key/value
                                  blocks + restart arrays, 5-byte block trailers, 48-byte footer)
                                  mapping "" -> BundleHeaderProto and every tensor name ->
                                  BundleEntryProto (dtype, shape, shard, offset, size, crc32c)
    prefix.data-00000-of-00001    the raw little-endian tensor bytes
    prefix.meta                   the MetaGraphDef (not needed when the graph is built by code;
                                  the reference's restore() only globs for it, :371-373)

`read_bundle` imports such a checkpoint into {name: ndarray} -- the published weights of
run.py:14 load with it, variable names and layouts already match this engine's -- and
`write_bundle` produces one that TensorFlow's Saver can restore.  Checksums are CRC-32C
(rsu_crc32c_host in librsu_b200.so), masked the LevelDB way.

Format follows tensorflow/core/lib/io/{format,block,table_builder}.cc and
tensorflow/core/util/tensor_bundle/tensor_bundle.cc + protobuf/tensor_bundle.proto as published
(TensorFlow r1.4).  No TensorFlow-written file exists in the reference tree or in this
environment, so the importer is pinned only by round trips through the writer and by the
format's published constants (table magic number, crc mask delta).
"""
import ctypes as C
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
CRC_MASK_DELTA = 0xa282ead8
BLOCK_SIZE = 256 * 1024   # tensorflow/core/lib/io/table_options.h
RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64 = 1, 2, 3, 4, 9
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_UINT8: np.uint8,
             DT_INT64: np.int64}
_DT_OF_NP = {np.dtype(v): k for k, v in _NP_OF_DT.items()}


# ------------------------------------------------------------------ checksums
def crc32c(data, crc=0):
    from . import _lib
    buf = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    arr = (C.c_char * len(buf)).from_buffer_copy(buf) if len(buf) else None
    return int(_lib.load().rsu_crc32c_host(C.c_uint(crc), arr, C.c_ulonglong(len(buf))))


def crc32c_array(a):
    from . import _lib
    a = np.ascontiguousarray(a)
    return int(_lib.load().rsu_crc32c_host(C.c_uint(0), C.c_void_p(a.ctypes.data), C.c_ulonglong(a.nbytes)))


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + CRC_MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / protobuf wire format
def _put_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _pb_fields(buf):
    """Yield (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _get_varint(buf, pos)
        elif wire == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        yield field, wire, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for field, wire, v in _pb_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:  # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
            for f2, _, dim in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, sv in _pb_fields(dim):
                        if f3 == 1:
                            size = _signed64(sv)
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = _signed64(v)
        elif field == 5:
            e["size"] = _signed64(v)
        elif field == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif field == 7:
            e["sliced"] = True
    return e


def _encode_entry(dtype, shape, offset, size, crc):
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d
                    for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype)
    out += b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


def _encode_header(num_shards=1, producer=1):
    # BundleHeaderProto { num_shards = 1; endianness = 2 (LITTLE = 0, omitted);
    #                     VersionDef version = 3 { producer = 1 } }
    # producer = kTensorBundleVersion (1) of tensor_bundle.cc; the reader accepts any producer >= 0
    ver = b"\x08" + _put_varint(producer)
    return b"\x08" + _put_varint(num_shards) + b"\x1a" + _put_varint(len(ver)) + ver


# ------------------------------------------------------------------ SSTable (LevelDB table) reader
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):  # overlapping copies are byte-serial by definition
            out.append(out[-off])
    assert len(out) == n, "corrupt snappy block"
    return bytes(out)


def _read_block(data, offset, size, verify):
    raw = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        got = mask_crc(crc32c(raw + trailer[:1]))
        if want != got:
            raise ValueError("checkpoint index block checksum mismatch")
    if trailer[0] == 0:
        return raw
    if trailer[0] == 1:
        return _snappy_uncompress(raw)
    raise ValueError("unknown block compression %d" % trailer[0])


def _block_entries(block):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = data[-48:]
    _, pos = _get_varint(footer, 0)          # metaindex handle: offset
    _, pos = _get_varint(footer, pos)        #                   size
    idx_off, pos = _get_varint(footer, pos)  # index handle
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p2 = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p2)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ------------------------------------------------------------------ SSTable writer
class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""

    def add(self, key, value):
        shared = 0
        if self.count % RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(path, items):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order."""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index = _BlockBuilder()
    cur = _BlockBuilder()
    prev = None
    for key, value in items:
        assert prev is None or key > prev, "keys must be added in increasing order"
        cur.add(key, value)
        prev = key
        if cur.size() >= BLOCK_SIZE:
            index.add(key, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count:
        index.add(prev, emit(cur.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


# ------------------------------------------------------------------ bundles
def is_bundle(prefix):
    return os.path.exists(prefix + ".index")


def read_bundle(prefix, verify=True):
    """{tensor name: ndarray} of the checkpoint `prefix` (prefix.index + prefix.data-*)."""
    entries, num_shards = {}, 1
    for key, value in read_table(prefix + ".index", verify):
        if key == b"":
            for field, _, v in _pb_fields(value):
                if field == 1:
                    num_shards = v
                elif field == 2 and v != 0:
                    raise ValueError("big-endian checkpoints are not supported")
            continue
        entries[key.decode()] = _parse_entry(value)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e["sliced"]:
            raise ValueError("partitioned variable %s is not supported" % name)
        if e["dtype"] not in _NP_OF_DT:
            continue  # strings etc. (none among the model's variables)
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        a = np.frombuffer(raw, dtype=_NP_OF_DT[e["dtype"]]).reshape(e["shape"]).copy()
        if verify and e["crc32c"] is not None and mask_crc(crc32c_array(a)) != e["crc32c"]:
            raise ValueError("tensor %s: checksum mismatch" % name)
        out[name] = a
    return out


def write_bundle(prefix, tensors):
    """Write {name: ndarray} as a single-shard V2 checkpoint (names sorted, as the Saver does)."""
    names = sorted(tensors)
    items = [(b"", _encode_header())]
    offset = 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            shape = np.asarray(tensors[name]).shape  # (ascontiguousarray turns 0-d into 1-d)
            a = np.ascontiguousarray(tensors[name])
            if a.dtype not in _DT_OF_NP:
                raise ValueError("tensor %s: unsupported dtype %s" % (name, a.dtype))
            f.write(a.tobytes())
            items.append((name.encode(), _encode_entry(_DT_OF_NP[a.dtype], shape, offset, a.nbytes,
                                                       mask_crc(crc32c_array(a)))))
            offset += a.nbytes
    write_table(prefix + ".index", items)
