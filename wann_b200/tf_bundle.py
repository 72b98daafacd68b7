"""TensorFlow V2 checkpoint bundle I/O (the on-disk format either side of the hot path:
`Saver.restore(sess, '../policy_net_model/model')`, policy_net_utils.pyx:68-73).

A bundle is `<prefix>.index` -- a LevelDB-format SSTable mapping tensor name -> BundleEntryProto
{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6} (first key "" = BundleHeaderProto) --
plus `<prefix>.data-00000-of-00001` holding the raw little-endian tensors.  The reference ships
`policy_net_model/model.index` but NOT the data shard (.MISSING_LARGE_BLOBS:1); this reader
verifies every tensor against the masked CRC32C stored in the index so that a real shard can be
dropped in and trusted.  No TensorFlow dependency: the two protobufs are parsed by hand."""
import os
import struct

import numpy as np

from . import _lib as L

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT = 1
_MASK_DELTA = 0xa282ead8


def crc32c(buf, seed=0):
    b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else \
        np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return int(L.load().wann_crc32c(b.ctypes.data if b.size else None, b.size, seed))


def mask_crc(crc):
    """tensorflow/core/lib/hash/crc32c.h Mask(): rotate right 15 and add a constant."""
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------ varint / protobuf (minimal)
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf):
    """protobuf wire format -> list of (field, wiretype, value)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _parse_entry(value):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for f, wt, v in _parse_fields(value):
        if f == 1:
            e["dtype"] = v
        elif f == 2:                                   # TensorShapeProto { repeated Dim dim = 2 {size=1} }
            dims = []
            for f2, _, v2 in _parse_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _parse_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v                            # fixed32, masked
    return e


# ------------------------------------------------------------------ SSTable (LevelDB table format)
def _read_block(data, offset, size):
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
        raise ValueError("SSTable block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise ValueError("compressed SSTable blocks (type %d) are not supported" % ctype)
    n_restarts = struct.unpack_from("<I", contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(contents, pos)
        unshared, pos = _varint(contents, pos)
        vlen, pos = _varint(contents, pos)
        key = key[:shared] + bytes(contents[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(contents[pos:pos + vlen])))
        pos += vlen
    return out


def read_index(index_path):
    """-> dict name -> entry {dtype, shape, shard_id, offset, size, crc32c}; '' header dropped."""
    data = open(index_path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TF bundle index (bad SSTable magic)" % index_path)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)       # index handle
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _read_block(data, ioff, isize):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for key, value in _read_block(data, boff, bsize):
            if key == b"":
                continue
            entries[key.decode()] = _parse_entry(value)
    return entries


def load_policy_weights(prefix, scope="", final_kernel=3, verify=True):
    """Read the 12 policy-net tensors from <prefix>.index / <prefix>.data-00000-of-00001.
    `scope` = '' for the shipped single net, 'white_net/' or 'black_net/' for the dual nets
    (policy_net_utils.pyx:22-25).  Raises if a tensor is missing, mis-shaped or fails its CRC."""
    from . import weights as W
    entries = read_index(prefix + ".index")
    shard = prefix + ".data-00000-of-00001"
    if not os.path.exists(shard):
        raise FileNotFoundError("%s is missing (the reference repository does not ship it; "
                                "expected %d bytes)" % (shard, max(e["offset"] + e["size"] for e in entries.values())))
    out = {}
    with open(shard, "rb") as f:
        for name, shape in W.shapes(final_kernel).items():
            key = scope + name
            if key not in entries:
                raise KeyError("tensor %r not in %s.index" % (key, prefix))
            e = entries[key]
            if e["dtype"] != DT_FLOAT or tuple(e["shape"]) != tuple(shape):
                raise ValueError("tensor %r: dtype %d shape %s, expected float32 %s" % (key, e["dtype"], e["shape"], shape))
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"] or e["size"] != 4 * int(np.prod(shape)):
                raise ValueError("tensor %r: short read / size mismatch" % key)
            if verify and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise ValueError("tensor %r fails its CRC32C (index %08x)" % (key, e["crc32c"]))
            out[name] = np.frombuffer(raw, dtype="<f4").reshape(shape).copy()
    return out


# ------------------------------------------------------------------ writer (round trips, fixtures)
def _block(entries):
    """One uncompressed SSTable block with restart interval 1 (no key prefix sharing)."""
    body, restarts = bytearray(), []
    for key, value in entries:
        restarts.append(len(body))
        body += _enc_varint(0) + _enc_varint(len(key)) + _enc_varint(len(value)) + key + value
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    trailer = bytes([0])
    return bytes(body) + trailer + struct.pack("<I", mask_crc(crc32c(bytes(body) + trailer)))


def _pb(field, wt, payload):
    key = _enc_varint((field << 3) | wt)
    if wt == 0:
        return key + _enc_varint(payload)
    if wt == 2:
        return key + _enc_varint(len(payload)) + payload
    if wt == 5:
        return key + struct.pack("<I", payload)
    raise ValueError(wt)


def save_policy_weights(prefix, weights, scope=""):
    """Write a V2 bundle (one data shard) readable by load_policy_weights."""
    names = sorted(weights)
    shard = bytearray()
    rows = [(b"", _pb(1, 0, 1) + _pb(3, 2, _pb(1, 0, 1)))]      # header: num_shards=1, version{producer=1}
    for name in names:
        a = np.ascontiguousarray(weights[name], dtype="<f4")
        raw = a.tobytes()
        shape = b"".join(_pb(2, 2, _pb(1, 0, d)) for d in a.shape)
        entry = _pb(1, 0, DT_FLOAT) + _pb(2, 2, shape) + _pb(4, 0, len(shard)) + _pb(5, 0, len(raw)) + \
            _pb(6, 5, mask_crc(crc32c(raw)))
        rows.append(((scope + name).encode(), entry))
        shard += raw
    rows.sort(key=lambda kv: kv[0])
    data_block = _block(rows)
    meta_block = _block([])
    meta_off = len(data_block)
    index_block = _block([(rows[-1][0] + b"\x00", _enc_varint(0) + _enc_varint(len(data_block) - 5))])
    index_off = meta_off + len(meta_block)
    footer = _enc_varint(meta_off) + _enc_varint(len(meta_block) - 5) + \
        _enc_varint(index_off) + _enc_varint(len(index_block) - 5)
    footer = footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(data_block + meta_block + index_block + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(shard))
