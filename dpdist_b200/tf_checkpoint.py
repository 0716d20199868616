"""Reader / writer of TensorFlow's V2 checkpoint format ("tensor bundle": `<prefix>.index` +
`<prefix>.data-00000-of-00001`), restricted to what the DPDist variables need: dense float32 tensors in a
single shard.  Lets a reference-trained `model.ckpt` (train_multi_gpu_pc_compare_dist.py:311,
`saver.save(sess, os.path.join(LOG_DIR, "model.ckpt"))`; consumed by
pcrnet-registration/iterative_PCRNet_ours.py:229) be loaded under its TF variable names without TensorFlow.

TensorFlow is not vendored in the reference ("TensorFlow >= 1.14", README.md:39) and cannot be installed here,
so the format is restated from its published definition and has only been round-trip tested against this
module's own writer, never against a TF-written file:
  * `.index` is a LevelDB-format table (tensorflow/core/lib/io/table): prefix-compressed key/value blocks with a
    restart array, each followed by a 1-byte compression tag (0 = none, 1 = snappy) and a masked CRC32C; a 48-byte
    footer = metaindex BlockHandle, index BlockHandle (varint64 offset, size), zero padding, magic 0xdb4775248b80fb57.
  * key ""  -> BundleHeaderProto {num_shards = 1, endianness = 2, version = 3}
    key <variable name> -> BundleEntryProto {dtype = 1, shape = 2 (TensorShapeProto: repeated Dim{size = 1}),
    shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32, masked), slices = 7}
  * `.data-*` holds the raw little-endian tensor bytes at [offset, offset + size).
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT = 1
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


# ---------------------------------------------------------------- varints / protobuf wire format
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> list of (field, wire_type, value); value = int (varint / fixed) or bytes (length-delimited)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _field(field, wt, payload):
    tag = _put_varint((field << 3) | wt)
    if wt == 0:
        return tag + _put_varint(payload)
    if wt == 2:
        return tag + _put_varint(len(payload)) + payload
    if wt == 5:
        return tag + struct.pack("<I", payload)
    raise ValueError(wt)


# ---------------------------------------------------------------- CRC32C (Castagnoli), masked as LevelDB / TF do
def _crc_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return np.array(tbl, dtype=np.uint32)


_CRC_TABLE = _crc_table()


def crc32c(data):
    data = bytes(data)
    if len(data) > 4096:     # the variables are megabytes: use the library's host routine
        from . import _lib
        return int(_lib.load().dpd_crc32c(data, len(data)))
    c = 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------- table blocks
def _read_block(buf, offset, size):
    tag = buf[offset + size]
    if tag == 1:
        raise NotImplementedError("snappy-compressed index blocks are not supported (TF writes them uncompressed)")
    if tag != 0:
        raise ValueError("unknown block compression tag %d" % tag)
    block = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _handle(buf, pos):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def _write_block(entries):
    """One block, no prefix compression (restart interval 1), + compression tag + masked crc."""
    body, restarts = bytearray(), []
    for k, v in entries:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    trailer = b"\x00"
    return bytes(body), trailer + struct.pack("<I", _mask(crc32c(bytes(body) + trailer)))


# ---------------------------------------------------------------- public API
def _prefix(path):
    for suf in (".index", ".meta"):
        if path.endswith(suf):
            return path[:-len(suf)]
    return path


def list_variables(path):
    """-> {name: (dtype, shape)} of a V2 checkpoint prefix (e.g. 'log/.../model.ckpt')."""
    return {n: (e["dtype"], e["shape"]) for n, e in _read_index(_prefix(path)).items()}


def _read_index(prefix):
    with open(prefix + ".index", "rb") as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s.index is not a TensorFlow V2 checkpoint index (bad table magic)" % prefix)
    footer = buf[-48:]
    _, _, p = _handle(footer, 0)                 # metaindex
    ioff, isize, _ = _handle(footer, p)          # index block
    entries = {}
    for _, hv in _read_block(buf, ioff, isize):
        boff, bsize, _ = _handle(hv, 0)
        for key, val in _read_block(buf, boff, bsize):
            if key == b"":
                hdr = dict((f, v) for f, _, v in _parse_proto(val))
                if hdr.get(1, 1) != 1:
                    raise NotImplementedError("multi-shard checkpoints are not supported")
                if hdr.get(2, 0) != 0:
                    raise NotImplementedError("big-endian checkpoints are not supported")
                continue
            e = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0, "crc": None, "sliced": False}
            for f, _, v in _parse_proto(val):
                if f == 1:
                    e["dtype"] = v
                elif f == 2:
                    e["shape"] = tuple(dict((ff, vv) for ff, _, vv in _parse_proto(d)).get(1, 0)
                                       for ff0, _, d in _parse_proto(v) if ff0 == 2)
                elif f == 3:
                    e["shard"] = v
                elif f == 4:
                    e["offset"] = v
                elif f == 5:
                    e["size"] = v
                elif f == 6:
                    e["crc"] = v
                elif f == 7:
                    e["sliced"] = True
            entries[key.decode()] = e
    return entries


def load_checkpoint(path, names=None, verify_crc=True):
    """-> {tf_variable_name: ndarray}.  `names`: optional iterable restricting what is read (optimizer slots such
    as '.../weights/Adam' are skipped that way)."""
    prefix = _prefix(path)
    index = _read_index(prefix)
    data_path = prefix + ".data-00000-of-00001"
    out = {}
    with open(data_path, "rb") as fh:
        for n, e in index.items():
            if names is not None and n not in names:
                continue
            if e["sliced"] or e["shard"] != 0:
                raise NotImplementedError("variable %s is partitioned / sharded" % n)
            if e["dtype"] not in _DTYPES:
                raise NotImplementedError("variable %s has unsupported dtype enum %d" % (n, e["dtype"]))
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("variable %s: data file truncated" % n)
            if verify_crc and e["crc"] is not None and _mask(crc32c(raw)) != e["crc"]:
                raise ValueError("variable %s: crc32c mismatch" % n)
            out[n] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def save_checkpoint(path, variables):
    """Write {name: float32 ndarray} as a single-shard V2 checkpoint at prefix `path`."""
    prefix = _prefix(path)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries, offset = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for n in sorted(variables):
            a = np.asarray(variables[n], dtype=np.float32)
            a = a if a.ndim == 0 else np.ascontiguousarray(a)
            raw = a.tobytes()
            fh.write(raw)
            shape = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in a.shape)
            val = (_field(1, 0, DT_FLOAT) + _field(2, 2, shape) + (_field(4, 0, offset) if offset else b"") +
                   _field(5, 0, len(raw)) + _field(6, 5, _mask(crc32c(raw))))
            entries.append((n.encode(), val))
            offset += len(raw)
    header = _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1))          # num_shards = 1, version {producer = 1}
    entries = [(b"", header)] + entries
    out = bytearray()
    body, tr = _write_block(entries)
    data_handle = _put_varint(0) + _put_varint(len(body))
    out += body + tr
    meta_off = len(out)
    body, tr = _write_block([])
    meta_handle = _put_varint(meta_off) + _put_varint(len(body))
    out += body + tr
    idx_off = len(out)
    body, tr = _write_block([(entries[-1][0] + b"\x00", data_handle)])   # index key >= last key of the data block
    idx_handle = _put_varint(idx_off) + _put_varint(len(body))
    out += body + tr
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))
    return prefix
