"""TF V2 checkpoint (tensor bundle) reader / writer used to load a reference-trained model.ckpt under its TF
variable names.  CPU only.  The format is restated from TensorFlow's published definition;
these tests pin the pieces with known answers (CRC32C check value, LevelDB prefix compression) and round trips."""
import os
import struct

import numpy as np
import pytest

from dpdist_b200 import _lib, tf_checkpoint as T

NAMES = ["pc_compare/dpdist_local/mapper_conv%d/%s" % (i, s) for i in (1, 2, 3, 4) for s in ("weights", "biases")]
SHAPES = [(1, 2503, 1, 1024), (1024,), (1, 1, 1024, 1024), (1024,), (1, 1, 1024, 1024), (1024,), (1, 1, 1024, 3), (3,)]


def test_crc32c_known_answers():
    # CRC-32C (Castagnoli) check value of "123456789" (RFC 3720 appendix B.4 vectors): 0xE3069283
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    lib = _lib.load()
    assert lib.dpd_crc32c(b"123456789", 9) == 0xE3069283
    big = bytes(np.random.default_rng(0).integers(0, 256, 100000, dtype=np.uint8))
    c = 0xFFFFFFFF
    for b in big[:5000]:
        c = int(T._CRC_TABLE[(c ^ b) & 0xFF]) ^ (c >> 8)
    assert lib.dpd_crc32c(big[:5000], 5000) == c ^ 0xFFFFFFFF          # library routine == table-driven python
    # LevelDB mask: rotate right by 15 and add 0xa282ead8
    assert T._mask(0) == 0xa282ead8


def test_round_trip_of_the_dpdist_variables(tmp_path):
    rng = np.random.default_rng(1)
    var = {n: rng.normal(size=s).astype(np.float32) for n, s in zip(NAMES, SHAPES)}
    var["beta1_power"] = np.float32(0.9)                                # scalars (optimizer state) survive too
    prefix = T.save_checkpoint(str(tmp_path / "model.ckpt"), var)
    assert sorted(os.listdir(tmp_path)) == ["model.ckpt.data-00000-of-00001", "model.ckpt.index"]
    listed = T.list_variables(prefix + ".index")
    assert listed["pc_compare/dpdist_local/mapper_conv1/weights"] == (T.DT_FLOAT, (1, 2503, 1, 1024))
    got = T.load_checkpoint(prefix)
    assert sorted(got) == sorted(var)
    for n in var:
        assert got[n].dtype == np.float32 and got[n].shape == np.shape(var[n]) and np.array_equal(got[n], var[n]), n
    only = T.load_checkpoint(prefix, names={NAMES[1]})
    assert list(only) == [NAMES[1]]


def test_corruption_is_detected(tmp_path):
    var = {"a/weights": np.arange(6, dtype=np.float32).reshape(2, 3)}
    prefix = T.save_checkpoint(str(tmp_path / "m.ckpt"), var)
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[5] ^= 0x40
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="crc32c"):
        T.load_checkpoint(prefix)
    assert T.load_checkpoint(prefix, verify_crc=False)["a/weights"].shape == (2, 3)
    open(prefix + ".index", "ab").write(b"x")
    with pytest.raises(ValueError, match="magic"):
        T.load_checkpoint(prefix)


def test_reader_handles_prefix_compressed_multi_block_tables(tmp_path):
    """A hand-built index in the layout TF's table builder emits: shared key prefixes (restart interval > 1), two
    data blocks, offsets omitted when zero."""
    vals = {"scope/conv1/biases": np.float32([1, 2, 3]), "scope/conv1/weights": np.float32([[4, 5], [6, 7]]),
            "scope/conv2/biases": np.float32([8])}
    names = sorted(vals)
    data, entries, off = b"", [], 0
    for n in names:
        raw = vals[n].tobytes()
        shape = b"".join(T._field(2, 2, T._field(1, 0, d)) for d in vals[n].shape)
        e = T._field(1, 0, 1) + T._field(2, 2, shape) + (T._field(4, 0, off) if off else b"") + T._field(5, 0, len(raw)) + \
            T._field(6, 5, T._mask(T.crc32c(raw)))
        entries.append((n.encode(), e))
        data += raw
        off += len(raw)
    open(tmp_path / "m.ckpt.data-00000-of-00001", "wb").write(data)

    def block(kvs):            # one restart point, prefix compression against the previous key
        body, prev = bytearray(), b""
        for k, v in kvs:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
            body += T._put_varint(shared) + T._put_varint(len(k) - shared) + T._put_varint(len(v)) + k[shared:] + v
            prev = k
        body += struct.pack("<II", 0, 1)
        return bytes(body) + b"\x00" + struct.pack("<I", T._mask(T.crc32c(bytes(body) + b"\x00")))

    header = (b"", T._field(1, 0, 1) + T._field(3, 2, T._field(1, 0, 1)))
    b0, b1 = block([header] + entries[:2]), block(entries[2:])
    h0 = T._put_varint(0) + T._put_varint(len(b0) - 5)
    h1 = T._put_varint(len(b0)) + T._put_varint(len(b1) - 5)
    meta = block([])
    index = block([(b"scope/conv1/x", h0), (b"scope/conv3", h1)])
    moff = len(b0) + len(b1)
    footer = T._put_varint(moff) + T._put_varint(len(meta) - 5) + T._put_varint(moff + len(meta)) + T._put_varint(len(index) - 5)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC)
    open(tmp_path / "m.ckpt.index", "wb").write(b0 + b1 + meta + index + footer)
    got = T.load_checkpoint(str(tmp_path / "m.ckpt"))
    assert sorted(got) == names
    for n in names:
        assert np.array_equal(got[n], vals[n]) and got[n].shape == vals[n].shape
