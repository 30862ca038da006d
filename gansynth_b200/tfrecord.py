"""TFRecord files of tf.train.Example records without TensorFlow (reference make_tfrecord.py:27-47 writes
them, dataset.py:18-26,46-48 reads them).

File format (tensorflow/core/lib/io/record_writer.cc): per record
    uint64 length | uint32 masked_crc32c(length bytes) | data[length] | uint32 masked_crc32c(data)
with masked(c) = rotr(c, 15) + 0xa282ead8 (mod 2^32), all little-endian.  The CRC runs in the native
library (gs_crc32c, csrc/io.cu).

Record payload: a serialised tf.train.Example (tensorflow/core/example/{example,feature}.proto)
    Example  { Features features = 1; }
    Features { map<string, Feature> feature = 1; }        // map entry: key = 1, value = 2
    Feature  { oneof { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
    *List    { repeated value = 1; }                      // floats / int64s packed or not
Only the wire subset those messages use is implemented (varint, 64-bit, length-delimited, 32-bit).
"""
import ctypes
import struct

from . import _lib

_MASK_DELTA = 0xA282EAD8


def crc32c(data):
    out = ctypes.c_uint(0)
    buf = bytes(data)
    _lib.host_call("gs_crc32c", buf, len(buf), ctypes.byref(out))
    return out.value


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ framing
def read_records(path, verify=True):
    """Yields the payload bytes of every record (tf.io.tf_record_iterator)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError("%s: truncated record header" % path)
            (length,), (len_crc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and masked_crc32c(head[:8]) != len_crc:
                raise IOError("%s: corrupted record length" % path)
            body = f.read(length + 4)
            if len(body) < length + 4:
                raise IOError("%s: truncated record" % path)
            data = body[:length]
            if verify and masked_crc32c(data) != struct.unpack("<I", body[length:])[0]:
                raise IOError("%s: corrupted record data" % path)
            yield data


class TFRecordWriter(object):
    """tf.io.TFRecordWriter(path): ``with TFRecordWriter(p) as w: w.write(record_bytes)``."""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, record):
        record = bytes(record)
        head = struct.pack("<Q", len(record))
        self._f.write(head + struct.pack("<I", masked_crc32c(head)) + record + struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------------ protobuf wire
def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    v &= (1 << 64) - 1          # negative int64 -> 10-byte two's complement, as protobuf does
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf):
    """Yields (field number, wire type, value) of one message; value is int or bytes."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf):
    for num, wt, v in _fields(buf):
        if num == 1:                               # BytesList
            return [bytes(x) for n, _, x in _fields(v) if n == 1]
        if num == 2:                               # FloatList
            out = []
            for n, w, x in _fields(v):
                if n == 1:
                    out.extend(struct.unpack("<%df" % (len(x) // 4), x) if w == 2 else struct.unpack("<f", x))
            return out
        if num == 3:                               # Int64List
            out = []
            for n, w, x in _fields(v):
                if n != 1:
                    continue
                if w == 2:                         # packed
                    p = 0
                    while p < len(x):
                        i, p = _varint(x, p)
                        out.append(_signed64(i))
                else:
                    out.append(_signed64(x))
            return out
    return []


def parse_example(record):
    """Serialised tf.train.Example -> {feature name: list of bytes / float / int}."""
    out = {}
    for num, _, feats in _fields(record):
        if num != 1:
            continue
        for n, _, entry in _fields(feats):
            if n != 1:
                continue
            key, value = None, b""
            for fn, _, x in _fields(entry):
                if fn == 1:
                    key = bytes(x).decode("utf-8")
                elif fn == 2:
                    value = x
            if key is not None:
                out[key] = _parse_feature(value)
    return out


def _ld(num, payload):
    return _put_varint((num << 3) | 2) + _put_varint(len(payload)) + payload


def serialize_example(features):
    """{name: bytes | str | int | float | list of one of those} -> serialised tf.train.Example
    (tf.train.Example(features=tf.train.Features(feature=...)).SerializeToString(); map entries are
    written in sorted key order, which is what deterministic serialisation produces)."""
    entries = b""
    for key in sorted(features):
        vals = features[key]
        if not isinstance(vals, (list, tuple)):
            vals = [vals]
        first = vals[0] if vals else b""
        if isinstance(first, (bytes, str)):
            body = b"".join(_ld(1, v.encode("utf-8") if isinstance(v, str) else bytes(v)) for v in vals)
            feat = _ld(1, body)
        elif isinstance(first, float):
            feat = _ld(2, _ld(1, struct.pack("<%df" % len(vals), *vals)))
        else:
            feat = _ld(3, _ld(1, b"".join(_put_varint(int(v)) for v in vals)))
        entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)
