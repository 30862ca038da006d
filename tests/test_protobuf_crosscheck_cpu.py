"""CPU: the hand-written protobuf wire code (tfrecord.py, tf_checkpoint.py) against Google's protobuf runtime.  The
message schemas are declared here from the published .proto files (tensorflow/core/example/{example,feature}.proto,
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/framework/{tensor_shape,versions}.proto: field numbers
and types as TensorFlow ships them); the encoder / decoder being compared against is protobuf's own."""
import struct

import numpy as np
import pytest

descriptor_pb2 = pytest.importorskip("google.protobuf.descriptor_pb2")
from google.protobuf import descriptor_pool, message_factory  # noqa: E402

from gansynth_b200 import tf_checkpoint as tfc  # noqa: E402
from gansynth_b200 import tfrecord  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, packed=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    return f


def _classes():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "gs_crosscheck.proto", "gsx", "proto3"
    m = fd.message_type.add(); m.name = "BytesList"
    _field(m, "value", 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    m = fd.message_type.add(); m.name = "FloatList"
    _field(m, "value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Int64List"
    _field(m, "value", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Feature"
    m.oneof_decl.add().name = "kind"
    for name, num, t in (("bytes_list", 1, ".gsx.BytesList"), ("float_list", 2, ".gsx.FloatList"), ("int64_list", 3, ".gsx.Int64List")):
        _field(m, name, num, F.TYPE_MESSAGE, type_name=t).oneof_index = 0
    m = fd.message_type.add(); m.name = "Features"
    e = m.nested_type.add(); e.name = "FeatureEntry"; e.options.map_entry = True
    _field(e, "key", 1, F.TYPE_STRING)
    _field(e, "value", 2, F.TYPE_MESSAGE, type_name=".gsx.Feature")
    _field(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".gsx.Features.FeatureEntry")
    m = fd.message_type.add(); m.name = "Example"
    _field(m, "features", 1, F.TYPE_MESSAGE, type_name=".gsx.Features")
    # tensor bundle
    m = fd.message_type.add(); m.name = "TensorShapeProto"
    d = m.nested_type.add(); d.name = "Dim"
    _field(d, "size", 1, F.TYPE_INT64)
    _field(d, "name", 2, F.TYPE_STRING)
    _field(m, "dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".gsx.TensorShapeProto.Dim")
    _field(m, "unknown_rank", 3, F.TYPE_BOOL)
    m = fd.message_type.add(); m.name = "VersionDef"
    _field(m, "producer", 1, F.TYPE_INT32)
    _field(m, "min_consumer", 2, F.TYPE_INT32)
    m = fd.message_type.add(); m.name = "BundleHeaderProto"
    _field(m, "num_shards", 1, F.TYPE_INT32)
    _field(m, "endianness", 2, F.TYPE_INT32)          # enum Endianness { LITTLE = 0; BIG = 1; }
    _field(m, "version", 3, F.TYPE_MESSAGE, type_name=".gsx.VersionDef")
    m = fd.message_type.add(); m.name = "BundleEntryProto"
    _field(m, "dtype", 1, F.TYPE_INT32)               # enum DataType
    _field(m, "shape", 2, F.TYPE_MESSAGE, type_name=".gsx.TensorShapeProto")
    _field(m, "shard_id", 3, F.TYPE_INT32)
    _field(m, "offset", 4, F.TYPE_INT64)
    _field(m, "size", 5, F.TYPE_INT64)
    _field(m, "crc32c", 6, F.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return {n: get(pool.FindMessageTypeByName("gsx." + n)) for n in ("Example", "BundleHeaderProto", "BundleEntryProto")}


def test_example_encoding_matches_protobuf():
    cls = _classes()
    ex = cls["Example"]()
    ex.features.feature["path"].bytes_list.value.append(b"nsynth-train/audio/guitar_acoustic_001-060-100.wav")
    ex.features.feature["pitch"].int64_list.value.append(60)
    ex.features.feature["source"].int64_list.value.append(0)
    want = ex.SerializeToString(deterministic=True)
    mine = tfrecord.serialize_example(dict(path=b"nsynth-train/audio/guitar_acoustic_001-060-100.wav", pitch=60, source=0))
    assert mine == want
    # protobuf reads what this package writes (negative int64, several floats) ...
    rec = tfrecord.serialize_example(dict(a=[-5, 7, 1 << 40], f=[0.25, -1.5], s=[b"x", b"yz"]))
    back = cls["Example"].FromString(rec)
    assert list(back.features.feature["a"].int64_list.value) == [-5, 7, 1 << 40]
    assert list(back.features.feature["f"].float_list.value) == [0.25, -1.5]
    assert list(back.features.feature["s"].bytes_list.value) == [b"x", b"yz"]
    # ... and this package reads what protobuf writes
    assert tfrecord.parse_example(want) == {"path": [b"nsynth-train/audio/guitar_acoustic_001-060-100.wav"],
                                            "pitch": [60], "source": [0]}


def test_bundle_protos_match_protobuf():
    cls = _classes()
    arr = np.zeros((3, 3, 32, 64), np.float32)
    mine = tfc._entry_proto(arr, 1234567, 0xDEADBEEF)
    e = cls["BundleEntryProto"].FromString(mine)
    assert (e.dtype, [d.size for d in e.shape.dim], e.shard_id, e.offset, e.size, e.crc32c) == \
        (1, [3, 3, 32, 64], 0, 1234567, arr.nbytes, 0xDEADBEEF)
    ref = cls["BundleEntryProto"](dtype=1, offset=1234567, size=arr.nbytes, crc32c=0xDEADBEEF)
    for s in arr.shape:
        ref.shape.dim.add().size = s
    assert ref.SerializeToString(deterministic=True) == mine
    parsed = tfc._parse_entry(ref.SerializeToString(deterministic=True))
    assert parsed["dtype"] == 1 and parsed["shape"] == [3, 3, 32, 64] and parsed["offset"] == 1234567
    assert parsed["size"] == arr.nbytes and parsed["crc"] == 0xDEADBEEF
    scalar = tfc._entry_proto(np.asarray(7, np.int64), 0, 1)
    s = cls["BundleEntryProto"].FromString(scalar)
    assert s.dtype == 9 and len(s.shape.dim) == 0 and s.size == 8 and s.offset == 0
    header = cls["BundleHeaderProto"](num_shards=1)
    header.version.producer = 1
    assert header.SerializeToString(deterministic=True) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
