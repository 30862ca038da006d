"""TensorFlow-1 checkpoints (tf.train.Saver V2 "tensor bundle") without TensorFlow: read AND write
`<prefix>.index` + `<prefix>.data-00000-of-00001` (+ the `checkpoint` state file), so weights trained with the
reference (models.py:120-130: SingularMonitoredSession(checkpoint_dir=model_dir) + CheckpointSaverHook) can be
loaded by variable name and checkpoints written here can be restored by the reference.

**Format provenance: unpinned.**  TensorFlow is not installed and the reference ships no checkpoint, so the
format below is restated from TensorFlow's sources as the author remembers them (tensorflow/core/util/
tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/lib/io/{table_builder,format,block}.cc,
tensorflow/core/protobuf/tensor_bundle.proto); the tests are round trips plus structural known answers
(footer magic, block trailer CRCs, masked CRC-32C of the tensor bytes), not a TF-written fixture.

.index -- a LevelDB-style sorted string table:
    [data block]* [metaindex block] [index block] [footer: metaindex handle, index handle, pad to 40 B, magic
    0xdb4775248b80fb57 LE];  every block is followed by a 5-byte trailer (compression type: 0 none, 1 snappy;
    masked CRC-32C of block + type);  block = prefix-compressed entries (varint shared, varint non_shared,
    varint value_len, key suffix, value) + uint32 restart offsets + uint32 restart count;  handles are
    (varint offset, varint size).  Key "" holds BundleHeaderProto {num_shards=1, endianness=0, version{producer=1}};
    every other key is a variable name with BundleEntryProto {dtype=1, shape=2 {dim=2 {size=1}}, shard_id=3,
    offset=4, size=5, crc32c=6 fixed32 (masked)}.
.data-00000-of-00001 -- the tensors' little-endian bytes back to back at those offsets.
"""
import os
import struct

import numpy as np

from . import tfrecord
from .tfrecord import _fields, _put_varint, _varint

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 6: np.int8, 9: np.int64, 10: np.bool_}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


def _masked_crc(data):
    return tfrecord.masked_crc32c(data)


# ------------------------------------------------------------------------------------------------ snappy (read only)
def _snappy_decompress(buf):
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
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
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):                             # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise IOError("snappy: length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------------ table
def _read_block(data, offset, size, verify):
    raw, ctype = data[offset:offset + size], data[offset + size]
    if verify:
        (crc,) = struct.unpack("<I", data[offset + size + 1:offset + size + 5])
        if _masked_crc(data[offset:offset + size + 1]) != crc:
            raise IOError("table block at %d: checksum mismatch" % offset)
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise IOError("table block at %d: unknown compression %d" % (offset, ctype))
    return raw


def _block_entries(raw):
    (num_restarts,) = struct.unpack("<I", raw[-4:])
    limit = len(raw) - 4 * (num_restarts + 1)
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(raw, pos)
        non_shared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        yield key, raw[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a sorted string table, in key order."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != _MAGIC:
        raise IOError("%s: not a sorted string table (bad magic)" % path)
    footer = data[-48:]
    _, p = _varint(footer, 0)
    _, p = _varint(footer, p)                           # metaindex handle (unused)
    ioff, p = _varint(footer, p)
    isize, p = _varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _varint(handle, 0)
        bsize, _ = _varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


def _build_block(items, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    return bytes(body) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def write_table(path, items, block_size=4096):
    """Writes (key, value) pairs (keys strictly increasing) as an uncompressed sorted string table."""
    out, index, block, nbytes = bytearray(), [], [], 0

    def emit(entries):
        raw = _build_block(entries)
        handle = _put_varint(len(out)) + _put_varint(len(raw))
        out.extend(raw + b"\x00" + struct.pack("<I", _masked_crc(raw + b"\x00")))
        return handle

    def flush():
        nonlocal block, nbytes
        if block:
            index.append((block[-1][0], emit(block)))   # the last key is a valid separator (>= every key in the block)
            block, nbytes = [], 0

    prev = None
    for k, v in items:
        if prev is not None and not prev < k:
            raise ValueError("table keys must be strictly increasing")
        prev = k
        block.append((k, v))
        nbytes += len(k) + len(v)
        if nbytes >= block_size:
            flush()
    flush()
    meta = emit([])
    idx = emit(index)
    footer = meta + idx
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------------------------------------ bundle
def _ld(num, payload):
    return _put_varint((num << 3) | 2) + _put_varint(len(payload)) + payload


def _vi(num, value):
    return _put_varint(num << 3) + _put_varint(value)


def _entry_proto(arr, offset, crc):
    shape = b"".join(_ld(2, _vi(1, int(d))) for d in arr.shape)
    msg = _vi(1, _DTYPE_ENUM[arr.dtype]) + _ld(2, shape)
    if offset:
        msg += _vi(4, offset)
    msg += _vi(5, arr.nbytes) + _put_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return msg


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, crc=None, sliced=False)
    for num, wt, v in _fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            e["shape"] = [next((x for n, _, x in _fields(d) if n == 1), 0) for n2, _, d in _fields(v) if n2 == 2]
        elif num == 3:
            e["shard"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            (e["crc"],) = struct.unpack("<I", v)
        elif num == 7:
            e["sliced"] = True
    return e


def save_bundle(prefix, tensors):
    """{name: array} -> `<prefix>.index`, `<prefix>.data-00000-of-00001` and the `checkpoint` state file."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _vi(1, 1) + _ld(3, _vi(1, 1))              # num_shards = 1, (endianness = LITTLE = 0), version.producer = 1
    items, offset = [(b"", header)], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda n: n.encode("utf-8")):
            arr = np.asarray(tensors[name])
            if not arr.flags.c_contiguous:              # (ascontiguousarray would turn a scalar into shape [1])
                arr = np.ascontiguousarray(arr)
            if arr.dtype not in _DTYPE_ENUM:
                raise TypeError("%s: unsupported dtype %s" % (name, arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"), _entry_proto(arr, offset, _masked_crc(raw))))
            offset += len(raw)
    write_table(prefix + ".index", items)
    # CheckpointState text proto, as tf.train.Saver maintains it: the new prefix becomes model_checkpoint_path and is
    # appended to the all_model_checkpoint_paths already listed (an existing state file is extended, not overwritten)
    base = os.path.basename(prefix)
    state_path = os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint")
    older = []
    if os.path.exists(state_path):
        for line in open(state_path):
            if line.startswith("all_model_checkpoint_paths:"):
                name = line.split(":", 1)[1].strip().strip('"')
                if name != base and name not in older:
                    older.append(name)
    tmp = state_path + ".tmp"
    with open(tmp, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % base)
        for name in older + [base]:
            f.write('all_model_checkpoint_paths: "%s"\n' % name)
    os.replace(tmp, state_path)


def load_bundle(prefix, verify=True):
    """`<prefix>.index` + data shards -> {name: numpy array}."""
    entries = read_table(prefix + ".index", verify=verify)
    num_shards = 1
    out, shards = {}, {}
    for key, value in entries:
        if key == b"":
            for num, _, v in _fields(value):
                if num == 1:
                    num_shards = v
                elif num == 2 and v != 0:
                    raise IOError("%s: big-endian bundles are not supported" % prefix)
            continue
        e = _parse_entry(value)
        if e["sliced"]:
            raise IOError("%s: partitioned variable %s is not supported" % (prefix, key.decode()))
        if e["dtype"] not in _DTYPES:
            raise IOError("%s: %s has unsupported dtype enum %d" % (prefix, key.decode(), e["dtype"]))
        if e["shard"] not in shards:
            shards[e["shard"]] = open("%s.data-%05d-of-%05d" % (prefix, e["shard"], num_shards), "rb")
        f = shards[e["shard"]]
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise IOError("%s: %s is truncated" % (prefix, key.decode()))
        if verify and e["crc"] is not None and _masked_crc(raw) != e["crc"]:
            raise IOError("%s: %s fails its checksum" % (prefix, key.decode()))
        out[key.decode("utf-8")] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")).reshape(e["shape"]).copy()
    for f in shards.values():
        f.close()
    return out


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: the prefix named by `<model_dir>/checkpoint`, or None."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    for line in open(state):
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            return name if os.path.isabs(name) else os.path.join(model_dir, name)
    return None


# ------------------------------------------------------------------------------------------------ GANSynth naming
def split_training_state(tensors, beta2=(0.99, 0.99)):
    """Checkpoint of the reference graph (models.py:67-89) -> dict(variables, global_step, optimizers).
    Adam slots are `<var>/Adam` (m) and `<var>/Adam_1` (v); the generator's optimiser is applied first and owns
    `beta1_power` / `beta2_power`, the discriminator's owns `beta1_power_1` / `beta2_power_1`; the step count of
    each is recovered from beta2_power = beta2 ** t."""
    variables, m, v = {}, {}, {}
    for name, arr in tensors.items():
        if name.endswith("/Adam"):
            m[name[:-len("/Adam")]] = arr
        elif name.endswith("/Adam_1"):
            v[name[:-len("/Adam_1")]] = arr
        elif name.startswith(("generator/", "discriminator/")):
            variables[name] = arr
    opt = {}
    for scope, suffix, b2 in (("generator", "", beta2[0]), ("discriminator", "_1", beta2[1])):
        t = None
        if "beta2_power" + suffix in tensors:
            power = float(tensors["beta2_power" + suffix])
            # the variable holds beta2 ** (t + 1); it underflows to 0 after ~9000 steps at beta2 = 0.99, and both
            # optimisers step once per global step (models.py:189-192), so the global step stands in then
            if 0.0 < power < 1.0:
                t = int(round(np.log(power) / np.log(b2))) - 1
            else:
                t = int(tensors["global_step"]) if "global_step" in tensors else 0
        opt[scope] = dict(m={n: a for n, a in m.items() if n.startswith(scope + "/")},
                          v={n: a for n, a in v.items() if n.startswith(scope + "/")}, t=t)
    step = int(tensors["global_step"]) if "global_step" in tensors else None
    return dict(variables=variables, global_step=step, optimizers=opt)


def join_training_state(variables, global_step, optimizers, beta1=(0.0, 0.0), beta2=(0.99, 0.99)):
    """Inverse of split_training_state: the tensor names a tf.train.Saver of the reference graph writes."""
    out = {n: np.asarray(a, np.float32) for n, a in variables.items()}
    out["global_step"] = np.asarray(global_step, np.int64)
    for (scope, suffix), b1, b2 in zip((("generator", ""), ("discriminator", "_1")), beta1, beta2):
        st = optimizers.get(scope)
        if st is None:
            continue
        for n, a in st["m"].items():
            out[n + "/Adam"] = np.asarray(a, np.float32)
        for n, a in st["v"].items():
            out[n + "/Adam_1"] = np.asarray(a, np.float32)
        out["beta1_power" + suffix] = np.asarray(b1 ** (st["t"] + 1), np.float32)
        out["beta2_power" + suffix] = np.asarray(b2 ** (st["t"] + 1), np.float32)
    return out
