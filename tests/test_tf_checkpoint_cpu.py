"""CPU: TensorFlow-1 Saver (tensor bundle) files written and read without TensorFlow.  Format provenance is
"unpinned" (no TF-written fixture exists here, see gansynth_b200/tf_checkpoint.py): these are round trips and
structural known answers."""
import os
import struct

import numpy as np
import pytest

from gansynth_b200 import tf_checkpoint as tfc
from gansynth_b200 import tfrecord


def test_sorted_table_round_trip_and_structure(tmp_path):
    p = str(tmp_path / "t.index")
    items = [(b"", b"header")] + [(("var/%04d/weight" % i).encode(), os.urandom(1 + i % 50)) for i in range(500)]
    tfc.write_table(p, items, block_size=512)
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    assert tfc.read_table(p) == items
    # first data block: trailer = type 0 + masked crc32c(block + type)
    footer = raw[-48:]
    _, q = tfrecord._varint(footer, 0)
    _, q = tfrecord._varint(footer, q)
    ioff, q = tfrecord._varint(footer, q)
    isize, _ = tfrecord._varint(footer, q)
    k0, h0 = next(tfc._block_entries(raw[ioff:ioff + isize]))
    boff, q = tfrecord._varint(h0, 0)
    bsize, _ = tfrecord._varint(h0, q)
    assert boff == 0 and raw[bsize] == 0
    assert struct.unpack("<I", raw[bsize + 1:bsize + 5])[0] == tfrecord.masked_crc32c(raw[:bsize + 1])
    # prefix compression is in use (keys share "var/0") and a flipped byte is caught
    assert bsize < sum(len(k) + len(v) + 3 for k, v in items[:20])
    bad = bytearray(raw)
    bad[10] ^= 0x40
    open(p, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        tfc.read_table(p)
    with pytest.raises(ValueError):
        tfc.write_table(p, [(b"b", b""), (b"a", b"")])


def test_snappy_blocks_are_readable():
    # literal "abcdabcdabcd" as literal(4) + copy(len 8, offset 4): tag 0b000011_00, then copy-1 tag
    comp = bytes([12, (3 << 2) | 0]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert tfc._snappy_decompress(comp) == b"abcdabcdabcd"


def test_bundle_round_trip_names_shapes_dtypes(tmp_path):
    g = np.random.default_rng(0)
    tensors = {
        "generator/conv_block_2x16/dense/weight": g.standard_normal((512, 64)).astype(np.float32),
        "generator/conv_block_2x16/dense/bias": np.zeros(64, np.float32),
        "generator/weight": g.standard_normal((61, 256)).astype(np.float32),
        "discriminator/conv_block_2x16/conv/weight": g.standard_normal((3, 3, 5, 4)).astype(np.float32),
        "global_step": np.asarray(1234, np.int64),
        "beta2_power": np.asarray(0.99 ** 3, np.float32),
    }
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.save_bundle(prefix, tensors)
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(a.nbytes for a in tensors.values())
    got = tfc.load_bundle(prefix)
    assert set(got) == set(tensors)
    for n, a in tensors.items():
        assert got[n].dtype == a.dtype and got[n].shape == a.shape and np.array_equal(got[n], a)
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix and tfc.latest_checkpoint(str(tmp_path / "none")) is None
    # the header entry and one BundleEntryProto, field by field
    entries = dict(tfc.read_table(prefix + ".index"))
    assert entries[b""] == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    e = tfc._parse_entry(entries[b"generator/weight"])
    assert e["dtype"] == 1 and e["shape"] == [61, 256] and e["size"] == 61 * 256 * 4
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert e["crc"] == tfrecord.masked_crc32c(data[e["offset"]:e["offset"] + e["size"]])
    # corruption of the data shard is detected through the per-tensor checksum
    bad = bytearray(data)
    bad[e["offset"] + 5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(IOError):
        tfc.load_bundle(prefix)


def test_training_state_naming_round_trip():
    v = {"generator/a/weight": np.ones((2, 3), np.float32), "discriminator/b/bias": np.zeros(4, np.float32)}
    opt = {"generator": dict(m={"generator/a/weight": np.full((2, 3), 0.5, np.float32)},
                             v={"generator/a/weight": np.full((2, 3), 0.25, np.float32)}, t=7),
           "discriminator": dict(m={"discriminator/b/bias": np.ones(4, np.float32)},
                                 v={"discriminator/b/bias": np.ones(4, np.float32)}, t=9)}
    flat = tfc.join_training_state(v, 42, opt)
    assert {"generator/a/weight/Adam", "generator/a/weight/Adam_1", "beta1_power", "beta2_power", "beta1_power_1",
            "beta2_power_1", "global_step"} <= set(flat)
    assert abs(float(flat["beta2_power"]) - 0.99 ** 8) < 1e-7 and float(flat["beta1_power"]) == 0.0
    back = tfc.split_training_state(flat)
    assert back["global_step"] == 42 and back["optimizers"]["generator"]["t"] == 7
    assert back["optimizers"]["discriminator"]["t"] == 9 and set(back["variables"]) == set(v)
    assert np.array_equal(back["optimizers"]["generator"]["v"]["generator/a/weight"], opt["generator"]["v"]["generator/a/weight"])


def test_model_export_then_import_into_a_fresh_model(emu, tmp_path):
    """GANSynth.export_tf_checkpoint / import_tf_checkpoint on the CPU emulation backend: a fresh model (no variables
    yet) is populated from the checkpoint alone -- the embedding's shape gives num_labels and latent_dim."""
    import torch
    import gansynth_b200.models as M
    import gansynth_b200.networks as N
    import gansynth_b200.ops as ops
    from common import HYPER, SMALL
    pg = N.PGGAN(growing_level=0.3, **SMALL)
    pg._ensure_variables("generator", 256, 61)
    pg._ensure_variables("discriminator", 0, 61)
    model = M.GANSynth(pg.generator, pg.discriminator, None, None, {}, HYPER, device="cpu")
    model.global_step.value = 17
    want = {n: v.clone() for n, v in emu.state().items()}
    prefix = model.export_tf_checkpoint(str(tmp_path / "ckpt" / "model.ckpt-17"))
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    # a new store, a new model, nothing created yet
    store2 = ops.set_default_store(ops.VariableStore(device="cpu", seed=123))
    M.reset_global_step()
    pg2 = N.PGGAN(growing_level=0.3, **SMALL)
    model2 = M.GANSynth(pg2.generator, pg2.discriminator, None, None, {}, HYPER, device="cpu")
    assert not store2.vars
    assert model2.import_tf_checkpoint(str(tmp_path / "ckpt")) == prefix
    assert int(model2.global_step.value) == 17 and set(store2.vars) == set(want)
    for n, v in store2.state().items():
        assert torch.equal(v, want[n]), n


def test_checkpoint_state_file_keeps_older_paths(tmp_path):
    """save_bundle extends `<dir>/checkpoint` the way tf.train.Saver does: model_checkpoint_path is the newest prefix,
    all_model_checkpoint_paths lists every prefix written so far (an existing state file is not clobbered)."""
    import numpy as np
    from gansynth_b200 import tf_checkpoint as tfc
    for step in (1000, 2000, 3000):
        tfc.save_bundle(str(tmp_path / ("model.ckpt-%d" % step)), {"v": np.full((2,), step, np.float32)})
    lines = open(tmp_path / "checkpoint").read().splitlines()
    assert lines[0] == 'model_checkpoint_path: "model.ckpt-3000"'
    assert lines[1:] == ['all_model_checkpoint_paths: "model.ckpt-%d"' % s for s in (1000, 2000, 3000)]
    assert tfc.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-3000")
    assert float(tfc.load_bundle(tfc.latest_checkpoint(str(tmp_path)))["v"][0]) == 3000.0
