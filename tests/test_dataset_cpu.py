"""CPU: input-pipeline host logic (reference dataset.py:12-91, make_tfrecord.py:27-47): CRC-32C known answers,
TFRecord framing, tf.train.Example wire format, native WAV decoding, and the shuffle / repeat / filter / batch
stage order.  The float conversion is the device half (tests/test_dataset_gpu.py)."""
import ctypes
import os
import struct

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from gansynth_b200 import _lib, dataset, tfrecord
from gansynth_b200.make_tfrecord import write_tfrecord


def test_crc32c_known_answers():
    """RFC 3720 B.4 test vectors + the classic check value."""
    assert tfrecord.crc32c(b"123456789") == 0xE3069283
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA
    assert tfrecord.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfrecord.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tfrecord.crc32c(b"") == 0


def test_example_wire_format_known_bytes():
    """Hand-assembled protobuf: Example{features{feature{"pitch": int64_list{[60]}}}} (Int64List is packed)."""
    want = bytes([0x0A, 0x10, 0x0A, 0x0E, 0x0A, 0x05]) + b"pitch" + bytes([0x12, 0x05, 0x1A, 0x03, 0x0A, 0x01, 0x3C])
    assert tfrecord.serialize_example(dict(pitch=60)) == want
    assert tfrecord.parse_example(want) == {"pitch": [60]}
    # unpacked repeated int64 (what older writers emit) and a negative value parse too
    unpacked = bytes([0x0A, 0x11, 0x0A, 0x0F, 0x0A, 0x05]) + b"pitch" + bytes([0x12, 0x06, 0x1A, 0x04, 0x08, 0x3C, 0x08, 0x3D])
    assert tfrecord.parse_example(unpacked) == {"pitch": [60, 61]}
    rec = tfrecord.serialize_example(dict(path=b"nsynth-train/audio/x.wav", pitch=-3, source=2, gain=[0.5, 1.5]))
    assert tfrecord.parse_example(rec) == {"path": [b"nsynth-train/audio/x.wav"], "pitch": [-3], "source": [2],
                                           "gain": [0.5, 1.5]}


def test_tfrecord_framing_round_trip_and_corruption(tmp_path):
    p = str(tmp_path / "a.tfrecord")
    recs = [b"", b"x", os.urandom(1000), tfrecord.serialize_example(dict(pitch=24))]
    with tfrecord.TFRecordWriter(p) as w:
        for r in recs:
            w.write(r)
    raw = open(p, "rb").read()
    assert len(raw) == sum(16 + len(r) for r in recs)
    # framing: uint64 length, masked crc of the length bytes
    (n0,), (c0,) = struct.unpack("<Q", raw[:8]), struct.unpack("<I", raw[8:12])
    crc = tfrecord.crc32c(raw[:8])
    assert n0 == 0 and c0 == (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    assert list(tfrecord.read_records(p)) == recs
    bad = bytearray(raw)
    bad[-10] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        list(tfrecord.read_records(p))
    assert len(list(tfrecord.read_records(p, verify=False))) == 4


def _write_wav(path, data, rate=16000):
    wavfile.write(path, rate, data)


def test_wav_decode_matches_decode_wav_semantics(tmp_path):
    g = np.random.default_rng(0)
    short = g.integers(-32768, 32767, 1000, dtype=np.int16)
    long_ = g.integers(-32768, 32767, 5000, dtype=np.int16)
    stereo = g.integers(-32768, 32767, (3000, 2), dtype=np.int16)
    for name, d in (("short", short), ("long", long_), ("stereo", stereo)):
        _write_wav(str(tmp_path / (name + ".wav")), d)
    # a file with an extra chunk between fmt and data (odd length -> pad byte)
    raw = open(str(tmp_path / "short.wav"), "rb").read()
    extra = raw[:36] + b"LIST" + struct.pack("<I", 5) + b"abcde\x00" + raw[36:]
    extra = extra[:4] + struct.pack("<I", len(extra) - 8) + extra[8:]
    open(str(tmp_path / "extra.wav"), "wb").write(extra)
    paths = [str(tmp_path / (n + ".wav")) for n in ("short", "long", "stereo", "extra")]
    out = dataset.decode_wav_files(paths, desired_samples=4000, threads=3).numpy()
    assert out.shape == (4, 4000) and out.dtype == np.int16
    assert np.array_equal(out[0, :1000], short) and not out[0, 1000:].any()       # zero-padded at the end
    assert np.array_equal(out[1], long_[:4000])                                    # cropped
    assert np.array_equal(out[2, :3000], stereo[:, 0]) and not out[2, 3000:].any() # channel 0
    assert np.array_equal(out[3], out[0])
    # single-buffer entry point reports rate and length
    dst = (ctypes.c_short * 10)()
    rate, n = ctypes.c_int(0), ctypes.c_int(0)
    _lib.host_call("gs_wav_decode_pcm16", raw, len(raw), dst, 10, ctypes.byref(rate), ctypes.byref(n))
    assert rate.value == 16000 and n.value == 1000 and list(dst) == list(short[:10])


def test_wav_decode_rejects_what_decode_wav_rejects(tmp_path):
    _write_wav(str(tmp_path / "f32.wav"), np.zeros(100, np.float32))
    _write_wav(str(tmp_path / "u8.wav"), np.zeros(100, np.uint8))
    open(str(tmp_path / "junk.wav"), "wb").write(b"not a wav file at all")
    for name in ("f32.wav", "u8.wav", "junk.wav", "missing.wav"):
        with pytest.raises(_lib.GansynthLibraryError):
            dataset.decode_wav_files([str(tmp_path / name)], desired_samples=64)


def _make_dataset(tmp_path, n=40, length=300):
    g = np.random.default_rng(1)
    examples = []
    for i in range(n):
        p = str(tmp_path / ("clip%03d.wav" % i))
        _write_wav(p, np.full(length + (i % 3) * 50, i + 1, np.int16))   # sample value identifies the clip
        examples.append((str(i), dict(path=p, pitch=int(20 + (i * 7) % 70), instrument_source=int(i % 3))))
    rec = str(tmp_path / "nsynth_test.tfrecord")
    write_tfrecord(rec, examples)
    return rec, examples


def test_pipeline_filter_batch_epochs(tmp_path):
    rec, examples = _make_dataset(tmp_path)
    pitches, sources = range(24, 85), [0]
    keep = [i for i, (_, v) in enumerate(examples) if 24 <= v["pitch"] <= 84 and v["instrument_source"] == 0]
    assert 4 < len(keep) < len(examples)
    pipe = dataset.NSynthPipeline([rec], batch_size=4, num_epochs=2, shuffle=False, pitches=pitches, sources=sources,
                                  device="cpu", waveform_length=320)
    batches = list(pipe)
    assert len(batches) == (2 * len(keep)) // 4                              # repeat before batch, remainder dropped
    ids = np.concatenate([b[0][:, 0].numpy() for b in batches]) - 1
    assert list(ids) == (keep + keep)[:len(ids)]                             # file order, epochs concatenated
    for wave, lab in batches:
        assert wave.dtype == torch.int16 and wave.shape == (4, 320) and lab.shape == (4, 61)
        for r in range(4):
            i = int(wave[r, 0]) - 1
            n = 300 + (i % 3) * 50
            assert bool((wave[r, :min(n, 320)] == i + 1).all()) and not bool(wave[r, min(n, 320):].any())
            assert int(lab[r].argmax()) == examples[i][1]["pitch"] - 24 and float(lab[r].sum()) == 1.0
    with pytest.raises(StopIteration):
        next(pipe)


def test_pipeline_shuffle_is_a_seeded_permutation_per_epoch(tmp_path):
    rec, examples = _make_dataset(tmp_path)
    def ids(seed, buffer_size=None):
        pipe = dataset.NSynthPipeline([rec], batch_size=5, num_epochs=2, shuffle=True, buffer_size=buffer_size,
                                      device="cpu", seed=seed, waveform_length=64)
        return np.concatenate([b[0][:, 0].numpy() for b in pipe]) - 1
    a, b, c = ids(0), ids(0), ids(1)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert sorted(a[:40]) == list(range(40)) and sorted(a[40:]) == list(range(40))   # each epoch = one permutation
    assert not np.array_equal(a[:40], a[40:])                                        # reshuffle_each_iteration
    small = ids(0, buffer_size=4)
    assert sorted(small[:40]) == list(range(40))
    assert max(np.arange(40) - small[:40]) <= 3 + 40 and all(small[i] <= i + 4 for i in range(40))  # bounded look-ahead


def test_main_parser_matches_reference_flags():
    from gansynth_b200.gan_synth_main import build_parser
    a = build_parser().parse_args([])
    assert (a.model_dir, a.filenames, a.batch_size, a.num_epochs, a.total_steps, a.growing_steps, a.classifier) == \
        ("gan_synth_model", "nsynth*.tfrecord", 8, None, 1000000, 1000000, "pitch_classifier.pb")
    assert not (a.train or a.evaluate or a.generate)


def test_pipeline_close_releases_the_prefetch_thread(tmp_path):
    rec, _ = _make_dataset(tmp_path)
    pipe = dataset.NSynthPipeline([rec], batch_size=2, num_epochs=None, shuffle=True, device="cpu", waveform_length=64)
    next(pipe)
    pipe.close()
    pipe._thread.join(timeout=5.0)
    assert not pipe._thread.is_alive()
    with pytest.raises(StopIteration):
        next(pipe)


def test_tfrecord_framing_against_tensorboard_stub(tmp_path):
    """Third-party pin of the TFRecord container: TensorBoard ships its own pure-Python TFRecord reader / writer and
    masked CRC-32C (tensorboard.compat.tensorflow_stub, written against TensorFlow's record_writer.cc)."""
    pywrap = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    from tensorboard.summary.writer.record_writer import RecordWriter
    recs = [b"", b"a", os.urandom(300), tfrecord.serialize_example(dict(path=b"x.wav", pitch=60, source=0))]
    for r in recs:
        assert tfrecord.masked_crc32c(r) == pywrap.masked_crc32c(r) and tfrecord.crc32c(r) == pywrap.crc32c(r)
    mine = str(tmp_path / "mine.tfrecord")
    with tfrecord.TFRecordWriter(mine) as w:
        for r in recs:
            w.write(r)
    reader = pywrap.PyRecordReader_New(mine)          # TensorBoard reads (and CRC-checks) this package's file
    got = []
    while True:
        try:
            reader.GetNext()
        except Exception:
            break
        got.append(reader.record())
    assert got == recs
    theirs = str(tmp_path / "theirs.tfrecord")
    with open(theirs, "wb") as f:                     # ... and this package reads TensorBoard's
        w = RecordWriter(f)
        for r in recs:
            w.write(r)
        w.flush()
    assert list(tfrecord.read_records(theirs)) == recs
    assert open(theirs, "rb").read() == open(mine, "rb").read()
