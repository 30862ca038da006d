"""NSynth JSON -> TFRecord-of-paths files (reference make_tfrecord.py:7-47), without TensorFlow.

Run from the directory that holds the `nsynth*/examples.json` + `nsynth*/audio/*.wav` trees:
    python -m gansynth_b200.make_tfrecord
writes `nsynth_train.tfrecord` (80 %) and `nsynth_test.tfrecord` (20 %) of records
{path: bytes, pitch: int64, source: int64 (= instrument_source)} that dataset.nsynth_input_fn reads --
byte-compatible with the files the reference script writes (same framing and Example encoding)."""
import json
import pathlib
import random

from .tfrecord import TFRecordWriter, serialize_example


def collect_examples(root="."):
    """make_tfrecord.py:9-16."""
    examples = {}
    for filename in pathlib.Path(root).glob("nsynth*/*.json"):
        with open(filename) as file:
            loaded = json.load(file)
        for key, value in loaded.items():
            value.update(dict(path=str(filename.parent / "audio" / ("%s.wav" % key))))
        examples.update(loaded)
    return list(examples.items())


def write_tfrecord(path, examples):
    """make_tfrecord.py:24-47."""
    with TFRecordWriter(path) as writer:
        for _, value in examples:
            writer.write(serialize_example(dict(path=value["path"].encode(), pitch=int(value["pitch"]),
                                                source=int(value["instrument_source"]))))


def main(root=".", seed=None):
    examples = collect_examples(root)
    random.Random(seed).shuffle(examples)
    cut = int(len(examples) * 0.8)
    for name, part in (("nsynth_train", examples[:cut]), ("nsynth_test", examples[cut:])):
        write_tfrecord("%s.tfrecord" % name, part)
    return len(examples)


if __name__ == "__main__":
    print("%d examples written" % main())
