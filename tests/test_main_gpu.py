"""GPU: the user path end to end with the reference's command line (gan_synth_main.py:25-36, 91-130): TFRecord-of-paths
+ WAV files -> nsynth_input_fn -> GANSynth.train (checkpoint) -> --generate (samples/*.wav)."""
import glob

import numpy as np
import pytest
import torch
from scipy.io import wavfile

pytestmark = pytest.mark.gpu


def test_train_then_generate_from_the_command_line(cuda_store, tmp_path, monkeypatch):
    from gansynth_b200 import dataset, gan_synth_main
    from gansynth_b200.make_tfrecord import write_tfrecord
    import gansynth_b200.models as pmodels
    rng = np.random.default_rng(0)
    examples = []
    for i in range(8):
        path = str(tmp_path / ("clip%d.wav" % i))
        t = np.arange(64000) / 16000.0
        tone = 0.3 * np.sin(2 * np.pi * 110.0 * (1 + i) * t) * np.exp(-2.0 * t) + 0.01 * rng.standard_normal(64000)
        wavfile.write(path, 16000, (tone * 32767).astype(np.int16))
        examples.append((str(i), dict(path=path, pitch=30 + 5 * i, instrument_source=0)))
    write_tfrecord(str(tmp_path / "nsynth_train.tfrecord"), examples)
    monkeypatch.chdir(tmp_path)
    dataset.reset_pipelines()
    common = ["--filenames", str(tmp_path / "nsynth*.tfrecord"), "--batch_size", "4", "--model_dir", str(tmp_path / "model")]
    try:
        gan_synth_main.main(["--train", "--total_steps", "2", "--growing_steps", "4", "--num_epochs", "2"] + common)
        assert int(pmodels.get_or_create_global_step().value) == 2
        assert glob.glob(str(tmp_path / "model" / "model.ckpt-2.pt"))
        dataset.reset_pipelines()
        gan_synth_main.main(["--generate", "--growing_steps", "4"] + common)
    finally:
        dataset.reset_pipelines()
    files = sorted(glob.glob(str(tmp_path / "samples" / "*.wav")))
    assert len(files) == 8                                   # one epoch of 8 clips in batches of 4
    for f in files:
        rate, data = wavfile.read(f)
        assert rate == 16000 and data.shape == (64000,) and data.dtype == np.float32 and np.isfinite(data).all()
    assert torch.cuda.is_available()
