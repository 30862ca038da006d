"""GPU: device half of the input pipeline (reference dataset.py:28-40): int16 -> float32 / 32768 bit-exact, and
nsynth_input_fn end to end (TFRecord -> native WAV decode -> pinned int16 -> device floats + one-hot labels)."""
import functools

import numpy as np
import pytest
import torch
from scipy.io import wavfile

pytestmark = pytest.mark.gpu


def test_pcm16_to_float_bit_exact():
    from gansynth_b200 import functional as F
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 8, 9, 64000, 8 * 64000 + 3):
        pcm = torch.randint(-32768, 32768, (n,), generator=g, dtype=torch.int32).to(torch.int16)
        pcm[:2] = torch.tensor([-32768, 32767], dtype=torch.int16)[:min(2, n)]
        got = F.K.pcm16_to_float(pcm.cuda())
        assert torch.equal(got.cpu(), pcm.to(torch.float32) / 32768.0)
    # an unaligned view takes the scalar path
    base = torch.randint(-32768, 32768, (1001,), generator=g, dtype=torch.int32).to(torch.int16).cuda()
    view = base[1:].contiguous()
    assert torch.equal(F.K.pcm16_to_float(view).cpu(), view.cpu().to(torch.float32) / 32768.0)


def test_nsynth_input_fn_end_to_end(tmp_path):
    from gansynth_b200 import dataset
    from gansynth_b200.make_tfrecord import write_tfrecord
    g = np.random.default_rng(0)
    clips, examples = [], []
    for i in range(10):
        d = g.integers(-32768, 32767, 64000 if i % 2 else 50000, dtype=np.int16)
        p = str(tmp_path / ("c%d.wav" % i))
        wavfile.write(p, 16000, d)
        clips.append(d)
        examples.append((str(i), dict(path=p, pitch=24 + 6 * i, instrument_source=0 if i != 3 else 1)))
    rec = str(tmp_path / "nsynth_x.tfrecord")
    write_tfrecord(rec, examples)
    dataset.reset_pipelines()
    fn = functools.partial(dataset.nsynth_input_fn, filenames=[rec], batch_size=4, num_epochs=1, shuffle=False,
                           pitches=range(24, 85), sources=[0])
    keep = [i for i in range(10) if i != 3]
    seen = 0
    for _ in range(2):
        wave, lab = fn()
        assert wave.is_cuda and wave.dtype == torch.float32 and wave.shape == (4, 64000) and lab.shape == (4, 61)
        for r in range(4):
            i = keep[seen + r]
            want = np.zeros(64000, np.float32)
            want[:len(clips[i])] = clips[i].astype(np.float32) / 32768.0
            assert np.array_equal(wave[r].cpu().numpy(), want)
            assert int(lab[r].argmax()) == 6 * i and float(lab[r].sum()) == 1.0
        seen += 4
    with pytest.raises(StopIteration):
        fn()
    dataset.reset_pipelines()
