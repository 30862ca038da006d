"""CPU: the reference's own `gan_synth_main.py`, byte for byte, runs on this package through the `tensorflow` stand-in
and the module aliases of `gansynth_b200/compat/` (north_star: "so gan_synth_main.py drops in unchanged").  The file is
copied from /root/reference into a temporary directory at test time (it must sit apart from the reference's sibling
modules, which Python would otherwise import first); skipped where the reference tree does not exist."""
import os
import shutil
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/gan_synth_main.py"

pytestmark = pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="reference tree not present")


def _run(tmp_path, *flags):
    shutil.copy(REF_MAIN, tmp_path / "gan_synth_main.py")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "gansynth_b200", "compat"), ROOT]))
    return subprocess.run([sys.executable, "gan_synth_main.py"] + list(flags), cwd=tmp_path, env=env, capture_output=True,
                          text=True, timeout=600)


def test_unmodified_main_builds_the_model(tmp_path):
    r = _run(tmp_path)                       # no action flag: parse arguments, build PGGAN + GANSynth, exit
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the CPU-box failure mode")
def test_unmodified_main_reaches_the_input_pipeline(tmp_path):
    """--generate walks into GANSynth.generate -> nsynth_input_fn -> the CUDA pipeline; without a GPU it must fail
    THERE (loudly), not earlier in the stand-in."""
    r = _run(tmp_path, "--generate", "--filenames", str(tmp_path / "none*.tfrecord"))
    assert r.returncode != 0
    assert os.path.join("gansynth_b200", "dataset.py") in r.stderr and "tensorflow" not in r.stderr.splitlines()[-1]


def test_stand_in_refuses_everything_else():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import tensorflow as tf; "
            "assert tf.cast(3, tf.float32) == 3 and tf.ConfigProto(a=1).a == 1; "
            "step = tf.train.create_global_step(); level = tf.divide(x=step, y=4); step.value = 2; "
            "assert abs(level() - 0.5) < 1e-12; "
            "ok = False\n"
            "try:\n    tf.nn\nexcept AttributeError:\n    ok = True\n"
            "assert ok" % (ROOT, os.path.join(ROOT, "gansynth_b200", "compat")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_unmodified_pitch_classifier_main_builds_the_model(tmp_path):
    """The reference's pitch_classifier_main.py, byte for byte: ResNet + PitchClassifier construction through the stand-in
    (tf.train.exponential_decay inside the learning-rate lambda included); no action flag, so it exits after the build."""
    ref = "/root/reference/pitch_classifier_main.py"
    shutil.copy(ref, tmp_path / "pitch_classifier_main.py")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "gansynth_b200", "compat"), ROOT]))
    r = subprocess.run([sys.executable, "pitch_classifier_main.py"], cwd=tmp_path, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
