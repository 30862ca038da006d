"""Command line of the reference's pitch_classifier_main.py (same flags and hyper-parameters): trains / evaluates the
ResNet pitch classifier whose features GANSynth.evaluate compares (models.py:253-410, networks.py:293-413).

    python -m gansynth_b200.pitch_classifier_main --train --filenames 'nsynth_train.tfrecord' --batch_size 64
    python -m gansynth_b200.gan_synth_main --evaluate --classifier pitch_classifier_model      # its model_dir
"""
import argparse
import functools
import glob

import torch

from .dataset import nsynth_input_fn
from .models import PitchClassifier, exponential_decay
from .networks import ResNet
from .utils import Dict


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--model_dir", type=str, default="pitch_classifier_model")
    parser.add_argument("--filenames", type=str, default="nsynth*.tfrecord")
    parser.add_argument("--batch_size", type=int, default=64)
    parser.add_argument("--num_epochs", type=int, default=100)
    parser.add_argument("--total_steps", type=int, default=50000)
    parser.add_argument("--train", action="store_true")
    parser.add_argument("--evaluate", action="store_true")
    return parser


def build(args, device="cuda"):
    """pitch_classifier_main.py:40-77."""
    resnet = ResNet.pitch_classifier(classes=len(range(24, 85)))
    return PitchClassifier(
        network=resnet,
        input_fn=functools.partial(nsynth_input_fn, filenames=glob.glob(args.filenames), batch_size=args.batch_size,
                                   num_epochs=args.num_epochs if args.train else 1, shuffle=bool(args.train),
                                   pitches=range(24, 85), sources=[0]),
        spectral_params=Dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75),
        hyper_params=Dict(weight_decay=1e-4,
                          learning_rate=lambda global_step: exponential_decay(
                              0.128 * args.batch_size / 256, global_step, 70000 * args.num_epochs / 4 / args.batch_size, 0.1),
                          momentum=0.9, use_nesterov=True),
        device=device)


def main(argv=None):
    args = build_parser().parse_args(argv)
    torch.cuda.set_device(0)
    classifier = build(args)
    if args.train:
        classifier.train(model_dir=args.model_dir, config=None, total_steps=args.total_steps, save_checkpoint_steps=1000,
                         save_summary_steps=100, log_tensor_steps=100)
    if args.evaluate:
        print(classifier.evaluate(model_dir=args.model_dir, config=None))


if __name__ == "__main__":
    main()
