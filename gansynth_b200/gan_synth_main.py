"""Entry point with the command line of the reference gan_synth_main.py (:25-36): same flags, same model
configuration (:44-50), same spectral / hyper parameters (:70-88).

    python -m gansynth_b200.gan_synth_main --train --filenames 'nsynth*.tfrecord' --batch_size 8
    python -m torch.distributed.run --nproc-per-node 8 -m gansynth_b200.gan_synth_main --train ...

TensorFlow is only scaffolding in the reference main (graph / session / global step / tf.random.normal);
here the global step is `get_or_create_global_step()`, latents come from torch.randn on the device, `config`
is unused.  Under torchrun every rank trains on its own shard of the shuffled records (rank-dependent
shuffle seed) and GANSynth all-reduces the gradients once per sub-step (--batch_size stays the PER-GPU batch;
the learning rate follows the global batch as :79-82 intends)."""
import argparse
import functools
import glob
import os

import torch

from . import dataset
from .models import GANSynth, get_or_create_global_step
from .networks import PGGAN
from .utils import Dict


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--model_dir", type=str, default="gan_synth_model")
    parser.add_argument("--filenames", type=str, default="nsynth*.tfrecord")
    parser.add_argument("--batch_size", type=int, default=8)
    parser.add_argument("--num_epochs", type=int, default=None)
    parser.add_argument("--total_steps", type=int, default=1000000)
    parser.add_argument("--growing_steps", type=int, default=1000000)
    parser.add_argument("--classifier", type=str, default="pitch_classifier.pb")
    parser.add_argument("--train", action="store_true")
    parser.add_argument("--evaluate", action="store_true")
    parser.add_argument("--generate", action="store_true")
    return parser


def build(args, device="cuda", world_size=1, rank=0):
    """gan_synth_main.py:40-89 -> (PGGAN, GANSynth)."""
    torch.manual_seed(rank)                                             # tf.set_random_seed(0), per-rank latents
    global_step = get_or_create_global_step()
    pggan = PGGAN(min_resolution=[2, 16], max_resolution=[128, 1024], min_channels=32, max_channels=256,
                  growing_level=global_step / args.growing_steps)       # re-evaluated at every call
    pipeline_args = dict(filenames=sorted(glob.glob(args.filenames)), batch_size=args.batch_size,
                         num_epochs=args.num_epochs if args.train else 1, shuffle=True if args.train else False,
                         pitches=range(24, 85), sources=[0])
    if world_size > 1:
        pipe = dataset.NSynthPipeline(device=device, seed=rank, **pipeline_args)
        real_input_fn = functools.partial(next, pipe)
    else:
        real_input_fn = functools.partial(dataset.nsynth_input_fn, **pipeline_args)
    global_batch = args.batch_size * world_size
    gan_synth = GANSynth(
        generator=pggan.generator, discriminator=pggan.discriminator, real_input_fn=real_input_fn,
        fake_input_fn=lambda: torch.randn(args.batch_size, 256, device=device),
        spectral_params=Dict(waveform_length=64000, sample_rate=16000, spectrogram_shape=[128, 1024], overlap=0.75),
        hyper_params=Dict(generator_learning_rate=8e-4 * global_batch / 8, generator_beta1=0.0, generator_beta2=0.99,
                          discriminator_learning_rate=8e-4 * global_batch / 8, discriminator_beta1=0.0,
                          discriminator_beta2=0.99, mode_seeking_loss_weight=0.1, real_gradient_penalty_weight=5.0,
                          fake_gradient_penalty_weight=0.0),
        device=device)
    return pggan, gan_synth


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.evaluate and str(args.classifier).endswith(".pb"):
        parser.error("--classifier %s is a frozen TensorFlow GraphDef, which cannot be executed without TensorFlow; pass the "
                     "CHECKPOINT of the pitch classifier instead (model_dir of the reference's pitch_classifier_main.py): "
                     "it is run by gansynth_b200.networks.ResNet" % args.classifier)
    world_size, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world_size > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl")
    _, gan_synth = build(args, "cuda", world_size, rank)
    if args.train:
        gan_synth.train(model_dir=args.model_dir, config=None, total_steps=args.total_steps,
                        save_checkpoint_steps=1000, save_summary_steps=100, log_tensor_steps=100)
    if args.evaluate:
        print(gan_synth.evaluate(model_dir=args.model_dir, config=None, classifier=args.classifier,
                                 input_name="images:0", output_names=["features:0", "logits:0"]))
    if args.generate:
        from scipy.io import wavfile
        os.makedirs("samples", exist_ok=True)
        num_waveforms = 0
        for waveforms in gan_synth.generate(model_dir=args.model_dir, config=None):
            for waveform in waveforms:
                wavfile.write("samples/%d.wav" % num_waveforms, rate=16000, data=waveform)
                num_waveforms += 1
        print("%d waveforms are generated in `samples` directory" % num_waveforms)


if __name__ == "__main__":
    main()
