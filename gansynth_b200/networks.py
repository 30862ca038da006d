"""Progressive-growing generator / discriminator on the CUDA kernels
(host mirror of reference networks.py:1-290: same class, constructor and method signatures).

`generator(latents [B,256], labels one-hot [B,61])` returns images NCHW [B, 2, H, W];
`discriminator(images NCHW, labels)` returns (features [B,256], logits [B,61]).
Feature maps are NHWC between the two boundaries.  `growing_level` may be a float, a callable or any
object with __float__ (the reference passes global_step / growing_steps): it is read on the host at
every call, so tf.cond (networks.py:126-152, 261-287) becomes a Python branch and only the active
sub-network launches kernels.
"""
import math

import numpy as np
import torch

from . import functional as F
from . import ops
from .ops import (batch_stddev, conv2d, conv2d_transpose, dense, downscale2d, embedding, group_normalization,
                  max_pooling2d, pixel_normalization, reduce_mean_spatial, upscale2d, variable_scope)


def log(x, base):
    """networks.py:6-7."""
    return math.log(x) / math.log(base)


def lerp(a, b, t):
    """networks.py:10-11: t * a + (1 - t) * b."""
    return F.Axpby.apply(a, b, float(t), 1.0 - float(t))


class PGGAN(object):

    def __init__(self, min_resolution, max_resolution, min_channels, max_channels, growing_level):
        # networks.py:16-29
        self.min_resolution = np.asanyarray(min_resolution)
        self.max_resolution = np.asanyarray(max_resolution)
        self.min_channels = min_channels
        self.max_channels = max_channels
        self.growing_level = growing_level

        def log2(x):
            return 0 if (x == 1).all() else 1 + log2(x >> 1)

        self.min_depth = log2(self.min_resolution // self.min_resolution)
        self.max_depth = log2(self.max_resolution // self.min_resolution)
        self._built = set()
        # progressive-growing blend weights (t, 1 - t) in device memory, maintained by GANSynth while it runs a
        # sub-step (so that the sub-step can be a replayed CUDA graph); None / inactive: lerp takes host floats
        self.lerp_coef = None
        self.device_lerp_active = False

    def structure_key(self):
        """What the kernel sequence of a forward pass depends on: the depth at which the blend happens
        (networks.py:126-152, 261-287: `growing_depth > depth` picks the branch, the fraction only scales)."""
        gd = self.growing_depth
        return ("grown",) if gd > self.max_depth else ("grow", int(math.ceil(gd)))

    def update_lerp_coef(self, device):
        """Writes lerp's (t, 1 - t) = (depth - growing_depth, 1 - t) for the current global step to `lerp_coef`."""
        gd = self.growing_depth
        t = float(math.ceil(gd) - gd) if 0.0 < gd <= self.max_depth else 0.0
        if self.lerp_coef is None or self.lerp_coef.device != torch.device(device):
            self.lerp_coef = torch.zeros(2, device=device, dtype=torch.float32)
        # pageable source: the driver stages it before returning, so the host value cannot be overwritten in flight
        self.lerp_coef.copy_(torch.tensor([t, 1.0 - t], dtype=torch.float32), non_blocking=True)

    def _lerp(self, a, b, t):
        a, b = F.plain(a), F.plain(b)
        if self.device_lerp_active and self.lerp_coef is not None:
            return F.AxpbyDev.apply(a, b, self.lerp_coef)
        return lerp(a, b, t)

    @property
    def growing_depth(self):
        """networks.py:29, evaluated on the host each time it is read."""
        level = self.growing_level() if callable(self.growing_level) else self.growing_level
        return log(1.0 + ((1 << (self.max_depth + 1)) - 1) * float(level), 2.0)

    def resolution(self, depth):
        return self.min_resolution << depth

    def channels(self, depth):
        return min(self.max_channels, self.min_channels << (self.max_depth - depth))

    # ------------------------------------------------------------------ variable creation
    def _ensure_variables(self, name, latent_dim, num_labels):
        """The reference builds BOTH branches of every tf.cond, so all variables of all depths exist
        from the start (and are in every checkpoint).  Create them eagerly, in the reference's order."""
        if name in self._built:
            return
        store = ops.default_store()
        with store.variable_scope(name):
            if name.startswith("generator"):
                ops.get_weight([num_labels, latent_dim], 1.0, True)
                for depth in range(self.min_depth, self.max_depth + 1):
                    ch = self.channels(depth)
                    with store.variable_scope("conv_block_{}x{}".format(*self.resolution(depth))):
                        if depth == self.min_depth:
                            units = ch * int(self.resolution(depth).prod())
                            with store.variable_scope("dense"):
                                ops.get_weight([2 * latent_dim, units], 2.0, True), ops.get_bias([units])
                        else:
                            with store.variable_scope("upscale_conv"):
                                ops.get_weight([3, 3, self.channels(depth - 1), ch], 2.0, True), ops.get_bias([ch])
                        with store.variable_scope("conv"):
                            ops.get_weight([3, 3, ch, ch], 2.0, True), ops.get_bias([ch])
                    with store.variable_scope("color_block_{}x{}".format(*self.resolution(depth))):
                        with store.variable_scope("conv"):
                            ops.get_weight([1, 1, ch, 2], 1.0, True), ops.get_bias([2])
            else:
                for depth in range(self.min_depth, self.max_depth + 1):
                    ch = self.channels(depth)
                    with store.variable_scope("conv_block_{}x{}".format(*self.resolution(depth))):
                        if depth == self.min_depth:
                            feat = self.channels(depth - 1)
                            with store.variable_scope("conv"):
                                ops.get_weight([3, 3, ch + 1, ch], 2.0, True), ops.get_bias([ch])
                            with store.variable_scope("dense"):
                                ops.get_weight([ch * int(self.resolution(depth).prod()), feat], 2.0, True)
                                ops.get_bias([feat])
                            with store.variable_scope("logits"):
                                ops.get_weight([feat, num_labels], 1.0, True), ops.get_bias([num_labels])
                        else:
                            with store.variable_scope("conv"):
                                ops.get_weight([3, 3, ch, ch], 2.0, True), ops.get_bias([ch])
                            with store.variable_scope("conv_downscale"):
                                ops.get_weight([3, 3, ch, self.channels(depth - 1)], 2.0, True)
                                ops.get_bias([self.channels(depth - 1)])
                    with store.variable_scope("color_block_{}x{}".format(*self.resolution(depth))):
                        with store.variable_scope("conv"):
                            ops.get_weight([1, 1, 2, ch], 2.0, True), ops.get_bias([ch])
        self._built.add(name)

    # ------------------------------------------------------------------ generator
    def generator(self, latents, labels, name="generator", reuse=None):
        """networks.py:31-161."""
        self._ensure_variables(name, latents.shape[1], labels.shape[1])
        growing_depth = self.growing_depth
        resolution, channels = self.resolution, self.channels

        def conv_block(inputs, depth):
            with variable_scope("conv_block_{}x{}".format(*resolution(depth))):
                if depth == self.min_depth:
                    inputs = pixel_normalization(inputs)
                    with variable_scope("dense"):
                        inputs = dense(inputs, units=channels(depth) * int(resolution(depth).prod()), use_bias=True,
                                       variance_scale=2.0, scale_weight=True, activation="leaky_relu")
                        # tf.reshape to [B, C, H, W] (row-major NCHW), then to the NHWC working layout
                        b = inputs.shape[0]
                        h, w = (int(r) for r in resolution(depth))
                        inputs = F.TransposeInner.apply(inputs.reshape(b, channels(depth), h * w))
                        inputs = inputs.reshape(b, h, w, channels(depth))
                        inputs = pixel_normalization(inputs)
                    with variable_scope("conv"):
                        # conv -> leaky_relu -> pixel_normalization (networks.py:57-68) as one fused layer
                        inputs = conv2d(inputs, filters=channels(depth), kernel_size=[3, 3], use_bias=True,
                                        variance_scale=2.0, scale_weight=True, activation="leaky_relu",
                                        pixel_norm_epsilon=1.0e-12, protocol=True)
                    return inputs
                with variable_scope("upscale_conv"):
                    inputs = conv2d_transpose(inputs, filters=channels(depth), kernel_size=[3, 3], strides=[2, 2],
                                              use_bias=True, variance_scale=2.0, scale_weight=True,
                                              activation="leaky_relu", pixel_norm_epsilon=1.0e-12, protocol=True)
                with variable_scope("conv"):
                    inputs = conv2d(inputs, filters=channels(depth), kernel_size=[3, 3], use_bias=True,
                                    variance_scale=2.0, scale_weight=True, activation="leaky_relu",
                                    pixel_norm_epsilon=1.0e-12, protocol=True)
                return inputs

        def color_block(inputs, depth):
            with variable_scope("color_block_{}x{}".format(*resolution(depth))):
                with variable_scope("conv"):
                    inputs = conv2d(inputs, filters=2, kernel_size=[1, 1], use_bias=True, variance_scale=1.0,
                                    scale_weight=True)
                    inputs = ops.tanh(inputs)
                return inputs

        def grow(feature_maps, depth):
            def high_resolution_images():
                return grow(conv_block(feature_maps, depth), depth + 1)

            def middle_resolution_images():
                return upscale2d(color_block(conv_block(feature_maps, depth), depth),
                                 factors=resolution(self.max_depth) // resolution(depth))

            def low_resolution_images():
                return upscale2d(color_block(feature_maps, depth - 1),
                                 factors=resolution(self.max_depth) // resolution(depth - 1))

            grown = growing_depth > depth
            if depth == self.min_depth:
                return high_resolution_images() if (grown and depth < self.max_depth) else middle_resolution_images()
            if depth == self.max_depth:
                if grown:
                    return middle_resolution_images()
                return self._lerp(low_resolution_images(), middle_resolution_images(), depth - growing_depth)
            if grown:
                return high_resolution_images()
            return self._lerp(low_resolution_images(), middle_resolution_images(), depth - growing_depth)

        with variable_scope(name):
            embedded = embedding(labels, units=latents.shape[1], variance_scale=1.0, scale_weight=True)
            images = grow(torch.cat([latents, embedded], dim=1), self.min_depth)
        return F.nhwc_to_nchw(F.plain(images))

    # ------------------------------------------------------------------ discriminator
    def discriminator(self, images, labels, name="discriminator", reuse=None):
        """networks.py:163-290."""
        self._ensure_variables(name, 0, labels.shape[1])
        growing_depth = self.growing_depth
        resolution, channels = self.resolution, self.channels
        images = F.nchw_to_nhwc(images)

        def conv_block(inputs, depth):
            with variable_scope("conv_block_{}x{}".format(*resolution(depth))):
                if depth == self.min_depth:
                    with variable_scope("conv"):
                        # conv2d(concat([x, batch_stddev(x)])) with the [3,3,C+1,C] variable of the reference
                        # (networks.py:174-184), evaluated as conv(x, W[:, :, :C]) + conv(stddev, W[:, :, C:]):
                        # the C-channel part stays on the vectorised / tensor-core kernels instead of
                        # dragging a 257-channel tensor through the generic path.
                        ch = channels(depth)
                        inputs = F.plain(inputs)
                        weight, alpha = ops.get_weight([3, 3, ch + 1, ch], 2.0, True)
                        bias = ops.get_bias([ch])
                        stddev = batch_stddev(inputs).contiguous()
                        main = F.ConvC.apply(inputs, weight[:, :, :ch, :].contiguous(), 3, 1, False, alpha)
                        extra = F.ConvC.apply(stddev, weight[:, :, ch:, :].contiguous(), 3, 1, False, alpha)
                        inputs = F.BiasAct.apply(F.Axpby.apply(main, extra, 1.0, 1.0), bias, F.ACT_LRELU)
                    with variable_scope("dense"):
                        # tf.layers.flatten of the NCHW tensor: index = c*H*W + h*W + w
                        b, h, w, c = inputs.shape
                        inputs = F.TransposeInner.apply(inputs.reshape(b, h * w, c)).reshape(b, c * h * w)
                        inputs = dense(inputs, units=channels(depth - 1), use_bias=True, variance_scale=2.0,
                                       scale_weight=True, activation="leaky_relu")
                        features = inputs
                    with variable_scope("logits"):
                        logits = dense(inputs, units=labels.shape[1], use_bias=True, variance_scale=1.0,
                                       scale_weight=True)
                    return features, logits
                # protocol=True: each activated output travels as functional.PreMasked, so that the next convolution's
                # backward applies this layer's leaky-relu mask in its epilogue
                with variable_scope("conv"):
                    inputs = conv2d(inputs, filters=channels(depth), kernel_size=[3, 3], use_bias=True,
                                    variance_scale=2.0, scale_weight=True, activation="leaky_relu", protocol=True)
                with variable_scope("conv_downscale"):
                    inputs = conv2d(inputs, filters=channels(depth - 1), kernel_size=[3, 3], strides=[2, 2],
                                    use_bias=True, variance_scale=2.0, scale_weight=True, activation="leaky_relu",
                                    protocol=True)
                return inputs

        def color_block(inputs, depth):
            with variable_scope("color_block_{}x{}".format(*resolution(depth))):
                with variable_scope("conv"):
                    inputs = conv2d(inputs, filters=channels(depth), kernel_size=[1, 1], use_bias=True,
                                    variance_scale=2.0, scale_weight=True, activation="leaky_relu", protocol=True)
                return inputs

        def grow(depth):
            def high_resolution_feature_maps():
                return conv_block(grow(depth + 1), depth)

            def middle_resolution_feature_maps():
                return conv_block(color_block(downscale2d(
                    images, factors=resolution(self.max_depth) // resolution(depth)), depth), depth)

            def low_resolution_feature_maps():
                return color_block(downscale2d(
                    images, factors=resolution(self.max_depth) // resolution(depth - 1)), depth - 1)

            grown = growing_depth > depth
            if depth == self.min_depth:
                return (high_resolution_feature_maps() if (grown and depth < self.max_depth)
                        else middle_resolution_feature_maps())
            if depth == self.max_depth:
                if grown:
                    return middle_resolution_feature_maps()
                return self._lerp(low_resolution_feature_maps(), middle_resolution_feature_maps(), depth - growing_depth)
            if grown:
                return high_resolution_feature_maps()
            return self._lerp(low_resolution_feature_maps(), middle_resolution_feature_maps(), depth - growing_depth)

        with variable_scope(name):
            return grow(self.min_depth)


class ResNet(object):
    """The pitch classifier `GANSynth.evaluate` takes its features from (reference networks.py:293-413: pre-activation
    ResNet v2 blocks without bottleneck, group normalisation, weight-standardised convolutions; constructor arguments as
    in pitch_classifier_main.py:42-53).  `__call__(images NCHW [B, 2, H, W]) -> (features [B, F], logits [B, classes])`;
    differentiable (every layer is an autograd Function over the CUDA kernels), trained by `models.PitchClassifier`.
    Variables carry the reference's names (`resnet/conv/weight`, `resnet/residual_block_0_0/conv_1st/weight`,
    `.../group_normalization_1st/gamma`, `resnet/logits/weight` ...), so a checkpoint of the reference's classifier
    loads by name (GANSynth.import_tf_checkpoint / VariableStore.load)."""

    def __init__(self, conv_param, pool_param, residual_params, groups, classes):
        self.conv_param = conv_param
        self.pool_param = pool_param
        self.residual_params = residual_params
        self.groups = groups
        self.classes = classes

    @staticmethod
    def _get(param, key):
        return param[key] if isinstance(param, dict) else getattr(param, key)

    def __call__(self, inputs, name="resnet", reuse=None):
        get = self._get

        def residual_block(inputs, filters, strides, projection_shortcut, groups):
            # networks.py:305-358
            shortcut = inputs
            with variable_scope("group_normalization_1st"):
                inputs = group_normalization(inputs, groups=groups, relu=True)
            if projection_shortcut:
                with variable_scope("projection_shortcut"):
                    shortcut = conv2d(inputs, filters=filters, kernel_size=[1, 1], strides=strides, use_bias=False,
                                      variance_scale=2.0, apply_weight_standardization=True)
            with variable_scope("conv_1st"):
                inputs = conv2d(inputs, filters=filters, kernel_size=[3, 3], strides=strides, use_bias=True,
                                variance_scale=2.0, apply_weight_standardization=True)
            with variable_scope("group_normalization_2nd"):
                inputs = group_normalization(inputs, groups=groups, relu=True)
            with variable_scope("conv_2nd"):
                inputs = conv2d(inputs, filters=filters, kernel_size=[3, 3], strides=[1, 1], use_bias=True,
                                variance_scale=2.0, apply_weight_standardization=True)
            return F.Axpby.apply(inputs, shortcut, 1.0, 1.0)

        with variable_scope(name):
            inputs = F.nchw_to_nhwc(inputs)
            if self.conv_param:
                with variable_scope("conv"):
                    inputs = conv2d(inputs, filters=get(self.conv_param, "filters"), kernel_size=get(self.conv_param, "kernel_size"),
                                    strides=get(self.conv_param, "strides"), use_bias=True, variance_scale=2.0,
                                    apply_weight_standardization=True)
            if self.pool_param:
                inputs = max_pooling2d(inputs, kernel_size=get(self.pool_param, "kernel_size"),
                                       strides=get(self.pool_param, "strides"))
            for i, rp in enumerate(self.residual_params):
                for j in range(get(rp, "blocks")):
                    with variable_scope("residual_block_{}_{}".format(i, j)):
                        inputs = residual_block(inputs, filters=get(rp, "filters"),
                                                strides=get(rp, "strides") if j == 0 else [1, 1],
                                                projection_shortcut=(j == 0), groups=self.groups)
            with variable_scope("group_normalization"):
                inputs = group_normalization(inputs, groups=self.groups, relu=True)
            features = reduce_mean_spatial(inputs)
            with variable_scope("logits"):
                logits = dense(features, units=self.classes, use_bias=True, variance_scale=1.0)
            return features, logits

    @classmethod
    def pitch_classifier(cls, classes=61):
        """The configuration of pitch_classifier_main.py:42-53 (ResNet-34 layout, 32 groups, 61 pitches)."""
        return cls(conv_param=dict(filters=64, kernel_size=[7, 7], strides=[2, 2]),
                   pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
                   residual_params=[dict(filters=64, strides=[1, 1], blocks=3), dict(filters=128, strides=[2, 2], blocks=4),
                                    dict(filters=256, strides=[2, 2], blocks=6), dict(filters=512, strides=[2, 2], blocks=3)],
                   groups=32, classes=classes)
