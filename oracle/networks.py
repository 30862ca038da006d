"""Oracle restatement of the progressive-growing generator / discriminator
(reference networks.py:1-290), PyTorch CPU, NCHW.  Test infrastructure only.

Variables live in a flat dict keyed by the TF variable names the reference's scopes produce
(networks.py:40,43,57,70,82,96,154,172,175,185,196,207,217,231; ops.py:157,176), e.g.
``generator/conv_block_2x16/dense/weight`` or ``discriminator/color_block_128x1024/conv/bias``.
``growing_depth`` is a host float, so tf.cond (networks.py:126-152, 261-287) is a Python branch.
"""
import math

import numpy as np
import torch

from . import ops


def lerp(a, b, t):
    """networks.py:10-11: t * a + (1 - t) * b  (t weights the first argument)."""
    return t * a + (1.0 - t) * b


class PGGAN(object):

    def __init__(self, min_resolution, max_resolution, min_channels, max_channels, growing_level):
        # networks.py:16-29
        self.min_resolution = np.asarray(min_resolution)
        self.max_resolution = np.asarray(max_resolution)
        self.min_channels = min_channels
        self.max_channels = max_channels
        self.growing_level = growing_level
        ratio = self.max_resolution // self.min_resolution
        self.min_depth = 0
        self.max_depth = int(round(math.log2(int(ratio[0]))))
        assert (ratio == (1 << self.max_depth)).all()

    @property
    def growing_depth(self):
        level = float(self.growing_level() if callable(self.growing_level) else self.growing_level)
        return math.log2(1.0 + ((1 << (self.max_depth + 1)) - 1) * level)

    def resolution(self, depth):
        return self.min_resolution << depth

    def channels(self, depth):
        return min(self.max_channels, self.min_channels << (self.max_depth - depth))

    def _block(self, kind, depth):
        return "{}_block_{}x{}".format(kind, *self.resolution(depth))

    # ------------------------------------------------------------------ variables
    def variable_shapes(self, latent_dim=256, num_labels=61):
        """Every variable the reference graph creates (all cond branches are built), in creation
        order per network: name -> (shape, variance_scale used for the He constant)."""
        g, d = {}, {}
        g["generator/weight"] = ((num_labels, latent_dim), 1.0)
        for depth in range(self.min_depth, self.max_depth + 1):
            cb = "generator/" + self._block("conv", depth)
            ch = self.channels(depth)
            if depth == self.min_depth:
                units = ch * int(self.resolution(depth).prod())
                g[cb + "/dense/weight"] = ((2 * latent_dim, units), 2.0)
                g[cb + "/dense/bias"] = ((units,), None)
                g[cb + "/conv/weight"] = ((3, 3, ch, ch), 2.0)
                g[cb + "/conv/bias"] = ((ch,), None)
            else:
                g[cb + "/upscale_conv/weight"] = ((3, 3, self.channels(depth - 1), ch), 2.0)
                g[cb + "/upscale_conv/bias"] = ((ch,), None)
                g[cb + "/conv/weight"] = ((3, 3, ch, ch), 2.0)
                g[cb + "/conv/bias"] = ((ch,), None)
            kb = "generator/" + self._block("color", depth)
            g[kb + "/conv/weight"] = ((1, 1, ch, 2), 1.0)
            g[kb + "/conv/bias"] = ((2,), None)
        for depth in range(self.min_depth, self.max_depth + 1):
            cb = "discriminator/" + self._block("conv", depth)
            ch = self.channels(depth)
            if depth == self.min_depth:
                d[cb + "/conv/weight"] = ((3, 3, ch + 1, ch), 2.0)
                d[cb + "/conv/bias"] = ((ch,), None)
                feat = self.channels(depth - 1)
                d[cb + "/dense/weight"] = ((ch * int(self.resolution(depth).prod()), feat), 2.0)
                d[cb + "/dense/bias"] = ((feat,), None)
                d[cb + "/logits/weight"] = ((feat, num_labels), 1.0)
                d[cb + "/logits/bias"] = ((num_labels,), None)
            else:
                d[cb + "/conv/weight"] = ((3, 3, ch, ch), 2.0)
                d[cb + "/conv/bias"] = ((ch,), None)
                d[cb + "/conv_downscale/weight"] = ((3, 3, ch, self.channels(depth - 1)), 2.0)
                d[cb + "/conv_downscale/bias"] = ((self.channels(depth - 1),), None)
            kb = "discriminator/" + self._block("color", depth)
            d[kb + "/conv/weight"] = ((1, 1, 2, ch), 2.0)
            d[kb + "/conv/bias"] = ((ch,), None)
        return g, d

    def init_variables(self, seed=3, latent_dim=256, num_labels=61, dtype=torch.float32, bias_std=0.0):
        """Weights ~ truncated N(0,1) resampled beyond 2 sigma (ops.py:159), biases zero (ops.py:178).
        ``bias_std`` > 0 gives non-zero biases so parity tests exercise the bias path."""
        gen = torch.Generator().manual_seed(seed)
        out = {}
        for table in self.variable_shapes(latent_dim, num_labels):
            for name, (shape, vs) in table.items():
                if vs is None:
                    t = torch.randn(shape, generator=gen, dtype=torch.float64) * bias_std
                else:
                    t = torch.empty(shape, dtype=torch.float64)
                    torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=gen)
                out[name] = t.to(dtype)
        return out

    # ------------------------------------------------------------------ generator
    def generator(self, params, latents, labels, name="generator"):
        """networks.py:31-161."""
        P = lambda n: params[name + "/" + n]
        gd = self.growing_depth

        def conv_block(x, depth):
            cb = self._block("conv", depth)
            if depth == self.min_depth:
                x = ops.pixel_normalization(x)
                x = ops.dense(x, P(cb + "/dense/weight"), P(cb + "/dense/bias"), 2.0)
                x = x.reshape(-1, self.channels(depth), *[int(r) for r in self.resolution(depth)])
                x = ops.pixel_normalization(ops.leaky_relu(x))
                x = ops.conv2d(x, P(cb + "/conv/weight"), P(cb + "/conv/bias"), (1, 1), 2.0)
                return ops.pixel_normalization(ops.leaky_relu(x))
            x = ops.conv2d_transpose(x, P(cb + "/upscale_conv/weight"), P(cb + "/upscale_conv/bias"), (2, 2), 2.0)
            x = ops.pixel_normalization(ops.leaky_relu(x))
            x = ops.conv2d(x, P(cb + "/conv/weight"), P(cb + "/conv/bias"), (1, 1), 2.0)
            return ops.pixel_normalization(ops.leaky_relu(x))

        def color_block(x, depth):
            kb = self._block("color", depth)
            return torch.tanh(ops.conv2d(x, P(kb + "/conv/weight"), P(kb + "/conv/bias"), (1, 1), 1.0))

        def grow(fm, depth):
            def high():
                return grow(conv_block(fm, depth), depth + 1)

            def middle():
                return ops.upscale2d(color_block(conv_block(fm, depth), depth),
                                     self.resolution(self.max_depth) // self.resolution(depth))

            def low():
                return ops.upscale2d(color_block(fm, depth - 1),
                                     self.resolution(self.max_depth) // self.resolution(depth - 1))

            grown = gd > depth
            if depth == self.min_depth:
                return high() if (grown and depth < self.max_depth) else middle()
            if depth == self.max_depth:
                return middle() if grown else lerp(low(), middle(), depth - gd)
            return high() if grown else lerp(low(), middle(), depth - gd)

        emb = ops.embedding(labels, P("weight"), 1.0)
        return grow(torch.cat([latents, emb], dim=1), self.min_depth)

    # ------------------------------------------------------------------ discriminator
    def discriminator(self, params, images, labels, name="discriminator"):
        """networks.py:163-290.  Returns (features, logits)."""
        P = lambda n: params[name + "/" + n]
        gd = self.growing_depth

        def conv_block(x, depth):
            cb = self._block("conv", depth)
            if depth == self.min_depth:
                x = torch.cat([x, ops.batch_stddev(x)], dim=1)
                x = ops.leaky_relu(ops.conv2d(x, P(cb + "/conv/weight"), P(cb + "/conv/bias"), (1, 1), 2.0))
                x = x.reshape(x.shape[0], -1)
                feats = ops.leaky_relu(ops.dense(x, P(cb + "/dense/weight"), P(cb + "/dense/bias"), 2.0))
                logits = ops.dense(feats, P(cb + "/logits/weight"), P(cb + "/logits/bias"), 1.0)
                return feats, logits
            x = ops.leaky_relu(ops.conv2d(x, P(cb + "/conv/weight"), P(cb + "/conv/bias"), (1, 1), 2.0))
            return ops.leaky_relu(ops.conv2d(x, P(cb + "/conv_downscale/weight"),
                                             P(cb + "/conv_downscale/bias"), (2, 2), 2.0))

        def color_block(x, depth):
            kb = self._block("color", depth)
            return ops.leaky_relu(ops.conv2d(x, P(kb + "/conv/weight"), P(kb + "/conv/bias"), (1, 1), 2.0))

        def grow(depth):
            def high():
                return conv_block(grow(depth + 1), depth)

            def middle():
                f = self.resolution(self.max_depth) // self.resolution(depth)
                return conv_block(color_block(ops.downscale2d(images, f), depth), depth)

            def low():
                f = self.resolution(self.max_depth) // self.resolution(depth - 1)
                return color_block(ops.downscale2d(images, f), depth - 1)

            grown = gd > depth
            if depth == self.min_depth:
                return high() if (grown and depth < self.max_depth) else middle()
            if depth == self.max_depth:
                return middle() if grown else lerp(low(), middle(), depth - gd)
            return high() if grown else lerp(low(), middle(), depth - gd)

        return grow(self.min_depth)


class ResNet(object):
    """networks.py:293-413 restated (forward): ResNet-v2 blocks, group normalisation, weight-standardised convolutions.
    `params` maps the reference's variable names (resnet/conv/weight, resnet/residual_block_i_j/..., resnet/logits/...)."""

    def __init__(self, conv_param, pool_param, residual_params, groups, classes):
        self.conv_param, self.pool_param, self.residual_params = conv_param, pool_param, residual_params
        self.groups, self.classes = groups, classes

    def variable_shapes(self, in_channels=2):
        d = {}
        c = in_channels
        if self.conv_param:
            k = self.conv_param["kernel_size"]
            d["resnet/conv/weight"] = (k[0], k[1], c, self.conv_param["filters"])
            d["resnet/conv/bias"] = (self.conv_param["filters"],)
            c = self.conv_param["filters"]
        for i, rp in enumerate(self.residual_params):
            for j in range(rp["blocks"]):
                b = "resnet/residual_block_%d_%d/" % (i, j)
                f = rp["filters"]
                d[b + "group_normalization_1st/beta"] = (c,)
                d[b + "group_normalization_1st/gamma"] = (c,)
                if j == 0:
                    d[b + "projection_shortcut/weight"] = (1, 1, c, f)
                d[b + "conv_1st/weight"] = (3, 3, c, f)
                d[b + "conv_1st/bias"] = (f,)
                d[b + "group_normalization_2nd/beta"] = (f,)
                d[b + "group_normalization_2nd/gamma"] = (f,)
                d[b + "conv_2nd/weight"] = (3, 3, f, f)
                d[b + "conv_2nd/bias"] = (f,)
                c = f
        d["resnet/group_normalization/beta"] = (c,)
        d["resnet/group_normalization/gamma"] = (c,)
        d["resnet/logits/weight"] = (c, self.classes)
        d["resnet/logits/bias"] = (self.classes,)
        return d

    def init_variables(self, seed=5, dtype=torch.float32):
        """Random values for EVERY variable (also gamma / beta / biases, which the reference starts at 1 / 0), so that
        parity tests exercise them."""
        gen = torch.Generator().manual_seed(seed)
        out = {}
        for name, shape in self.variable_shapes().items():
            t = torch.randn(shape, generator=gen, dtype=torch.float64)
            if name.endswith("gamma"):
                t = 1.0 + 0.3 * t
            elif name.endswith(("beta", "bias")):
                t = 0.2 * t
            else:
                fan_in = 1
                for s_ in shape[:-1]:
                    fan_in *= s_
                t = t * (2.0 / fan_in) ** 0.5
            out[name] = t.to(dtype)
        return out

    def __call__(self, params, images):
        P = lambda n: params["resnet/" + n]
        x = images
        if self.conv_param:
            x = ops.conv2d_plain(x, P("conv/weight"), P("conv/bias"), self.conv_param["strides"])
        if self.pool_param:
            x = ops.max_pooling2d(x, self.pool_param["kernel_size"], self.pool_param["strides"])
        for i, rp in enumerate(self.residual_params):
            for j in range(rp["blocks"]):
                b = "residual_block_%d_%d/" % (i, j)
                strides = rp["strides"] if j == 0 else [1, 1]
                shortcut = x
                x = torch.relu(ops.group_normalization(x, P(b + "group_normalization_1st/gamma"),
                                                       P(b + "group_normalization_1st/beta"), self.groups))
                if j == 0:
                    shortcut = ops.conv2d_plain(x, P(b + "projection_shortcut/weight"), None, strides)
                x = ops.conv2d_plain(x, P(b + "conv_1st/weight"), P(b + "conv_1st/bias"), strides)
                x = torch.relu(ops.group_normalization(x, P(b + "group_normalization_2nd/gamma"),
                                                       P(b + "group_normalization_2nd/beta"), self.groups))
                x = ops.conv2d_plain(x, P(b + "conv_2nd/weight"), P(b + "conv_2nd/bias"), [1, 1])
                x = x + shortcut
        x = torch.relu(ops.group_normalization(x, P("group_normalization/gamma"), P("group_normalization/beta"), self.groups))
        features = x.mean(dim=(2, 3))
        logits = features @ P("logits/weight") + P("logits/bias")
        return features, logits
