"""TEST INFRASTRUCTURE: a torch-CPU emulation of gansynth_b200.kernels.CudaBackend's primitive API.

It lets the CPU test-suite check the host-side logic of the product (autograd Function composition in
functional.py, network wiring, loss, optimiser bookkeeping, sparse spectral constants) against the
oracle without a GPU.  It is never imported by the product; the product has no CPU path.
Each method restates the contract of the C-ABI entry point of the same name (include/gansynth_b200.h).
"""
import math

import torch
import torch.nn.functional as TF


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _act(y, act):
    return TF.leaky_relu(y, 0.2) if act == 1 else y


class EmuBackend(object):

    def register_parameters(self, flats):
        pass

    def weight_cache_refresh(self, flat=None):
        pass

    def weight_cache_reset(self):
        pass

    def _w(self, w, wswap):
        return w.permute(0, 1, 3, 2) if wswap else w   # -> [k, k, ci, co]

    def conv_c(self, x, w, bias, ksize, stride, wswap, alpha, act, precise=False, mask_src=None):
        if mask_src is not None:
            return self.mask_mul(EmuBackend.conv_c(self, x, w, bias, ksize, stride, wswap, alpha, act), mask_src)
        wt = self._w(w, wswap).permute(3, 2, 0, 1)
        pb = max(ksize - stride, 0) // 2                 # TF SAME on sizes divisible by the stride
        pa = max(ksize - stride, 0) - pb
        xp = TF.pad(_nchw(x), (pb, pa, pb, pa))
        y = TF.conv2d(xp, wt, stride=stride) * alpha
        if bias is not None:
            y = y + bias.view(1, -1, 1, 1)
        return _nhwc(_act(y, act))

    def conv_pn(self, x, w, bias, form, ksize, stride, wswap, alpha, eps):
        fn = EmuBackend.conv_c if form == "c" else EmuBackend.conv_t     # class functions: one kernel, one call
        return self.pn_fwd(fn(self, x, w, bias, ksize, stride, wswap, alpha, 1), eps)

    def pn_bwd_mask_y(self, y, r, dy, want_colsum=False):
        return self.pn_bwd_mask(y / r.unsqueeze(-1), r, dy, want_colsum)

    def pn_bwd_mask_second_y(self, y, r, dy, u):
        return self.pn_bwd_mask_second(y / r.unsqueeze(-1), r, dy, u)

    def conv_t(self, dy, w, bias, ksize, stride, wswap, alpha, act, precise=False, mask_src=None):
        if mask_src is not None:
            return self.mask_mul(EmuBackend.conv_t(self, dy, w, bias, ksize, stride, wswap, alpha, act), mask_src)
        wt = self._w(w, wswap).permute(3, 2, 0, 1)      # [co, ci, k, k]
        pb = max(ksize - stride, 0) // 2
        n, oh, ow, co = dy.shape
        full = TF.conv_transpose2d(_nchw(dy), wt, stride=stride)
        h, wd = oh * stride, ow * stride
        full = TF.pad(full, (0, max(0, pb + wd - full.shape[3]), 0, max(0, pb + h - full.shape[2])))
        dx = full[:, :, pb:pb + h, pb:pb + wd] * alpha
        if bias is not None:
            dx = dx + bias.view(1, -1, 1, 1)
        return _nhwc(_act(dx, act))

    def conv_w(self, x, dy, ksize, stride, wswap, alpha, bias_of=None, out=None):
        if out is not None:
            out[0].add_(EmuBackend.conv_w(self, x, dy, ksize, stride, wswap, alpha))
            if out[1] is not None:
                src = x if bias_of == "x" else dy
                out[1].add_(src.reshape(-1, src.shape[-1]).sum(0))
            return None
        if bias_of is not None:
            src = x if bias_of == "x" else dy
            return EmuBackend.conv_w(self, x, dy, ksize, stride, wswap, alpha), src.reshape(-1, src.shape[-1]).sum(0)
        pb = max(ksize - stride, 0) // 2
        pa = max(ksize - stride, 0) - pb
        xp = TF.pad(_nchw(x), (pb, pa, pb, pa))
        ci, co = x.shape[3], dy.shape[3]
        # (torch.nn.grad.conv2d_weight aborts on 1x1 stride-2 shapes whose last row / column is unused: differentiate instead)
        # -- and crop the rows / columns no window reaches (1x1 stride 2), which corrupt the heap in torch's CPU kernel
        oh, ow = dy.shape[1], dy.shape[2]
        xp = xp[:, :, :(oh - 1) * stride + ksize, :(ow - 1) * stride + ksize].contiguous()
        w0 = torch.zeros(co, ci, ksize, ksize, dtype=x.dtype, requires_grad=True)
        with torch.enable_grad():
            (dw,) = torch.autograd.grad(TF.conv2d(xp.detach(), w0, stride=stride), w0, _nchw(dy).contiguous().detach())
        dw = dw.permute(2, 3, 1, 0) * alpha             # [k, k, ci, co]
        return (dw.permute(0, 1, 3, 2) if wswap else dw).contiguous()

    def dense_fwd(self, x, w, alpha):
        return (x @ w) * alpha

    def dense_dgrad(self, dy, w, alpha):
        return (dy @ w.t()) * alpha

    def dense_wgrad(self, x, dy, alpha):
        return (x.t() @ dy) * alpha

    def embedding_fwd(self, table, idx, alpha):
        return table[idx] * alpha

    def embedding_bwd(self, dy, idx, rows, alpha):
        out = torch.zeros(rows, dy.shape[1], dtype=dy.dtype)
        return out.index_add_(0, idx, dy * alpha)

    def lrelu(self, x):
        return TF.leaky_relu(x, 0.2)

    def group_norm(self, x, gamma, beta, groups, eps, relu):
        n, c = x.shape[0], x.shape[-1]
        v = x.reshape(n, -1, groups, c // groups)
        mean = v.mean(dim=(1, 3), keepdim=True)
        var = v.var(dim=(1, 3), unbiased=False, keepdim=True)
        y = ((v - mean) / torch.sqrt(var + eps)).reshape(x.shape) * gamma + beta
        cnt = v.shape[1] * v.shape[3]
        stats = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)          # [n, groups, 2] as the kernel keeps them
        return (torch.relu(y) if relu else y), stats

    def group_norm_bwd(self, x, y, dy, stats, gamma, groups, eps, relu):
        xd, gd = x.detach().clone().requires_grad_(True), gamma.detach().clone().requires_grad_(True)
        bd = torch.zeros_like(gamma).requires_grad_(True)
        if relu:
            dy = dy * (y > 0).to(dy.dtype)          # the mask comes from the forward OUTPUT (which includes beta)
        with torch.enable_grad():
            out, _ = EmuBackend.group_norm(self, xd, gd, bd, groups, eps, False)
            dx, dg, db = torch.autograd.grad(out, (xd, gd, bd), dy)
        return dx, dg, db

    def max_pool_bwd(self, x, y, dy, ksize, stride):
        xd = x.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (dx,) = torch.autograd.grad(EmuBackend.max_pool(self, xd, ksize, stride), xd, dy)
        return dx

    def spatial_mean_bwd(self, dy, shape):
        hw = shape[1] * shape[2]
        return (dy / hw).view(shape[0], 1, 1, shape[3]).expand(*shape).contiguous()

    def adam_slice(self, n, rank, world):
        quads = n // 4
        per = -(-quads // world)
        return 4 * min(per * rank, quads), 4 * min(per * (rank + 1), quads)

    def momentum_step(self, p, g, accum, wd, lr, momentum, nesterov, grad_scale=1.0):
        with torch.no_grad():
            gi = g * grad_scale + (wd * p if wd is not None else 0.0)
            accum.mul_(momentum).add_(gi)
            p.sub_(lr * (gi + momentum * accum if nesterov else accum))

    def max_pool(self, x, ksize, stride):
        h, w = x.shape[1], x.shape[2]
        oh, ow = -(-h // stride), -(-w // stride)
        ph, pw = max((oh - 1) * stride + ksize - h, 0), max((ow - 1) * stride + ksize - w, 0)
        xp = TF.pad(_nchw(x), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
        return _nhwc(TF.max_pool2d(xp, ksize, stride))

    def spatial_mean(self, x):
        return x.mean(dim=(1, 2))

    def mask_mul(self, v, y):
        return v * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.2))

    def tanh_fwd(self, x):
        return torch.tanh(x)

    def tanh_bwd(self, y, dy):
        return dy * (1 - y * y)

    def tanh_bwd2(self, y, dy, u):
        return -2 * y * dy * u

    def axpby(self, a, b, alpha, beta):
        return alpha * a if b is None else alpha * a + beta * b

    def axpby_dev(self, a, b, coef, ia, ib):
        return coef[ia] * a if b is None else coef[ia] * a + coef[ib] * b

    def bias_act(self, x, bias, act):
        return _act(x if bias is None else x + bias, act)

    def row_broadcast(self, s, lead_shape):
        return s.expand(*lead_shape, s.numel()).contiguous()

    def mask_mul_colsum(self, v, y):
        out = self.mask_mul(v, y)
        return out, self.col_sum(out)

    def pn_bwd_mask(self, a, r, dy, want_colsum):
        dz = self.mask_mul(self.pn_bwd(a, r, dy), a)
        return dz, (self.col_sum(dz) if want_colsum else None)

    def pn_bwd_mask_second(self, a, r, dy, u):
        mu = self.mask_mul(u, a)
        return self.mask_mul(self.pn_bwd2(a, r, dy, mu), a), self.pn_bwd(a, r, mu)

    def col_sum(self, v):
        return v.reshape(-1, v.shape[-1]).sum(0)

    def pn_fwd(self, a, eps):
        r = 1.0 / torch.sqrt((a * a).mean(-1) + eps)
        return a * r.unsqueeze(-1), r

    def pn_bwd(self, a, r, dy):
        c = a.shape[-1]
        dot = (a * dy).sum(-1, keepdim=True)
        r = r.unsqueeze(-1)
        return r * dy - (r ** 3 / c) * dot * a

    def pn_bwd2(self, a, r, dy, u):
        c = a.shape[-1]
        r = r.unsqueeze(-1)
        ud = (u * dy).sum(-1, keepdim=True)
        ad = (a * dy).sum(-1, keepdim=True)
        ua = (u * a).sum(-1, keepdim=True)
        return -(r ** 3 / c) * (ud * a + ad * u + ua * dy) + 3 * (r ** 5 / c ** 2) * ua * ad * a

    def _sd(self, x, groups, eps):
        b, e = x.shape
        g = x.reshape(groups, b // groups, e)
        mu = g.mean(0, keepdim=True)
        s = torch.sqrt(((g - mu) ** 2).mean(0) + eps)
        return g, mu, s

    def stddev_fwd(self, x, groups, eps):
        _, _, s = self._sd(x, groups, eps)
        return s.mean(1)

    def stddev_bwd(self, x, df, groups, eps):
        g, mu, s = self._sd(x, groups, eps)
        e = x.shape[1]
        return (df.view(1, -1, 1) * (g - mu) / (e * groups * s)).reshape(x.shape)

    def stddev_bwd2(self, x, df, u, groups, eps):
        g, mu, s = self._sd(x, groups, eps)
        e = x.shape[1]
        ug = u.reshape(g.shape)
        ubar = ug.mean(0, keepdim=True)
        ux = (ug * (g - mu)).sum(0)
        cm = df.view(1, -1, 1) / (e * groups)
        gx = cm * ((ug - ubar) / s - (ux / (groups * s ** 3)) * (g - mu))
        q = (ux / s).sum(1) / (e * groups)
        return gx.reshape(x.shape), q

    def upscale(self, x, fh, fw, scale):
        return (x.repeat_interleave(fh, 1).repeat_interleave(fw, 2) * scale).contiguous()

    def pool(self, x, fh, fw, scale):
        n, hh, ww, c = x.shape
        return x.reshape(n, hh // fh, fh, ww // fw, fw, c).sum((2, 4)) * scale

    def transpose_inner(self, x):
        return x.transpose(1, 2).contiguous()

    def row_dot(self, a, b):
        return (a * b).sum(1)

    def row_scale(self, a, s, alpha=1.0):
        return a * (alpha * s).unsqueeze(1)

    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, t, grad_scale=1.0):
        lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
        with torch.no_grad():
            gi = g * grad_scale
            m.mul_(beta1).add_(gi, alpha=1.0 - beta1)
            v.mul_(beta2).add_(gi * gi, alpha=1.0 - beta2)
            p.sub_(lr_t * m / (torch.sqrt(v) + eps))

    # spectral: same sparse constants as the kernels, dense torch arithmetic
    def spectrogram_fwd(self, wave, consts, time_steps, frames_per_run):
        b, wave_len = wave.shape
        nsamp = 512 * (time_steps - 1) + 2048
        x = TF.pad(wave, (nsamp - wave_len, 0))
        frames = x.unfold(-1, 2048, 512) * consts["hann"]
        s = torch.fft.rfft(frames, n=2048)[..., 1:]
        mag, ph = torch.abs(s), torch.atan2(s.imag + 0.0, s.real + 0.0)
        k0, w = consts["mel_k0"].long(), consts["mel_w"]
        mm = torch.zeros(b, time_steps, 1024)
        pp = torch.zeros(b, time_steps, 1024)
        for i in range(w.shape[0]):
            idx = (k0 + i).clamp(max=1023)
            mm = mm + mag[..., idx] * w[i]
            pp = pp + ph[..., idx] * w[i]
        logmel = (torch.log(mm + 1e-6) + 3.76) / 10.05
        pi = torch.tensor(math.pi, dtype=torch.float32)
        d = pp[:, 1:] - pp[:, :-1]
        md = torch.remainder(d + pi, 2 * pi) - pi
        md = torch.where((md == -pi) & (d > 0), pi.expand_as(md), md)
        inst = torch.cat([pp[:, :1], md], 1) / pi
        return logmel, inst

    def spectrogram_generic(self, wave, consts, time_steps, bins, frame_step):
        b, wave_len = wave.shape
        n_fft = 2 * bins
        nsamp = frame_step * (time_steps - 1) + n_fft
        x = TF.pad(wave, (nsamp - wave_len, 0))
        s = torch.fft.rfft(x.unfold(-1, n_fft, frame_step) * consts["hann"], n=n_fft)[..., 1:]
        mag, ph = torch.abs(s), torch.atan2(s.imag + 0.0, s.real + 0.0)
        mm, pp = mag @ consts["mel"], ph @ consts["mel"]
        pi = torch.tensor(math.pi, dtype=torch.float32)
        d = pp[:, 1:] - pp[:, :-1]
        md = torch.remainder(d + pi, 2 * pi) - pi
        md = torch.where((md == -pi) & (d > 0), pi.expand_as(md), md)
        return (torch.log(mm + 1e-6) + 3.76) / 10.05, torch.cat([pp[:, :1], md], 1) / pi

    def waveform_generic(self, logmel, inst, consts, wave_len, bins, frame_step):
        b, t, _ = logmel.shape
        n_fft = 2 * bins
        mm = torch.exp(logmel * 10.05 - 3.76)
        pp = torch.cumsum(inst * torch.tensor(math.pi, dtype=torch.float32), 1)
        mag, ph = mm @ consts["pinv"], pp @ consts["pinv"]
        s = TF.pad(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), (1, 0))
        frames = torch.fft.irfft(s, n=n_fft) * consts["synth_window"]
        nsamp = frame_step * (t - 1) + n_fft
        out = torch.zeros(b, nsamp)
        for k in range(t):
            out[:, frame_step * k:frame_step * k + n_fft] += frames[:, k]
        return out[:, nsamp - wave_len:]

    def waveform_fwd(self, logmel, inst, consts, wave_len, frames_per_segment=None):
        b, t, _ = logmel.shape
        mm = torch.exp(logmel * 10.05 - 3.76)
        pp = torch.cumsum(inst * torch.tensor(math.pi, dtype=torch.float32), 1)
        j0, cnt, w = consts["pb_j0"].long(), consts["pb_cnt"], consts["pb_w"]
        mag = torch.zeros(b, t, 1024)
        ph = torch.zeros(b, t, 1024)
        for i in range(w.shape[0]):
            idx = (j0 + i).clamp(max=1023)
            mag = mag + mm[..., idx] * w[i]
            ph = ph + pp[..., idx] * w[i]
        s = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))
        s = TF.pad(s, (1, 0))
        frames = torch.fft.irfft(s, n=2048) * consts["synth_window"]
        nsamp = 512 * (t - 1) + 2048
        out = torch.zeros(b, nsamp)
        for k in range(t):
            out[:, 512 * k:512 * k + 2048] += frames[:, k]
        return out[:, nsamp - wave_len:]
