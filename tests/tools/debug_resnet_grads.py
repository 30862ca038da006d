"""Per-variable gradient parity of networks.ResNet (CUDA) against the oracle (debug helper)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gansynth_b200.ops as ops, gansynth_b200.networks as pnet
from oracle import networks as onet
store = ops.set_default_store(ops.VariableStore(device="cuda", seed=0))
cfg = dict(conv_param=dict(filters=8, kernel_size=[7, 7], strides=[2, 2]), pool_param=dict(kernel_size=[3, 3], strides=[2, 2]),
           residual_params=[dict(filters=8, strides=[1, 1], blocks=2), dict(filters=16, strides=[2, 2], blocks=2)], groups=4, classes=11)
o = onet.ResNet(**cfg); params = o.init_variables(seed=5); net = pnet.ResNet(**cfg)
images = torch.randn(4, 2, 32, 64, generator=torch.Generator().manual_seed(1))
net(images.cuda()); store.load(params)
for v in store.vars.values():
    v.requires_grad_(True)
f, l = net(images.cuda()); loss = (l * l).sum()
names = list(store.vars)
g = torch.autograd.grad(loss, [store.vars[n] for n in names], allow_unused=True)
po = {n: p.double().requires_grad_(True) for n, p in params.items()}
fo, lo = o(po, images.double()); go = torch.autograd.grad((lo * lo).sum(), [po[n] for n in names])
print("logits err", float((l.cpu().double() - lo).abs().max() / lo.abs().max()))
for n, a, b in zip(names, g, go):
    e = float((a.cpu().double() - b).abs().max() / (b.abs().max() + 1e-30)) if a is not None else None
    print("%-60s %s" % (n, "None" if e is None else "%.2e" % e))
