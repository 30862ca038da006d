import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gansynth_b200.functional as F
import gansynth_b200.models as pmodels
import gansynth_b200.networks as pnet
import gansynth_b200.ops as ops
from common import HYPER, SMALL
from oracle import networks as onet
g = torch.Generator().manual_seed(5)
batches = [(0.5 * torch.randn(4, 512, generator=g), torch.nn.functional.one_hot(torch.randint(0, 61, (4,), generator=g), 61).float(),
            torch.randn(4, 256, generator=g), torch.randn(4, 256, generator=g)) for _ in range(6)]
F.K.impl = int(os.environ.get("IMPL", "4"))
for use_graphs in (False, True):
    store = ops.set_default_store(ops.VariableStore(device="cuda", seed=0))
    pmodels.reset_global_step()
    opg = onet.PGGAN(growing_level=1.0, **SMALL)
    params = opg.init_variables(seed=3, bias_std=0.1)
    ppg = pnet.PGGAN(growing_level=1.0, **SMALL)
    ppg._ensure_variables("generator", 256, 61); ppg._ensure_variables("discriminator", 0, 61)
    store.load(params)
    model = pmodels.GANSynth(ppg.generator, ppg.discriminator, None, None, {}, HYPER)
    model.use_cuda_graphs = use_graphs
    model.real_images_from_waveforms = lambda w: w.reshape(4, 2, 16, 16)
    out = []
    for w, lab, z1, z2 in batches:
        d = model.discriminator_step(w.cuda(), lab.cuda(), z1.cuda())
        d_now = float(d) if os.environ.get("READ_NOW") else None
        gl = model.generator_step(lab.cuda(), z2.cuda())
        out.append("%.6f(%s)/%.6f" % (float(d), "-" if d_now is None else "%.6f" % d_now, float(gl)))
    print("graphs" if use_graphs else "eager ", " ".join(out))
