"""Where does the step-0 d_loss deviation come from?  Compares mine(fp32) with the fp32 and fp64 oracles."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-modality-minipig-gan_b200"))
from mpgan import GAN  # noqa: E402
from oracle.gan import GANOracle, lightning_step, synthetic_batch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
B, S = 2, 64
torch.manual_seed(0)
o32 = GANOracle("final", dims=2, spatial=S)
o64 = copy.deepcopy(o32).double()
init = copy.deepcopy(o32.state_dict())
batch = synthetic_batch(B, 2, S, seed=1)
b64 = {k: v.double() for k, v in batch.items()}
dbatch = {k: v.cuda() for k, v in batch.items()}

for prec in ("fp32", "bf16"):
    mine = GAN(1, S, S, precision=prec)
    mine.load_state_dict(init)
    o32.load_state_dict(init)
    o64.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in init.items()})
    op32, _ = o32.configure_optimizers()
    op64, _ = o64.configure_optimizers()
    l32 = lightning_step(o32, op32, batch, 0)
    l64 = lightning_step(o64, op64, b64, 0)
    lm = mine.fit_batch(dbatch, 0)
    print(prec, "losses mine", [float(x) for x in lm], "o32", [float(x) for x in l32], "o64", [float(x) for x in l64])
    # G parameters after the step
    def pdiff(a, b):
        num = den = 0.0
        for (n1, p1), (n2, p2) in zip(a, b):
            num += float((p1.detach().double().cpu() - p2.detach().double().cpu()).norm()) ** 2
            den += float(p2.detach().double().norm()) ** 2
        return (num / den) ** 0.5
    print("  G params after step: mine vs o64", pdiff(mine.generator.named_parameters(), o64.generator.named_parameters()),
          " o32 vs o64", pdiff(o32.generator.named_parameters(), o64.generator.named_parameters()))
    print("  D params after step: mine vs o64", pdiff(mine.discriminator.named_parameters(), o64.discriminator.named_parameters()),
          " o32 vs o64", pdiff(o32.discriminator.named_parameters(), o64.discriminator.named_parameters()))
    # isolate the D pass: load the oracle's post-G-step state (all of it) into mine and redo only the D-pass forward
    mine2 = GAN(1, S, S, precision=prec)
    o32b = GANOracle("final", dims=2, spatial=S)
    o32b.load_state_dict(init)
    opb, _ = o32b.configure_optimizers()
    for i, net in enumerate((o32b.generator, o32b.discriminator)):
        for p in net.parameters():
            p.requires_grad_(i == 0)
    loss = o32b.training_step(batch, 0, 0)
    loss.backward()
    opb[0].step(), opb[0].zero_grad()
    for p in o32b.parameters():
        p.requires_grad_(True)
    mine2.load_state_dict(o32b.state_dict())
    with torch.no_grad():
        d_ref = o32b.training_step(batch, 0, 1)
        d_mine = mine2.training_step(dbatch, 0, 1)
    print("  D-pass forward only from identical post-G-step weights: mine", float(d_mine), "o32", float(d_ref))
