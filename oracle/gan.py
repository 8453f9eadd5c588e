"""Oracle GAN step (TEST INFRASTRUCTURE ONLY): CPU/fp32 restatement of the reference's LightningModule.

Follows
* /root/reference/code/GAN/GAN_final.py:212-308   (``GAN``: BCE + L1, full-image D)      -> variant="final"
* /root/reference/test_runs/GAN.py:236-447        (patch D + perceptual + L1 on patches) -> variant="perceptual"
* pytorch-lightning==1.2.1 two-optimizer loop (/root/reference/REQUIREMENTS_updated.txt:131; call sites
  GAN_final.py:480-492): per batch, for opt_idx in (0, 1): toggle_optimizer -> training_step -> backward ->
  optimizer.step -> zero_grad -> untoggle.
* monai==0.4.0 ``RandSpatialCropSamplesd`` (call site test_runs/GAN.py:263-272,320): per sample, per spatial
  dim in order, ``origin = R.randint(0, size - roi + 1)``; same origin applied to every key.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .nets import CasNetGenerator, Discriminator, PatchDiscriminator


def sample_patch_origins(rng, batch, num_samples, spatial_shape, roi):
    """MONAI RandSpatialCropd.randomize order: for each volume, for each sample, for each dim."""
    out = np.empty((batch, num_samples, len(spatial_shape)), dtype=np.int64)
    for b in range(batch):
        for s in range(num_samples):
            for d, size in enumerate(spatial_shape):
                out[b, s, d] = rng.randint(0, size - roi + 1)
    return out


def gather_patches(vol, origins, roi):
    """vol (B,C,*S) -> (B*num_samples, C, roi, ...) by plain slicing (exact copy), volume-major order."""
    patches = []
    for b in range(vol.shape[0]):
        for o in origins[b]:
            sl = (slice(None),) + tuple(slice(int(x), int(x) + roi) for x in o)
            patches.append(vol[b][sl].unsqueeze(0))
    return torch.cat(patches, dim=0)


class GANOracle(nn.Module):
    def __init__(self, variant="final", dims=2, spatial=256, n_unet_blocks=None, d_lr=5e-4, g_lr=5e-4,
                 lr=2e-4, b1=0.5, b2=0.999, one_sided_label_value=0.9, roi=16, num_samples=128,
                 channels=None, strides=None):
        super().__init__()
        self.variant, self.dims, self.spatial = variant, dims, spatial
        self.b1, self.b2, self.one_sided = b1, b2, one_sided_label_value
        self.roi, self.num_samples = roi, num_samples
        shape = (1,) + (spatial,) * dims
        if variant == "final":
            self.g_lr, self.d_lr = g_lr, d_lr
            self.generator = CasNetGenerator(shape, n_unet_blocks or 6, dims, channels or (16, 32, 64, 128),
                                             strides or (2, 2, 2))
            self.discriminator = Discriminator(shape, dims=dims, spatial=spatial)
        else:
            self.g_lr = self.d_lr = lr
            self.generator = CasNetGenerator(shape, n_unet_blocks or 4, dims, channels or (32, 64, 128, 256),
                                             strides or (2, 2, 2, 2))
            self.discriminator = PatchDiscriminator(shape, dims=dims, spatial=roi)
        self.logged = {}

    def forward(self, x):
        return self.generator(x)

    # -- losses: GAN_final.py:244-248, test_runs/GAN.py:281-298
    def adversarial_loss(self, y_hat, y):
        return F.binary_cross_entropy(y_hat, y)

    def reconstruction_loss(self, y_hat, y):
        return F.l1_loss(y_hat, y)

    def perceptual_loss(self, y_hat_acts, y_acts):
        assert set(y_acts.keys()) == set(y_hat_acts.keys())
        total = torch.zeros(1, dtype=y_hat_acts[0].dtype)
        for k in y_acts.keys():
            total = total + F.l1_loss(y_acts[k], y_hat_acts[k]) / y_acts[k].numel()
        return total

    def log(self, name, value):
        self.logged[name] = value.detach().clone()

    def training_step(self, batch, batch_idx, optimizer_idx, patch_origins=None):
        t1, t2 = batch["t1w"], batch["t2w"]
        n = t1.shape[0]
        if self.variant == "final":
            if optimizer_idx == 0:
                gen = self(t1)
                self.generated_imgs = gen
                g_adv = self.adversarial_loss(self.discriminator(gen), torch.ones(n, 1).type_as(t1))
                g_rec = self.reconstruction_loss(gen, t2)
                g_loss = g_adv + g_rec
                self.log("g_adv_loss", g_adv), self.log("g_recon_loss", g_rec), self.log("g_loss", g_loss)
                return g_loss
            real = self.adversarial_loss(self.discriminator(t2), (torch.ones(n, 1) * self.one_sided).type_as(t1))
            fake = self.adversarial_loss(self.discriminator(self(t1).detach()), torch.zeros(n, 1).type_as(t1))
            d_loss = (real + fake) / 2
            self.log("d_loss", d_loss)
            return d_loss
        # perceptual variant: prologue runs for both optimizer indices (test_runs/GAN.py:308-337)
        gen = self(t1)
        self.generated_imgs = gen
        fake_p = gather_patches(gen, patch_origins, self.roi)
        real_p = gather_patches(t2, patch_origins, self.roi)
        m = fake_p.shape[0]
        if optimizer_idx == 0:
            out_f, acts_f = self.discriminator(fake_p)
            _, acts_r = self.discriminator(real_p)
            g_perc = self.perceptual_loss(acts_f, acts_r)
            g_adv = self.adversarial_loss(out_f, torch.ones(m, 1).type_as(t1))
            g_rec = self.reconstruction_loss(fake_p, real_p)
            g_loss = g_adv + g_rec + g_perc
            self.log("g_perceptual_loss", g_perc), self.log("g_adv_loss", g_adv)
            self.log("g_recon_loss", g_rec), self.log("g_loss", g_loss)
            return g_loss
        real = self.adversarial_loss(self.discriminator(real_p)[0], (torch.ones(m, 1) * self.one_sided).type_as(t1))
        fake = self.adversarial_loss(self.discriminator(fake_p)[0], torch.zeros(m, 1).type_as(t1))
        d_loss = (real + fake) / 2
        self.log("d_loss", d_loss)
        return d_loss

    def configure_optimizers(self):
        opt_g = torch.optim.Adam(self.generator.parameters(), lr=self.g_lr, betas=(self.b1, self.b2))
        opt_d = torch.optim.Adam(self.discriminator.parameters(), lr=self.d_lr, betas=(self.b1, self.b2))
        return [opt_g, opt_d], []


def lightning_step(model, optimizers, batch, batch_idx, patch_origins=None, keep_grads=None):
    """One batch of the Lightning 1.2.1 multi-optimizer loop.  ``keep_grads`` (dict) receives a clone of every
    parameter gradient right before the optimizer consumes it (name -> tensor)."""
    nets = (model.generator, model.discriminator)
    losses = []
    for opt_idx, opt in enumerate(optimizers):
        # toggle_optimizer: only the current optimizer's parameters require grad
        for i, net in enumerate(nets):
            for p in net.parameters():
                p.requires_grad_(i == opt_idx)
        kw = {"patch_origins": patch_origins} if model.variant != "final" else {}
        loss = model.training_step(batch, batch_idx, opt_idx, **kw)
        loss.backward()
        if keep_grads is not None:
            prefix = "generator." if opt_idx == 0 else "discriminator."
            for name, p in nets[opt_idx].named_parameters():
                keep_grads[prefix + name] = p.grad.detach().clone()
        opt.step()
        opt.zero_grad()
        for net in nets:  # untoggle
            for p in net.parameters():
                p.requires_grad_(True)
        losses.append(loss.detach().clone())
    return losses


def synthetic_batch(batch, dims, spatial, seed=1):
    """SURVEY.md §8d synthetic inputs: uniform [-1,1], T1 then T2 from one seeded generator."""
    g = torch.Generator().manual_seed(seed)
    shape = (batch, 1) + (spatial,) * dims
    t1 = torch.rand(shape, generator=g) * 2 - 1
    t2 = torch.rand(shape, generator=g) * 2 - 1
    return {"t1w": t1, "t2w": t2}
