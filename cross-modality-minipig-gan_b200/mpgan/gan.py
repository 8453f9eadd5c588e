"""``GAN``: the reference's LightningModule surface on top of the B200 kernel plans.

Mirrors /root/reference/code/GAN/GAN_final.py:212-308 (variant="final": BCE + L1, full-image discriminator) and
/root/reference/test_runs/GAN.py:236-447 (variant="perceptual": patch discriminator, discriminator-feature
"perceptual" loss, L1 on patches).  Two ways to drive it:

* the reference's own protocol -- ``training_step(batch, batch_idx, optimizer_idx)`` returns a loss tensor, the
  caller runs ``loss.backward()`` / ``optimizer.step()`` / ``zero_grad()`` (``fit_batch`` below is the
  pytorch-lightning 1.2.1 two-optimizer loop, with ``toggle_optimizer`` semantics);
* ``fused_step(batch)`` -- the same arithmetic as one static sequence of kernels with no autograd and no host
  synchronisation, which ``capture()`` records into a CUDA graph.  ``bench.py`` times this path.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .nets import CasNetGenerator, Discriminator, PatchDiscriminator, DEFAULT_PRECISION, remap_monai_keys
from .runtime import FlatAdam


class _Hparams(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


try:   # the reference's base class (GAN_final.py:212) when pytorch-lightning is installed; a plain nn.Module otherwise
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001  (absent in this image: SURVEY.md section 7)
    _pl = None
    _Base = nn.Module
    HAVE_LIGHTNING = False


class _BCE(torch.autograd.Function):
    """F.binary_cross_entropy(y_hat, y) (mean) with torch's log clamp (GAN_final.py:244-245)."""

    @staticmethod
    def forward(ctx, prob, target):
        p = prob.detach().float().contiguous()
        t = target.detach().float().contiguous().expand_as(p).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=p.device)
        ops.bce_fwd(p, t, 1.0, loss)
        ctx.save_for_backward(p, t)
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        d = ops.bce_bwd(p, t, 1.0, g.contiguous(), torch.empty_like(p))
        return d, None


class _L1(torch.autograd.Function):
    """F.l1_loss(y_hat, y) (mean) (GAN_final.py:247-248)."""

    @staticmethod
    def forward(ctx, a, b):
        ac, bc = a.detach().float().contiguous(), b.detach().float().contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=ac.device)
        ops.l1_fwd(ac, bc, 1.0, loss)
        ctx.save_for_backward(ac, bc)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        da = ops.l1_bwd(a, b, 1.0, g, torch.empty_like(a), False) if ctx.needs_input_grad[0] else None
        db = ops.l1_bwd(b, a, 1.0, g, torch.empty_like(b), False) if ctx.needs_input_grad[1] else None
        return da, db


class GAN(_Base):
    """``pl.LightningModule`` when pytorch-lightning is importable (so ``pl.Trainer.fit`` drives it as it drives the
    reference's class: ``training_step(batch, batch_idx, optimizer_idx)``, ``configure_optimizers`` returning two real
    ``torch.optim.Optimizer``s, ``on_epoch_end``), ``nn.Module`` with the same surface otherwise (``fit_batch`` is the
    Lightning 1.2.1 two-optimizer loop)."""

    def __init__(self, channels, width, height, depth=None, latent_dim: int = 100, d_lr: float = 0.0005,
                 g_lr: float = 0.0005, b1: float = 0.5, b2: float = 0.999, batch_size: int = 64,
                 example_data=None, one_sided_label_value=0.9, variant="final", lr=None, precision=None,
                 n_unet_blocks=None, num_samples=128, roi=16, **kwargs):
        super().__init__()
        assert variant in ("final", "perceptual")
        self.variant = variant
        if variant == "perceptual":  # test_runs/GAN.py:238-250: a single lr (2e-4) for both optimisers
            g_lr = d_lr = 0.0002 if lr is None else lr
        self._ctor_args = dict(channels=channels, width=width, height=height, depth=depth, variant=variant, lr=lr,
                               precision=precision, n_unet_blocks=n_unet_blocks, num_samples=num_samples, roi=roi)
        if HAVE_LIGHTNING:   # GAN_final.py:231 (names are looked up in this frame's locals)
            self.save_hyperparameters("latent_dim", "g_lr", "d_lr", "b1", "b2", "batch_size", "one_sided_label_value")
        else:
            self.hparams = _Hparams(latent_dim=latent_dim, g_lr=g_lr, d_lr=d_lr, b1=b1, b2=b2, batch_size=batch_size,
                                    one_sided_label_value=one_sided_label_value)
        data_shape = (channels, width, height) if depth is None else (channels, width, height, depth)
        self.dims = len(data_shape) - 1
        self.precision = precision or DEFAULT_PRECISION
        self.num_samples, self.roi = num_samples, roi
        if variant == "final":
            self.generator = CasNetGenerator(img_shape=data_shape, n_unet_blocks=n_unet_blocks or 6,
                                             precision=self.precision)
            # the reference hard-codes the Linear fan-in for 128^3 inputs (GAN_final.py:201); it is derived from
            # `width` here so that other (and 2-D) sizes work -- identical at the reference's own 128^3
            self.discriminator = Discriminator(img_shape=data_shape, precision=self.precision, spatial=width)
        else:
            self.generator = CasNetGenerator(img_shape=data_shape, n_unet_blocks=n_unet_blocks or 4,
                                             channels=(32, 64, 128, 256), strides=(2, 2, 2, 2),
                                             precision=self.precision)
            self.discriminator = PatchDiscriminator(img_shape=data_shape, spatial=roi, precision=self.precision)
        if example_data is not None and variant == "final":
            self.example_input_array_test = example_data[0]["t1w"]
            self.example_input_array_train = example_data[1]["t1w"]
        self.logged = {}
        self._rng = np.random.RandomState()  # RandSpatialCropSamplesd's unseeded RandomState
        self._opt = None
        self._graph = None
        self._const_cache = {}
        self.comm = None  # set by mpgan.ddp.attach()
        # fused_step: gradient all-reduces overlapped with compute (mpgan/ddp.py); MPGAN_NO_COMM_OVERLAP=1 = blocking form
        self.overlap_comm = os.environ.get("MPGAN_NO_COMM_OVERLAP", "0") != "1"

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts both MONAI namings of the generator's BatchNorm / PReLU keys (``remap_monai_keys``)."""
        return super().load_state_dict(remap_monai_keys(state_dict, "legacy"), strict=strict, **kw)

    def reference_state_dict(self, key_style="adn"):
        """State dict under the reference's parameter names.  ``key_style``: "adn" (``unitN.adn.N.* / adn.A.*``, MONAI
        with the acti-norm-dropout block -- the 0.4.0 changelog lists it, and the reference pins monai==0.4.0) or
        "legacy" (``unitN.norm.* / act.*``)."""
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}
        return remap_monai_keys(sd, key_style) if key_style == "adn" else sd

    # ---------------------------------------------------------------- Lightning checkpoint wire format (8f, N3)
    _CTOR_KEYS = ("channels", "width", "height", "depth", "latent_dim", "d_lr", "g_lr", "b1", "b2", "batch_size",
                  "one_sided_label_value", "variant", "lr", "precision", "n_unet_blocks", "num_samples", "roi")

    def save_checkpoint(self, path, epoch=0, global_step=0, key_style="adn", optimizers=None):
        """Write a pytorch-lightning 1.2.1 style ``.ckpt`` (a ``torch.save``d dict with ``state_dict`` under the
        reference's parameter names, ``hyper_parameters``, ``optimizer_states`` in ``torch.optim.Adam``'s layout and an
        empty ``lr_schedulers`` list), the file ``ModelCheckpoint`` produces at GAN_final.py:448-472 and
        ``load_from_checkpoint`` reads at inferrence.py:97-106.  Logical (NC[D]HW / OI[D]HW) tensors, so the reference
        can load it back.  ``key_style``: see ``reference_state_dict``."""
        hp = dict(self.hparams)
        hp.update(self._ctor_args)
        opts = optimizers if optimizers is not None else self._opt
        ckpt = {"epoch": int(epoch), "global_step": int(global_step), "pytorch-lightning_version": "1.2.1",
                "state_dict": self.reference_state_dict(key_style),
                "hyper_parameters": hp,
                "optimizer_states": [o.state_dict() for o in opts] if opts else [],
                "lr_schedulers": []}
        torch.save(ckpt, path)
        return path

    def load_optimizer_states(self, states, device=None):
        """Resume: restore Adam's moments / step counters (``ckpt["optimizer_states"]``) into the flat buffers."""
        if self._opt is None:
            self._opt = self.configure_optimizers()[0]
        dev = device or next(self.discriminator.parameters()).device
        for net in (self.generator, self.discriminator):
            net.runtime.ensure(dev)
        for opt, st in zip(self._opt, states):
            opt.load_state_dict(st)
        return self._opt

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams_file=None, strict=True, **kwargs):
        """``LightningModule.load_from_checkpoint`` as the reference calls it (inferrence.py:97-106): constructor
        arguments come from the checkpoint's ``hyper_parameters``, then ``hparams_file`` (yaml), then ``kwargs``;
        unknown keys (``img_shape`` ...) are ignored like the reference's ``**kwargs``."""
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        args = dict(ckpt.get("hyper_parameters", {}) or {})
        if hparams_file is not None:
            import yaml
            with open(hparams_file) as f:
                args.update(yaml.safe_load(f) or {})
        args.update(kwargs)
        ctor = {k: args[k] for k in cls._CTOR_KEYS if k in args}
        for need in ("channels", "width", "height"):
            if need not in ctor:
                raise RuntimeError(f"checkpoint carries no '{need}': pass it to load_from_checkpoint like the reference does")
        model = cls(**ctor)
        state = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        result = model.load_state_dict(state, strict=strict)
        for net in (model.generator, model.discriminator):
            if net._runtime is not None:
                net._runtime.mark_dirty()
        model.load_result = result
        model.optimizer_states = ckpt.get("optimizer_states", [])   # restored by load_optimizer_states() on the device
        return model

    def freeze(self):
        """LightningModule.freeze (inferrence.py:109): eval mode, no parameter gradients."""
        for p in self.parameters():
            p.requires_grad_(False)
        return self.eval()

    # ---------------------------------------------------------------- reference surface
    def forward(self, x):
        return self.generator(x)

    def adversarial_loss(self, y_hat, y):
        return _BCE.apply(y_hat, y)

    def reconstruction_loss(self, y_hat, y):
        return _L1.apply(y_hat, y)

    def perceptual_loss(self, y_hat_activations, y_activations):
        assert set(y_activations.keys()) == set(y_hat_activations.keys())
        running_sum = torch.zeros(1, dtype=torch.float32, device=y_hat_activations[0].device)
        for key in y_activations.keys():
            running_sum = running_sum + _L1.apply(y_activations[key], y_hat_activations[key]) / y_activations[key].numel()
        return running_sum

    def log(self, name, value, **kw):
        self.logged[name] = value.detach()
        if HAVE_LIGHTNING and getattr(self, "trainer", None) is not None:
            super().log(name, value, **kw)

    def sample_patch_origins(self, batch, spatial):
        o = np.empty((batch, self.num_samples, len(spatial)), dtype=np.int64)
        for b in range(batch):
            for s in range(self.num_samples):
                for d, size in enumerate(spatial):
                    o[b, s, d] = self._rng.randint(0, size - self.roi + 1)
        return o

    def _patches(self, vol, origins_dev):
        """(B,1,*S) -> (B*num_samples,1,roi..) via the gather kernel (C == 1: NCHW and channels-last coincide)."""
        b, sp = vol.shape[0], tuple(vol.shape[2:])
        v = vol.contiguous().reshape((b,) + sp + (1,))
        return _PatchGather.apply(v, origins_dev, self.num_samples, self.roi).reshape(
            (b * self.num_samples, 1) + (self.roi,) * len(sp))

    def training_step(self, batch, batch_idx, optimizer_idx, patch_origins=None):
        t1, t2 = batch["t1w"], batch["t2w"]
        n = t1.shape[0]
        dev = t1.device
        if self.variant == "final":
            if optimizer_idx == 0:
                gen = self(t1)
                self.generated_imgs = gen
                valid = torch.ones(n, 1, device=dev)
                g_adv = self.adversarial_loss(self.discriminator(gen), valid)
                self.log("g_adv_loss", g_adv)
                g_rec = self.reconstruction_loss(gen, t2)
                self.log("g_recon_loss", g_rec)
                g_loss = g_adv + g_rec
                self.log("g_loss", g_loss)
                return g_loss
            valid = torch.ones(n, 1, device=dev) * self.hparams.one_sided_label_value
            real_loss = self.adversarial_loss(self.discriminator(t2), valid)
            fake = torch.zeros(n, 1, device=dev)
            fake_loss = self.adversarial_loss(self.discriminator(self(t1).detach()), fake)
            d_loss = (real_loss + fake_loss) / 2
            self.log("d_loss", d_loss)
            return d_loss
        # ---- perceptual variant (test_runs/GAN.py:300-438): the prologue runs for both optimizer indices
        gen = self(t1)
        self.generated_imgs = gen
        if patch_origins is None:
            patch_origins = self.sample_patch_origins(n, tuple(t1.shape[2:]))
        o_dev = torch.as_tensor(np.asarray(patch_origins).reshape(-1, self.dims), dtype=torch.int32, device=dev)
        fake_p, real_p = self._patches(gen, o_dev), self._patches(t2, o_dev)
        m = fake_p.shape[0]
        if optimizer_idx == 0:
            out_f, acts_f = self.discriminator(fake_p)
            _, acts_r = self.discriminator(real_p)
            g_perc = self.perceptual_loss(acts_f, acts_r)
            self.log("g_perceptual_loss", g_perc)
            g_adv = self.adversarial_loss(out_f, torch.ones(m, 1, device=dev))
            self.log("g_adv_loss", g_adv)
            g_rec = self.reconstruction_loss(fake_p, real_p)
            self.log("g_recon_loss", g_rec)
            g_loss = g_adv + g_rec + g_perc
            self.log("g_loss", g_loss)
            return g_loss
        valid = torch.ones(m, 1, device=dev) * self.hparams.one_sided_label_value
        real_loss = self.adversarial_loss(self.discriminator(real_p)[0], valid)
        fake_loss = self.adversarial_loss(self.discriminator(fake_p)[0], torch.zeros(m, 1, device=dev))
        d_loss = (real_loss + fake_loss) / 2
        self.log("d_loss", d_loss)
        return d_loss

    def configure_optimizers(self):
        hp = self.hparams
        opt_g = FlatAdam(self.generator, lr=hp.g_lr, betas=(hp.b1, hp.b2))
        opt_d = FlatAdam(self.discriminator, lr=hp.d_lr, betas=(hp.b1, hp.b2))
        self._opt = [opt_g, opt_d]
        return [opt_g, opt_d], []

    def on_epoch_end(self):
        """GAN_final.py:310-317: two extra train-mode generator forwards (they update the BN running stats);
        returns the images instead of writing them to TensorBoard."""
        outs = []
        for name in ("example_input_array_test", "example_input_array_train"):
            x = getattr(self, name, None)
            if x is not None:
                outs.append(self(x.to(self.discriminator.model_conv[0].weight.device).float()))
        return outs

    # ---------------------------------------------------------------- Lightning 1.2.1 loop
    def fit_batch(self, batch, batch_idx=0, optimizers=None, patch_origins=None):
        """toggle_optimizer -> training_step -> backward -> step -> zero_grad -> untoggle, for opt_idx 0 then 1."""
        if optimizers is None:
            if self._opt is None:
                self._opt = self.configure_optimizers()[0]
            optimizers = self._opt
        nets = (self.generator, self.discriminator)
        losses = []
        for opt_idx, opt in enumerate(optimizers):
            for i, net in enumerate(nets):
                for p in net.parameters():
                    p.requires_grad_(i == opt_idx)
            kw = {"patch_origins": patch_origins} if self.variant != "final" else {}
            loss = self.training_step(batch, batch_idx, opt_idx, **kw)
            loss.backward()
            if self.comm is not None:
                self.comm.allreduce(nets[opt_idx].runtime.grad)
            opt.step()
            opt.zero_grad()
            for net in nets:
                for p in net.parameters():
                    p.requires_grad_(True)
            losses.append(loss.detach())
        return losses

    # ---------------------------------------------------------------- fused static step
    def _consts(self, n, dev):
        """Label tensors (ones / one-sided 0.9 / zeros) per (batch, device).  Never evicted: a captured CUDA graph holds
        raw pointers to the entry it was recorded with, and an eager step with another batch size must not free it."""
        key = (n, str(dev), float(self.hparams.one_sided_label_value))
        c = self._const_cache.get(key)
        if c is None:
            ones = torch.ones(n, 1, device=dev)
            c = (ones, ones * self.hparams.one_sided_label_value, torch.zeros(n, 1, device=dev))
            self._const_cache[key] = c
        return c

    def fused_step(self, batch, logs=None, grad_probe=None):
        """One full two-optimizer training step (variant "final").  Returns ``logs`` (device fp32):
        [g_adv, g_recon, d_real/2, d_fake/2];  g_loss = logs[0]+logs[1], d_loss = logs[2]+logs[3]."""
        if self.variant != "final":
            return self._fused_step_perceptual(batch, logs, grad_probe)
        t1, t2 = batch["t1w"], batch["t2w"]
        G, D, hp = self.generator, self.discriminator, self.hparams
        n, dev = t1.shape[0], t1.device
        ones, soft, zeros = self._consts(n, dev)
        if logs is None:
            logs = torch.zeros(4, dtype=torch.float32, device=dev)
        else:
            logs.zero_()
        t2c = t2.contiguous()
        # ---- optimizer 0: generator (discriminator frozen: data gradient only)
        gen, gplan = G.run_forward(t1, save=True, need_wgrad=True)
        p, dplan = D.run_forward(gen, save=True, need_wgrad=False)
        ops.bce_fwd(p, ones, 1.0, logs[0:1])
        ops.l1_fwd(gen, t2c, 1.0, logs[1:2])
        dprob = ops.bce_bwd(p, ones, 1.0, None, torch.empty_like(p))
        dgen = D.run_backward(dplan, dprob, need_dx=True)
        ops.l1_bwd(gen, t2c, 1.0, None, dgen, True)
        G.run_backward(gplan, dgen, need_dx=False)
        if grad_probe is not None:  # test hook: look at the gradients before the optimiser consumes them
            grad_probe("generator", G)
        # data parallel: G's bucket reduces on the NCCL stream UNDER the discriminator's real-batch forward, which
        # depends on neither the generator's gradients nor its update
        overlap = self.comm is not None and self.overlap_comm
        h_g = self.comm.start(G.runtime.grad) if overlap else None
        if self.comm is not None and not overlap:
            self.comm.allreduce(G.runtime.grad)
        # ---- optimizer 1: discriminator (generator frozen: forward only, train-mode BN)
        p_real, plan_r = D.run_forward(t2, save=True, need_wgrad=True)
        ops.bce_fwd(p_real, soft, 0.5, logs[2:3])
        if overlap:
            self.comm.finish(h_g)
        G.runtime.adam_step(hp.g_lr, hp.b1, hp.b2)
        G.runtime.zero_grad()
        gen2, _ = G.run_forward(t1, save=False, need_wgrad=False)
        p_fake, plan_f = D.run_forward(gen2, save=True, need_wgrad=True)
        ops.bce_fwd(p_fake, zeros, 0.5, logs[3:4])
        D.run_backward(plan_f, ops.bce_bwd(p_fake, zeros, 0.5, None, torch.empty_like(p_fake)), need_dx=False)
        handles = []
        if overlap and grad_probe is None:
            # D's bucket is split at layer 3 (flat order: D1, BN1, D2, BN2 | D3, BN3, D4, BN4, Linear): the upper part is
            # final once layer 3's weight gradient of THIS (last) backward is enqueued, and reduces under layers 2 / 1
            split = D.runtime.offsets[id(D.model_conv[6].weight)]
            grad = D.runtime.grad

            def layer_done(i, wgrad_stream):
                if i == 2:
                    handles.append(self.comm.start(grad[split:], stream=wgrad_stream))

            D.run_backward(plan_r, ops.bce_bwd(p_real, soft, 0.5, None, torch.empty_like(p_real)), need_dx=False,
                           on_layer_done=layer_done)
            handles.append(self.comm.start(grad[:split]))
            for h in handles:
                self.comm.finish(h)
        else:
            D.run_backward(plan_r, ops.bce_bwd(p_real, soft, 0.5, None, torch.empty_like(p_real)), need_dx=False)
            if grad_probe is not None:
                grad_probe("discriminator", D)
            if self.comm is not None:
                self.comm.allreduce(D.runtime.grad)
        D.runtime.adam_step(hp.d_lr, hp.b1, hp.b2)
        D.runtime.zero_grad()
        return logs

    def _fused_step_perceptual(self, batch, logs=None, grad_probe=None):
        """The test_runs/GAN.py:300-438 two-optimizer step as one static kernel sequence.  ``batch`` carries the patch
        origins as ``batch["origins"]`` (int32 device tensor (B*num_samples, dims); sampled on the host like MONAI's
        RandSpatialCropSamplesd when absent).  Returns ``logs`` (device fp32):
        [g_adv, g_recon, g_perceptual, d_real/2, d_fake/2];  g_loss = sum of the first three, d_loss = last two.
        The discriminator's 16 activations stay in their internal channels-last layout: the feature-matching loss is
        an element-wise L1, so no NCHW copies are made and its gradients enter the backward plan in place."""
        t1, t2 = batch["t1w"], batch["t2w"]
        G, D, hp = self.generator, self.discriminator, self.hparams
        n, dev = t1.shape[0], t1.device
        sp = tuple(t1.shape[2:])
        o_dev = batch.get("origins")
        if o_dev is None:
            o_dev = torch.as_tensor(np.asarray(self.sample_patch_origins(n, sp)).reshape(-1, self.dims), dtype=torch.int32,
                                    device=dev)
        m = n * self.num_samples
        ones, soft, zeros = self._consts(m, dev)
        if logs is None:
            logs = torch.zeros(5, dtype=torch.float32, device=dev)
        else:
            logs.zero_()
        patch_shape = (m, 1) + (self.roi,) * self.dims

        def gather(vol):   # (B,1,*S) fp32 -> (m,1,roi..) fp32 patches (one channel: NCHW == channels-last)
            v = vol.contiguous().reshape((n,) + sp + (1,))
            return ops.patch_gather(v, o_dev, self.num_samples, self.roi).reshape(patch_shape)

        real_p = gather(t2)
        # ---- optimizer 0: generator (discriminator frozen)
        gen, gplan = G.run_forward(t1, save=True, need_wgrad=True)
        fake_p = gather(gen)
        out_f, plan_f = D.run_forward(fake_p, save=True, need_wgrad=False, want_acts=True, logical_acts=False)
        _, plan_r = D.run_forward(real_p, save=False, need_wgrad=False, want_acts=True, logical_acts=False)
        acts_f, acts_r = plan_f.extra["acts_raw"], plan_r.extra["acts_raw"]
        dacts = {}
        perc64 = torch.zeros(1, dtype=torch.float64, device=dev)   # the 16 terms span 4 decades: summed in fp64
        for k, (af, ar) in enumerate(zip(acts_f, acts_r)):   # sum_k l1_loss(y[k], y_hat[k]) / numel_k
            # neither the loss terms nor their gradients are materialised here: the discriminator's backward accumulates
            # each gradient into its own tensor (and the loss term into perc64) with ONE pass over the activation pair
            dacts[k] = ops.LazyL1(af.contiguous(), ar.contiguous(), 1.0 / af.numel(), perc64)
        ops.bce_fwd(out_f, ones, 1.0, logs[0:1])
        ops.l1_fwd(fake_p, real_p, 1.0, logs[1:2])
        dprob = ops.bce_bwd(out_f, ones, 1.0, None, torch.empty_like(out_f))
        dfake = D.run_backward(plan_f, dprob, need_dx=True, act_grads=dacts, internal=True)
        logs[2:3].copy_(perc64)
        dfake = dfake.contiguous()
        ops.l1_bwd(fake_p, real_p, 1.0, None, dfake.reshape(patch_shape), True)
        dgen = torch.zeros((n,) + sp + (1,), dtype=torch.float32, device=dev)
        ops.patch_scatter_add(dfake.reshape((m,) + (self.roi,) * self.dims + (1,)), o_dev, self.num_samples, self.roi, dgen)
        G.run_backward(gplan, dgen.reshape(gen.shape), need_dx=False)
        if grad_probe is not None:
            grad_probe("generator", G)
        if self.comm is not None:
            self.comm.allreduce(G.runtime.grad)
        G.runtime.adam_step(hp.g_lr, hp.b1, hp.b2)
        G.runtime.zero_grad()
        # ---- optimizer 1: discriminator (the prologue re-runs the updated generator, forward only)
        gen2, _ = G.run_forward(t1, save=False, need_wgrad=False)
        fake_p2 = gather(gen2)
        p_real, plan_r2 = D.run_forward(real_p, save=True, need_wgrad=True, want_acts=False)
        ops.bce_fwd(p_real, soft, 0.5, logs[3:4])
        p_fake, plan_f2 = D.run_forward(fake_p2, save=True, need_wgrad=True, want_acts=False)
        ops.bce_fwd(p_fake, zeros, 0.5, logs[4:5])
        D.run_backward(plan_f2, ops.bce_bwd(p_fake, zeros, 0.5, None, torch.empty_like(p_fake)), need_dx=False)
        D.run_backward(plan_r2, ops.bce_bwd(p_real, soft, 0.5, None, torch.empty_like(p_real)), need_dx=False)
        if grad_probe is not None:
            grad_probe("discriminator", D)
        if self.comm is not None:
            self.comm.allreduce(D.runtime.grad)
        D.runtime.adam_step(hp.d_lr, hp.b1, hp.b2)
        D.runtime.zero_grad()
        return logs

    def capture(self, batch):
        """Record ``fused_step`` on static input buffers into a CUDA graph.  Returns (graph, static_batch, logs)."""
        dev = batch["t1w"].device
        static = {k: batch[k].clone() for k in ("t1w", "t2w")}
        if self.variant != "final":   # the patch origins are a static device tensor the caller refreshes per step
            o = batch.get("origins")
            if o is None:
                n, sp = static["t1w"].shape[0], tuple(static["t1w"].shape[2:])
                o = torch.as_tensor(np.asarray(self.sample_patch_origins(n, sp)).reshape(-1, self.dims), dtype=torch.int32,
                                    device=dev)
            static["origins"] = o.clone()
        logs = torch.zeros(4 if self.variant == "final" else 5, dtype=torch.float32, device=dev)
        self.generator.runtime.ensure(dev)
        self.discriminator.runtime.ensure(dev)
        self._consts(static["t1w"].shape[0] * (1 if self.variant == "final" else self.num_samples), dev)
        snap = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):  # warm-up (lazy module loading, cudaFuncSetAttribute) outside capture
            self.fused_step(static, logs)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        calls0 = _lib.ABI_CALLS
        with torch.cuda.graph(graph):
            self.fused_step(static, logs)
        self.abi_calls_per_step = _lib.ABI_CALLS - calls0
        self._restore(snap)  # capture does not execute, but keep state exactly as before either way
        self._graph = (graph, static, logs)
        return self._graph

    def _snapshot(self):
        st = []
        for net in (self.generator, self.discriminator):
            rt = net.runtime
            st.append((rt.flat.clone(), rt.exp_avg.clone(), rt.exp_avg_sq.clone(), rt.adam_state.clone(),
                       [b.clone() for b in net.buffers()]))
        return st

    def _restore(self, snap):
        for net, (flat, m, v, a, bufs) in zip((self.generator, self.discriminator), snap):
            rt = net.runtime
            rt.flat.copy_(flat), rt.exp_avg.copy_(m), rt.exp_avg_sq.copy_(v), rt.adam_state.copy_(a)
            rt.grad.zero_()
            for b, s in zip(net.buffers(), bufs):
                b.copy_(s)
            rt.refresh_shadows(force=True)


class HostFedStep:
    """Drives the captured training step (``GAN.capture``) from HOST batches, the way a data loader feeds it: the pinned
    host -> device copies of step i run on a copy stream into one of two device staging sets while step i - 1's graph is
    still executing; the step then copies the staging set into the graph's static inputs (device to device), replays
    the graph and sends the loss scalars to pinned host memory.  Nothing synchronises the host: ``step`` returns the
    pinned ``logs_host`` tensor, valid after ``torch.cuda.synchronize()`` (or an event the caller records)."""

    def __init__(self, model, example_batch):
        if model._graph is None:
            model.capture(example_batch)
        self.model = model
        self.graph, self.static, self.logs = model._graph
        self.keys = [k for k in ("t1w", "t2w") if k in self.static]
        self.copy_stream = torch.cuda.Stream()
        self.stage = [{k: torch.empty_like(self.static[k]) for k in self.keys} for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.logs_host = torch.zeros(self.logs.numel()).pin_memory()
        self.i = 0

    def step(self, host_batch):
        """``host_batch``: {"t1w", "t2w"} pinned fp32 host tensors of the captured shape."""
        j = self.i & 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[j])       # staging set j was consumed by step i - 2
            for k in self.keys:
                self.stage[j][k].copy_(host_batch[k], non_blocking=True)
            self.ready[j].record(self.copy_stream)
        cur.wait_event(self.ready[j])
        for k in self.keys:
            self.static[k].copy_(self.stage[j][k], non_blocking=True)
        self.free[j].record(cur)
        self.graph.replay()
        self.logs_host.copy_(self.logs, non_blocking=True)
        self.i += 1
        return self.logs_host


class _PatchGather(torch.autograd.Function):
    """RandSpatialCropSamplesd + torch.cat (test_runs/GAN.py:313-337): bit-exact gather, deterministic scatter-add."""

    @staticmethod
    def forward(ctx, vol, origins, num_samples, roi):
        ctx.save_for_backward(origins)
        ctx.meta = (tuple(vol.shape), num_samples, roi)
        return ops.patch_gather(vol.detach().contiguous(), origins, num_samples, roi)

    @staticmethod
    def backward(ctx, g):
        (origins,) = ctx.saved_tensors
        shape, num_samples, roi = ctx.meta
        dvol = torch.zeros(shape, dtype=g.dtype, device=g.device)
        ops.patch_scatter_add(g.contiguous(), origins, num_samples, roi, dvol)
        return dvol, None, None, None
