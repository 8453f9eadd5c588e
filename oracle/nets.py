"""Oracle networks (TEST INFRASTRUCTURE ONLY): rank-generic CPU/fp32 restatement of the reference nets.

* ``CasNetGenerator``        <- /root/reference/code/GAN/GAN_final.py:92-122 (6x UNet(16,32,64,128) + Tanh)
                                 /root/reference/test_runs/GAN.py:94-129     (4x UNet(32,64,128,256) + Tanh)
* ``Discriminator``          <- /root/reference/code/GAN/GAN_final.py:159-209 (full-image D, Linear -> 1)
* ``PatchDiscriminator``     <- /root/reference/test_runs/GAN.py:136-198      (patch D, returns (validity, 16 clones))

The reference is 3-D only (``dimensions=3``, ``Conv3d``); ``dims=2`` builds the "2-D twin" SURVEY.md M2
prescribes for the BASELINE configs: the same layer list with ``Conv3d -> Conv2d`` and the Linear fan-in
recomputed from the input size (256*61*61 at 256^2).  ``dims=3`` with ``spatial=128`` (final D) or ``16``
(patch D) is layer-for-layer the reference class; ``oracle/make_golden.py`` pins that equivalence by loading
one state-dict into both and comparing outputs bit-for-bit.
"""
import torch
import torch.nn as nn

from .monai_unet import UNet

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_BN = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class CasNetGenerator(nn.Module):
    def __init__(self, img_shape, n_unet_blocks=6, dims=3, channels=(16, 32, 64, 128), strides=(2, 2, 2)):
        super().__init__()
        self.img_shape = img_shape
        blocks = [UNet(dimensions=dims, in_channels=1, out_channels=1, channels=channels, strides=strides,
                       num_res_units=2, norm="batch") for _ in range(n_unet_blocks)]
        blocks.append(nn.Tanh())
        self.model = nn.Sequential(*blocks)

    def forward(self, x):
        return self.model(x)


def _valid_out(size, k, s):
    return (size - k) // s + 1


class Discriminator(nn.Module):
    """GAN_final.py:159-209.  conv(k3,s1) 1->64->128, conv(k4,s2) 128->256->256, all valid padding."""

    LAYERS = ((1, 64, 3, 1), (64, 128, 3, 1), (128, 256, 4, 2), (256, 256, 4, 2))

    def __init__(self, img_shape, use_perceptual=True, dims=3, spatial=128):
        super().__init__()
        self.use_perceptual = use_perceptual
        mods, s = [], spatial
        for cin, cout, k, st in self.LAYERS:
            mods += [_CONV[dims](cin, cout, kernel_size=k, stride=st), _BN[dims](cout),
                     nn.LeakyReLU(0.2, inplace=True)]
            s = _valid_out(s, k, st)
        self.model_conv = nn.Sequential(*mods)
        self.model_linear = nn.Sequential(nn.Flatten(), nn.Linear(256 * s ** dims, 1), nn.Sigmoid())

    def forward(self, img):
        return self.model_linear(self.model_conv(img))


class PatchDiscriminator(nn.Module):
    """test_runs/GAN.py:136-198.  4x conv(k3,s1,valid) 1->64->128->256->512, Linear(512*8^d, 64), Linear(64,1)."""

    LAYERS = ((1, 64, 3, 1), (64, 128, 3, 1), (128, 256, 3, 1), (256, 512, 3, 1))

    def __init__(self, img_shape, use_perceptual=True, dims=3, spatial=16):
        super().__init__()
        self.use_perceptual = use_perceptual
        mods, s = [], spatial
        for cin, cout, k, st in self.LAYERS:
            mods += [_CONV[dims](cin, cout, kernel_size=k, stride=st), _BN[dims](cout),
                     nn.LeakyReLU(0.2, inplace=True)]
            s = _valid_out(s, k, st)
        self.model_conv = nn.Sequential(*mods)
        self.model_linear = nn.Sequential(nn.Flatten(), nn.Linear(512 * s ** dims, 64), nn.Linear(64, 1),
                                          nn.Sigmoid())

    def forward(self, x):
        # every sub-module output is cloned *before* the next (in-place LeakyReLU) module overwrites it,
        # test_runs/GAN.py:183-198
        acts, idx = {}, 0
        for module in list(self.model_conv) + list(self.model_linear):
            x = module(x)
            if self.use_perceptual:
                acts[idx] = x.clone()
                idx += 1
        return x, acts
