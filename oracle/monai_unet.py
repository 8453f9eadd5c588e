"""Restatement of ``monai==0.4.0`` ``monai.networks.nets.UNet`` (TEST INFRASTRUCTURE ONLY).

The reference builds its generator from MONAI's UNet
(``/root/reference/code/GAN/GAN_final.py:106-114``,
``/root/reference/test_runs/GAN.py:112-120``) but does not vendor MONAI, and
MONAI is pinned to 0.4.0 (``/root/reference/REQUIREMENTS.txt:71``).  MONAI is
not installable in this image, so its published 0.4.0 algorithm is restated
here (pre-``ADN`` refactor: ``conv -> norm -> act`` registered as
``conv``/``norm``/``act`` sub-modules).  PARITY UNPINNED for this file alone:
there is no MONAI wheel to run it against; the invariants the reference itself
states are checked in ``tests/test_oracle.py`` (shape preservation
``generator_test.py:84-88``; 75 parameter tensors / 402 442 parameters per
2-D UNet(16,32,64,128)).

Module/attribute names reproduce the 0.4.0 state-dict keys
(``model.0.conv.unit0.conv.weight`` ...), see SURVEY.md §8c.
"""
import numpy as np
import torch
import torch.nn as nn

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_BN = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class Convolution(nn.Sequential):
    """monai.networks.blocks.Convolution (0.4.0): conv [-> norm(BATCH) -> PReLU]."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3,
                 conv_only=False, is_transposed=False):
        super().__init__()
        padding = (kernel_size - 1) // 2  # same_padding(kernel_size, dilation=1)
        if is_transposed:
            conv = _CONVT[dimensions](in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                      padding=padding, output_padding=strides - 1, bias=True)
        else:
            conv = _CONV[dimensions](in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                     padding=padding, bias=True)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("norm", _BN[dimensions](out_channels))
            self.add_module("act", nn.PReLU())


class ResidualUnit(nn.Module):
    """monai.networks.blocks.ResidualUnit (0.4.0): conv stack + (conv | 1x1 conv | identity) residual."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3, subunits=2,
                 last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        padding = (kernel_size - 1) // 2
        schannels, sstrides = in_channels, strides
        subunits = max(1, subunits)
        for su in range(subunits):
            conv_only = last_conv_only and su == (subunits - 1)
            unit = Convolution(dimensions, schannels, out_channels, strides=sstrides, kernel_size=kernel_size,
                               conv_only=conv_only)
            self.conv.add_module(f"unit{su:d}", unit)
            schannels, sstrides = out_channels, 1
        if np.prod(strides) != 1 or in_channels != out_channels:
            rkernel, rpad = kernel_size, padding
            if np.prod(strides) == 1:  # channel change only: 1x1 conv, no padding
                rkernel, rpad = 1, 0
            self.residual = _CONV[dimensions](in_channels, out_channels, rkernel, strides, rpad, bias=True)

    def forward(self, x):
        res = self.residual(x)
        cx = self.conv(x)
        return cx + res


class SkipConnection(nn.Module):
    """monai.networks.layers.SkipConnection (0.4.0): cat([x, submodule(x)], dim=1)."""

    def __init__(self, submodule, cat_dim=1):
        super().__init__()
        self.submodule = submodule
        self.cat_dim = cat_dim

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], self.cat_dim)


class UNet(nn.Module):
    """monai.networks.nets.UNet (0.4.0) with act=PRELU, norm=BATCH, dropout=0 (the reference's settings)."""

    def __init__(self, dimensions, in_channels, out_channels, channels, strides, kernel_size=3,
                 up_kernel_size=3, num_res_units=0, norm="batch", **_ignored):
        super().__init__()
        assert str(norm).lower().endswith("batch"), "reference uses Norm.BATCH only"
        self.dimensions = dimensions
        self.in_channels, self.out_channels = in_channels, out_channels
        self.channels, self.strides = tuple(channels), tuple(strides)
        self.kernel_size, self.up_kernel_size = kernel_size, up_kernel_size
        self.num_res_units = num_res_units

        def _create_block(inc, outc, channels, strides, is_top):
            c, s = channels[0], strides[0]
            if len(channels) > 2:
                subblock = _create_block(c, c, channels[1:], strides[1:], False)
                upc = c * 2
            else:
                subblock = self._get_bottom_layer(c, channels[1])
                upc = c + channels[1]
            down = self._get_down_layer(inc, c, s, is_top)
            up = self._get_up_layer(upc, outc, s, is_top)
            return nn.Sequential(down, SkipConnection(subblock), up)

        self.model = _create_block(in_channels, out_channels, self.channels, self.strides, True)

    def _get_down_layer(self, in_channels, out_channels, strides, is_top):
        if self.num_res_units > 0:
            return ResidualUnit(self.dimensions, in_channels, out_channels, strides=strides,
                                kernel_size=self.kernel_size, subunits=self.num_res_units)
        return Convolution(self.dimensions, in_channels, out_channels, strides=strides,
                           kernel_size=self.kernel_size)

    def _get_bottom_layer(self, in_channels, out_channels):
        return self._get_down_layer(in_channels, out_channels, 1, False)

    def _get_up_layer(self, in_channels, out_channels, strides, is_top):
        conv = Convolution(self.dimensions, in_channels, out_channels, strides=strides,
                           kernel_size=self.up_kernel_size,
                           conv_only=is_top and self.num_res_units == 0, is_transposed=True)
        if self.num_res_units > 0:
            ru = ResidualUnit(self.dimensions, out_channels, out_channels, strides=1,
                              kernel_size=self.kernel_size, subunits=1, last_conv_only=is_top)
            conv = nn.Sequential(conv, ru)
        return conv

    def forward(self, x):
        return self.model(x)
