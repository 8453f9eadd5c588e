"""mpgan: B200-native (sm_100a) implementation of the GAN train/inference hot path of
mbrzus/Cross-Modality-Minipig-Gan behind the reference's own Python classes.

    from mpgan import GAN, CasNetGenerator, Discriminator, PatchDiscriminator

Arithmetic runs in libmpgan_sm100.so (hand-written CUDA, C ABI in include/mpgan.h); torch provides device
memory, streams and torch.distributed only.  There is no CPU fallback.
"""
from . import _lib, ops  # noqa: F401
from .gan import GAN, HostFedStep  # noqa: F401
from .nets import CasNetGenerator, Discriminator, PatchDiscriminator, UNet, DEFAULT_PRECISION  # noqa: F401
from .runtime import FlatAdam, Runtime  # noqa: F401
from . import inference, transforms  # noqa: F401,E402

__all__ = ["GAN", "HostFedStep", "CasNetGenerator", "Discriminator", "PatchDiscriminator", "UNet", "FlatAdam", "Runtime"]
