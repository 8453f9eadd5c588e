"""CPU-side checks: the C-ABI library loads, exports every symbol include/mpgan.h declares, fails loudly without
a GPU, and the host-side mirror reproduces the reference's module tree / initialisation."""
import ctypes
import subprocess

import pytest
import torch

from mpgan import _lib, GAN, CasNetGenerator, Discriminator, PatchDiscriminator
from oracle.gan import GANOracle
from oracle.nets import CasNetGenerator as OGen, Discriminator as ODis, PatchDiscriminator as OPatch


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mpgan.h but not exported"
    assert set(declared) == set(_lib._SIGS), "ctypes signature table out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    assert lib.mpgan_version() == 100


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped binary (B200_PROFILING.md: UTCHMMA, UTMALDG, LDTM)."""
    try:
        out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic
    assert "HMMA.16816" not in out  # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.mpgan_device_ok() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()
    g = CasNetGenerator((1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(1, 1, 32, 32))
    # a raw compute entry point must report an error, not succeed silently
    x = torch.zeros(64)
    rc = lib.mpgan_cast(0, x.data_ptr(), 0, x.data_ptr(), 64, None)
    assert rc != 0 and _lib.last_error()


def test_shape_errors_are_reported():
    lib = _lib.load()
    g = _lib.ConvGeom()
    g.rank = 4
    rc = lib.mpgan_conv_fprop(ctypes.byref(g), 0, None, 1, None, None, None, 1, None)
    assert rc == -1 and "rank" in _lib.last_error()


@pytest.mark.parametrize("dims,size", [(2, 64), (3, 16)])
def test_module_tree_and_init_match_reference(dims, size):
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    g, d = CasNetGenerator(shape), Discriminator(shape, spatial=size)
    torch.manual_seed(0)
    og, od = OGen(shape, 6, dims), ODis(shape, dims=dims, spatial=size)
    for mine, ref in ((g, og), (d, od)):
        sd, rsd = mine.state_dict(), ref.state_dict()
        assert list(sd.keys()) == list(rsd.keys())
        assert all(torch.equal(sd[k], rsd[k]) for k in sd)
        assert [n for n, _ in mine.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_gan_surface_matches_reference():
    torch.manual_seed(0)
    m = GAN(1, 64, 64)
    torch.manual_seed(0)
    o = GANOracle("final", dims=2, spatial=64)
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), o.state_dict().values()))
    assert m.hparams.g_lr == 5e-4 and m.hparams.d_lr == 5e-4 and m.hparams.b1 == 0.5 and m.hparams.one_sided_label_value == 0.9
    opts, sched = m.configure_optimizers()
    assert len(opts) == 2 and sched == []
    for name in ("forward", "adversarial_loss", "reconstruction_loss", "perceptual_loss", "training_step",
                 "configure_optimizers", "on_epoch_end"):
        assert callable(getattr(m, name))
    # literal 3-D reference signature GAN(channels, width, height, depth)
    m3 = GAN(1, 128, 128, 128)
    assert m3.discriminator.model_linear[1].in_features == 256 * 29 ** 3
    p = GAN(1, 32, 32, variant="perceptual")
    assert isinstance(p.discriminator, PatchDiscriminator) and p.hparams.g_lr == 2e-4
    torch.manual_seed(0)
    pd = PatchDiscriminator((1, 16, 16, 16))
    torch.manual_seed(0)
    opd = OPatch((1, 16, 16, 16), dims=3, spatial=16)
    assert all(torch.equal(a, b) for a, b in zip(pd.state_dict().values(), opd.state_dict().values()))
