"""Host-side logic that needs no GPU: convolution geometry of the plan builder, percentile rank selection, checkpoint
hyper-parameter resolution."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from mpgan import ops
from mpgan import transforms as T
from oracle import transforms as ot


@pytest.mark.parametrize("k,s,p,size", [(3, 1, 0, 256), (3, 1, 1, 128), (3, 2, 1, 256), (4, 2, 0, 252), (4, 2, 0, 125), (1, 1, 0, 32)])
def test_conv_spec_matches_torch_output_sizes(k, s, p, size):
    conv = nn.Conv2d(2, 3, k, stride=s, padding=p)
    spec = ops.ConvSpec.from_module(conv)
    y = conv(torch.zeros(1, 2, size, size))
    assert spec.y_of_x((size, size)) == tuple(y.shape[2:]) and not spec.transposed and (spec.cx, spec.cy) == (2, 3)
    g = spec.geom(5, (size, size), tuple(y.shape[2:]))
    assert (g.rank, g.n, g.cx, g.cy) == (2, 5, 2, 3) and list(g.xs) == [1, size, size] and list(g.k) == [1, k, k]
    if s == 2 and k == 3:   # MONAI's up path: ConvTranspose(k3, s2, p1, output_padding 1) doubles the size
        ct = nn.ConvTranspose2d(3, 2, k, stride=s, padding=p, output_padding=s - 1)
        st = ops.ConvSpec.from_module(ct)
        out = ct(torch.zeros(1, 3, size // 2, size // 2))
        assert st.transposed and (st.cx, st.cy) == (2, 3) and st.x_of_y((size // 2, size // 2)) == tuple(out.shape[2:])


def test_conv_spec_rank3():
    conv = nn.Conv3d(1, 64, 3)
    spec = ops.ConvSpec.from_module(conv)
    assert spec.rank == 3 and spec.taps == 27 and spec.y_of_x((128, 128, 128)) == (126, 126, 126)   # GAN_final.py:167-169


@pytest.mark.parametrize("n", [1, 2, 5, 2097152])
@pytest.mark.parametrize("q", [0, 1.0, 37.5, 99.0, 100])
def test_percentile_rank_selection_matches_numpy_positions(n, q):
    lo, hi, frac = T.percentile_ranks(n, q)
    assert (lo, hi, frac) == ot.percentile_ranks(n, q)
    pos = q / 100.0 * (n - 1)
    assert lo == int(np.floor(pos)) and 0 <= lo <= hi <= n - 1 and hi - lo in (0, 1) and abs(lo + frac - pos) < 1e-9
    if q in (0, 100) or n == 1:
        assert hi == lo and frac == 0          # exact percentiles need one order statistic (min / max fast path)


def test_load_from_checkpoint_resolves_hparams_like_the_reference(tmp_path):
    """inferrence.py:97-106 passes channels/width/height/depth, an hparams.yaml and img_shape next to the checkpoint."""
    import mpgan
    torch.manual_seed(0)
    m = mpgan.GAN(1, 32, 32, g_lr=1e-3)
    ckpt = str(tmp_path / "m.ckpt")
    m.save_checkpoint(ckpt)
    yml = tmp_path / "hparams.yaml"
    yml.write_text("d_lr: 0.002\nb1: 0.4\nlatent_dim: 100\n")
    back = mpgan.GAN.load_from_checkpoint(channels=1, width=32, height=32, checkpoint_path=ckpt, hparams_file=str(yml),
                                          img_shape=(32, 32), strict=False)
    assert back.hparams.g_lr == 1e-3 and back.hparams.d_lr == 0.002 and back.hparams.b1 == 0.4
    opts, scheds = back.configure_optimizers()
    assert len(opts) == 2 and scheds == []
    with pytest.raises(RuntimeError, match="channels"):
        torch.save({"state_dict": m.state_dict()}, str(tmp_path / "bare.ckpt"))
        mpgan.GAN.load_from_checkpoint(str(tmp_path / "bare.ckpt"))
