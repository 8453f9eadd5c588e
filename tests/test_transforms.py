"""Intensity transforms / volume metrics / inference / checkpoint wire format.

CPU tests pin the numpy oracle (oracle/transforms.py) against numpy's own percentile and hand-computed vectors and
check the Lightning-style checkpoint round trip; ``-m gpu`` tests compare the CUDA kernels with the oracle --
bit-exact for the order statistics and for the rounded 0..255 volumes."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import transforms as ot

gpu = pytest.mark.gpu


def _volume(shape, seed, background=0.4):
    """Synthetic MRI-like volume: smooth-ish positive intensities with a large exactly-zero background."""
    rng = np.random.RandomState(seed)
    v = rng.gamma(2.0, 150.0, size=shape).astype(np.float32)
    v[rng.rand(*shape) < background] = 0.0
    return v


# ------------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_percentile_matches_numpy():
    v = _volume((17, 19, 23), 0)
    for q in (0, 1, 37.5, 50, 99, 100):
        assert ot.percentile_f32(v, q) == np.float32(np.percentile(v.astype(np.float64), q))
    lo, hi, frac = ot.percentile_ranks(2097152, 1.0)
    assert (lo, hi) == (20971, 20972) and abs(frac - 0.51) < 1e-9


def test_oracle_scale_intensity_known_answers():
    x = np.array([0, 1, 2, 3, 4], dtype=np.float32)
    # 0/100 percentiles = min/max -> [0, 255]
    np.testing.assert_array_equal(ot.to_display_range(x), np.array([0, 64, 128, 191, 255], dtype=np.float32))  # 63.75->64, 127.5->128 (half to even), 191.25->191
    # 1/99 percentiles of 0..4 are 0.04 / 3.96; values outside clip to [-1, 1]
    y = ot.scale_intensity_range_percentiles(x, 1.0, 99.0, -1.0, 1.0, clip=True)
    assert y[0] == -1.0 and y[-1] == 1.0 and abs(float(y[2])) < 1e-6
    # constant image: a_max == a_min -> img - a_min + b_min
    c = np.full(7, 3.0, dtype=np.float32)
    np.testing.assert_array_equal(ot.scale_intensity_range_percentiles(c, 1, 99, -1, 1, clip=True), np.full(7, -1.0, np.float32))
    # relative=True rescales the target range by the percentiles
    r = ot.scale_intensity_range_percentiles(x, 0, 100, 0, 200, clip=False, relative=True)
    np.testing.assert_allclose(r, x / 4 * 200, rtol=1e-6)
    assert ot.mean_absolute_error(x, x[::-1]) == pytest.approx(2.4) and ot.mean_squared_error(x, x[::-1]) == pytest.approx(8.0)


def test_checkpoint_round_trip_and_reference_key_names(tmp_path):
    import mpgan
    from oracle.gan import GANOracle
    torch.manual_seed(0)
    model = mpgan.GAN(1, 64, 64)
    # default key style "adn": MONAI's acti-norm-dropout naming (unitN.adn.N.* / unitN.adn.A.*); both namings load back
    adn = model.save_checkpoint(str(tmp_path / "adn.ckpt"), epoch=3, global_step=12)
    adn_keys = list(torch.load(adn, weights_only=False)["state_dict"].keys())
    assert "generator.model.0.model.0.conv.unit0.adn.N.running_mean" in adn_keys
    assert "generator.model.0.model.0.conv.unit0.adn.A.weight" in adn_keys
    assert "generator.model.0.model.2.0.adn.N.weight" in adn_keys and not any(".norm." in k or ".act." in k for k in adn_keys)
    assert [k for k in adn_keys if k.startswith("discriminator.")] == [k for k in model.state_dict() if k.startswith("discriminator.")]
    back_adn = mpgan.GAN.load_from_checkpoint(adn)
    assert len(back_adn.load_result.missing_keys) == 0 and len(back_adn.load_result.unexpected_keys) == 0
    for k, v in model.state_dict().items():
        assert torch.equal(v, back_adn.state_dict()[k]), k
    path = model.save_checkpoint(str(tmp_path / "gen_epoch=3-g_loss=1.00-d_loss=0.50.ckpt"), epoch=3, global_step=12,
                                 key_style="legacy")
    ckpt = torch.load(path, weights_only=False)
    assert ckpt["pytorch-lightning_version"] == "1.2.1" and ckpt["epoch"] == 3 and ckpt["global_step"] == 12
    assert ckpt["optimizer_states"] == [] and ckpt["lr_schedulers"] == []     # no optimizer was ever configured
    # the reference's (oracle = reference classes restated, pinned in tests/golden) module tree loads it strictly
    ora = GANOracle("final", dims=2, spatial=64)
    assert list(ora.state_dict().keys()) == list(ckpt["state_dict"].keys())
    ora.load_state_dict(ckpt["state_dict"], strict=True)
    # and back: load_from_checkpoint with the reference's extra kwargs (inferrence.py:97-106)
    back = mpgan.GAN.load_from_checkpoint(path, img_shape=(64, 64), strict=False)
    for k, v in model.state_dict().items():
        assert torch.equal(v, back.state_dict()[k]), k
    assert back.hparams.g_lr == model.hparams.g_lr and back.freeze().training is False
    assert not any(p.requires_grad for p in back.parameters())
    # a checkpoint written by the reference side (plain state_dict under "state_dict") loads too
    torch.save({"state_dict": ora.state_dict(), "hyper_parameters": {"channels": 1, "width": 64, "height": 64}},
               str(tmp_path / "ref.ckpt"))
    again = mpgan.GAN.load_from_checkpoint(str(tmp_path / "ref.ckpt"))
    assert torch.equal(again.state_dict()["discriminator.model_linear.1.weight"], ora.state_dict()["discriminator.model_linear.1.weight"])


def test_transforms_refuse_cpu_tensors():
    from mpgan import transforms as T
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.ScaleIntensityRangePercentiles(1, 99, -1, 1, clip=True)(torch.zeros(8))
    with pytest.raises(ValueError):
        T.ScaleIntensityRangePercentiles(-1, 99, -1, 1)


# ------------------------------------------------------------------------------------------------- CUDA kernels
@gpu
@pytest.mark.parametrize("shape,seed", [((128, 128, 128), 1), ((61, 67, 7), 2), ((5,), 3), ((1,), 4)])
def test_order_statistics_exact(shape, seed):
    from mpgan import transforms as T
    v = _volume(shape, seed)
    if v.size > 3:
        v.ravel()[1] = -3.5          # negatives and ties around zero order correctly
        v.ravel()[2] = -0.0
    flat = np.sort(v.ravel())
    n = flat.size
    ranks = sorted({0, n - 1, n // 2, (n - 1) // 100, min(n - 1, (n - 1) // 100 + 1), (99 * (n - 1)) // 100})
    got = T.order_statistics(torch.from_numpy(v).cuda(), ranks)
    assert [np.float32(g) for g in got] == [flat[r] for r in ranks]


@gpu
@pytest.mark.parametrize("lower,upper,b_min,b_max,rnd", [(1.0, 99.0, -1.0, 1.0, False), (0, 100, 0, 255, True)])
def test_scale_intensity_range_percentiles_bit_exact(lower, upper, b_min, b_max, rnd):
    from mpgan import transforms as T
    v = _volume((96, 128, 128), 5)
    x = torch.from_numpy(v).cuda()
    y = T.ScaleIntensityRangePercentiles(lower, upper, b_min, b_max, clip=True)(x, round_half_even=rnd)
    ref = ot.scale_intensity_range_percentiles(v, lower, upper, b_min, b_max, clip=True)
    if rnd:
        ref = np.round(ref)
    assert np.array_equal(y.cpu().numpy(), ref)                       # bit-exact, including the rounded bytes
    assert T.percentiles(x, [lower, upper]) == [float(ot.percentile_f32(v, lower)), float(ot.percentile_f32(v, upper))]
    d = T.ScaleIntensityRangePercentilesd(["t1w"], lower, upper, b_min, b_max, clip=True)({"t1w": x, "other": 1})
    assert d["other"] == 1 and (rnd or torch.equal(d["t1w"], y))


@gpu
def test_rescale_edge_cases_and_metrics():
    from mpgan import transforms as T
    c = torch.full((1000,), 3.0, device="cuda")
    assert torch.equal(T.ScaleIntensityRangePercentiles(1, 99, -1, 1, clip=True)(c), torch.full_like(c, -1.0))
    h = T.rescale_intensity(torch.arange(5, device="cuda").float(), 0, 4, 0, 255, clip=(0, 255), round_half_even=True,
                            out_dtype=torch.float16)
    assert h.dtype == torch.float16 and h.tolist() == [0, 64, 128, 191, 255]
    a, b = _volume((64, 64, 33), 6), _volume((64, 64, 33), 7)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    assert T.mean_absolute_error(ta, tb) == pytest.approx(ot.mean_absolute_error(a, b), rel=1e-9)
    assert T.mean_squared_error(ta, tb) == pytest.approx(ot.mean_squared_error(a, b), rel=1e-9)
    with pytest.raises(RuntimeError):
        T.order_statistics(torch.zeros(0, device="cuda"), [0])
    with pytest.raises(RuntimeError):
        T.error_sums(ta, tb[:, :, :3])


@gpu
def test_inference_volume_matches_oracle_and_batching_is_invisible():
    """BASELINE config 5 in miniature: generator-only eval forward over a stack of 512x512 slices, batched, then the
    reference's display-range post-processing and MAE."""
    import mpgan
    from mpgan import inference
    from oracle.gan import GANOracle
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=512).eval()
    model = mpgan.GAN(1, 512, 512, precision="fp32")
    model.load_state_dict(ora.state_dict())
    model.cuda().freeze()
    g = torch.Generator().manual_seed(3)
    t1 = torch.rand((3, 1, 512, 512), generator=g) * 2 - 1
    t2 = torch.rand((3, 1, 512, 512), generator=g) * 2 - 1
    with torch.no_grad():
        ref = ora.generator(t1)
    out = inference.infer_volume(model, t1.cuda(), batch=2)            # 2 + 1: ragged last batch
    assert out.shape == t1.shape and rel_l2(out, ref) <= 1e-4           # fp32 mode tolerance (north_star)
    one = inference.infer_volume(model.generator, t1.cuda(), batch=64)
    assert torch.equal(one, out)
    gg = inference.GraphedGenerator(model, (2, 1, 512, 512), torch.device("cuda"))
    assert torch.equal(inference.infer_volume(model, t1.cuda(), batch=2, graphed=gg), out)   # graph replay + eager tail                                        # eval-mode BN: batching cannot change a bit
    disp = inference.to_display_range(out)
    assert np.array_equal(disp.cpu().numpy(), ot.to_display_range(out.cpu().numpy()))
    m = inference.evaluate(out, t2.cuda())
    assert m["mae"] == pytest.approx(ot.mean_absolute_error(out.cpu().numpy(), t2.numpy()), rel=1e-9)
    model.set_precision("bf16") if hasattr(model, "set_precision") else None
    model.generator.set_precision("bf16")
    out16 = inference.infer_volume(model, t1.cuda(), batch=3)
    assert rel_l2(out16, ref) <= 1e-2                                   # bf16 tolerance (north_star)
