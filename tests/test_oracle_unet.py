"""Cross-checks of the oracle's composite blocks against torch.nn.functional primitives and
the committed golden vectors.  No reference tests exist to mirror."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle.config import CDCConfig
from oracle.sampler import OracleDecoder
from oracle.unet import RB, Attn, Up, sinusoidal, unet_flops
from oracle.weights import build_codec, build_unet, synthetic_cond, synthetic_image, synthetic_init

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CFG = CDCConfig()


def test_flops_match_survey():
    assert abs(unet_flops(CFG, 1, 256, 256) / 1e9 - 163.84) < 0.01
    assert abs(unet_flops(CFG, 1, 512, 768) / 1e9 - 985.04) < 0.01
    assert abs(unet_flops(CFG, 16, 256, 256) / 1e9 - 2621.46) < 0.01
    assert abs(unet_flops(CFG, 1, 2048, 2048) / 1e9 - 10756.21) < 0.01


def test_weights_bf16_exact_and_param_count():
    net = build_unet(CFG)
    n = sum(p.numel() for p in net.parameters())
    assert abs(n / 1e6 - 17.9) < 0.05
    for p in net.parameters():
        assert torch.equal(p, p.bfloat16().float())


def test_attn_matches_sdpa():
    torch.manual_seed(0)
    a = Attn(256, 32, 4, 64).eval()
    x = torch.randn(2, 256, 6, 5)
    with torch.no_grad():
        got = a(x)
        n = F.group_norm(x, 32, a.gn.weight, a.gn.bias, 1e-5)
        qkv = F.conv2d(n, a.qkv.weight, a.qkv.bias)
        q, k, v = [t.reshape(2, 4, 64, 30).transpose(2, 3) for t in qkv.chunk(3, dim=1)]
        o = F.scaled_dot_product_attention(q, k, v).transpose(2, 3).reshape(2, 256, 6, 5)
        want = x + F.conv2d(o, a.proj.weight, a.proj.bias)
    assert (got - want).abs().max() < 1e-5


def test_rb_matches_functional():
    torch.manual_seed(1)
    rb = RB(96, 64, 32, 256).eval()
    x, te = torch.randn(2, 96, 8, 8), torch.randn(2, 256)
    with torch.no_grad():
        got = rb(x, te)
        h = F.conv2d(x, rb.conv1.weight, rb.conv1.bias, padding=1)
        ss = F.linear(F.silu(te), rb.film.weight, rb.film.bias)
        s, sh = ss[:, :64], ss[:, 64:]
        h = F.group_norm(h, 32, rb.gn1.weight, rb.gn1.bias, 1e-5) * (1 + s[:, :, None, None]) + sh[:, :, None, None]
        h = F.conv2d(F.silu(h), rb.conv2.weight, rb.conv2.bias, padding=1)
        h = F.silu(F.group_norm(h, 32, rb.gn2.weight, rb.gn2.bias, 1e-5))
        want = h + F.conv2d(x, rb.res.weight, rb.res.bias)
    assert (got - want).abs().max() < 1e-5


def test_up_is_nearest_then_conv():
    torch.manual_seed(2)
    up = Up(8, 4).eval()
    x = torch.randn(1, 8, 3, 5)
    with torch.no_grad():
        want = F.conv2d(x.repeat_interleave(2, 2).repeat_interleave(2, 3), up.up.weight, up.up.bias, padding=1)
        assert (up(x) - want).abs().max() < 1e-6


def test_sinusoidal_layout():
    e = sinusoidal(torch.tensor([0, 7]), 64)
    assert e.shape == (2, 64)
    assert torch.allclose(e[0, :32], torch.zeros(32)) and torch.allclose(e[0, 32:], torch.ones(32))
    assert abs(float(e[1, 0]) - np.sin(7.0)) < 1e-6 and abs(float(e[1, 32]) - np.cos(7.0)) < 1e-6


def test_cfg1_step_matches_golden():
    g = np.load(os.path.join(GOLD, "cfg1_step.npz"))
    dec = OracleDecoder(CFG, build_unet(CFG, seed=0))
    dec.set_sample_schedule(17)
    x = synthetic_init(1, 256, 256)
    cond = synthetic_cond(CFG, 1, 256, 256)
    for t in (999, 0):
        x0 = dec.predict_x0(x, t, cond)[:, :, ::8, ::8].numpy()
        xp = dec.denoise_step(x, t, cond)[:, :, ::8, ::8].numpy()
        assert np.abs(x0 - g[f"x0_t{t}"]).max() < 2e-4
        assert np.abs(xp - g[f"xprev_t{t}"]).max() < 2e-4
    try:
        dec.denoise_step(x, 998, cond)
        assert False, "t outside the schedule must raise"
    except ValueError:
        pass


def test_codec_matches_golden():
    g = np.load(os.path.join(GOLD, "codec_128.npz"))
    codec = build_codec(CFG, seed=1)
    enc = codec.encode(synthetic_image(1, 128, 128))
    assert np.abs(enc["y"].numpy() - g["y"]).max() < 1e-4
    # integers: identical wherever y-mu is not within float noise of a rounding boundary
    d = (enc["q"].numpy() != g["q"]).mean()
    assert d < 1e-3
    with torch.no_grad():
        ctx = codec.context(torch.from_numpy(g["q"]).float() + torch.from_numpy(g["mu"]))
    assert np.abs(ctx[3].numpy() - g["c3"]).max() < 1e-4
    assert ctx[0].shape == (1, 64, 128, 128)
