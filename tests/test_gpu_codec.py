"""Codec side of the decode loop on a real B200: analysis encoder, hyper-encoder,
hyper-decoder (5x5 stride-2 convs, 5x5 stride-2 transposed convs as four output-parity convs, LeakyReLU epilogue) on the
tcgen05 conv kernels, against oracle/codec.py -- teacher-forced stage by stage, against the committed golden vectors,
bitwise run-to-run, and chained into the integer kernels and the decoder."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
TOL = 1e-2  # relative to max(1, |ref|): the north star's per-step tolerance, applied per stage

_cache = {}


def _setup():
    if "s" in _cache:
        return _cache["s"]
    from cdc_b200 import CDCConfig, Codec, Decoder
    from oracle.config import CDCConfig as OCfg
    from oracle.weights import build_codec, build_unet
    torch.set_num_threads(os.cpu_count())
    ocfg = OCfg()
    net = build_unet(ocfg, seed=0)
    oc = build_codec(ocfg, seed=1)
    w = dict(net.state_dict())
    for k, v in oc.state_dict().items():
        w[("context." + k[len("context."):]) if k.startswith("context.") else ("codec." + k)] = v
    dec = Decoder(CDCConfig(), w, device=DEV)
    gauss, fact = oc.tables()
    codec = Codec(dec, gauss_tables=gauss, fact_tables=fact, median=oc.prior.median)
    _cache["s"] = (codec, oc, dec, w)
    return _cache["s"]


def _rel(a, b):
    return ((a.cpu() - b).abs() / b.abs().clamp(min=1.0)).max().item()


@pytest.mark.parametrize("shape", [(1, 128, 128), (2, 192, 256)])
def test_encoder_and_hyper_networks_match_the_oracle_stage_by_stage(shape):
    from oracle.weights import synthetic_image
    codec, oc, _, _ = _setup()
    B, H, W = shape
    img = synthetic_image(B, H, W, index=3)
    with torch.no_grad():
        y_ref = oc.encoder(2.0 * img - 1.0)
        z_ref = oc.hyper_enc(y_ref)
        z_hat_ref = torch.round(z_ref)  # (median 0)
        mu_ref, sg_ref = oc.hyper_dec(z_hat_ref)
    y = codec.analysis(img)
    e_y = _rel(y, y_ref)
    z = codec.hyper_encode(y_ref)            # teacher-forced: the oracle's y
    e_z = _rel(z, z_ref)
    mu, sg = codec.hyper_decode(z_hat_ref)   # teacher-forced: the oracle's z_hat
    e_mu, e_sg = _rel(mu, mu_ref), _rel(sg, sg_ref)
    print(f"{shape}: y {e_y:.5f} (max {y_ref.abs().max().item():.2f}), z {e_z:.5f} (max {z_ref.abs().max().item():.2f}), "
          f"mu {e_mu:.5f}, sigma {e_sg:.5f} (rel. to max(1, |ref|))")
    assert y.shape == y_ref.shape and z.shape == z_ref.shape and mu.shape == mu_ref.shape
    assert max(e_y, e_z, e_mu, e_sg) <= TOL
    assert (sg >= np.float32(0.11)).all()
    # with untrained weights z rounds to 0 and hyper_dec(0) only exercises the biases: drive both hyper networks with
    # inputs of realistic magnitude as well (5x5 stride-2 taps, the four transposed-conv parities, LeakyReLU both ways)
    g = torch.Generator().manual_seed(17)
    y_big = (8.0 * torch.randn(y_ref.shape, generator=g)).bfloat16().float()
    zh_big = torch.round(4.0 * torch.randn(z_ref.shape, generator=g))
    with torch.no_grad():
        z_big_ref = oc.hyper_enc(y_big)
        mu_big_ref, sg_big_ref = oc.hyper_dec(zh_big)
    e_zb = _rel(codec.hyper_encode(y_big), z_big_ref)
    mu_b, sg_b = codec.hyper_decode(zh_big)
    e_mb, e_sb = _rel(mu_b, mu_big_ref), _rel(sg_b, sg_big_ref)
    print(f"{shape}: driven inputs: z {e_zb:.5f} (max {z_big_ref.abs().max().item():.2f}), mu {e_mb:.5f} "
          f"(max {mu_big_ref.abs().max().item():.2f}), sigma {e_sb:.5f} (max {sg_big_ref.abs().max().item():.2f}, "
          f"{(sg_big_ref > 0.11).float().mean().item():.2f} above the clamp)")
    assert max(e_zb, e_mb, e_sb) <= TOL


def test_codec_golden_vectors():
    """tests/golden/codec_128.npz (oracle/make_golden.py): y of the 128 x 128 synthetic image; mu, sigma of its z_hat."""
    from oracle.weights import synthetic_image
    codec, _, _, _ = _setup()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "codec_128.npz"))
    y = codec.analysis(synthetic_image(1, 128, 128))
    mu, sg = codec.hyper_decode(torch.from_numpy(g["qz"].astype(np.float32)))
    e = [_rel(y, torch.from_numpy(g["y"])), _rel(mu, torch.from_numpy(g["mu"])), _rel(sg, torch.from_numpy(g["sigma"]))]
    print(f"golden codec_128: y {e[0]:.5f} mu {e[1]:.5f} sigma {e[2]:.5f}")
    assert max(e) <= TOL


def test_encode_is_deterministic_and_integer_stages_are_bit_exact():
    """Same image -> the same bits, run after run (fixed accumulation order): what lets encoder and decoder agree on
    (mu, sigma).  Given the GPU's fp32 y / mu / sigma, rounding and CDF lookup equal the oracle's integer arithmetic."""
    from oracle import entropy as oe
    from oracle.weights import synthetic_image
    codec, oc, _, _ = _setup()
    img = synthetic_image(2, 128, 192, index=5)
    a, b = codec.encode(img), codec.encode(img)
    for k in ("y", "z", "mu", "sigma", "q", "y_hat", "qz", "z_hat"):
        assert torch.equal(a[k], b[k]), k
    for x, y_ in zip(a["y_sym"] + a["z_sym"], b["y_sym"] + b["z_sym"]):
        assert torch.equal(x, y_)
    gauss, fact = oc.tables()
    q_ref, yh_ref = oe.quantize_symbols(a["y"].cpu(), a["mu"].cpu())
    assert torch.equal(a["q"].cpu(), q_ref) and torch.equal(a["y_hat"].cpu(), yh_ref)
    for x, r in zip(a["y_sym"], oe.cdf_lookup(q_ref, a["sigma"].cpu(), gauss)):
        assert torch.equal(x.cpu(), r)
    med = oc.prior.median.detach()[None, :, None, None].expand_as(a["z"].cpu()).contiguous()
    qz_ref, _ = oe.quantize_symbols(a["z"].cpu(), med)
    assert torch.equal(a["qz"].cpu(), qz_ref)
    ch = torch.arange(256, dtype=torch.int32)[None, :, None, None].expand_as(qz_ref).contiguous()
    for x, r in zip(a["z_sym"], oe.lookup_rows(qz_ref, ch, fact)):
        assert torch.equal(x.cpu(), r)
    bits = oe.estimated_bits(a["y_sym"][2].cpu(), a["y_sym"][3].cpu(), a["y_sym"][4].cpu())
    print(f"estimated bits for y: {bits:.0f} ({bits / (2 * 128 * 192):.3f} bpp); escapes {(a['y_sym'][4] > 0).sum().item()}")


def test_decoder_side_rederives_the_latent_bitwise_and_decodes():
    """The decoder side gets only (qz, q): z_hat -> cdc_hyper_decode -> y_hat must equal the encoder's y_hat bit for bit,
    and the diffusion decoder turns it into an image (J1 pipeline at a test size)."""
    from cdc_b200.synthetic import init_noise
    from oracle.weights import synthetic_image
    codec, oc, dec, _ = _setup()
    img = synthetic_image(1, 128, 128, index=7)
    enc = codec.encode(img)
    y_hat, mu, sigma = codec.latent_from_symbols(enc["qz"].cpu(), enc["q"].cpu())
    assert torch.equal(y_hat, enc["y_hat"]) and torch.equal(mu, enc["mu"]) and torch.equal(sigma, enc["sigma"])
    out = codec.decompress(enc["qz"], enc["q"], 5, init=init_noise(1, 128, 128).to(DEV))
    assert out.shape == (1, 3, 128, 128) and torch.isfinite(out).all() and 0.0 <= out.min().item() and out.max().item() <= 1.0
    # the same latent through the oracle's fp32 decoder (context net + 5 DDIM steps): free-running, reported and loosely gated
    from oracle.sampler import OracleDecoder
    from oracle.weights import build_unet
    from oracle.config import CDCConfig as OCfg
    orc = OracleDecoder(OCfg(), build_unet(OCfg(), seed=0), context_net=oc.context)
    ref = orc.decode(y_hat.cpu(), 5, init=init_noise(1, 128, 128))
    e = (out.cpu() - ref).abs().max().item()
    print(f"decompress vs oracle decode of the same y_hat (5 free-running steps): max-abs {e:.5f}")
    assert e < 0.05


def test_codec_requires_its_weights():
    from cdc_b200 import CDCConfig, Codec, Decoder
    from cdc_b200.synthetic import random_weights
    dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device=DEV)
    with pytest.raises(RuntimeError):
        Codec(dec)
    d2 = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True, with_codec=True), device=DEV)
    w = random_weights(CDCConfig(), seed=0, with_context=True, with_codec=True)
    c2 = Codec(d2, fact_tables=Codec.prior_tables(w))
    from cdc_b200.synthetic import image
    enc = c2.encode(image(1, 64, 64))
    assert enc["q"].dtype == torch.int32 and enc["z_sym"] is not None
    with pytest.raises(ValueError):
        c2.analysis(torch.zeros(1, 3, 100, 128))
