"""Random-init weights and synthetic inputs of the right architecture for benchmarking and smoke runs
(there are no trained checkpoints: the reference ships none and there is no network).  Names and shapes
follow the state dict the C ABI expects (include/cdc_b200.h, cdc_load_weights); values are
PyTorch-default-like (uniform +-1/sqrt(fan_in)) and rounded to bf16-representable fp32."""
import math

import torch

from .config import CDCConfig


def _u(g, shape, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return ((torch.rand(shape, generator=g) * 2 - 1) * b).bfloat16().float()


def _conv(d, g, name, cout, cin, k):
    d[name + ".weight"] = _u(g, (cout, cin, k, k), cin * k * k)
    d[name + ".bias"] = _u(g, (cout,), cin * k * k)


def _rb(d, g, name, cin, cout, temb):
    _conv(d, g, name + ".conv1", cout, cin, 3)
    _conv(d, g, name + ".conv2", cout, cout, 3)
    for gn in (".gn1", ".gn2"):
        d[name + gn + ".weight"] = (1.0 + 0.1 * torch.randn(cout, generator=g)).bfloat16().float()
        d[name + gn + ".bias"] = (0.1 * torch.randn(cout, generator=g)).bfloat16().float()
    if temb:
        d[name + ".film.weight"] = _u(g, (2 * cout, temb), temb)
        d[name + ".film.bias"] = _u(g, (2 * cout,), temb)
    if cin != cout:
        _conv(d, g, name + ".res", cout, cin, 1)


def _convt(d, g, name, cin, cout, k):
    d[name + ".weight"] = _u(g, (cin, cout, k, k), cout * k * k)  # ConvTranspose2d layout [in, out, k, k]
    d[name + ".bias"] = _u(g, (cout,), cout * k * k)


def codec_weights(cfg: CDCConfig = CDCConfig(), seed: int = 1):
    """Random-init analysis encoder, hyper-encoder / hyper-decoder and factorised prior under the names the C ABI expects
    ("codec." + the oracle Codec.state_dict() keys)."""
    g = torch.Generator().manual_seed(seed + 100)
    C, c, d = cfg.channels, cfg.latent_ch, {}
    _conv(d, g, "codec.encoder.stem", C[0], 3, 3)
    prev = C[0]
    for i, ch in enumerate(C):
        _rb(d, g, f"codec.encoder.rbs.{i}", prev, ch, 0)
        _conv(d, g, f"codec.encoder.downs.{i}", ch, ch, 3)
        prev = ch
    _conv(d, g, "codec.hyper_enc.c1", c, c, 3)
    _conv(d, g, "codec.hyper_enc.c2", c, c, 5)
    _conv(d, g, "codec.hyper_enc.c3", c, c, 5)
    _convt(d, g, "codec.hyper_dec.t1", c, c, 5)
    _convt(d, g, "codec.hyper_dec.t2", c, c, 5)
    _conv(d, g, "codec.hyper_dec.c3", 2 * c, c, 3)
    # factorised prior (Balle-style cumulative MLP, filters (3,3,3,3), untrained init): host-side only (CDF tables)
    f = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1.0 / 5)
    for i in range(5):
        init = math.log(math.expm1(1.0 / scale / f[i + 1]))
        d[f"codec.prior.mats.{i}"] = torch.full((c, f[i + 1], f[i]), init).bfloat16().float()
        d[f"codec.prior.biases.{i}"] = (torch.rand((c, f[i + 1], 1), generator=g) - 0.5).bfloat16().float()
        if i < 4:
            d[f"codec.prior.factors.{i}"] = torch.zeros(c, f[i + 1], 1)
    d["codec.prior.median"] = torch.zeros(c)
    return d


def random_weights(cfg: CDCConfig = CDCConfig(), seed: int = 0, with_context: bool = True, with_codec: bool = False):
    g = torch.Generator().manual_seed(seed)
    C, te, d = cfg.channels, cfg.temb, {}
    d["temb.lin1.weight"], d["temb.lin1.bias"] = _u(g, (te, 64), 64), _u(g, (te,), 64)
    d["temb.lin2.weight"], d["temb.lin2.bias"] = _u(g, (te, te), te), _u(g, (te,), te)
    _conv(d, g, "stem", C[0], 3 + C[0], 3)
    for i, c in enumerate(C):
        cin = C[0] if i == 0 else C[i - 1] + C[i]
        _rb(d, g, f"down.{i}.rb1", cin, c, te)
        _rb(d, g, f"down.{i}.rb2", c, c, te)
        _conv(d, g, f"down.{i}.down", c, c, 3)
    _rb(d, g, "mid.rb1", C[-1], C[-1], te)
    d["mid.attn.gn.weight"] = (1.0 + 0.1 * torch.randn(C[-1], generator=g)).bfloat16().float()
    d["mid.attn.gn.bias"] = (0.1 * torch.randn(C[-1], generator=g)).bfloat16().float()
    _conv(d, g, "mid.attn.qkv", 3 * C[-1], C[-1], 1)
    _conv(d, g, "mid.attn.proj", C[-1], C[-1], 1)
    _rb(d, g, "mid.rb2", C[-1], C[-1], te)
    prev = C[-1]
    for i in reversed(range(len(C))):
        _conv(d, g, f"up.{i}.up.up", C[i], prev, 3)
        _rb(d, g, f"up.{i}.rb1", 2 * C[i], C[i], te)
        _rb(d, g, f"up.{i}.rb2", C[i], C[i], te)
        prev = C[i]
    _conv(d, g, "final", 3, C[0], 3)
    if with_context:
        prev = cfg.latent_ch
        for i in reversed(range(len(C))):
            _conv(d, g, f"context.ups.{i}.up", C[i], prev, 3)
            _rb(d, g, f"context.rbs.{i}", C[i], C[i], 0)
            prev = C[i]
    if with_codec:
        d.update(codec_weights(cfg, seed + 1))
    return d

def image(B, H, W, index=0):
    """Synthetic image in [0,1]: uniform noise, low-pass filtered twice with a 3x3 box."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1234 + index)
    x = torch.rand(B, 3, H, W, generator=g)
    k = torch.ones(3, 1, 3, 3) / 9.0
    for _ in range(2):
        x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), k, groups=3)
    return x


def latent(B, H, W, index=0, ch=256):
    """y_hat = rint(4 * randn) fp32 [B, ch, H/16, W/16] (SURVEY.md section 8d synthetic inputs)."""
    g = torch.Generator().manual_seed(1000 + index)
    return torch.round(4.0 * torch.randn(B, ch, H // 16, W // 16, generator=g))


def init_noise(B, H, W, index=0, gamma=0.8):
    g = torch.Generator().manual_seed(2000 + index)
    return gamma * torch.randn(B, 3, H, W, generator=g)


def entropy_inputs(n, seed=3000):
    """(y, mu, sigma) of the integer path: sigma = exp(U[ln .05, ln 300]) -- beyond both ends of the
    scale table --, y - mu ~ sigma * randn plus 0.1 % outliers at +-(5..50) sigma (escapes)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g)
    sigma = torch.exp(math.log(0.05) + u * (math.log(300.0) - math.log(0.05)))
    mu = 3.0 * torch.randn(n, generator=g)
    r = sigma * torch.randn(n, generator=g)
    outlier = torch.rand(n, generator=g) < 1e-3
    mag = (5.0 + 45.0 * torch.rand(n, generator=g)) * sigma
    sgn = torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    r = torch.where(outlier, sgn * mag, r)
    return (mu + r).float(), mu.float(), sigma.float()


def gaussian_tables():
    from .entropy_tables import gaussian_tables as build
    return build()
