"""Seeded bf16-exact weights and synthetic inputs shared by the oracle and the CUDA
path (SURVEY.md Appendix A.6 and section 8d "Synthetic inputs").

Reference file:line: none -- /root/reference/README.md is 0 bytes (no weights ship).
Test infrastructure only; see oracle/__init__.py.  bench.py / tests use it to make
the state dict handed to the product through its public `Decoder(weights=...)`.
"""
import torch
import torch.nn.functional as F

from .codec import Codec
from .config import CDCConfig
from .unet import UNet


def _bf16_exact_(module):
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.bfloat16().float())
    return module


def build_unet(cfg: CDCConfig = CDCConfig(), seed: int = 0) -> UNet:
    """torch.manual_seed(seed); default inits; every parameter rounded to bf16-representable fp32.
    GroupNorm affine params and the FiLM linears are perturbed away from their (1,0)/(default)
    inits so that parity tests exercise gamma/beta/scale/shift rather than identities."""
    torch.manual_seed(seed)
    net = UNet(cfg)
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if ".gn" in name and name.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif ".gn" in name and name.endswith("bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return _bf16_exact_(net).eval()


def build_codec(cfg: CDCConfig = CDCConfig(), seed: int = 1) -> Codec:
    torch.manual_seed(seed)
    return _bf16_exact_(Codec(cfg)).eval()


def synthetic_image(B, H, W, index=0):
    """x ~ U[0,1], low-pass filtered (3x3 box, twice), Generator(seed=1234+index)."""
    g = torch.Generator().manual_seed(1234 + index)
    x = torch.rand(B, 3, H, W, generator=g)
    k = torch.ones(3, 1, 3, 3) / 9.0
    for _ in range(2):
        x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), k, groups=3)
    return x


def synthetic_latent(B, H, W, index=0, ch=256):
    """y_hat = rint(4*randn) fp32 [B,ch,H/16,W/16], seed 1000+index."""
    g = torch.Generator().manual_seed(1000 + index)
    return torch.round(4.0 * torch.randn(B, ch, H // 16, W // 16, generator=g))


def synthetic_cond(cfg: CDCConfig, B, H, W, index=0):
    """Stand-in context maps for kernel-level tests that do not want to run the context
    net: c_i ~ 0.5*randn rounded to bf16, [B, C[i], H/2^i, W/2^i], seed 1500+index."""
    g = torch.Generator().manual_seed(1500 + index)
    out = []
    for i, c in enumerate(cfg.channels):
        t = 0.5 * torch.randn(B, c, H >> i, W >> i, generator=g)
        out.append(t.bfloat16().float())
    return tuple(out)


def synthetic_init(B, H, W, index=0, gamma=0.8):
    """x_T = gamma * randn, seed 2000+index."""
    g = torch.Generator().manual_seed(2000 + index)
    return gamma * torch.randn(B, 3, H, W, generator=g)


def synthetic_entropy_inputs(n, seed=3000):
    """sigma = exp(U[ln .05, ln 300]); y-mu ~ sigma*randn + 0.1% outliers at +-(5..50) sigma."""
    g = torch.Generator().manual_seed(seed)
    import math
    u = torch.rand(n, generator=g)
    sigma = torch.exp(math.log(0.05) + u * (math.log(300.0) - math.log(0.05)))
    mu = 3.0 * torch.randn(n, generator=g)
    r = sigma * torch.randn(n, generator=g)
    out = torch.rand(n, generator=g) < 1e-3
    mag = (5.0 + 45.0 * torch.rand(n, generator=g)) * sigma
    sgn = torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    r = torch.where(out, sgn * mag, r)
    y = mu + r
    return y.float(), mu.float(), sigma.float()
