"""Oracle conditional denoising UNet, fp32 eager.

Reference file:line: none -- /root/reference/README.md is 0 bytes.  The block
structure follows BASELINE.json `north_star` (GroupNorm+SiLU, FiLM time
conditioning, 3x3/1x1 convs, softmax self-attention at 1/16 resolution, context
maps concatenated at the entry of every down level).

Test infrastructure only; see oracle/__init__.py.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import CDCConfig


def conv(cin, cout, k, s=1):
    return nn.Conv2d(cin, cout, k, stride=s, padding=k // 2, bias=True)


class RB(nn.Module):
    """Time-conditioned post-norm ResBlock (A.2).  `film=False` gives RBn (codec side)."""

    def __init__(self, cin, cout, groups, temb=None):
        super().__init__()
        self.conv1 = conv(cin, cout, 3)
        self.gn1 = nn.GroupNorm(groups, cout, eps=1e-5, affine=True)
        self.conv2 = conv(cout, cout, 3)
        self.gn2 = nn.GroupNorm(groups, cout, eps=1e-5, affine=True)
        self.film = nn.Linear(temb, 2 * cout) if temb is not None else None
        self.res = conv(cin, cout, 1) if cin != cout else None

    def forward(self, x, te=None):
        h = self.conv1(x)
        h = self.gn1(h)
        if self.film is not None:
            s, sh = self.film(F.silu(te)).chunk(2, dim=1)
            h = h * (1.0 + s[:, :, None, None]) + sh[:, :, None, None]
        h = F.silu(h)
        h = self.conv2(h)
        h = F.silu(self.gn2(h))
        r = x if self.res is None else self.res(x)
        return h + r


class Up(nn.Module):
    """nearest x2 then conv3x3 (A.2)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up = conv(cin, cout, 3)

    def forward(self, x):
        return self.up(F.interpolate(x, scale_factor=2, mode="nearest"))


class Attn(nn.Module):
    """x + proj(softmax(q k^T / sqrt(d)) v) over HW tokens, GN first (A.2)."""

    def __init__(self, c, groups, heads, head_dim):
        super().__init__()
        assert heads * head_dim == c
        self.heads, self.head_dim = heads, head_dim
        self.gn = nn.GroupNorm(groups, c, eps=1e-5, affine=True)
        self.qkv = conv(c, 3 * c, 1)
        self.proj = conv(c, c, 1)

    def forward(self, x):
        B, C, H, W = x.shape
        n = self.gn(x)
        q, k, v = self.qkv(n).chunk(3, dim=1)  # each [B, C, H, W]; head h = channels h*d..h*d+d-1
        q = q.reshape(B, self.heads, self.head_dim, H * W)
        k = k.reshape(B, self.heads, self.head_dim, H * W)
        v = v.reshape(B, self.heads, self.head_dim, H * W)
        s = torch.einsum("bhdi,bhdj->bhij", q, k) * (1.0 / math.sqrt(self.head_dim))
        p = torch.softmax(s, dim=-1)
        o = torch.einsum("bhij,bhdj->bhdi", p, v).reshape(B, C, H, W)
        return x + self.proj(o)


def sinusoidal(t, dim=64):
    """e = [sin(t f_j), cos(t f_j)], f_j = exp(-ln(1e4) j / (dim/2)) (A.2)."""
    half = dim // 2
    j = torch.arange(half, dtype=torch.float32, device=t.device)
    f = torch.exp(-math.log(10000.0) * j / half)
    a = t.to(torch.float32)[:, None] * f[None, :]
    return torch.cat([torch.sin(a), torch.cos(a)], dim=1)


class TimeEmbed(nn.Module):
    def __init__(self, temb):
        super().__init__()
        self.lin1 = nn.Linear(64, temb)
        self.lin2 = nn.Linear(temb, temb)

    def forward(self, t):
        return self.lin2(F.silu(self.lin1(sinusoidal(t, 64))))


class DownLevel(nn.Module):
    def __init__(self, cin, c, groups, temb):
        super().__init__()
        self.rb1 = RB(cin, c, groups, temb)
        self.rb2 = RB(c, c, groups, temb)
        self.down = conv(c, c, 3, 2)


class UpLevel(nn.Module):
    def __init__(self, prev, c, groups, temb):
        super().__init__()
        self.up = Up(prev, c)
        self.rb1 = RB(2 * c, c, groups, temb)
        self.rb2 = RB(c, c, groups, temb)


class Mid(nn.Module):
    def __init__(self, c, cfg):
        super().__init__()
        self.rb1 = RB(c, c, cfg.groups, cfg.temb)
        self.attn = Attn(c, cfg.groups, cfg.heads, cfg.head_dim)
        self.rb2 = RB(c, c, cfg.groups, cfg.temb)


class UNet(nn.Module):
    """predict_x0(x_t, t, cond) (A.3).  Module construction order is the weight-init order (A.6)."""

    def __init__(self, cfg: CDCConfig = CDCConfig()):
        super().__init__()
        self.cfg = cfg
        C = cfg.channels
        self.temb = TimeEmbed(cfg.temb)
        self.stem = conv(3 + C[0], C[0], 3)
        downs = []
        for i, c in enumerate(C):
            cin = C[0] if i == 0 else C[i - 1] + C[i]
            downs.append(DownLevel(cin, c, cfg.groups, cfg.temb))
        self.down = nn.ModuleList(downs)
        self.mid = Mid(C[-1], cfg)
        ups = {}
        prev = C[-1]
        for i in reversed(range(len(C))):
            ups[str(i)] = UpLevel(prev, C[i], cfg.groups, cfg.temb)
            prev = C[i]
        self.up = nn.ModuleDict(ups)  # keyed by level: "3","2","1","0"
        self.final = conv(C[0], 3, 3)

    def forward(self, x_t, t, cond):
        """x_t [B,3,H,W] fp32; t int64 [B] training indices; cond = (c0,c1,c2,c3) -> x0_hat [B,3,H,W]."""
        te = self.temb(t)
        h = self.stem(torch.cat([x_t, cond[0]], dim=1))
        skips = []
        for i, lvl in enumerate(self.down):
            hin = h if i == 0 else torch.cat([h, cond[i]], dim=1)
            h = lvl.rb1(hin, te)
            h = lvl.rb2(h, te)
            skips.append(h)
            h = lvl.down(h)
        h = self.mid.rb1(h, te)
        h = self.mid.attn(h)
        h = self.mid.rb2(h, te)
        for i in reversed(range(len(self.down))):
            lvl = self.up[str(i)]
            h = lvl.up(h)
            h = lvl.rb1(torch.cat([h, skips[i]], dim=1), te)
            h = lvl.rb2(h, te)
        return self.final(h)


def unet_flops(cfg: CDCConfig, B: int, H: int, W: int) -> float:
    """Algorithmic FLOPs of one predict_x0 (SURVEY.md section 8d rule): conv 2*M*N*K with
    true N, K; attention 4*N_tok^2*d per head; linears 2*B*in*out."""
    C = cfg.channels
    fl = 0.0

    def cv(m, n, k):
        return 2.0 * m * n * k

    def rb(m, cin, cout):
        f = cv(m, cout, 9 * cin) + cv(m, cout, 9 * cout)
        if cin != cout:
            f += cv(m, cout, cin)
        return f + 2.0 * B * cfg.temb * 2 * cout

    M = B * H * W
    fl += cv(M, C[0], 9 * (3 + C[0]))
    m = M
    for i, c in enumerate(C):
        cin = C[0] if i == 0 else C[i - 1] + C[i]
        fl += rb(m, cin, c) + rb(m, c, c)
        m //= 4
        fl += cv(m, c, 9 * c)
    c = C[-1]
    fl += rb(m, c, c) * 2
    ntok = m // B
    fl += cv(m, 3 * c, c) + cv(m, c, c) + B * cfg.heads * 4.0 * ntok * ntok * cfg.head_dim
    prev = c
    for i in reversed(range(len(C))):
        m *= 4
        fl += cv(m, C[i], 9 * prev) + rb(m, 2 * C[i], C[i]) + rb(m, C[i], C[i])
        prev = C[i]
    fl += cv(M, 3, 9 * C[0])
    fl += 2.0 * B * (64 * cfg.temb + cfg.temb * cfg.temb)
    return fl
