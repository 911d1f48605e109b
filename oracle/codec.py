"""Oracle codec around the decode loop: analysis encoder,
hyper encoder/decoder, factorised prior, context net.  These are the "next" rows
(section 8 f1/f2); only `ContextNet` feeds the hot path (it produces `cond`).

Reference file:line: none -- /root/reference/README.md is 0 bytes.
Test infrastructure only; see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import CDCConfig
from .entropy import build_gaussian_tables, build_tables_from_pmfs, cdf_lookup, lookup_rows, quantize_symbols
from .unet import RB, Up, conv


class Encoder(nn.Module):
    """Conv3(3->64), then per level RBn(.->C[i]), Down(C[i])  =>  y [B,256,H/16,W/16]."""

    def __init__(self, cfg: CDCConfig):
        super().__init__()
        C = cfg.channels
        self.stem = conv(3, C[0], 3)
        self.rbs = nn.ModuleList()
        self.downs = nn.ModuleList()
        prev = C[0]
        for c in C:
            self.rbs.append(RB(prev, c, cfg.groups, None))
            self.downs.append(conv(c, c, 3, 2))
            prev = c

    def forward(self, x):
        h = self.stem(x)
        for rb, d in zip(self.rbs, self.downs):
            h = d(rb(h))
        return h


class HyperEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.c1 = conv(c, c, 3)
        self.c2 = nn.Conv2d(c, c, 5, stride=2, padding=2)
        self.c3 = nn.Conv2d(c, c, 5, stride=2, padding=2)

    def forward(self, y):
        h = F.leaky_relu(self.c1(y), 0.2)
        h = F.leaky_relu(self.c2(h), 0.2)
        return self.c3(h)


class HyperDecoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.t1 = nn.ConvTranspose2d(c, c, 5, stride=2, padding=2, output_padding=1)
        self.t2 = nn.ConvTranspose2d(c, c, 5, stride=2, padding=2, output_padding=1)
        self.c3 = conv(c, 2 * c, 3)

    def forward(self, z_hat):
        h = F.leaky_relu(self.t1(z_hat), 0.2)
        h = F.leaky_relu(self.t2(h), 0.2)
        mu, sigma_raw = self.c3(h).chunk(2, dim=1)
        return mu, torch.clamp(sigma_raw, min=0.11)


class FactorizedPrior(nn.Module):
    """Balle-style per-channel cumulative MLP, filters (3,3,3,3); untrained init, so the
    per-channel median is the `median` parameter (0) and the PMF support is [-10, 10]."""

    def __init__(self, c, filters=(3, 3, 3, 3), init_scale=10.0):
        super().__init__()
        self.c = c
        f = (1,) + tuple(filters) + (1,)
        scale = init_scale ** (1.0 / (len(filters) + 1))
        self.mats, self.biases, self.factors = nn.ParameterList(), nn.ParameterList(), nn.ParameterList()
        for i in range(len(filters) + 1):
            init = float(np.log(np.expm1(1.0 / scale / f[i + 1])))
            self.mats.append(nn.Parameter(torch.full((c, f[i + 1], f[i]), init)))
            self.biases.append(nn.Parameter(torch.empty(c, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(filters):
                self.factors.append(nn.Parameter(torch.zeros(c, f[i + 1], 1)))
        self.median = nn.Parameter(torch.zeros(c))
        self.support = 10

    def logits_cumulative(self, x):  # x [c, 1, n]
        h = x
        for i in range(len(self.mats)):
            h = torch.matmul(F.softplus(self.mats[i]), h) + self.biases[i]
            if i < len(self.factors):
                h = h + torch.tanh(self.factors[i]) * torch.tanh(h)
        return h

    @torch.no_grad()
    def build_tables(self):
        s = torch.arange(-self.support, self.support + 1, dtype=torch.float32)
        samples = (self.median[:, None] + s[None, :])[:, None, :].double()
        dbl = FactorizedPrior.__new__(FactorizedPrior)  # float64 evaluation for a stable table
        lower = self._logits_double(samples - 0.5)
        upper = self._logits_double(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:]))[:, 0]
        del dbl
        pm = [pmf[i].numpy() for i in range(self.c)]
        return build_tables_from_pmfs(pm, [-self.support] * self.c, tail.numpy())

    def _logits_double(self, x):
        h = x
        for i in range(len(self.mats)):
            h = torch.matmul(F.softplus(self.mats[i].double()), h) + self.biases[i].double()
            if i < len(self.factors):
                h = h + torch.tanh(self.factors[i].double()) * torch.tanh(h)
        return h


class ContextNet(nn.Module):
    """y_hat -> (c0, c1, c2, c3): for i in 3..0: h = Up(prev->C[i])(h); h = RBn(C[i])(h); c_i = h."""

    def __init__(self, cfg: CDCConfig):
        super().__init__()
        C = cfg.channels
        self.ups, self.rbs = nn.ModuleDict(), nn.ModuleDict()
        prev = cfg.latent_ch
        for i in reversed(range(len(C))):
            self.ups[str(i)] = Up(prev, C[i])
            self.rbs[str(i)] = RB(C[i], C[i], cfg.groups, None)
            prev = C[i]
        self.n = len(C)

    def forward(self, y_hat):
        h = y_hat
        out = [None] * self.n
        for i in reversed(range(self.n)):
            h = self.rbs[str(i)](self.ups[str(i)](h))
            out[i] = h
        return tuple(out)


class Codec(nn.Module):
    """Encoder + hyperprior + context net (A.5).  Module order is the weight-init order (A.6)."""

    def __init__(self, cfg: CDCConfig = CDCConfig()):
        super().__init__()
        self.cfg = cfg
        c = cfg.latent_ch
        self.encoder = Encoder(cfg)
        self.hyper_enc = HyperEncoder(c)
        self.prior = FactorizedPrior(c)
        self.hyper_dec = HyperDecoder(c)
        self.context = ContextNet(cfg)
        self._gauss = None
        self._fact = None

    def tables(self):
        if self._gauss is None:
            self._gauss = build_gaussian_tables()
            self._fact = self.prior.build_tables()
        return self._gauss, self._fact

    @torch.no_grad()
    def encode(self, img01):
        """img in [0,1] -> dict with y, z, mu, sigma, q (int32), y_hat, and the entropy-coder symbols."""
        gauss, fact = self.tables()
        x = 2.0 * img01 - 1.0
        y = self.encoder(x)
        z = self.hyper_enc(y)
        med = self.prior.median[None, :, None, None].expand_as(z).contiguous()
        qz, z_hat = quantize_symbols(z, med)
        ch = torch.arange(z.shape[1], dtype=torch.int32)[None, :, None, None].expand_as(z).contiguous()
        z_sym = lookup_rows(qz, ch, fact)
        mu, sigma = self.hyper_dec(z_hat)
        q, y_hat = quantize_symbols(y, mu)
        y_sym = cdf_lookup(q, sigma, gauss)
        return dict(y=y, z=z, mu=mu, sigma=sigma, q=q, y_hat=y_hat, qz=qz, z_hat=z_hat, y_sym=y_sym, z_sym=z_sym)
