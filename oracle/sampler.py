"""Oracle schedule + X-parameterised DDIM sampler.

Reference file:line: none -- /root/reference/README.md is 0 bytes.  The cosine
schedule, the integer-only step-index formula and the fused update
x_prev = c0*clamp(x0_hat) + c1*x_t are pinned by SURVEY.md A.4 and checked by the
known-answer tests in tests/test_oracle_sampler.py.

Test infrastructure only; see oracle/__init__.py.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from .config import CDCConfig


def alphas_cumprod(T: int = 1000) -> np.ndarray:
    """Cosine schedule, float64: f(u)=cos^2((u/T+0.008)/1.008*pi/2), beta_t=clip(1-f(t+1)/f(t),0,0.999)."""
    u = np.arange(T + 1, dtype=np.float64)
    f = np.cos((u / T + 0.008) / 1.008 * math.pi / 2.0) ** 2
    beta = np.clip(1.0 - f[1:] / f[:-1], 0.0, 0.999)
    return np.cumprod(1.0 - beta)


def step_indices(K: int, T: int = 1000):
    """idx_k = ((K-1-k)*(T-1) + (K-1)//2) // (K-1), integer only (K=17 -> 999,937,...,500,...,62,0)."""
    if K == 1:
        return [T - 1]
    return [((K - 1 - k) * (T - 1) + (K - 1) // 2) // (K - 1) for k in range(K)]


@dataclass
class Schedule:
    K: int
    idx: list           # training index per step k
    c0: np.ndarray      # float32 [K]
    c1: np.ndarray      # float32 [K]


def make_schedule(K: int, T: int = 1000) -> Schedule:
    ab = alphas_cumprod(T)
    idx = step_indices(K, T)
    c0 = np.zeros(K, dtype=np.float64)
    c1 = np.zeros(K, dtype=np.float64)
    for k in range(K):
        a_t = ab[idx[k]]
        a_p = ab[idx[k + 1]] if k + 1 < K else 1.0
        c1[k] = math.sqrt(1.0 - a_p) / math.sqrt(1.0 - a_t)
        c0[k] = math.sqrt(a_p) - c1[k] * math.sqrt(a_t)
    return Schedule(K, idx, c0.astype(np.float32), c1.astype(np.float32))


def ddim_update(x_t: torch.Tensor, x0_hat: torch.Tensor, c0: float, c1: float) -> torch.Tensor:
    """x_prev = c0*clamp(x0_hat,-1,1) + c1*x_t in fp32 (A.4)."""
    return float(c0) * x0_hat.clamp(-1.0, 1.0) + float(c1) * x_t


class OracleDecoder:
    """Same surface as the CUDA `Decoder`: set_sample_schedule,
    predict_x0, denoise_step, decode, quantize_symbols, cdf_lookup."""

    def __init__(self, cfg: CDCConfig, unet, context_net=None, tables=None):
        self.cfg = cfg
        self.unet = unet.eval()
        self.context_net = context_net
        self.tables = tables
        self.sched = None

    def set_sample_schedule(self, steps: int):
        self.sched = make_schedule(steps, self.cfg.T)
        return self.sched

    def _k_of(self, t: int) -> int:
        assert self.sched is not None, "call set_sample_schedule first"
        if int(t) not in self.sched.idx:
            raise ValueError(f"t={t} is not in the active {self.sched.K}-step schedule")
        return self.sched.idx.index(int(t))

    @torch.no_grad()
    def predict_x0(self, x_t, t, cond):
        tt = torch.full((x_t.shape[0],), int(t), dtype=torch.int64)
        return self.unet(x_t, tt, cond)

    @torch.no_grad()
    def denoise_step(self, x_t, t, cond):
        k = self._k_of(t)
        x0 = self.predict_x0(x_t, t, cond)
        return ddim_update(x_t, x0, self.sched.c0[k], self.sched.c1[k])

    @torch.no_grad()
    def decode(self, latent, steps, *, init=None, gamma=0.8, seed=0, cond=None, trajectory=None):
        """latent y_hat fp32 [B,256,H/16,W/16] -> image fp32 [B,3,H,W] in [0,1]."""
        self.set_sample_schedule(steps)
        if cond is None:
            cond = self.context_net(latent)
        B, _, h, w = latent.shape
        if init is None:
            g = torch.Generator().manual_seed(seed)
            x = gamma * torch.randn(B, 3, h * 16, w * 16, generator=g)
        else:
            x = init.clone()
        for k in range(steps):
            if trajectory is not None:
                trajectory.append(x.clone())
            x = self.denoise_step(x, self.sched.idx[k], cond)
        return (x.clamp(-1.0, 1.0) + 1.0) / 2.0

    def quantize_symbols(self, y, mu):
        from .entropy import quantize_symbols
        return quantize_symbols(y, mu)

    def cdf_lookup(self, q, sigma):
        from .entropy import cdf_lookup
        return cdf_lookup(q, sigma, self.tables)
