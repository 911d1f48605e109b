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
    e0: np.ndarray = None   # float32 [K]: x0 = e0 * x_t + e1 * out  (X-parameterisation: 0, 1)
    e1: np.ndarray = None
    sg: np.ndarray = None   # float32 [K]: DDIM sigma_k (0 when eta = 0)


def make_schedule(K: int, T: int = 1000, eta: float = 0.0, pred: str = "x") -> Schedule:
    """Per-step coefficients of the generalised DDIM update (SURVEY.md A.4 + section 8 row f4):

        x0     = e0 * x_t + e1 * out                      pred "x": out = x0_hat (e0 = 0, e1 = 1)
                                                          pred "eps": out = eps_hat, x0 = (x_t - sqrt(1-a_t) eps)/sqrt(a_t)
        sigma  = eta * sqrt((1-a_p)/(1-a_t)) * sqrt(1 - a_t/a_p)      (0 on the last step, a_p = 1)
        x_prev = sqrt(a_p) x0c + sqrt(1-a_p-sigma^2) * (x_t - sqrt(a_t) x0c)/sqrt(1-a_t) + sigma z,   x0c = clamp(x0)
               = c0 * x0c + c1 * x_t + sigma * z

    eta = 0, pred "x" is the deterministic X-parameterised sampler of A.4 (c0, c1 unchanged)."""
    assert pred in ("x", "eps")
    ab = alphas_cumprod(T)
    idx = step_indices(K, T)
    c0, c1, e0, e1, sg = (np.zeros(K, dtype=np.float64) for _ in range(5))
    for k in range(K):
        a_t = ab[idx[k]]
        a_p = ab[idx[k + 1]] if k + 1 < K else 1.0
        s = eta * math.sqrt((1.0 - a_p) / (1.0 - a_t)) * math.sqrt(1.0 - a_t / a_p) if (eta > 0.0 and a_p < 1.0) else 0.0
        c1[k] = math.sqrt(max(1.0 - a_p - s * s, 0.0)) / math.sqrt(1.0 - a_t)
        c0[k] = math.sqrt(a_p) - c1[k] * math.sqrt(a_t)
        sg[k] = s
        e0[k] = 1.0 / math.sqrt(a_t) if pred == "eps" else 0.0
        e1[k] = -math.sqrt(1.0 - a_t) / math.sqrt(a_t) if pred == "eps" else 1.0
    f = lambda a: a.astype(np.float32)
    return Schedule(K, idx, f(c0), f(c1), f(e0), f(e1), f(sg))


# ---- counter-based sampler noise (shared definition with csrc/sampler.cuh) ---------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al., Random123) on uint32 numpy arrays; returns four uint32 arrays.
    KAT: counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & m32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & m32, lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox_normal(seed: int, step: int, B: int, H: int, W: int) -> torch.Tensor:
    """z [B,3,H,W] fp32: the sampler noise of `step`.  Pixel p = (b*H + h)*W + w: counter (p_lo, p_hi, step, 0), key =
    seed; words 0/1 -> Box-Muller pair (channels 0, 1), words 2/3 -> channel 2:
        u = ((w >> 9) + 0.5) / 2^23,  t = (w >> 8) / 2^24,  z = sqrt(-2 ln u) * (cos | sin)(2 pi t)."""
    p = np.arange(B * H * W, dtype=np.uint64)
    w0, w1, w2, w3 = philox4x32_10(p & np.uint64(0xFFFFFFFF), p >> np.uint64(32), np.full(p.shape, step, dtype=np.uint64),
                                   np.zeros(p.shape, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u0 = ((w0 >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    u2 = ((w2 >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    t1 = (w1 >> np.uint32(8)).astype(np.float64) / 16777216.0
    t3 = (w3 >> np.uint32(8)).astype(np.float64) / 16777216.0
    r0, r2 = np.sqrt(-2.0 * np.log(u0)), np.sqrt(-2.0 * np.log(u2))
    z = np.stack([r0 * np.cos(2.0 * np.pi * t1), r0 * np.sin(2.0 * np.pi * t1), r2 * np.cos(2.0 * np.pi * t3)], axis=0)
    z = z.reshape(3, B, H, W).transpose(1, 0, 2, 3)
    return torch.from_numpy(np.ascontiguousarray(z).astype(np.float32))


def ddim_update(x_t: torch.Tensor, out: torch.Tensor, c0: float, c1: float, e0: float = 0.0, e1: float = 1.0,
                sg: float = 0.0, z: torch.Tensor = None) -> torch.Tensor:
    """x0 = e0*x_t + e1*out; x_prev = c0*clamp(x0,-1,1) + c1*x_t (+ sg*z), all fp32 (A.4; defaults: X-param, eta = 0)."""
    x0 = out if (float(e0) == 0.0 and float(e1) == 1.0) else float(e1) * out + float(e0) * x_t
    xp = float(c0) * x0.clamp(-1.0, 1.0) + float(c1) * x_t
    if float(sg) != 0.0:
        xp = xp + float(sg) * z
    return xp


class OracleDecoder:
    """Same surface as the CUDA `Decoder`: set_sample_schedule,
    predict_x0, denoise_step, decode, quantize_symbols, cdf_lookup."""

    def __init__(self, cfg: CDCConfig, unet, context_net=None, tables=None):
        self.cfg = cfg
        self.unet = unet.eval()
        self.context_net = context_net
        self.tables = tables
        self.sched = None

    def set_sample_schedule(self, steps: int, eta: float = 0.0, pred: str = "x", seed: int = 0):
        self.sched = make_schedule(steps, self.cfg.T, eta, pred)
        self.eta, self.pred, self.seed = float(eta), pred, int(seed)
        return self.sched

    def _k_of(self, t: int) -> int:
        assert self.sched is not None, "call set_sample_schedule first"
        if int(t) not in self.sched.idx:
            raise ValueError(f"t={t} is not in the active {self.sched.K}-step schedule")
        return self.sched.idx.index(int(t))

    @torch.no_grad()
    def network_out(self, x_t, t, cond):
        """The UNet's raw output: x0_hat (pred "x") or eps_hat (pred "eps")."""
        tt = torch.full((x_t.shape[0],), int(t), dtype=torch.int64)
        return self.unet(x_t, tt, cond)

    @torch.no_grad()
    def predict_x0(self, x_t, t, cond):
        out = self.network_out(x_t, t, cond)
        k = self._k_of(t) if self.sched is not None and int(t) in self.sched.idx else None
        if k is None or getattr(self, "pred", "x") == "x":
            return out
        return float(self.sched.e1[k]) * out + float(self.sched.e0[k]) * x_t

    @torch.no_grad()
    def denoise_step(self, x_t, t, cond):
        k = self._k_of(t)
        out = self.network_out(x_t, t, cond)
        sc = self.sched
        z = None
        if float(sc.sg[k]) != 0.0:
            B, _, H, W = x_t.shape
            z = philox_normal(self.seed, k, B, H, W)
        return ddim_update(x_t, out, sc.c0[k], sc.c1[k], sc.e0[k], sc.e1[k], sc.sg[k], z)

    @torch.no_grad()
    def decode(self, latent, steps, *, init=None, gamma=0.8, seed=0, cond=None, trajectory=None):
        """latent y_hat fp32 [B,256,H/16,W/16] -> image fp32 [B,3,H,W] in [0,1]."""
        if self.sched is None or self.sched.K != steps:
            self.set_sample_schedule(steps, getattr(self, "eta", 0.0), getattr(self, "pred", "x"), getattr(self, "seed", 0))
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
