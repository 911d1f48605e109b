"""Architecture config of the decode path (mirrors the oracle's CDCConfig; the
reference ships no config: /root/reference/README.md is 0 bytes)."""
from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class CDCConfig:
    base: int = 64
    mults: Tuple[int, ...] = (1, 2, 3, 4)
    groups: int = 32
    heads: int = 4
    head_dim: int = 64
    temb: int = 256
    T: int = 1000
    latent_ch: int = 256
    gn_eps: float = 1e-5

    @property
    def channels(self) -> Tuple[int, ...]:
        return tuple(self.base * m for m in self.mults)
