"""Oracle integer path: latent rounding, scale->CDF-row index, 16-bit CDF table
build and (lo, hi) interval lookup (SURVEY.md Appendix A.5, rows a8/a9 of section 8).

Reference file:line: none -- /root/reference/README.md is 0 bytes.  The table
conventions restate the published CompressAI `GaussianConditional` algorithm
(64-level log-spaced scale table 0.11..256, 16-bit quantised CDFs, tail mass,
`pmf_to_quantized_cdf` zero-bin stealing); compressai is NOT installed here and
no version is pinned by the reference, so this is a restatement from the paper
trail, pinned by the KATs in tests/test_oracle_entropy.py.

All table arithmetic is numpy int64 / float64; lookups are exact integers.
Test infrastructure only; see oracle/__init__.py.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
from scipy.special import ndtr
from scipy.stats import norm

PRECISION = 16
SCALE_MIN = 0.11
SCALE_MAX = 256.0
LEVELS = 64
TAIL_MASS = 1e-9


def scale_table() -> np.ndarray:
    """table_j = exp(ln 0.11 + j (ln 256 - ln 0.11)/63), j = 0..63, stored fp32."""
    j = np.arange(LEVELS, dtype=np.float64)
    t = np.exp(math.log(SCALE_MIN) + j * (math.log(SCALE_MAX) - math.log(SCALE_MIN)) / (LEVELS - 1))
    return t.astype(np.float32)


def pmf_to_quantized_cdf(pmf: np.ndarray, precision: int = PRECISION) -> np.ndarray:
    """Integer PMF -> CDF with zero-width bins repaired by stealing one count from the
    smallest bin of width > 1 and shifting the entries in between (A.5)."""
    n = len(pmf)
    total_target = 1 << precision
    f = np.rint(np.asarray(pmf, dtype=np.float64) * total_target).astype(np.int64)
    s = int(f.sum())
    f = (total_target * f) // s
    cdf = np.zeros(n + 1, dtype=np.int64)
    cdf[1:] = np.cumsum(f)
    cdf[-1] = total_target
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            freq = cdf[1:] - cdf[:-1]
            cand = np.where(freq > 1, freq, np.iinfo(np.int64).max)
            best = int(np.argmin(cand))  # first smallest bin with width > 1
            assert cand[best] != np.iinfo(np.int64).max, "cannot repair zero-width bin"
            if best < i:
                cdf[best + 1:i + 1] -= 1
            else:
                cdf[i + 1:best + 1] += 1
    return cdf


@dataclass
class CDFTables:
    """Flattened rows: row r occupies cdf[row_start[r] : row_start[r] + cdf_length[r]]."""
    cdf: np.ndarray          # int32 [total]
    row_start: np.ndarray    # int32 [rows]
    cdf_length: np.ndarray   # int32 [rows]
    offset: np.ndarray       # int32 [rows]
    scale_table: np.ndarray  # float32 [64] (Gaussian tables only; zeros otherwise)

    @property
    def rows(self):
        return len(self.row_start)


def _pack(rows, offsets, table):
    lens = np.array([len(r) for r in rows], dtype=np.int32)
    start = np.zeros(len(rows), dtype=np.int32)
    start[1:] = np.cumsum(lens)[:-1]
    flat = np.concatenate(rows).astype(np.int32)
    return CDFTables(flat, start, lens, np.asarray(offsets, dtype=np.int32), table)


def build_gaussian_tables() -> CDFTables:
    """One quantised CDF row per scale level (A.5)."""
    table = scale_table()
    m = -norm.ppf(TAIL_MASS / 2.0)
    rows, offsets = [], []
    for j in range(LEVELS):
        sj = float(table[j])
        center = int(math.ceil(sj * m))
        s = np.arange(2 * center + 1, dtype=np.float64)
        a = np.abs(s - center)
        pmf = ndtr((0.5 - a) / sj) - ndtr((-0.5 - a) / sj)
        tail = 2.0 * ndtr((-0.5 - center) / sj)
        cdf = pmf_to_quantized_cdf(np.concatenate([pmf, [tail]]))
        assert len(cdf) == 2 * center + 3
        rows.append(cdf)
        offsets.append(-center)
    return _pack(rows, offsets, table)


def build_tables_from_pmfs(pmfs, offsets, tails) -> CDFTables:
    """Factorised-prior rows (index = channel): same quantiser, caller supplies the PMFs."""
    rows = [pmf_to_quantized_cdf(np.concatenate([p, [t]])) for p, t in zip(pmfs, tails)]
    return _pack(rows, offsets, np.zeros(LEVELS, dtype=np.float32))


def quantize_symbols(y: torch.Tensor, mu: torch.Tensor):
    """q = rint(y - mu) (half-to-even) -> int32; y_hat = q + mu (A.5)."""
    q = torch.round(y.float() - mu.float())
    return q.to(torch.int32), q + mu.float()


def build_indexes(sigma: torch.Tensor, table: np.ndarray) -> torch.Tensor:
    """idx = 63 - #{j in [0,62] : max(sigma, 0.11) <= table_j}; fp32 compare, ties count as <=."""
    s = torch.clamp(sigma.float(), min=float(np.float32(SCALE_MIN)))
    t = torch.from_numpy(table[:LEVELS - 1].copy())
    cnt = (s.reshape(-1, 1) <= t.reshape(1, -1)).sum(dim=1).to(torch.int32)
    return (LEVELS - 1 - cnt).reshape(sigma.shape)


def lookup_rows(q: torch.Tensor, idx: torch.Tensor, tables: CDFTables):
    """v = q - offset[idx]; escape outside [0, max_v); lo = cdf[idx][v], hi = cdf[idx][v+1] (A.5)."""
    qn = q.reshape(-1).numpy().astype(np.int64)
    ix = idx.reshape(-1).numpy().astype(np.int64)
    max_v = tables.cdf_length[ix].astype(np.int64) - 2
    v = qn - tables.offset[ix].astype(np.int64)
    raw = np.zeros_like(v)
    neg = v < 0
    big = v >= max_v
    raw[neg] = -2 * v[neg] - 1
    raw[big] = 2 * (v[big] - max_v[big])
    v = np.where(neg | big, max_v, v)
    base = tables.row_start[ix].astype(np.int64)
    lo = tables.cdf[base + v]
    hi = tables.cdf[base + v + 1]
    shp = q.shape
    mk = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).reshape(shp)
    return (mk(ix, np.int32), mk(v, np.int32), mk(lo, np.int32), mk(hi, np.int32), mk(raw, np.int32))


def cdf_lookup(q: torch.Tensor, sigma: torch.Tensor, tables: CDFTables):
    """(idx, v, lo, hi, raw), all int32 (lo/hi hold values <= 65536)."""
    idx = build_indexes(sigma, tables.scale_table)
    return lookup_rows(q, idx, tables)


def estimated_bits(lo: torch.Tensor, hi: torch.Tensor, raw: torch.Tensor) -> float:
    """sum -log2((hi-lo)/65536) plus an Exp-Golomb(0) cost for escapes; reported, not gated."""
    w = (hi - lo).double().clamp(min=1.0)
    bits = float((-(w / 65536.0).log2()).sum())
    r = raw[raw > 0].double()
    if r.numel():
        bits += float((2.0 * torch.floor(torch.log2(r + 1.0)) + 1.0).sum())
    return bits
