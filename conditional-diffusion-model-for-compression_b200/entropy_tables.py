"""Host-side construction of the 16-bit CDF tables the integer kernels look up: built once
in numpy int64 / float64 and uploaded (`DeviceTables`); only the LOOKUP runs on the device and must be bit-exact.
tests/test_oracle_entropy.py checks these tables against the oracle's, entry for entry (the product may not import the
oracle, so the construction is restated here)."""
import math
from dataclasses import dataclass

import numpy as np

PRECISION, LEVELS, SCALE_MIN, SCALE_MAX, TAIL_MASS = 16, 64, 0.11, 256.0, 1e-9


@dataclass
class Tables:
    cdf: np.ndarray          # int32, rows back to back
    row_start: np.ndarray    # int32 [rows]
    cdf_length: np.ndarray   # int32 [rows]
    offset: np.ndarray       # int32 [rows]
    scale_table: np.ndarray  # float32 [64] (zeros for index-by-channel tables)

    @property
    def rows(self):
        return int(self.row_start.shape[0])


def scale_table():
    step = (math.log(SCALE_MAX) - math.log(SCALE_MIN)) / (LEVELS - 1)
    return np.exp(math.log(SCALE_MIN) + step * np.arange(LEVELS, dtype=np.float64)).astype(np.float32)


def quantize_pmf(pmf, precision=PRECISION):
    """PMF (float64, tail mass last) -> CDF of length len + 1 ending at 2^precision with no empty bin: counts are
    round(p * 2^precision) renormalised by integer division; an empty bin takes one count from the first narrowest
    bin that can spare it, the entries in between shift by one."""
    total = 1 << precision
    cnt = np.rint(np.asarray(pmf, dtype=np.float64) * total).astype(np.int64)
    cnt = (cnt * total) // int(cnt.sum())
    cdf = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    cdf[-1] = total
    for i in range(len(cnt)):
        if cdf[i + 1] != cdf[i]:
            continue
        width = np.diff(cdf)
        width = np.where(width > 1, width, np.iinfo(np.int64).max)
        j = int(np.argmin(width))
        if width[j] == np.iinfo(np.int64).max:
            raise ValueError("CDF cannot be repaired: no bin wider than one count")
        if j < i:
            cdf[j + 1:i + 1] -= 1
        else:
            cdf[i + 1:j + 1] += 1
    return cdf


def _flatten(rows, offsets, table):
    length = np.array([len(r) for r in rows], dtype=np.int32)
    start = np.concatenate([[0], np.cumsum(length)[:-1]]).astype(np.int32)
    return Tables(np.concatenate(rows).astype(np.int32), start, length, np.asarray(offsets, dtype=np.int32), table)


def gaussian_tables():
    """One row per scale level j: support [-c_j, c_j], c_j = ceil(table_j * m), m = -Phi^-1(tail/2); bin mass
    Phi((.5 - |s|)/sigma) - Phi((-.5 - |s|)/sigma), then the two-sided tail mass as the escape bin."""
    from scipy.special import ndtr
    from scipy.stats import norm
    tab = scale_table()
    m = -norm.ppf(TAIL_MASS / 2.0)
    rows, offs = [], []
    for sj in tab.astype(np.float64):
        c = int(math.ceil(sj * m))
        a = np.abs(np.arange(-c, c + 1, dtype=np.float64))
        mass = ndtr((0.5 - a) / sj) - ndtr((-0.5 - a) / sj)
        rows.append(quantize_pmf(np.append(mass, 2.0 * ndtr((-0.5 - c) / sj))))
        offs.append(-c)
    return _flatten(rows, offs, tab)


def tables_from_pmfs(pmfs, offsets, tails):
    """Rows indexed by channel (factorised prior): the caller supplies each row's PMF, support offset and tail mass."""
    rows = [quantize_pmf(np.append(np.asarray(p, dtype=np.float64), t)) for p, t in zip(pmfs, tails)]
    return _flatten(rows, offsets, np.zeros(LEVELS, dtype=np.float32))


def factorized_tables(mats, biases, factors, median, support=10):
    """Rows indexed by channel for the factorised prior of z (Balle-style cumulative, oracle/codec.py FactorizedPrior):
    logits(x) = chain of  h <- softplus(M_i) h + b_i ; h <- h + tanh(f_i) tanh(h)  per channel, evaluated in float64 at
    median + s -+ 0.5 for s in [-support, support];  pmf = |sigmoid(sgn * upper) - sigmoid(sgn * lower)|,
    sgn = -sign(lower + upper);  tail = sigmoid(lower at the first bin) + sigmoid(-upper at the last bin).
    mats / biases / factors: lists of float arrays [c, f_out, f_in] / [c, f_out, 1] / [c, f_out, 1]; median [c]."""
    mats = [np.asarray(m, dtype=np.float64) for m in mats]
    biases = [np.asarray(b, dtype=np.float64) for b in biases]
    factors = [np.asarray(f, dtype=np.float64) for f in factors]
    med = np.asarray(median, dtype=np.float64)
    c = med.shape[0]

    def softplus(x):
        return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))

    def sigmoid(x):
        return np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x))))

    def logits(x):  # x [c, 1, n]
        h = x
        for i, (m, b) in enumerate(zip(mats, biases)):
            h = np.matmul(softplus(m), h) + b
            if i < len(factors):
                h = h + np.tanh(factors[i]) * np.tanh(h)
        return h

    s = np.arange(-support, support + 1, dtype=np.float32).astype(np.float64)
    samples = (med[:, None].astype(np.float32) + s[None, :].astype(np.float32)).astype(np.float64)[:, None, :]
    lower, upper = logits(samples - 0.5), logits(samples + 0.5)
    sgn = -np.sign(lower + upper)
    pmf = np.abs(sigmoid(sgn * upper) - sigmoid(sgn * lower))[:, 0, :]
    tail = (sigmoid(lower[:, 0, :1]) + sigmoid(-upper[:, 0, -1:]))[:, 0]
    return tables_from_pmfs([pmf[i] for i in range(c)], [-support] * c, tail)
