"""KATs + golden vectors for the oracle's integer path (rows a8/a9)."""
import os

import numpy as np
import torch

from oracle.entropy import (CDFTables, build_gaussian_tables, build_indexes, cdf_lookup, lookup_rows,
                            pmf_to_quantized_cdf, quantize_symbols, scale_table)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
_TB = None


def tables():
    global _TB
    if _TB is None:
        _TB = build_gaussian_tables()
    return _TB


def test_rint_half_to_even():
    y = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 3.5])
    q, yh = quantize_symbols(y, torch.zeros_like(y))
    assert q.tolist() == [0, 2, 2, 0, -2, 0, 4] and q.dtype == torch.int32
    q2, yh2 = quantize_symbols(torch.tensor([2.75]), torch.tensor([0.25]))
    assert q2.tolist() == [2] and yh2.tolist() == [2.25]


def test_scale_table_endpoints():
    t = scale_table()
    assert t.dtype == np.float32 and len(t) == 64
    assert abs(float(t[0]) - 0.11) < 1e-7 and abs(float(t[63]) - 256.0) < 1e-3
    assert np.all(np.diff(t) > 0)


def test_cdf_rows_monotone_and_total():
    tb = tables()
    assert tb.cdf.shape == (27256,) and int(tb.cdf_length.max()) == 3133
    for r in range(tb.rows):
        row = tb.cdf[tb.row_start[r]: tb.row_start[r] + tb.cdf_length[r]]
        assert row[0] == 0 and row[-1] == 65536
        assert np.all(np.diff(row) > 0)
        assert tb.cdf_length[r] == -2 * tb.offset[r] + 3


def test_build_indexes_ties_and_clamps():
    t = scale_table()
    sig = torch.tensor([0.0, 0.05, 0.11, float(t[1]), float(np.nextafter(t[1], np.float32(10))), float(t[62]), 255.0, 256.0, 1e6])
    idx = build_indexes(sig, t).tolist()
    assert idx[:3] == [0, 0, 0]          # clamped to table[0]; tie counts as <=
    assert idx[3] == 1 and idx[4] == 2   # exact threshold vs one ulp above
    assert idx[5] == 62 and idx[6] == 63 and idx[7] == 63 and idx[8] == 63


def test_pmf_to_quantized_cdf_steals():
    pmf = np.array([0.5, 0.0, 0.25, 0.0, 0.25])
    cdf = pmf_to_quantized_cdf(pmf)
    assert cdf[0] == 0 and cdf[-1] == 65536 and np.all(np.diff(cdf) > 0)
    assert np.diff(cdf).sum() == 65536


def test_lookup_escapes():
    tb = tables()
    # row 0: offset -1, cdf_length 5 -> max_v 3; symbols -1,0,1 in range, others escape
    q = torch.tensor([-1, 0, 1, 2, -2, 7], dtype=torch.int32)
    idx, v, lo, hi, raw = lookup_rows(q, torch.zeros(6, dtype=torch.int32), tb)
    assert v.tolist() == [0, 1, 2, 3, 3, 3]
    assert raw.tolist() == [0, 0, 0, 0, 1, 10]  # v=3>=max_v: 2*(3-3)=0 ; v=-1: 1 ; v=8: 2*(8-3)=10
    row = tb.cdf[:5]
    assert lo.tolist() == [row[0], row[1], row[2], row[3], row[3], row[3]]
    assert hi.tolist() == [row[1], row[2], row[3], row[4], row[4], row[4]]


def test_entropy_matches_golden_bit_exact():
    g = np.load(os.path.join(GOLD, "entropy.npz"))
    tb = tables()
    for k in ("cdf", "row_start", "cdf_length", "offset", "scale_table"):
        assert np.array_equal(g[k], getattr(tb, k)), k
    y, mu, sigma = (torch.from_numpy(g[k]) for k in ("y", "mu", "sigma"))
    q, _ = quantize_symbols(y, mu)
    assert np.array_equal(q.numpy(), g["q"])
    idx, v, lo, hi, raw = cdf_lookup(q, sigma, tb)
    for name, t in (("idx", idx), ("v", v), ("lo", lo), ("hi", hi), ("raw", raw)):
        assert np.array_equal(t.numpy(), g[name]), name
    assert (g["raw"] > 0).sum() > 0, "synthetic inputs must exercise the escape path"
    assert g["idx"].min() == 0 and g["idx"].max() == 63


def test_product_table_builder_equals_the_oracle():
    """The product restates the table construction (it may not import the oracle): identical entry for entry."""
    from cdc_b200.entropy_tables import gaussian_tables, tables_from_pmfs
    from cdc_b200.synthetic import entropy_inputs
    from oracle import entropy as oe
    from oracle.weights import build_codec, synthetic_entropy_inputs
    a, b = gaussian_tables(), oe.build_gaussian_tables()
    for f in ("cdf", "row_start", "cdf_length", "offset", "scale_table"):
        assert np.array_equal(getattr(a, f), getattr(b, f)) and getattr(a, f).dtype == getattr(b, f).dtype, f
    rng = np.random.default_rng(0)
    pm = [rng.dirichlet(np.ones(21) * 0.3) * 0.999 for _ in range(8)]
    pm[3][5] = 0.0  # an empty bin to repair
    c, d = tables_from_pmfs(pm, [-10] * 8, [1e-3] * 8), oe.build_tables_from_pmfs(pm, [-10] * 8, [1e-3] * 8)
    assert np.array_equal(c.cdf, d.cdf) and np.array_equal(c.row_start, d.row_start)
    for x, y in zip(entropy_inputs(4096), synthetic_entropy_inputs(4096)):
        assert torch.equal(x, y)
