"""Known-answer tests pinning the oracle's schedule and DDIM update.
The reference has no tests to mirror (/root/reference/README.md is 0 bytes)."""
import math
import os

import numpy as np
import torch

from oracle.sampler import alphas_cumprod, ddim_update, make_schedule, step_indices

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_step_indices_k17_and_k100():
    assert step_indices(17) == [999, 937, 874, 812, 749, 687, 624, 562, 500, 437, 375, 312, 250, 187, 125, 62, 0]
    i100 = step_indices(100)
    assert i100[:3] == [999, 989, 979] and i100[-3:] == [20, 10, 0]
    assert len(set(i100)) == 100 and sorted(i100, reverse=True) == i100


def test_alphas_cumprod_kat():
    ab = alphas_cumprod(1000)
    assert abs(ab[0] - 0.99995872) < 1e-8
    assert abs(ab[500] - 0.49228517) < 1e-8
    assert abs(ab[999] - 2.4288e-9) < 1e-12
    assert np.all(np.diff(ab) < 0)


def test_coefficients_kat():
    s = make_schedule(17)
    assert abs(float(s.c0[0]) - 0.096425) < 1e-6 and abs(float(s.c1[0]) - 0.995336) < 1e-6
    assert float(s.c0[16]) == 1.0 and float(s.c1[16]) == 0.0
    assert s.c0.dtype == np.float32 and s.c1.dtype == np.float32


def test_ddim_identity():
    """x_t = sqrt(ab_t) x0 + sqrt(1-ab_t) eps with x0_hat = x0 must land on the same (x0, eps) line at t_prev."""
    ab = alphas_cumprod(1000)
    s = make_schedule(17)
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(2, 3, 8, 8, generator=g, dtype=torch.float64) * 2 - 1
    eps = torch.randn(2, 3, 8, 8, generator=g, dtype=torch.float64)
    for k in range(17):
        a_t = ab[s.idx[k]]
        a_p = ab[s.idx[k + 1]] if k < 16 else 1.0
        x_t = math.sqrt(a_t) * x0 + math.sqrt(1 - a_t) * eps
        want = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
        got = ddim_update(x_t.float(), x0.float(), s.c0[k], s.c1[k]).double()
        assert (got - want).abs().max() < 1e-5
    # last step returns clamp(x0_hat) exactly
    x0c = (2 * x0).float()
    assert torch.equal(ddim_update(x_t.float(), x0c, s.c0[16], s.c1[16]), x0c.clamp(-1, 1))


def test_schedule_matches_golden():
    g = np.load(os.path.join(GOLD, "cfg1_step.npz"))
    s17, s100 = make_schedule(17), make_schedule(100)
    assert list(g["idx17"]) == s17.idx and list(g["idx100"]) == s100.idx
    assert np.array_equal(g["c0_17"], s17.c0) and np.array_equal(g["c1_17"], s17.c1)
    assert np.array_equal(g["c0_100"], s100.c0) and np.array_equal(g["c1_100"], s100.c1)
