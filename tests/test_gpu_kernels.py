"""Kernel-level parity on a real B200, every call through the C ABI (include/cdc_b200.h, kernel-level entry points in include/cdc_b200_tools.h).
Checker: torch fp32 ops for the floating-point kernels, the CPU oracle for the integer path."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False  # the torch checker must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"


def _lib():
    from cdc_b200 import _ffi
    return _ffi.lib()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _act_dtype():
    """Storage / MMA-operand dtype the library was built with (csrc/act.cuh): fp16 by default."""
    return torch.float16 if _lib().cdc_act_dtype() == 1 else torch.bfloat16


def _nhwc_act(t):
    return t.permute(0, 2, 3, 1).contiguous().to(_act_dtype())


def _mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(*shape, generator=g)).bfloat16().float().to(DEV)


def _check(o, ref, what):
    err = (o - ref).abs()
    # one output rounding: 2^-9 relative for bf16, 2^-12 for fp16 (+ fp32 accumulation-order noise)
    tol = (8e-3 if _act_dtype() == torch.bfloat16 else 1.5e-3) * ref.abs().clamp(min=1.0) + 1e-3
    bad = (err > tol).float().mean().item()
    assert torch.isfinite(o).all(), f"{what}: non-finite output"
    assert bad == 0.0, f"{what}: {bad:.4%} elements off, max err {err.max().item():.4f}, ref max {ref.abs().max().item():.3f}"


@pytest.mark.parametrize("C_,HW,silu,res,film", [(64, 4096, 1, 0, 1), (128, 1000, 1, 1, 0), (192, 640, 0, 0, 0),
                                                 (256, 256, 1, 1, 1)])
def test_groupnorm_apply(C_, HW, silu, res, film):
    L = _lib()
    B = 2
    x = _mk((B, HW, C_), 30, 2.0) + 0.5
    x = x.to(_act_dtype())
    r = _mk((B, HW, C_), 31).to(_act_dtype()) if res else None
    gamma, beta = _mk((C_,), 32) + 1.0, _mk((C_,), 33)
    fl = _mk((2 * C_,), 34, 0.3) if film else None
    y = torch.empty_like(x)
    rc = L.cdc_test_gn(_ptr(x), _ptr(r), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(fl), B, HW, C_, silu, 1e-5,
                       C.c_void_p(0))
    assert rc == 0
    xf = x.float().permute(0, 2, 1)  # [B, C, HW]
    ref = F.group_norm(xf, 32, gamma, beta, 1e-5)
    if film:
        ref = ref * (1 + fl[:C_])[None, :, None] + fl[C_:][None, :, None]
    if silu:
        ref = F.silu(ref)
    if res:
        ref = ref + r.float().permute(0, 2, 1)
    _check(y.float().permute(0, 2, 1), ref, "gn_apply")


@pytest.mark.parametrize("B,N", [(1, 256), (2, 64), (1, 1536), (1, 48), (1, 200)])
def test_attention_matches_sdpa(B, N):
    L = _lib()
    qkv = _mk((B, N, 768), 40, 1.5).to(_act_dtype())
    out = torch.empty(B, N, 256, device=DEV, dtype=_act_dtype())
    assert L.cdc_test_attention(_ptr(qkv), _ptr(out), B, N, 4, C.c_void_p(0)) == 0
    torch.cuda.synchronize()
    q, k, v = [t.float().reshape(B, N, 4, 64).transpose(1, 2) for t in qkv.split(256, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, 256)
    err = (out.float() - ref).abs().max().item()
    assert err < (2e-2 if _act_dtype() == torch.bfloat16 else 4e-3), f"attention max err {err}"


def test_quantize_and_cdf_lookup_bit_exact():
    from cdc_b200 import cdf_lookup, quantize_symbols
    from oracle import entropy as oe
    from oracle.weights import synthetic_entropy_inputs
    tb = oe.build_gaussian_tables()
    for n in (0, 1, 1000, 98304, 1 << 20):
        y, mu, sigma = synthetic_entropy_inputs(n, seed=3000 + n)
        if n == 0:
            continue
        q_ref, yh_ref = oe.quantize_symbols(y, mu)
        q, yh = quantize_symbols(y, mu, device=DEV)
        assert torch.equal(q.cpu(), q_ref) and torch.equal(yh.cpu(), yh_ref)
        ref = oe.cdf_lookup(q_ref, sigma, tb)
        got = cdf_lookup(q, sigma, tb, device=DEV)
        for a, b_, nm in zip(got, ref, ("idx", "v", "lo", "hi", "raw")):
            assert torch.equal(a.cpu(), b_), nm
    # half-to-even and exact table thresholds
    y = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 3.5, -2.5])
    q, _ = quantize_symbols(y, torch.zeros_like(y), device=DEV)
    assert q.cpu().tolist() == [0, 2, 2, 0, -2, 4, -2]
    t = torch.from_numpy(tb.scale_table)
    sig = torch.cat([t, torch.nextafter(t, torch.tensor(1e9)), torch.nextafter(t, torch.tensor(0.0)),
                     torch.tensor([0.0, 1e-3, 1e6])])
    qq = torch.zeros(sig.numel(), dtype=torch.int32)
    ref = oe.cdf_lookup(qq, sig, tb)
    got = cdf_lookup(qq, sig, tb, device=DEV)
    for a, b_ in zip(got, ref):
        assert torch.equal(a.cpu(), b_)


def test_golden_entropy_vectors_on_gpu():
    from cdc_b200 import cdf_lookup, quantize_symbols
    from oracle import entropy as oe
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "entropy.npz"))
    tb = oe.CDFTables(g["cdf"], g["row_start"], g["cdf_length"], g["offset"], g["scale_table"])
    q, _ = quantize_symbols(torch.from_numpy(g["y"]), torch.from_numpy(g["mu"]), device=DEV)
    assert np.array_equal(q.cpu().numpy(), g["q"])
    got = cdf_lookup(q, torch.from_numpy(g["sigma"]), tb, device=DEV)
    for a, nm in zip(got, ("idx", "v", "lo", "hi", "raw")):
        assert np.array_equal(a.cpu().numpy(), g[nm]), nm


def test_factorised_prior_lookup_by_channel():
    from cdc_b200 import cdf_lookup, quantize_symbols
    from oracle import entropy as oe
    from oracle.weights import build_codec
    codec = build_codec()
    _, fact = codec.tables()
    g = torch.Generator().manual_seed(5)
    z = 6.0 * torch.randn(2, 256, 4, 6, generator=g)
    med = codec.prior.median.detach()
    q_ref, zh_ref = oe.quantize_symbols(z, med[None, :, None, None].expand_as(z))
    q, zh = quantize_symbols(z, med, device=DEV, per_channel=True)
    assert torch.equal(q.cpu(), q_ref) and torch.equal(zh.cpu(), zh_ref)
    ch = torch.arange(256, dtype=torch.int32)[None, :, None, None].expand_as(q_ref).contiguous()
    ref = oe.lookup_rows(q_ref, ch, fact)
    got = cdf_lookup(q, None, fact, device=DEV)
    for a, b_ in zip(got, ref):
        assert torch.equal(a.cpu(), b_)
