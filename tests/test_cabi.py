"""CPU-side checks of the boundary: the library loads and exports every symbol the header declares;
the product has no import path into oracle/ and fails loudly without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "conditional-diffusion-model-for-compression_b200")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not all(os.path.exists(os.path.join(PKG, n)) for n in ("libcdc_b200.so", "libcdc_b200_bf16.so", "libcdc_b200_tools.so")):
        g.build()
    return True


def _declared(header):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"#ifdef CDC_TOOLS.*?#endif", "", hdr, flags=re.S)  # tools-build-only declarations
    return set(re.findall(r"\b(cdc_[a-z0-9_]+)\s*\(", hdr)) - {"cdc_b200"}


def test_header_symbols_are_exported(built):
    """The product library exports exactly what the two headers declare: the drop-in boundary (cdc_b200.h) and the
    test / profiling entry points (cdc_b200_tools.h); the measurement hook that corrupts results is NOT in it."""
    from cdc_b200 import _ffi
    assert _declared("cdc_b200.h") == set(_ffi.SYMBOLS), _declared("cdc_b200.h") ^ set(_ffi.SYMBOLS)
    assert _declared("cdc_b200_tools.h") == set(_ffi.TOOLS_SYMBOLS), _declared("cdc_b200_tools.h") ^ set(_ffi.TOOLS_SYMBOLS)
    L = _ffi.lib()
    for s in _ffi.SYMBOLS + _ffi.TOOLS_SYMBOLS:
        assert hasattr(L, s), s
    assert L.cdc_abi_version() == 2
    assert L.cdc_act_dtype() == 1
    assert not hasattr(L, "cdc_debug_graph_skip")


def test_variant_builds_load(built):
    """build() also produces the bf16 variant (precision tests) and the tools variant (A/B switches, graph-skip hook)."""
    from cdc_b200 import _ffi
    assert _ffi.lib("bf16").cdc_act_dtype() == 0
    T = _ffi.lib("tools")
    assert T.cdc_act_dtype() == 1 and hasattr(T, "cdc_debug_graph_skip")


def test_product_reads_no_environment(built):
    """Environment switches exist only under -DCDC_TOOLS: the product sources may not call getenv outside such blocks."""
    csrc = os.path.join(PKG, "csrc")
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh")):
            continue
        src = open(os.path.join(csrc, f)).read()
        depth, tools = [], False
        for ln in src.splitlines():
            t = ln.strip()
            if t.startswith("#if"):
                depth.append(t.startswith("#ifdef CDC_TOOLS"))
            elif t.startswith("#else") and depth:
                depth[-1] = False if depth[-1] else depth[-1]
            elif t.startswith("#endif") and depth:
                depth.pop()
            tools = any(depth)
            if "getenv(" in ln and not ln.lstrip().startswith("//"):
                assert tools, f"{f}: getenv outside #ifdef CDC_TOOLS: {ln.strip()}"


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/_ref" not in src, f


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cdc_b200 import CDCConfig, Decoder
    with pytest.raises(RuntimeError):
        Decoder(CDCConfig(), {}, device="cuda:0")


def test_synthetic_weights_match_the_oracle_state_dict():
    """The product's random-weight generator (bench / smoke) must produce exactly the tensor names and
    shapes of the oracle's state dicts, which is what cdc_load_weights consumes."""
    from cdc_b200.synthetic import random_weights
    from oracle.config import CDCConfig
    from oracle.weights import build_codec, build_unet
    w = random_weights(with_codec=True)
    ref = dict(build_unet(CDCConfig()).state_dict())
    for k, v in build_codec(CDCConfig()).state_dict().items():
        ref[k if k.startswith("context.") else "codec." + k] = v
    assert set(w) == set(ref), set(w) ^ set(ref)
    assert all(w[k].shape == ref[k].shape for k in w)
    assert set(random_weights()) == {k for k in ref if not k.startswith("codec.")}


def test_product_factorised_table_builder_equals_the_oracle():
    import numpy as np
    from cdc_b200.codec import Codec
    from oracle.config import CDCConfig
    from oracle.weights import build_codec
    oc = build_codec(CDCConfig(), seed=1)
    t = Codec.prior_tables({"codec." + k: v for k, v in oc.state_dict().items()})
    _, fact = oc.tables()
    for f in ("cdf", "row_start", "cdf_length", "offset"):
        assert np.array_equal(getattr(t, f), getattr(fact, f)), f
