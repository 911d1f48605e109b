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
    if not os.path.exists(os.path.join(PKG, "libcdc_b200.so")):
        g.build()
    return True


def test_header_symbols_are_exported(built):
    from cdc_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "cdc_b200.h")).read()
    declared = set(re.findall(r"\b(cdc_[a-z0-9_]+)\s*\(", hdr)) - {"cdc_b200"}
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    L = _ffi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.cdc_abi_version() == 1


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
    shapes of the oracle's state dict, which is what cdc_load_weights consumes."""
    from cdc_b200.synthetic import random_weights
    from oracle.config import CDCConfig
    from oracle.weights import build_codec, build_unet
    w = random_weights()
    ref = dict(build_unet(CDCConfig()).state_dict())
    ref.update({"context." + k: v for k, v in build_codec(CDCConfig()).context.state_dict().items()})
    assert set(w) == set(ref)
    assert all(w[k].shape == ref[k].shape for k in w)
