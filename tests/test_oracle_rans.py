"""CPU tests of the oracle's rANS bitstream (oracle/rans.py): exact round trips including
escapes and ragged streams, the coded size against the entropy bound, the container layout, corruption detection."""
import struct

import numpy as np
import pytest
import torch

from oracle import entropy as oe
from oracle import rans
from oracle.weights import synthetic_entropy_inputs


def _symbols(n, tb, seed):
    y, mu, sigma = synthetic_entropy_inputs(n, seed=seed)
    q, _ = oe.quantize_symbols(y, mu)
    return q.numpy(), [t.numpy() for t in oe.cdf_lookup(q, sigma, tb)]


def test_streams_per_channel_rule():
    assert [rans.streams_per_channel(h) for h in (1, 63, 64, 127, 128, 1536, 2048, 16384, 10 ** 6)] == [1, 1, 1, 1, 2, 16, 32, 32, 32]


@pytest.mark.parametrize("n_chan,hw", [(6, 100), (4, 1536), (3, 64), (2, 1), (5, 130), (1, 4099)])
def test_round_trip_is_exact_and_near_the_entropy_bound(n_chan, hw):
    tb = oe.build_gaussian_tables()
    q, (idx, v, lo, hi, raw) = _symbols(n_chan * hw, tb, seed=11 + hw)
    data = rans.encode(idx, v, lo, hi, raw, tb.cdf_length, n_chan, hw)
    assert np.array_equal(rans.decode(data, idx, tb), q)
    nc, h, spc, sizes, off = rans.parse(data)
    assert (nc, h, spc) == (n_chan, hw, rans.streams_per_channel(hw)) and len(sizes) == n_chan * spc
    assert np.all(sizes >= 4) and np.all(sizes % 2 == 0) and off[0] == 24 + 4 * len(sizes)
    esc = int((v == tb.cdf_length[idx] - 2).sum())
    payload_bits = 8 * (len(data) - 24 - 4 * len(sizes))
    ideal = rans.ideal_bits(lo, hi)
    # every stream pays its 32-bit final state (of which ~16 bits are information) and escapes pay 5 + nb bits
    assert ideal <= payload_bits <= ideal + 32 * len(sizes) + 40 * esc + 16


def test_escapes_of_every_size_round_trip():
    tb = oe.build_gaussian_tables()
    qs = np.array([0, 1, -1, 40, -40, 1000, -1000, 70000, -70000, 2 ** 20, -(2 ** 24), 2 ** 30, -(2 ** 30), 3], dtype=np.int32)
    idx = np.array([0, 0, 0, 0, 0, 5, 5, 10, 10, 63, 63, 63, 0, 63], dtype=np.int32)
    ix, v, lo, hi, raw = [t.numpy() for t in oe.lookup_rows(torch.from_numpy(qs), torch.from_numpy(idx), tb)]
    assert (v == tb.cdf_length[ix] - 2).sum() >= 10
    for n_chan, hw in ((1, 14), (2, 7), (14, 1)):
        data = rans.encode(ix, v, lo, hi, raw, tb.cdf_length, n_chan, hw)
        assert np.array_equal(rans.decode(data, ix, tb), qs)


def test_known_answer_single_stream():
    """Hand-checkable: two symbols of a toy row.  Row cdf = [0, 16384, 65536] (+ tail bin handled as a normal bin here):
    encoding s0 = bin 1 then... -- computed with the update rule of the module docstring."""
    # stream of symbols (start, freq): A = (0, 16384), B = (16384, 49152); decode order A, B => encode B first, then A
    x = rans.RANS_L
    x = ((x // 49152) << 16) + (x % 49152) + 16384      # push B: 65536 // 49152 = 1, 65536 % 49152 = 16384
    assert x == (1 << 16) + 16384 + 16384
    assert x < (16384 << 16)                             # no renormalisation before A
    x = ((x // 16384) << 16) + (x % 16384) + 0          # push A
    assert x == (6 << 16) + 0
    # the module must produce exactly this state for that two-symbol stream
    class T:  # one row: bins [0,16384), [16384,65536) and an (unused) empty-ish tail is not allowed, so use 3 bins
        cdf = np.array([0, 16384, 65535, 65536], dtype=np.int32)
        row_start = np.array([0], dtype=np.int32)
        cdf_length = np.array([4], dtype=np.int32)
        offset = np.array([0], dtype=np.int32)
        rows = 1
    idx = np.zeros(2, dtype=np.int32)
    v = np.array([0, 1], dtype=np.int32)
    lo = np.array([0, 16384], dtype=np.int32)
    hi = np.array([16384, 65535], dtype=np.int32)
    data = rans.encode(idx, v, lo, hi, np.zeros(2, dtype=np.int32), T.cdf_length, 1, 2)
    xb = rans.RANS_L
    xb = ((xb // 49151) << 16) + (xb % 49151) + 16384
    xb = ((xb // 16384) << 16) + (xb % 16384)
    assert struct.unpack_from("<I", data, 24 + 4)[0] == xb and len(data) == 24 + 4 + 4
    assert np.array_equal(rans.decode(data, idx, T), np.array([0, 1], dtype=np.int32))


def test_truncated_or_corrupt_streams_are_detected():
    tb = oe.build_gaussian_tables()
    q, (idx, v, lo, hi, raw) = _symbols(3 * 200, tb, seed=5)
    data = bytearray(rans.encode(idx, v, lo, hi, raw, tb.cdf_length, 3, 200))
    with pytest.raises(AssertionError):
        rans.decode(bytes(data[:-2]), idx, tb)
    bad = bytearray(data)
    bad[-1] ^= 0x55
    try:
        out = rans.decode(bytes(bad), idx, tb)
        assert not np.array_equal(out, q)  # a flipped payload byte cannot decode to the same symbols silently
    except (AssertionError, IndexError):
        pass
