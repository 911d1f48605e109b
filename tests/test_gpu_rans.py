"""rANS bitstream on a real B200: csrc/rans.cu must produce the oracle's container byte for
byte and decode it back exactly; at the cfg5 symbol count (4 194 304) the round trip and the entropy bound are checked as
size-independent properties."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _symbols(n, seed):
    from oracle import entropy as oe
    from oracle.weights import synthetic_entropy_inputs
    tb = oe.build_gaussian_tables()
    y, mu, sigma = synthetic_entropy_inputs(n, seed=seed)
    q, _ = oe.quantize_symbols(y, mu)
    return tb, q, oe.cdf_lookup(q, sigma, tb)


@pytest.mark.parametrize("n_chan,hw", [(6, 100), (4, 1536), (3, 64), (2, 1), (5, 130), (512, 96), (1, 4099), (256, 1536)])
def test_container_is_byte_exact_with_the_oracle(n_chan, hw):
    from cdc_b200.bitstream import rans_decode, rans_encode, streams_per_channel
    from cdc_b200.decoder import DeviceTables
    from oracle import rans
    tb, q, sym = _symbols(n_chan * hw, seed=21 + hw)
    assert streams_per_channel(hw) == rans.streams_per_channel(hw)
    ref = rans.encode(*[t.numpy() for t in sym], tb.cdf_length, n_chan, hw)
    dt = DeviceTables(tb, DEV)
    got = rans_encode([t.to(DEV) for t in sym], dt, n_chan, hw, DEV)
    assert got.dtype == torch.uint8 and bytes(got.cpu().numpy().tobytes()) == ref, (len(ref), got.numel())
    out = rans_decode(ref, sym[0], dt, DEV)            # the oracle's bytes through the GPU decoder
    assert torch.equal(out.cpu(), q.reshape(-1))
    assert torch.equal(rans_decode(got, sym[0].to(DEV), dt, DEV).cpu(), q.reshape(-1))
    if n_chan * hw <= 8192:
        assert np.array_equal(rans.decode(bytes(got.cpu().numpy().tobytes()), sym[0].numpy(), tb), q.numpy().reshape(-1))


def test_huge_escapes_byte_exact():
    from cdc_b200.bitstream import rans_decode, rans_encode
    from cdc_b200.decoder import DeviceTables
    from oracle import entropy as oe
    from oracle import rans
    tb = oe.build_gaussian_tables()
    qs = torch.tensor([0, 1, -1, 40, -40, 1000, -1000, 70000, -70000, 2 ** 20, -(2 ** 24), 2 ** 30, -(2 ** 30), 3], dtype=torch.int32)
    idx = torch.tensor([0, 0, 0, 0, 0, 5, 5, 10, 10, 63, 63, 63, 0, 63], dtype=torch.int32)
    sym = oe.lookup_rows(qs, idx, tb)
    dt = DeviceTables(tb, DEV)
    for n_chan, hw in ((1, 14), (2, 7), (14, 1)):
        ref = rans.encode(*[t.numpy() for t in sym], tb.cdf_length, n_chan, hw)
        got = rans_encode(sym, dt, n_chan, hw, DEV)
        assert bytes(got.cpu().numpy().tobytes()) == ref
        assert torch.equal(rans_decode(got, idx, dt, DEV).cpu(), qs)


def test_truncated_and_corrupt_containers_raise():
    from cdc_b200.bitstream import rans_decode, rans_encode
    from cdc_b200.decoder import DeviceTables
    tb, q, sym = _symbols(8 * 300, seed=9)
    dt = DeviceTables(tb, DEV)
    data = rans_encode(sym, dt, 8, 300, DEV).cpu()
    with pytest.raises(ValueError):
        rans_decode(data[:-2].clone(), sym[0], dt, DEV)
    with pytest.raises(ValueError):
        rans_decode(data[:20].clone(), sym[0], dt, DEV)
    bad = data.clone()
    bad[24] = 255  # first stream's size field: now points far outside the buffer
    bad[25] = 255
    bad[26] = 255
    with pytest.raises(ValueError):
        rans_decode(bad, sym[0], dt, DEV)
    bad = data.clone()
    bad[0] = 0
    with pytest.raises(ValueError):
        rans_decode(bad, sym[0], dt, DEV)
    flip = data.clone()
    flip[-1] ^= 0x55
    try:
        out = rans_decode(flip, sym[0], dt, DEV)
        assert not torch.equal(out.cpu(), q.reshape(-1))
    except ValueError:
        pass


def test_cfg5_symbol_count_round_trip_and_entropy_bound():
    """4 194 304 symbols (a 2048 x 2048 image's y: 256 channel rows of 16384): exact round trip; payload within the
    per-stream overhead of the entropy bound."""
    import time
    from cdc_b200 import cdf_lookup, quantize_symbols
    from cdc_b200.bitstream import rans_decode, rans_encode
    from cdc_b200.decoder import DeviceTables
    from cdc_b200.synthetic import entropy_inputs, gaussian_tables
    n_chan, hw = 256, 16384
    y, mu, sigma = (t.to(DEV) for t in entropy_inputs(n_chan * hw))
    dt = DeviceTables(gaussian_tables(), DEV)
    q, _ = quantize_symbols(y, mu, device=DEV)
    sym = cdf_lookup(q, sigma, dt, device=DEV)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    data = rans_encode(sym, dt, n_chan, hw, DEV)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = rans_decode(data, sym[0], dt, DEV)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    assert torch.equal(out, q.reshape(-1))
    ideal = float(-(torch.log2((sym[3] - sym[2]).double() / 65536.0)).sum().item())
    esc = int((sym[1] == dt.cdf_length[sym[0].long()] - 2).sum().item())
    ns = n_chan * 32
    payload = 8 * (data.numel() - 24 - 4 * ns)
    print(f"4194304 symbols: {data.numel()} bytes ({8 * data.numel() / (n_chan * hw):.3f} bits/symbol; entropy bound "
          f"{ideal / (n_chan * hw):.3f}), {esc} escapes; encode {1e3 * (t1 - t0):.2f} ms, decode {1e3 * (t2 - t1):.2f} ms (incl. allocations)")
    assert ideal <= payload <= ideal + 32 * ns + 40 * esc + 16


def test_codec_compress_decompress_bytes_round_trip():
    """J1 at a test size: image -> bytes -> (qz, q) recovered exactly -> image identical to decompress(qz, q)."""
    from tests.test_gpu_codec import _setup
    from cdc_b200.synthetic import init_noise
    from oracle.weights import synthetic_image
    codec, oc, dec, _ = _setup()
    img = synthetic_image(2, 128, 192, index=2)
    data, enc = codec.compress(img)
    qz, q = codec.decode_symbols(data)
    assert torch.equal(qz, enc["qz"]) and torch.equal(q, enc["q"])
    x = init_noise(2, 128, 192).to(DEV)
    a = codec.decompress_bytes(data, 4, init=x)
    b = codec.decompress(enc["qz"], enc["q"], 4, init=x)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    print(f"2 x 128 x 192: {len(data)} bytes = {8 * len(data) / (2 * 128 * 192):.3f} bpp (untrained weights)")
    with pytest.raises(ValueError):
        codec.decode_symbols(data[:-3])
