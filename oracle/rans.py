"""Oracle rANS bitstream for the quantised latents: consumes the (idx, v, lo, hi, raw)
symbols of `cdf_lookup` / `lookup_rows` and produces the on-wire bytes; the decoder turns the bytes back into q.

Reference file:line: none -- /root/reference/README.md is 0 bytes (upstream only estimates bits).  The coder is the
published range-ANS construction (Duda 2013; word-wise renormalisation as in Giesen's "rans_word"): 32-bit state in
[2^16, 2^32), 16-bit probability precision (the CDF tables' 65536 total), 16-bit words.  The FORMAT is pinned here and
shared with csrc/rans.cu:

  streams   : every channel row (image b, channel c) of `hw` symbols is cut into `spc` interleaved streams; stream j
              codes the symbols i = j, j + spc, j + 2 spc, ... (so that the 32 lanes of a GPU warp read consecutive
              symbols).  spc = streams_per_channel(hw): the largest power of two <= max(1, hw / 64), at most 32.
  symbol    : (start, freq) = (lo, hi - lo) of the CDF row, total 2^16.
  escape    : a symbol that fell outside its row's support (v == cdf_length[idx] - 2) codes the row's tail bin and
              then raw >= 0 as  nb = bit_length(raw + 1) - 1  in 5 uniform bits, followed by the low nb bits of
              raw + 1 (over 16: the high nb - 16 bits, then the low 16) as uniform symbols.
  encode    : symbols are pushed in REVERSE order:  if x >= freq << 16: emit(x & 0xffff), x >>= 16;
              x = ((x // freq) << 16) + x % freq + start,  starting from x = 2^16.
  stream    : u32 final state (little endian), then the emitted u16 words in reverse emission order (= decode order).
  container : "CDCR", u32 version = 1, u32 n_chan (= B * C), u32 hw, u32 spc, u32 0, u32 size[n_chan * spc] in bytes,
              then the streams back to back in (channel row, j) order.
  decode    : x = state;  per symbol: c = x & 0xffff, find v with cdf[v] <= c < cdf[v+1],
              x = freq * (x >> 16) + c - start;  if x < 2^16: x = x << 16 | next word.

Vectorised over streams with numpy (all streams advance in lockstep); exact integers throughout.
Test infrastructure only; see oracle/__init__.py.
"""
import struct

import numpy as np

MAGIC = b"CDCR"
RANS_L = 1 << 16
PREC = 16


def streams_per_channel(hw: int) -> int:
    s = 1
    while s * 2 <= max(1, hw // 64) and s < 32:
        s *= 2
    return s


def _sub_symbols(lo, hi, raw, esc):
    """Per element the (start, freq) of up to four symbols in DECODE order: main, nb, high bits, low bits;
    freq 0 = absent.  All arrays int64 [n]."""
    n = lo.shape[0]
    st = np.zeros((4, n), dtype=np.int64)
    fr = np.zeros((4, n), dtype=np.int64)
    st[0], fr[0] = lo, hi - lo
    r1 = raw.astype(np.int64) + 1
    nb = np.zeros(n, dtype=np.int64)
    nz = esc & (r1 > 0)
    nb[nz] = np.floor(np.log2(r1[nz].astype(np.float64))).astype(np.int64)
    # (float log2 is exact for < 2^53; guard the boundary anyway)
    nb = np.where((1 << np.minimum(nb + 1, 62)) <= r1, nb + 1, nb)
    nb = np.where((1 << nb) > r1, nb - 1, nb)
    nb = np.where(esc, nb, 0)
    bits = r1 - (1 << nb)  # low nb bits of raw + 1
    st[1] = np.where(esc, nb << 11, 0)
    fr[1] = np.where(esc, 1 << 11, 0)
    hi_n = np.maximum(nb - 16, 0)
    has_hi = esc & (nb > 16)
    st[2] = np.where(has_hi, (bits >> 16) << (16 - hi_n), 0)
    fr[2] = np.where(has_hi, 1 << (16 - hi_n), 0)
    lo_n = np.minimum(nb, 16)
    has_lo = esc & (nb > 0)
    st[3] = np.where(has_lo, (bits & ((1 << lo_n) - 1)) << (16 - lo_n), 0)
    fr[3] = np.where(has_lo, 1 << (16 - lo_n), 0)
    return st, fr


def encode(idx, v, lo, hi, raw, cdf_length, n_chan: int, hw: int, spc: int = None) -> bytes:
    """Symbols as flat int arrays of n_chan * hw elements (channel-row major) -> container bytes."""
    spc = spc or streams_per_channel(hw)
    idx, v, lo, hi, raw = (np.asarray(a).reshape(-1).astype(np.int64) for a in (idx, v, lo, hi, raw))
    assert idx.size == n_chan * hw
    esc = v == (np.asarray(cdf_length).astype(np.int64)[idx] - 2)
    st, fr = _sub_symbols(lo, hi, raw, esc)
    ns = n_chan * spc
    kmax = -(-hw // spc)
    chan = np.repeat(np.arange(n_chan), spc)
    j = np.tile(np.arange(spc), n_chan)
    x = np.full(ns, RANS_L, dtype=np.int64)
    words = np.zeros((ns, 4 * kmax + 2), dtype=np.uint16)
    cnt = np.zeros(ns, dtype=np.int64)
    for k in range(kmax - 1, -1, -1):
        i = j + k * spc
        live = i < hw
        e = chan * hw + np.minimum(i, hw - 1)
        for sub in (3, 2, 1, 0):  # reverse of the decode order
            f = np.where(live, fr[sub][e], 0)
            s = st[sub][e]
            act = f > 0
            fs = np.where(act, f, 1)
            emit = act & (x >= (fs << 16))
            words[np.nonzero(emit)[0], cnt[emit]] = (x[emit] & 0xFFFF).astype(np.uint16)
            cnt += emit
            x = np.where(emit, x >> 16, x)
            x = np.where(act, ((x // fs) << 16) + (x % fs) + s, x)
    sizes = (4 + 2 * cnt).astype(np.uint32)
    out = [MAGIC, struct.pack("<5I", 1, n_chan, hw, spc, 0), sizes.astype("<u4").tobytes()]
    for s_ in range(ns):
        out.append(struct.pack("<I", int(x[s_])))
        out.append(words[s_, :cnt[s_]][::-1].astype("<u2").tobytes())
    return b"".join(out)


def parse(data: bytes):
    assert data[:4] == MAGIC
    ver, n_chan, hw, spc, _ = struct.unpack_from("<5I", data, 4)
    assert ver == 1
    ns = n_chan * spc
    sizes = np.frombuffer(data, dtype="<u4", count=ns, offset=24).astype(np.int64)
    off = 24 + 4 * ns + np.concatenate([[0], np.cumsum(sizes)[:-1]])
    assert off[-1] + sizes[-1] == len(data) if ns else True
    return n_chan, hw, spc, sizes, off


def decode(data: bytes, idx, tables) -> np.ndarray:
    """Container bytes + the CDF row index of every element (known to the decoder from sigma / the channel) -> q int32
    flat [n_chan * hw]."""
    n_chan, hw, spc, sizes, off = parse(data)
    idx = np.asarray(idx).reshape(-1).astype(np.int64)
    assert idx.size == n_chan * hw
    ns = n_chan * spc
    buf = np.frombuffer(data, dtype=np.uint8)
    rows = tables.rows
    maxlen = int(tables.cdf_length.max())
    mat = np.full((rows, maxlen), 1 << 30, dtype=np.int64)  # padded CDF rows
    for r in range(rows):
        mat[r, :tables.cdf_length[r]] = tables.cdf[tables.row_start[r]:tables.row_start[r] + tables.cdf_length[r]]
    clen = tables.cdf_length.astype(np.int64)
    offs = tables.offset.astype(np.int64)

    def u16(pos):
        return buf[pos].astype(np.int64) | (buf[pos + 1].astype(np.int64) << 8)

    x = np.zeros(ns, dtype=np.int64)
    for b in range(4):
        x |= buf[off + b].astype(np.int64) << (8 * b)
    pos = off + 4
    end = off + sizes
    chan = np.repeat(np.arange(n_chan), spc)
    j = np.tile(np.arange(spc), n_chan)
    q = np.zeros(n_chan * hw, dtype=np.int64)
    kmax = -(-hw // spc)

    def pop(x, pos, start, freq, act):
        c = x & 0xFFFF
        xn = freq * (x >> 16) + c - start
        need = act & (xn < RANS_L)
        assert np.all(pos[need] + 2 <= end[need]), "truncated stream"
        w = np.where(need, u16(np.minimum(pos, len(buf) - 2)), 0)
        xn = np.where(need, (xn << 16) | w, xn)
        return np.where(act, xn, x), pos + 2 * need

    for k in range(kmax):
        i = j + k * spc
        live = i < hw
        e = chan * hw + np.minimum(i, hw - 1)
        r = idx[e]
        c = x & 0xFFFF
        v = (mat[r] <= c[:, None]).sum(axis=1) - 1
        start = mat[r, v]
        freq = mat[r, v + 1] - start
        x, pos = pop(x, pos, start, np.where(live, freq, 1), live)
        max_v = clen[r] - 2
        esc = live & (v == max_v)
        # escape payload: nb (5 bits), then the low nb bits of raw + 1
        nb = np.where(esc, (x & 0xFFFF) >> 11, 0)
        x, pos = pop(x, pos, nb << 11, 1 << 11, esc)
        hi_n = np.maximum(nb - 16, 0)
        has_hi = esc & (nb > 16)
        hv = np.where(has_hi, (x & 0xFFFF) >> (16 - hi_n), 0)
        x, pos = pop(x, pos, hv << (16 - hi_n), 1 << (16 - hi_n), has_hi)
        lo_n = np.minimum(nb, 16)
        has_lo = esc & (nb > 0)
        lv = np.where(has_lo, (x & 0xFFFF) >> (16 - lo_n), 0)
        x, pos = pop(x, pos, lv << (16 - lo_n), 1 << (16 - lo_n), has_lo)
        raw = (1 << nb) + (hv << 16) + lv - 1
        vv = np.where(esc, np.where(raw & 1, -(raw + 1) // 2, raw // 2 + max_v), v)
        q[e[live]] = (vv + offs[r])[live]
    assert np.all(x == RANS_L) and np.all(pos == end), "stream not consumed exactly"
    return q.astype(np.int32)


def ideal_bits(lo, hi) -> float:
    w = (np.asarray(hi).astype(np.float64) - np.asarray(lo).astype(np.float64))
    return float(-np.log2(w / 65536.0).sum())
