"""rANS bitstream of the quantised latents over the C ABI (include/cdc_b200.h cdc_rans_*).
The format is pinned by oracle/rans.py and produced byte for byte by csrc/rans.cu: 32-bit-state rANS, 16-bit precision,
every channel row cut into interleaved streams, container "CDCR" | version | n_chan | hw | spc | 0 | sizes | streams."""
import ctypes as C
import struct

import torch

from . import _ffi
from .decoder import DeviceTables, _stream_ptr

MAGIC = b"CDCR"


def streams_per_channel(hw: int) -> int:
    return int(_ffi.lib().cdc_rans_streams_per_channel(int(hw)))


def rans_encode(sym, tables: DeviceTables, n_chan: int, hw: int, device="cuda:0", spc=None) -> torch.Tensor:
    """sym = (idx, v, lo, hi, raw) int32 device tensors of n_chan * hw elements (channel-row major, i.e. NCHW order with
    n_chan = B * C) -> the container as a uint8 DEVICE tensor (exact length)."""
    L = _ffi.lib()
    dev = torch.device(device)
    spc = spc or streams_per_channel(hw)
    s = [t.to(device=dev, dtype=torch.int32).contiguous().reshape(-1) for t in sym]
    if any(t.numel() != n_chan * hw for t in s):
        raise ValueError("symbol arrays must hold n_chan * hw elements")
    with torch.cuda.device(dev):
        scratch = torch.empty(L.cdc_rans_scratch_bytes(n_chan, hw, spc), dtype=torch.uint8, device=dev)
        cap = L.cdc_rans_max_bytes(n_chan, hw, spc)
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        nbytes = torch.zeros(1, dtype=torch.int64, device=dev)
        rc = L.cdc_rans_encode(*[C.c_void_p(t.data_ptr()) for t in s], C.c_void_p(tables.cdf_length.data_ptr()), n_chan, hw, spc,
                               C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), cap, C.c_void_p(nbytes.data_ptr()),
                               _stream_ptr())
        if rc:
            raise RuntimeError(f"cdc_rans_encode failed ({rc})")
        return out[:int(nbytes.item())].clone()


def parse_header(data) -> tuple:
    head = bytes(data[:24].cpu().numpy().tobytes()) if isinstance(data, torch.Tensor) else bytes(data[:24])
    if len(head) < 24 or head[:4] != MAGIC:
        raise ValueError("not a CDCR container")
    ver, n_chan, hw, spc, _ = struct.unpack_from("<5I", head, 4)
    if ver != 1 or spc < 1 or spc > 32 or n_chan < 1 or hw < 1:
        raise ValueError("unsupported CDCR container header")
    return n_chan, hw, spc


def rans_decode(data, idx, tables: DeviceTables, device="cuda:0") -> torch.Tensor:
    """container (bytes, or a uint8 tensor) + the CDF row index of every element -> q int32 [n_chan * hw] (device).
    Raises ValueError if the container is truncated or corrupt."""
    L = _ffi.lib()
    dev = torch.device(device)
    if not isinstance(data, torch.Tensor):
        data = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    n_chan, hw, spc = parse_header(data)
    d = data.to(dev).contiguous()
    ix = idx.to(device=dev, dtype=torch.int32).contiguous().reshape(-1)
    if ix.numel() != n_chan * hw:
        raise ValueError(f"idx holds {ix.numel()} elements, the container {n_chan * hw}")
    if d.numel() < 24 + 4 * n_chan * spc:
        raise ValueError("truncated CDCR container")
    with torch.cuda.device(dev):
        q = torch.zeros(n_chan * hw, dtype=torch.int32, device=dev)
        scratch = torch.empty(8 * n_chan * spc + 8, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        rc = L.cdc_rans_decode(C.c_void_p(d.data_ptr()), d.numel(), C.c_void_p(ix.data_ptr()), C.c_void_p(tables.cdf.data_ptr()),
                               C.c_void_p(tables.row_start.data_ptr()), C.c_void_p(tables.cdf_length.data_ptr()),
                               C.c_void_p(tables.offset.data_ptr()), tables.rows, n_chan, hw, spc, C.c_void_p(scratch.data_ptr()),
                               C.c_void_p(q.data_ptr()), C.c_void_p(status.data_ptr()), _stream_ptr())
        if rc:
            raise RuntimeError(f"cdc_rans_decode failed ({rc})")
        if int(status.item()) != 0:
            raise ValueError("CDCR container is truncated or corrupt")
    return q
