"""Host-side mirror of the oracle's decoder interface over the C ABI.

Same surface as oracle/sampler.py OracleDecoder (the reference ships no code, so the oracle
*is* the reference interface for this path):
    set_sample_schedule(steps), predict_x0(x_t, t, cond), denoise_step(x_t, t, cond),
    decode(latent, steps, *, init=None, gamma=0.8, seed=0), quantize_symbols, cdf_lookup.
Tensors at this surface are NCHW fp32 torch tensors; CUDA tensors are used in place, CPU tensors
are staged through pinned memory by the library (cdc_decode_host).
"""
import ctypes as C

import torch

from . import _ffi
from .config import CDCConfig


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


class Decoder:
    def __init__(self, cfg: CDCConfig, weights, device="cuda:0", tables=None):
        """weights: dict name -> tensor (oracle UNet.state_dict(); optional 'context.*' entries from
        oracle ContextNet.state_dict()) or a path to a safetensors file holding the same."""
        if not torch.cuda.is_available():
            raise RuntimeError("cdc_b200.Decoder needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device)
        self.L = _ffi.lib()
        if isinstance(weights, str):
            from safetensors.torch import load_file
            weights = load_file(weights)
        cc = _ffi.CdcConfig(cfg.base, (C.c_int32 * 4)(*cfg.mults), cfg.groups, cfg.heads, cfg.head_dim, cfg.temb,
                            cfg.T, cfg.latent_ch, cfg.gn_eps)
        self.ctx = C.c_void_p()
        rc = self.L.cdc_create(C.byref(cc), self.device.index or 0, C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError(f"cdc_create failed ({rc}): {self.L.cdc_last_error(None).decode()}")
        with torch.cuda.device(self.device):
            for name, t in weights.items():
                d = _f32c(t.detach(), self.device)
                shape = (C.c_int64 * d.dim())(*d.shape)
                _ffi.check(self.ctx, self.L.cdc_load_weights(self.ctx, name.encode(), C.c_void_p(d.data_ptr()), shape,
                                                             d.dim()), f"cdc_load_weights({name})")
            _ffi.check(self.ctx, self.L.cdc_finalize_weights(self.ctx), "cdc_finalize_weights")
        self.has_context_net = bool(self.L.cdc_has_context_net(self.ctx))
        self.tables = tables
        self.steps = 0
        self.shape = None

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.L.cdc_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- schedule ------------------------------------------------------------------------------
    def set_sample_schedule(self, steps: int):
        _ffi.check(self.ctx, self.L.cdc_set_schedule(self.ctx, int(steps)), "cdc_set_schedule")
        self.steps = int(steps)
        self.idx = [self.L.cdc_schedule_index(self.ctx, k) for k in range(self.steps)]
        return self.idx

    def coeffs(self, k):
        a, b = C.c_float(), C.c_float()
        _ffi.check(self.ctx, self.L.cdc_schedule_coeffs(self.ctx, k, C.byref(a), C.byref(b)), "cdc_schedule_coeffs")
        return a.value, b.value

    def _k_of(self, t):
        if not self.steps:
            raise RuntimeError("call set_sample_schedule first")
        if int(t) not in self.idx:
            raise ValueError(f"t={t} is not in the active {self.steps}-step schedule")
        return self.idx.index(int(t))

    # ---- binding -------------------------------------------------------------------------------
    def bind(self, B, H, W):
        if self.shape != (B, H, W):
            _ffi.check(self.ctx, self.L.cdc_bind_io(self.ctx, B, H, W), "cdc_bind_io")
            self.shape = (B, H, W)

    def set_cond(self, cond):
        c = [_f32c(t, self.device) for t in cond]
        B, _, H, W = c[0].shape
        self.bind(B, H, W)
        _ffi.check(self.ctx, self.L.cdc_set_cond(self.ctx, *[C.c_void_p(t.data_ptr()) for t in c], _stream_ptr()),
                   "cdc_set_cond")

    def set_latent(self, latent):
        y = _f32c(latent, self.device)
        B, _, h, w = y.shape
        self.bind(B, h * 16, w * 16)
        _ffi.check(self.ctx, self.L.cdc_set_latent(self.ctx, C.c_void_p(y.data_ptr()), _stream_ptr()), "cdc_set_latent")

    def get_cond(self):
        raise NotImplementedError

    def _set_x(self, x):
        xd = _f32c(x, self.device)
        _ffi.check(self.ctx, self.L.cdc_set_x(self.ctx, C.c_void_p(xd.data_ptr()), _stream_ptr()), "cdc_set_x")

    def _get_x(self, to_image=False):
        B, H, W = self.shape
        out = torch.empty(B, 3, H, W, device=self.device, dtype=torch.float32)
        _ffi.check(self.ctx, self.L.cdc_get_x(self.ctx, C.c_void_p(out.data_ptr()), 1 if to_image else 0,
                                              _stream_ptr()), "cdc_get_x")
        return out

    def _get_x0(self):
        B, H, W = self.shape
        out = torch.empty(B, 3, H, W, device=self.device, dtype=torch.float32)
        _ffi.check(self.ctx, self.L.cdc_get_x0(self.ctx, C.c_void_p(out.data_ptr()), _stream_ptr()), "cdc_get_x0")
        return out

    # ---- the hot path --------------------------------------------------------------------------
    @torch.no_grad()
    def denoise_step(self, x_t, t, cond=None):
        """x_prev = c0_k * clamp(unet(x_t, t, cond)) + c1_k * x_t.  cond=None reuses the bound cond."""
        k = self._k_of(t)
        with torch.cuda.device(self.device):
            if cond is not None:
                self.set_cond(cond)
            self._set_x(x_t)
            _ffi.check(self.ctx, self.L.cdc_denoise_step(self.ctx, k, _stream_ptr()), "cdc_denoise_step")
            return self._get_x()

    @torch.no_grad()
    def predict_x0(self, x_t, t, cond=None):
        self.denoise_step(x_t, t, cond)
        with torch.cuda.device(self.device):
            return self._get_x0()

    @torch.no_grad()
    def decode(self, latent, steps, *, init=None, gamma=0.8, seed=0, cond=None, out=None):
        """latent y_hat fp32 [B,256,H/16,W/16] -> image fp32 [B,3,H,W] in [0,1] (K steps = one graph launch).
        Host tensors take the cdc_decode_host path (H2D + decode + D2H inside the library); page-locked inputs and a
        page-locked `out` (optional, host path only: fp32 [B,3,H,W], returned) are copied from / to directly."""
        if steps != self.steps:
            self.set_sample_schedule(steps)
        B, _, h, w = latent.shape
        H, W = h * 16, w * 16
        if init is None:
            g = torch.Generator().manual_seed(seed)
            init = gamma * torch.randn(B, 3, H, W, generator=g)
        with torch.cuda.device(self.device):
            self.bind(B, H, W)
            if cond is None and not latent.is_cuda and not init.is_cuda:
                # host buffers: pinned staging + H2D/D2H inside the library
                lat = latent.float().contiguous()
                x0 = init.float().contiguous()
                if out is None:
                    out = torch.empty(B, 3, H, W, dtype=torch.float32)
                elif out.shape != (B, 3, H, W) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
                    raise ValueError("out must be a contiguous fp32 host tensor of shape [B,3,H,W]")
                _ffi.check(self.ctx, self.L.cdc_decode_host(self.ctx, C.c_void_p(lat.data_ptr()),
                                                            C.c_void_p(x0.data_ptr()), C.c_void_p(out.data_ptr()),
                                                            _stream_ptr()), "cdc_decode_host")
                return out
            if cond is not None:
                self.set_cond(cond)
            else:
                self.set_latent(latent)
            self._set_x(init)
            _ffi.check(self.ctx, self.L.cdc_decode(self.ctx, _stream_ptr()), "cdc_decode")
            return self._get_x(to_image=True)

    def decode_resident(self):
        """Replay the K-step graph on whatever x / cond are bound (used by bench.py's device-timed leg)."""
        _ffi.check(self.ctx, self.L.cdc_decode(self.ctx, _stream_ptr()), "cdc_decode")

    # ---- introspection -------------------------------------------------------------------------
    def launches_per_step(self):
        return self.L.cdc_launches_per_step(self.ctx)

    def flops_per_step(self):
        return self.L.cdc_flops_per_step(self.ctx)

    def step_ops(self):
        n = self.L.cdc_num_step_ops(self.ctx)
        return [(self.L.cdc_step_op_name(self.ctx, i).decode(), self.L.cdc_step_op_flops(self.ctx, i),
                 self.L.cdc_step_op_bytes(self.ctx, i)) for i in range(n)]

    def run_step_op(self, i, k=0):
        _ffi.check(self.ctx, self.L.cdc_run_step_op(self.ctx, i, k, _stream_ptr()), "cdc_run_step_op")

    def profile_step(self, k=0, warm=1):
        """In-stream device time (us) of every op of step k (launched back to back, CUDA events in the library)."""
        import ctypes as C
        n = self.L.cdc_num_step_ops(self.ctx)
        buf = (C.c_float * n)()
        _ffi.check(self.ctx, self.L.cdc_profile_step(self.ctx, k, warm, buf, _stream_ptr()), "cdc_profile_step")
        return list(buf)

    # ---- integer path (same names as the oracle) -----------------------------------------------
    def quantize_symbols(self, y, mu):
        return quantize_symbols(y, mu, device=self.device)

    def cdf_lookup(self, q, sigma):
        return cdf_lookup(q, sigma, self.tables, device=self.device)


def quantize_symbols(y, mu, device="cuda:0", per_channel=False):
    """q = rint(y - mu) int32 (half-to-even), y_hat = q + mu.  per_channel: mu is a [C] vector for NCHW y."""
    L = _ffi.lib()
    dev = torch.device(device)
    yd = _f32c(y, dev)
    md = _f32c(mu, dev)
    q = torch.empty(yd.shape, dtype=torch.int32, device=dev)
    yh = torch.empty_like(yd)
    n = yd.numel()
    inner, mod = 0, 0
    if per_channel:
        mod = yd.shape[1]
        inner = yd[0, 0].numel()
    elif md.numel() != n:
        raise ValueError("mu must match y elementwise (or pass per_channel=True)")
    with torch.cuda.device(dev):
        rc = L.cdc_quantize(C.c_void_p(yd.data_ptr()), C.c_void_p(md.data_ptr()), C.c_void_p(q.data_ptr()),
                            C.c_void_p(yh.data_ptr()), n, inner, mod, _stream_ptr())
    if rc:
        raise RuntimeError(f"cdc_quantize failed ({rc})")
    return q, yh


class DeviceTables:
    """CDF tables (oracle/entropy.py CDFTables layout) resident on the device."""

    def __init__(self, tables, device="cuda:0"):
        dev = torch.device(device)
        self.cdf = torch.as_tensor(tables.cdf, dtype=torch.int32).to(dev)
        self.row_start = torch.as_tensor(tables.row_start, dtype=torch.int32).to(dev)
        self.cdf_length = torch.as_tensor(tables.cdf_length, dtype=torch.int32).to(dev)
        self.offset = torch.as_tensor(tables.offset, dtype=torch.int32).to(dev)
        self.scale_table = torch.as_tensor(tables.scale_table, dtype=torch.float32).to(dev)
        self.rows = int(self.row_start.numel())


def cdf_lookup(q, sigma, tables, device="cuda:0"):
    """(idx, v, lo, hi, raw) int32.  sigma=None: idx = channel of an NCHW q (factorised prior)."""
    L = _ffi.lib()
    dev = torch.device(device)
    if not isinstance(tables, DeviceTables):
        tables = DeviceTables(tables, dev)
    qd = q.to(device=dev, dtype=torch.int32).contiguous()
    n = qd.numel()
    outs = [torch.empty(qd.shape, dtype=torch.int32, device=dev) for _ in range(5)]
    inner = 1
    sp = C.c_void_p(0)
    sd = None
    if sigma is not None:
        sd = _f32c(sigma, dev)
        sp = C.c_void_p(sd.data_ptr())
    else:
        inner = qd[0, 0].numel()
    with torch.cuda.device(dev):
        rc = L.cdc_cdf_lookup(C.c_void_p(qd.data_ptr()), sp, C.c_void_p(tables.cdf.data_ptr()),
                              C.c_void_p(tables.row_start.data_ptr()), C.c_void_p(tables.cdf_length.data_ptr()),
                              C.c_void_p(tables.offset.data_ptr()),
                              C.c_void_p(tables.scale_table.data_ptr()) if sigma is not None else C.c_void_p(0),
                              tables.rows, inner, *[C.c_void_p(o.data_ptr()) for o in outs], n, _stream_ptr())
    if rc:
        raise RuntimeError(f"cdc_cdf_lookup failed ({rc})")
    return tuple(outs)
