"""cdc_b200 -- B200-native decode hot path of a CDC-style conditional-diffusion codec.

Host-side mirror of the oracle's decoder interface (`decode(latent, steps)`,
`denoise_step(x_t, t, cond)`) over the C ABI in include/cdc_b200.h.  PyTorch is used for device
memory, streams and torch.distributed only; all compute runs in libcdc_b200.so (hand-written
sm_100a CUDA).  There is no CPU fallback: importing works anywhere, but every compute call
raises if the shared library or an sm_100 device is missing.
"""
from .config import CDCConfig  # noqa: F401
from .decoder import Decoder, cdf_lookup, quantize_symbols  # noqa: F401
from .codec import Codec  # noqa: F401
from . import _ffi, dp  # noqa: F401

__all__ = ["CDCConfig", "Decoder", "Codec", "quantize_symbols", "cdf_lookup"]
