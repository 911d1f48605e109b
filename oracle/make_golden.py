"""Generate tests/golden/*.npz from the oracle (run: python -m oracle.make_golden).

The reference ships no fixtures (/root/reference/README.md is 0 bytes), so these
vectors pin the ORACLE ITSELF against silent drift; they are produced by this
script, in this container, and committed with it.
"""
import hashlib
import os

import numpy as np
import torch

from .codec import Codec
from .config import CDCConfig
from .entropy import build_gaussian_tables, cdf_lookup, quantize_symbols
from .sampler import OracleDecoder, make_schedule
from .weights import (build_codec, build_unet, synthetic_cond, synthetic_entropy_inputs, synthetic_image,
                      synthetic_init, synthetic_latent)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.use_deterministic_algorithms(True)
    cfg = CDCConfig()

    # --- cfg1: one denoise step, 1x256x256, fp32 (BASELINE.json configs[0]) ---
    net = build_unet(cfg, seed=0)
    dec = OracleDecoder(cfg, net)
    dec.set_sample_schedule(17)
    x = synthetic_init(1, 256, 256)
    cond = synthetic_cond(cfg, 1, 256, 256)
    out = {}
    for t in (999, 500, 0):
        x0 = dec.predict_x0(x, t, cond)
        xp = dec.denoise_step(x, t, cond)
        out[f"x0_t{t}"] = x0[:, :, ::8, ::8].numpy()
        out[f"xprev_t{t}"] = xp[:, :, ::8, ::8].numpy()
        out[f"sha_x0_t{t}"] = np.array(sha(x0.numpy()))
    s17, s100 = make_schedule(17), make_schedule(100)
    out.update(idx17=np.array(s17.idx), c0_17=s17.c0, c1_17=s17.c1, idx100=np.array(s100.idx),
               c0_100=s100.c0, c1_100=s100.c1)
    np.savez_compressed(os.path.join(OUT, "cfg1_step.npz"), **out)

    # --- integer path: tables + lookups on the section-8d synthetic distribution ---
    tb = build_gaussian_tables()
    y, mu, sigma = synthetic_entropy_inputs(8192)
    q, yhat = quantize_symbols(y, mu)
    idx, v, lo, hi, raw = cdf_lookup(q, sigma, tb)
    np.savez_compressed(os.path.join(OUT, "entropy.npz"), cdf=tb.cdf, row_start=tb.row_start,
                        cdf_length=tb.cdf_length, offset=tb.offset, scale_table=tb.scale_table,
                        y=y.numpy(), mu=mu.numpy(), sigma=sigma.numpy(), q=q.numpy(), idx=idx.numpy(),
                        v=v.numpy(), lo=lo.numpy(), hi=hi.numpy(), raw=raw.numpy())

    # --- codec side (next rows f1/f2): context net + encode on a 128x128 image ---
    codec = build_codec(cfg, seed=1)
    img = synthetic_image(1, 128, 128)
    enc = codec.encode(img)
    with torch.no_grad():
        ctx = codec.context(enc["y_hat"])
    np.savez_compressed(os.path.join(OUT, "codec_128.npz"), y=enc["y"].numpy(), q=enc["q"].numpy(),
                        mu=enc["mu"].numpy(), sigma=enc["sigma"].numpy(), qz=enc["qz"].numpy(),
                        c3=ctx[3].numpy(), c0_sub=ctx[0][:, :, ::8, ::8].numpy())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
