"""Decode the same inputs with GroupNorm 1 fused into conv2 for up to n N tiles (plan option FUSE_APPLY = n) and unfused
(= 0): the results must be bit-identical.  Usage: fuse_check.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import CDCConfig, Decoder, _ffi
from cdc_b200.synthetic import init_noise, latent, random_weights
n = sys.argv[1] if len(sys.argv) > 1 else "2"
w = random_weights(CDCConfig(), seed=0, with_context=True)
for B, H, W in ((1, 512, 768), (2, 128, 192), (3, 64, 128), (1, 256, 384)):
    lat, x = latent(B, H, W, index=3), init_noise(B, H, W, index=3)
    outs = []
    for fuse in (n, "0"):
        d = Decoder(CDCConfig(), w, device="cuda:0")
        d.set_plan_option(_ffi.OPT_FUSE_APPLY, int(fuse))
        outs.append(d.decode(lat, 3, init=x).clone())
        k = sum(1 for o in d.step_ops() if o[0].endswith("+gn_in"))
        del d
        print(f"B={B} {H}x{W} FUSE_APPLY={fuse}: {k} convs with the input GroupNorm fused")
    print("   bit-identical:", torch.equal(outs[0], outs[1]), " finite:", bool(torch.isfinite(outs[0]).all()))
