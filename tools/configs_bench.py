"""Decode throughput of the other BASELINE.json configs on one GPU (not bench lines: reported in DESIGN.md).
cfg3: batch 16 x 256x256, 100-step DDIM.  cfg5 (decode part): one 2048x2048 image, 17 steps, attention over 16384 tokens."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import CDCConfig, Decoder
from cdc_b200.synthetic import init_noise, latent, random_weights

dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device="cuda:0")
for name, B, H, W, K in (("cfg3", 16, 256, 256, 100), ("cfg5-decode", 1, 2048, 2048, 17), ("cfg2", 1, 512, 768, 17)):
    dec.set_sample_schedule(K)
    lat, x = latent(B, H, W, index=0).cuda(), init_noise(B, H, W, index=0).cuda()
    for _ in range(2):
        out = dec.decode(lat, K, init=x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = dec.decode(lat, K, init=x)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[1]
    fl = dec.flops_per_step() * K
    assert torch.isfinite(out).all()
    print(f"{name}: B={B} {H}x{W} K={K}: {ms:9.2f} ms per batch, {B / ms * 1e3:8.2f} images/s, {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
