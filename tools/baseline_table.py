#!/usr/bin/env python
"""Numbers for BASELINE.md section 5 that bench.py does not print: cfg1 (one step, 1x256x256) on the host CPU oracle and
on the GPU, cfg3 (16 x 256x256, 100 steps) and the cfg5 decode on one GPU, and the integer path on the CPU oracle at the
cfg5 symbol count.  Run under gpurun; prints one JSON object."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cdc_b200 import CDCConfig, Decoder  # noqa: E402
from cdc_b200.synthetic import init_noise, latent, random_weights  # noqa: E402

out = {"host_cores": os.cpu_count()}
dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device="cuda:0")


def gpu_decode(B, H, W, K, reps=3):
    dec.set_sample_schedule(K)
    lat, x = latent(B, H, W).cuda(), init_noise(B, H, W).cuda()
    for _ in range(2):
        dec.decode(lat, K, init=x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dec.decode(lat, K, init=x)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    return ms, dec.flops_per_step() * K / (ms * 1e-3) / 1e12


ms, tf = gpu_decode(1, 256, 256, 1, reps=7)
out["cfg1_gpu"] = {"ms_per_step_incl_context_net": ms, "note": "K = 1 decode: context net + one graph-launched step + image conversion"}
ms, tf = gpu_decode(16, 256, 256, 100)
out["cfg3_gpu"] = {"ms_per_batch": ms, "images_per_s": 16 / ms * 1e3, "tflops_algorithmic": tf}
ms, tf = gpu_decode(1, 2048, 2048, 17)
out["cfg5_decode_gpu"] = {"ms_per_image": ms, "images_per_s": 1e3 / ms, "tflops_algorithmic": tf}
ms, tf = gpu_decode(1, 512, 768, 17, reps=5)
out["cfg2_gpu_unflushed"] = {"ms_per_image": ms, "images_per_s": 1e3 / ms, "tflops_algorithmic": tf}

# ---- CPU oracle (checker / baseline only) ----
from oracle import entropy as oe  # noqa: E402
from oracle.config import CDCConfig as OCfg  # noqa: E402
from oracle.sampler import OracleDecoder  # noqa: E402
from oracle.weights import build_unet, synthetic_cond, synthetic_entropy_inputs, synthetic_init  # noqa: E402

torch.set_num_threads(os.cpu_count())
ocfg = OCfg()
orc = OracleDecoder(ocfg, build_unet(ocfg).to(memory_format=torch.channels_last))
orc.set_sample_schedule(17)
x, cond = synthetic_init(1, 256, 256), synthetic_cond(ocfg, 1, 256, 256)
ts = []
for _ in range(6):
    t0 = time.perf_counter()
    orc.denoise_step(x, 999, cond)
    ts.append(time.perf_counter() - t0)
out["cfg1_cpu"] = {"s_per_step": statistics.median(ts[1:]), "cores": os.cpu_count()}
n = 4194304
y, mu, sigma = synthetic_entropy_inputs(n)
tb = oe.build_gaussian_tables()
t0 = time.perf_counter()
q, _ = oe.quantize_symbols(y, mu)
t1 = time.perf_counter()
oe.cdf_lookup(q, sigma, tb)
t2 = time.perf_counter()
out["int_cpu"] = {"symbols": n, "quantize_s": t1 - t0, "quantize_gbs": 16 * n / (t1 - t0) / 1e9, "cdf_lookup_s": t2 - t1,
                  "cdf_lookup_gbs": 28 * n / (t2 - t1) / 1e9, "cores": os.cpu_count()}
print(json.dumps(out))
