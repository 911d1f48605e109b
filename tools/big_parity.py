"""One-off teacher-forced parity at sizes the GPU test suite does not cover (run under gpurun): a 1024x1536 image and the
cfg3 batch (16 x 256x256), one DDIM step each against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import CDCConfig, Decoder
from oracle.config import CDCConfig as OCfg
from oracle.sampler import OracleDecoder
from oracle.weights import build_unet, synthetic_cond, synthetic_init

ocfg = OCfg()
net = build_unet(ocfg, seed=0)
dec = Decoder(CDCConfig(), dict(net.state_dict()), device="cuda:0")
orc = OracleDecoder(ocfg, net)
dec.set_sample_schedule(17)
orc.set_sample_schedule(17)
torch.set_num_threads(os.cpu_count())
for B, H, W in ((1, 1024, 1536), (16, 256, 256), (2, 192, 320)):
    x, cond = synthetic_init(B, H, W), synthetic_cond(ocfg, B, H, W)
    with torch.no_grad():
        ref = orc.denoise_step(x, 500, cond)
    got = dec.denoise_step(x, 500, cond).cpu()
    e = (got - ref).abs().max().item()
    print(f"B={B} {H}x{W} t=500: max|x_prev - oracle| = {e:.5f}", "OK" if e < 1e-2 else "FAIL")
