"""One eager denoise step at BASELINE.json configs[1] size (1x512x768) for ncu: a warm-up step, then
one profiled step launched op by op through the C ABI (cdc_run_step_op)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cdc_b200 import CDCConfig, Decoder  # noqa: E402
from oracle.config import CDCConfig as OCfg  # noqa: E402
from oracle.weights import build_unet, synthetic_cond, synthetic_init  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 768)
ocfg = OCfg()
dec = Decoder(CDCConfig(), dict(build_unet(ocfg).state_dict()), device="cuda:0")
dec.set_sample_schedule(17)
x, cond = synthetic_init(1, H, W), synthetic_cond(ocfg, 1, H, W)
dec.denoise_step(x, 999, cond)  # warm-up (also binds the shape)
torch.cuda.synchronize()
n = len(dec.step_ops())
for i in range(n):
    dec.run_step_op(i, 8)
torch.cuda.synchronize()
print("profiled", n, "ops")
