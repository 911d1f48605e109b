"""SM clock actually sustained while the 17-step decode graph replays: one warp samples (globaltimer, clock64) every
50 us on a side stream (tools/clockmon.cu -> tools/bin/libclockmon.so) while the graph runs back to back."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cdc_b200 import CDCConfig, Decoder
from cdc_b200.synthetic import init_noise, latent, random_weights
M = C.CDLL(os.path.join(ROOT, "tools/bin/libclockmon.so"))
M.clockmon_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong]
dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device="cuda:0")
dec.set_sample_schedule(17)
lat, x = latent(1, 512, 768, index=0).cuda(), init_noise(1, 512, 768, index=0).cuda()
for _ in range(3): dec.decode(lat, 17, init=x)
torch.cuda.synchronize()
n, period = 4000, 50_000  # 200 ms
buf = torch.zeros(n, 2, dtype=torch.int64, device="cuda")
side = torch.cuda.Stream()
def run(reps, label):
    buf.zero_()
    torch.cuda.synchronize()
    M.clockmon_launch(C.c_void_p(side.cuda_stream), C.c_void_p(buf.data_ptr()), n, period)
    for _ in range(reps): dec.L.cdc_decode(dec.ctx, None)
    torch.cuda.synchronize()
    b = buf.cpu()
    t, c = b[:, 0].double(), b[:, 1].double()
    mhz = (c[1:] - c[:-1]) / (t[1:] - t[:-1]) * 1e3
    k = 200  # 10 ms windows
    wins = [float(mhz[i:i + k].mean()) for i in range(0, n - 1 - k, k)]
    print(label, "SM MHz per 10 ms window:", " ".join(f"{w:.0f}" for w in wins))
run(0, "idle         ")
run(7, "graph x7     ")
run(7, "graph x7 (2) ")
