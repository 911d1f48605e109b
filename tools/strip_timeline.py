"""Print the kh-fused conv kernel's issuer / epilogue timeline for one conv (CDC_STRIP_DEBUG=1) and time the launch with CUDA
events (second, warm call).  Usage: strip_timeline.py [cin] [cout] [H] [W] [mode]   (mode 1 = stride 2, 2 = nearest-x2 input)"""
import ctypes as C, os, sys
os.environ.setdefault("CDC_STRIP_DEBUG", "1")
os.environ.setdefault("CDC_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conditional-diffusion-model-for-compression_b200", "libcdc_b200_tools.so"))  # env switches live in the tools build (-DCDC_TOOLS)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import _ffi
L = _ffi.lib()
a = [int(v) for v in sys.argv[1:]]
cin, cout = (a + [64, 64])[:2] if len(a) < 2 else a[:2]
H, W = (a[2], a[3]) if len(a) >= 4 else (512, 768)
mode = a[4] if len(a) >= 5 else 0
OH, OW = (2 * H, 2 * W) if mode == 2 else (H // 2, W // 2) if mode == 1 else (H, W)
B = 1
dt = torch.float16 if L.cdc_act_dtype() == 1 else torch.bfloat16
x = (torch.randn(B, H, W, cin, device="cuda") * float(os.environ.get("XSCALE", "1"))).to(dt)
w = (torch.randn(cout, cin, 3, 3, device="cuda") / 24 * float(os.environ.get("XSCALE", "1"))).float()
b = torch.zeros(cout, device="cuda")
out = torch.empty(B, OH, OW, (cout + 63) // 64 * 64, device="cuda", dtype=dt)
st = torch.zeros(B, 32, 2, device="cuda", dtype=torch.int64)
for it in range(2):
    rc = L.cdc_test_conv(0, C.c_void_p(x.data_ptr()), cin, None, 0, B, H, W, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                         cout, 3, mode, 0, None, C.c_void_p(out.data_ptr()), None if mode != 0 else C.c_void_p(st.data_ptr()), None)
    assert rc == 0, L.cdc_last_error(None)
