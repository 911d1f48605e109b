"""Print the strip kernel's issuer timeline for one level-0-shaped conv (CDC_STRIP_DEBUG=1)."""
import ctypes as C, os, sys
os.environ.setdefault("CDC_STRIP_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import _ffi
L = _ffi.lib()
B, H, W, cin, cout = 1, 512, 768, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dt = torch.float16 if L.cdc_act_dtype() == 1 else torch.bfloat16
x = torch.randn(B, H, W, cin, device="cuda").to(dt)
w = (torch.randn(cout, cin, 3, 3, device="cuda") / 24).float()
b = torch.zeros(cout, device="cuda")
out = torch.empty(B, H, W, (cout + 63) // 64 * 64, device="cuda", dtype=dt)
st = torch.zeros(B * (H * W // 64 + 64) * 64, device="cuda")
pt = C.c_int(0)
for it in range(2):
    rc = L.cdc_test_conv(0, C.c_void_p(x.data_ptr()), cin, None, 0, B, H, W, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                         cout, 3, 0, 0, None, C.c_void_p(out.data_ptr()), C.c_void_p(st.data_ptr()), C.byref(pt), None)
    assert rc == 0, L.cdc_last_error(None)
