"""Time one conv launch through cdc_test_conv (CDC_TEST_CONV_REPS flushed repetitions, CUDA events inside the library).
Usage: conv_bench.py H W cin cout ksize mode force_bn [stats] [cin2]   (no arguments: the sweep below)"""
import ctypes as C, os, sys
os.environ.setdefault("CDC_TEST_CONV_REPS", "7")
os.environ.setdefault("CDC_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conditional-diffusion-model-for-compression_b200", "libcdc_b200_tools.so"))  # env switches live in the tools build (-DCDC_TOOLS)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cdc_b200 import _ffi
L = _ffi.lib()


def bench(H, W, cin, cout, ks, mode, fbn, stats=1, cin2=0, B=1):
    dt = torch.float16 if L.cdc_act_dtype() == 1 else torch.bfloat16
    x = torch.randn(B, H, W, cin, device="cuda").to(dt)
    x2 = torch.randn(B, H, W, cin2, device="cuda").to(dt) if cin2 else None
    w = (torch.randn(cout, cin + cin2, ks, ks, device="cuda") / 24).float()
    b = torch.zeros(cout, device="cuda")
    OH, OW = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 2 else (H, W))
    out = torch.empty(B, OH, OW, (cout + 63) // 64 * 64, device="cuda", dtype=dt)
    st = torch.zeros(B, 32, 2, device="cuda", dtype=torch.int64) if stats else None
    print(f"H{H} W{W} cin{cin}+{cin2} cout{cout} k{ks} mode{mode} force_bn{fbn}: ", end="", flush=True)
    rc = L.cdc_test_conv(0, C.c_void_p(x.data_ptr()), cin, C.c_void_p(x2.data_ptr()) if cin2 else None, cin2, B, H, W,
                         C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), cout, ks, mode, fbn, None,
                         C.c_void_p(out.data_ptr()), C.c_void_p(st.data_ptr()) if stats else None, None)
    assert rc == 0, L.cdc_last_error(None)
    torch.cuda.synchronize()


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    if a:
        bench(*a)
    else:
        for fbn in (0, 64, 128, 256):
            bench(64, 96, 256, 256, 3, 0, fbn)      # level 3
        for fbn in (0, 64, 128, 256):
            bench(32, 48, 256, 256, 3, 0, fbn)      # mid
        for fbn in (0, 64, 192):
            bench(128, 192, 192, 192, 3, 0, fbn)    # level 2
        for fbn in (0, 64, 128):
            bench(256, 384, 256, 128, 3, 0, fbn)    # up1.rb1.conv1
        bench(512, 768, 64, 64, 3, 0, 0)            # level 0 (kf)
        bench(256, 384, 128, 128, 3, 0, 0)          # level 1 (kf, two N tiles)
