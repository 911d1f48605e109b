"""CPU emulation of the CUDA path's bf16 storage points, to budget the per-step error against the
fp32 oracle (tolerance 1e-2 max-abs, BASELINE.json north_star).  Usage: python tools/precision_study.py [size]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle.config import CDCConfig
from oracle.weights import build_unet, synthetic_cond, synthetic_init

def bf(x, on=True):
    return x.bfloat16().float() if on else x

class Emu:
    """flags: x_in (bf16 x_t into stem), conv_out, act (GN-apply output), rb_out (RB output), attn"""
    def __init__(self, net, **fl):
        self.n = net; self.f = dict(x_in=True, conv_out=True, act=True, rb_out=True, attn=True, x_split=False); self.f.update(fl)
    def gn(self, h_fp32, h_st, gn, film=None):
        B, C = h_fp32.shape[:2]
        g = h_fp32.reshape(B, 32, -1)
        mean = g.mean(-1); var = g.var(-1, unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        cpg = C // 32
        mean = mean.repeat_interleave(cpg, 1)[:, :, None, None]; rstd = rstd.repeat_interleave(cpg, 1)[:, :, None, None]
        a = gn.weight[None, :, None, None] * rstd; b = gn.bias[None, :, None, None] - mean * a
        if film is not None:
            s, sh = film
            a = a * (1 + s[:, :, None, None]); b = b * (1 + s[:, :, None, None]) + sh[:, :, None, None]
        return a * h_st + b
    def rb(self, m, x, te):
        f = self.f
        h = m.conv1(x); hs = bf(h, f['conv_out'])
        film = None
        if m.film is not None:
            film = m.film(F.silu(te)).chunk(2, dim=1)
        h = bf(F.silu(self.gn(h, hs, m.gn1, film)), f['act'])
        h2 = m.conv2(h); h2s = bf(h2, f['conv_out'])
        r = x if m.res is None else bf(m.res(x), f['conv_out'])
        return bf(F.silu(self.gn(h2, h2s, m.gn2)) + r, f['rb_out'])
    def attn(self, a, x):
        f = self.f
        B, C, H, W = x.shape
        n = bf(self.gn(x, x, a.gn), f['attn'])
        qkv = bf(a.qkv(n), f['attn'])
        q, k, v = [t.reshape(B, 4, 64, H * W).transpose(2, 3) for t in qkv.chunk(3, 1)]
        o = F.scaled_dot_product_attention(q, k, v)
        o = bf(o.transpose(2, 3).reshape(B, C, H, W), f['attn'])
        return bf(x + a.proj(o), f['attn'])
    @torch.no_grad()
    def __call__(self, x_t, t, cond):
        n, f = self.n, self.f
        te = n.temb(t)
        xin = bf(x_t, f['x_in'])
        if f['x_split']:
            xin = bf(x_t) + bf(x_t - bf(x_t))
        h = bf(n.stem(torch.cat([xin, cond[0]], 1)), f['conv_out'])
        skips = []
        for i, lvl in enumerate(n.down):
            hin = h if i == 0 else torch.cat([h, cond[i]], 1)
            h = self.rb(lvl.rb1, hin, te); h = self.rb(lvl.rb2, h, te); skips.append(h)
            h = bf(lvl.down(h), f['conv_out'])
        h = self.rb(n.mid.rb1, h, te); h = self.attn(n.mid.attn, h); h = self.rb(n.mid.rb2, h, te)
        for i in reversed(range(4)):
            lvl = n.up[str(i)]
            h = bf(lvl.up(h), f['conv_out'])
            h = self.rb(lvl.rb1, torch.cat([h, skips[i]], 1), te); h = self.rb(lvl.rb2, h, te)
        return n.final(h)

if __name__ == "__main__":
    cfg = CDCConfig(); net = build_unet(cfg)
    H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    x = synthetic_init(1, H, W); cond = synthetic_cond(cfg, 1, H, W)
    for t in (999, 0):
        tt = torch.tensor([t])
        with torch.no_grad():
            ref = net(x, tt, cond)
        def run(**fl):
            out = Emu(net, **fl)(x, tt, cond)
            return (out - ref).abs().max().item(), (out - ref).pow(2).mean().sqrt().item()
        print(f"t={t} ref std {ref.std():.3f} max {ref.abs().max():.3f}")
        print("  none (sanity)         max %.5f rms %.5f" % run(x_in=False, conv_out=False, act=False, rb_out=False, attn=False))
        print("  all bf16              max %.5f rms %.5f" % run())
        for k in ("x_in", "conv_out", "act", "rb_out", "attn"):
            fl = dict(x_in=False, conv_out=False, act=False, rb_out=False, attn=False); fl[k] = True
            print(f"  only {k:9s}        max %.5f rms %.5f" % run(**fl))
        print("  all but x_in (split)  max %.5f rms %.5f" % run(x_split=True))
        print("  all but conv_out      max %.5f rms %.5f" % run(conv_out=False))
        print("  all but act           max %.5f rms %.5f" % run(act=False))
        print("  all but rb_out        max %.5f rms %.5f" % run(rb_out=False))
