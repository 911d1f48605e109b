"""In-graph cost of classes of ops by DIFFERENCING: replay time of the captured 17-step graph with the class left out
(CDC_GRAPH_SKIP) subtracted from the full graph's.  One subprocess per variant (the variable is read at capture time).
Round 2 replaced this by per-kernel globaltimer stamps (bench.py --ops-out / Decoder.profile_graph), which do not perturb
the caches; kept as a cross-check.  The switch exists only in the TOOLS build (libcdc_b200_tools.so, -DCDC_TOOLS)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r)
from cdc_b200 import CDCConfig, Decoder
from cdc_b200.synthetic import init_noise, latent, random_weights
dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device="cuda:0")
dec.set_sample_schedule(17)
lat, x = latent(1, 512, 768, index=0).cuda(), init_noise(1, 512, 768, index=0).cuda()
for _ in range(3): dec.decode(lat, 17, init=x)
torch.cuda.synchronize()
L = dec.L
ts = []
for _ in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); L.cdc_decode(dec.ctx, None); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
print("GRAPH_MS", ts[len(ts)//2])
''' % ROOT
variants = [("full", ""), ("no apply", "apply"),
            ("no attention", "sdpa,attn.gn.stats"), ("no res 1x1", ".res"), ("no up convs", ".up"), ("no down convs", ".down"),
            ("no level-0 3x3", "stem,down.0.rb,up.0.rb1.conv,up.0.rb2.conv,final"), ("no level-1 3x3", "down.1.rb1.conv,down.1.rb2.conv,up.1.rb1.conv,up.1.rb2.conv"),
            ("no level-2 3x3", "down.2.rb1.conv,down.2.rb2.conv,up.2.rb1.conv,up.2.rb2.conv"),
            ("no level-3 3x3", "down.3.rb1.conv,down.3.rb2.conv,up.3.rb1.conv,up.3.rb2.conv"), ("no mid", "mid.")]
if len(sys.argv) > 1:
    variants = [v for v in variants if v[0] == "full"] + [(a, a) for a in sys.argv[1:]]
base = None
for name, skip in variants:
    env = dict(os.environ)
    env["CDC_LIB_PATH"] = os.path.join(ROOT, "conditional-diffusion-model-for-compression_b200", "libcdc_b200_tools.so")
    if skip:
        env["CDC_GRAPH_SKIP"] = skip
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    ms = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("GRAPH_MS")]
    if not ms:
        print(name, "FAILED", out.stderr[-400:])
        continue
    if base is None:
        base = ms[0]
        print(f"{name:24s} graph {ms[0]:8.3f} ms  ({ms[0] / 17 * 1e3:7.1f} us / step)")
    else:
        print(f"{name:24s} graph {ms[0]:8.3f} ms  -> class costs {(base - ms[0]) / 17 * 1e3:7.1f} us / step ({100 * (base - ms[0]) / base:4.1f} %)")
