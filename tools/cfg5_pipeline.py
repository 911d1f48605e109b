#!/usr/bin/env python
"""BASELINE.json configs[4] (judge row J1): end-to-end codec on every GPU -- analysis encoder + hyperprior (latent
rounding, CDF-table lookup: bit-exact integer kernels) + rANS bitstream, then the decoder side: bitstream -> (qz, q) ->
hyper-decoder -> context net -> 17-step 2048 x 2048 diffusion decode with attention over 16384 tokens.  One image per GPU
in flight, images sharded over the ranks (cdc_b200.dp), no collective on the data path.

    python tools/cfg5_pipeline.py --images 2                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/cfg5_pipeline.py --images 16 > profiles/cfg5_r2.log

Prints ONE JSON line (rank 0): whole-job images/s (encode + entropy coding + decode), the per-stage device times of the
slowest rank, bytes per image, and the checks every image passed: symbols recovered bit-exactly from the bytes, the
decoder side re-deriving y_hat bit for bit, finite image in [0, 1]."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdc_b200 import CDCConfig, Codec, Decoder, dp  # noqa: E402
from cdc_b200.synthetic import image, init_noise, random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2, help="images in the job (sharded over the ranks)")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=17)
    a = ap.parse_args()
    rank, world = dp.env_rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cfg = CDCConfig()
    w = random_weights(cfg, seed=0, with_context=True, with_codec=True)
    dec = Decoder(cfg, w, device=dev)
    codec = Codec(dec, fact_tables=Codec.prior_tables(w))
    dec.set_sample_schedule(a.steps)
    H = W = a.size
    mine = dp.shard_indices(a.images, rank, world)
    stage_ms = {"encode": 0.0, "entropy_encode": 0.0, "entropy_decode": 0.0, "diffusion_decode": 0.0}
    stats = {"bytes": 0, "ok": 0}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    # inputs are pre-staged on the device (the job measures the codec, not the synthetic image generator)
    staged = {i: (image(1, H, W, index=i).to(dev), init_noise(1, H, W, index=i).to(dev)) for i in mine}

    def one(i, timed=True):
        img, x_t = staged[i]
        e0 = ev()
        enc = codec.encode(img)                       # analysis + hyper nets, rounding, CDF lookups
        e1 = ev()
        data, _ = None, None
        from cdc_b200.bitstream import rans_encode
        B, c, h, ww = enc["q"].shape
        zs = rans_encode(enc["z_sym"], codec.fact, B * c, (h // 4) * (ww // 4), dev)
        ys = rans_encode(enc["y_sym"], codec.gauss, B * c, h * ww, dev)
        import struct
        data = b"CDC5" + struct.pack("<3IQ", B, H, W, zs.numel()) + zs.cpu().numpy().tobytes() + ys.cpu().numpy().tobytes()
        e2 = ev()
        qz, q = codec.decode_symbols(data)            # bytes -> symbols (hyper-decoder inside: sigma -> CDF rows)
        e3 = ev()
        out = codec.decompress(qz, q, a.steps, init=x_t)
        e4 = ev()
        torch.cuda.synchronize()
        ok = (torch.equal(qz, enc["qz"]) and torch.equal(q, enc["q"]) and bool(torch.isfinite(out).all())
              and float(out.min()) >= 0.0 and float(out.max()) <= 1.0)
        y_hat, _, _ = codec.latent_from_symbols(qz, q)
        ok = ok and torch.equal(y_hat, enc["y_hat"])
        if timed:
            for k, (s, e) in zip(stage_ms, ((e0, e1), (e1, e2), (e2, e3), (e3, e4))):
                stage_ms[k] += s.elapsed_time(e)
            stats["bytes"] += len(data)
            stats["ok"] += int(ok)
        return out

    if mine:
        one(mine[0], timed=False)  # warm-up: graph capture, allocations
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n, secs = dp.decode_sharded(one, a.images, rank, world)
    table = dp.gather_metrics([n, secs, stats["bytes"], stats["ok"]] + [stage_ms[k] for k in stage_ms], device=dev)
    if rank == 0:
        slow = int(table[:, 1].argmax())
        per = max(float(table[slow, 0]), 1.0)
        print(json.dumps({
            "config": "cfg5: encoder + hyperprior rounding + CDF lookup (bit-exact) + rANS + %dx%d decode, attention at 1/16" % (H, W),
            "n_gpus": world, "images": a.images, "ddim_steps": a.steps,
            "images_per_s": dp.aggregate_throughput(table[:, :2]),
            "seconds_per_rank": [round(float(x), 3) for x in table[:, 1]],
            "stage_ms_per_image_slowest_rank": {k: round(float(table[slow, 4 + j]) / per, 2) for j, k in enumerate(stage_ms)},
            "bytes_per_image": float(table[:, 2].sum() / max(float(table[:, 0].sum()), 1.0)),
            "bpp": float(8.0 * table[:, 2].sum() / max(float(table[:, 0].sum()), 1.0) / (H * W)),
            "images_verified": int(table[:, 3].sum()),
            "checks": "symbols decoded from the bytes == encoder's (bit-exact); y_hat re-derived on the decoder side == encoder's; image finite in [0,1]",
            "saturation_events": dec.saturation_count(), "data": "synthetic image, random-init weights",
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
