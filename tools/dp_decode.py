#!/usr/bin/env python
"""BASELINE.json configs[3]: image-sharded data-parallel decode of N x 768x512 synthetic images, 17-step DDIM, through
cdc_b200.dp (rank r decodes images i = r mod world; no collective on the data path; one all_gather of (images, seconds)
per rank at the end).  Latents and x_T are pre-staged on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_decode.py --images 1024 > profiles/cfg4_r2.log
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdc_b200 import CDCConfig, Decoder, dp  # noqa: E402
from cdc_b200.synthetic import init_noise, latent, random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=17)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=768)
    a = ap.parse_args()
    rank, world = dp.env_rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    dec = Decoder(CDCConfig(), random_weights(CDCConfig(), seed=0, with_context=True), device=dev)
    dec.set_sample_schedule(a.steps)
    mine = dp.shard_indices(a.images, rank, world)
    # pre-stage: a handful of distinct inputs per rank, cycled (1024 x 11 MB would only measure the allocator)
    pool = 8
    lat = [latent(1, a.height, a.width, index=i).to(dev) for i in mine[:pool]]
    x = [init_noise(1, a.height, a.width, index=i).to(dev) for i in mine[:pool]]
    slot = {i: j % len(lat) for j, i in enumerate(mine)}
    checksum = torch.zeros((), device=dev, dtype=torch.float64)

    def decode(i):
        return dec.decode(lat[slot[i]], a.steps, init=x[slot[i]])

    def keep(i, img):
        checksum.add_(img.double().sum())

    for i in mine[:2]:
        decode(i)  # warm-up (graph capture)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n, secs = dp.decode_sharded(decode, a.images, rank, world, on_result=keep)
    table = dp.gather_metrics([n, secs, float(checksum.item())], device=dev)
    if rank == 0:
        print(json.dumps({"config": "cfg4: image-sharded DP decode", "images": a.images, "image": [a.height, a.width],
                          "ddim_steps": a.steps, "n_gpus": world, "images_per_rank": [int(r[0]) for r in table.tolist()],
                          "seconds_per_rank": [round(r[1], 3) for r in table.tolist()],
                          "images_per_s": dp.aggregate_throughput(table[:, :2]), "wall_s": round(time.perf_counter() - t0, 3),
                          "sum_of_pixels": float(table[:, 2].sum()), "saturation_events": dec.saturation_count()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
