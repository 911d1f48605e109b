"""World-size-2 gloo test (CPU) of the image-sharded DP driver: index arithmetic covers every image
exactly once, the metric gather works, and the aggregate is total images / slowest rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cdc_b200 import dp


def test_shards_partition_the_images():
    for n in (0, 1, 7, 8, 1024):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in dp.shard_indices(n, r, world))
            assert seen == list(range(n))
            sizes = [len(dp.shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert dp.shard_indices(1024, 3, 8)[:3] == [3, 11, 19]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    done = []

    def fake_decode(i):  # stands in for Decoder.decode on CPU; the DP logic does not depend on the kernel
        return torch.full((1, 3, 4, 4), float(i))

    n, secs = dp.decode_sharded(fake_decode, n_images, rank, world, on_result=lambda i, t: done.append((i, float(t.mean()))))
    table = dp.gather_metrics([n, max(secs, 1e-6), sum(i for i, _ in done)])
    dist.barrier()
    q.put((rank, done, table.tolist(), dp.aggregate_throughput(table)))
    dist.destroy_process_group()


def test_two_rank_gloo_decode_and_gather():
    world, n_images = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    decoded = sorted(i for _, done, _, _ in results for i, v in done if v == float(i))
    assert decoded == list(range(n_images))
    for rank, done, table, agg in results:
        assert [row[0] for row in table] == [5.0, 4.0]
        assert sum(row[2] for row in table) == sum(range(n_images))
        assert agg > 0
