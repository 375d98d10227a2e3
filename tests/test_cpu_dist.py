"""world_size-2 gloo test of the multi-GPU host logic: tile sharding + film all-gather.
The collective moves bytes only, so every rank must end with the exact full film."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rayn_b200.dist import gather_film_arrays, pack_tiles_numpy, shard_tiles, unpack_tiles_numpy
        from rayn_b200.film import tile_grid
        tile = (16, 16)
        shard = shard_tiles(*tile_grid(w, h, *tile), rank, world)
        rng = np.random.default_rng(42)  # same "full film" on every rank
        full = {"color": rng.random(3 * w * h, dtype=np.float32), "alpha": rng.random(w * h, dtype=np.float32),
                "background": rng.random(3 * w * h, dtype=np.float32), "normal": rng.random(3 * w * h, dtype=np.float32)}
        mine = {k: np.zeros_like(v) for k, v in full.items()}  # what this rank "rendered": only its own tiles
        unpack_tiles_numpy(pack_tiles_numpy(full, w, h, tile, shard), mine, w, h, tile, shard)

        def all_gather(vec):
            t = torch.from_numpy(vec)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return [o.numpy() for o in outs]

        got = gather_film_arrays(mine, w, h, tile, rank, world, all_gather)
        ok = all(np.array_equal(got[k].view(np.uint32), full[k].view(np.uint32)) for k in full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_film_gather_is_exact():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 104, 72, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
