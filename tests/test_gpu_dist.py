"""Multi-GPU contract T4: the film gathered from N tile-sharded ranks (NCCL all-gather of tile
slabs) is bit-identical to the 1-GPU film.  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import small_config
        from rayn_b200 import configs
        from rayn_b200.dist import DistFilm, device_frame_desc
        from rayn_b200.film import Renderer
        c, inp = small_config(3, (176, 104), 2, 3)  # 104 % 16 = 8: clipped tiles in the grid
        dev = torch.device("cuda", rank)
        inputs_dev = [torch.from_numpy(a).to(dev) for a in inp.arrays()]
        r = Renderer(rank)
        r.upload_scene(c["world"], c["camera"])
        film = DistFilm(r, 176, 104, (16, 16), rank, world)
        fd = device_frame_desc(inputs_dev, 176, 104, (16, 16), c["samples"], c["integrator"], 1, configs.frame_time_range(1),
                               (inp.sets_1d, inp.sets_2d), film.tile_list)
        film.render(fd)
        film.gather()
        got = film.to_host()
        ok = True
        if rank == 0:
            full = r.render_host(inp, (16, 16), c["integrator"], configs.frame_time_range(1))
            ok = all(np.array_equal(got[k].view(np.uint32), full[k].view(np.uint32)) for k in full) and float(full["color"].sum()) > 0
        q.put((rank, bool(ok)))
        r.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_film_is_bit_identical_to_one_gpu():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
