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
        film.render(fd)       # explicit two-step form: shard render, then rayn_b200_film_gather
        film.gather()
        got = film.to_host()
        # one-call form with HOST planes: rayn_b200_render_frame_sharded renders the shard, gathers and copies out
        from rayn_b200 import _lib as L
        from rayn_b200.film import make_frame_desc
        npx = 176 * 104
        host = {k: np.zeros(n * npx, np.float32) for k, n in (("color", 3), ("alpha", 1), ("background", 3), ("normal", 3))}
        hp = L.RaynFilmPlanes(host["color"].ctypes.data, host["alpha"].ctypes.data, host["background"].ctypes.data, host["normal"].ctypes.data, L.MEM_HOST)
        hd = make_frame_desc(176, 104, (16, 16), c["samples"], c["integrator"], 1, configs.frame_time_range(1),
                             tuple(a.ctypes.data for a in inp.arrays()), L.MEM_HOST, 0, 1, (inp.sets_1d, inp.sets_2d))
        film.render_gathered(hd, hp)
        full = r.render_host(inp, (16, 16), c["integrator"], configs.frame_time_range(1))  # every rank checks: all ranks must hold the film
        ok = all(np.array_equal(got[k].view(np.uint32), full[k].view(np.uint32)) for k in full) and float(full["color"].sum()) > 0
        ok = ok and all(np.array_equal(host[k].view(np.uint32), full[k].view(np.uint32)) for k in full)
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


def _multi_worker(q):
    """ONE process driving two GPUs through rayn_b200_comm_init_all + rayn_b200_render_frame_multi (the shape a Rust host
    that owns the Film would use)."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import small_config
    from rayn_b200 import _lib as L
    from rayn_b200 import configs
    from rayn_b200.film import Renderer, make_frame_desc
    c, inp = small_config(2, (176, 104), 2, 3)
    tr = configs.frame_time_range(1)
    rs = [Renderer(0), Renderer(1)]
    try:
        for r in rs:
            r.upload_scene(c["world"], c["camera"])
        lib = L.lib()
        ctxs = (C.c_void_p * 2)(rs[0].ctx, rs[1].ctx)
        L.check(lib.rayn_b200_comm_init_all(ctxs, 2), rs[0].ctx)
        npx = 176 * 104
        host = {k: np.zeros(n * npx, np.float32) for k, n in (("color", 3), ("alpha", 1), ("background", 3), ("normal", 3))}
        hp = L.RaynFilmPlanes(host["color"].ctypes.data, host["alpha"].ctypes.data, host["background"].ctypes.data, host["normal"].ctypes.data, L.MEM_HOST)
        hd = make_frame_desc(176, 104, (16, 16), c["samples"], c["integrator"], 1, tr, tuple(a.ctypes.data for a in inp.arrays()), L.MEM_HOST, 0, 1,
                             (inp.sets_1d, inp.sets_2d))
        L.check(lib.rayn_b200_render_frame_multi(ctxs, 2, C.byref(hd), C.byref(hp)), rs[0].ctx)
        paths = [r.stats().paths for r in rs]
        for r in rs:
            L.check(lib.rayn_b200_comm_destroy(r.ctx), r.ctx)
        full = rs[1].render_host(inp, (16, 16), c["integrator"], tr)
        ok = all(np.array_equal(host[k].view(np.uint32), full[k].view(np.uint32)) for k in full) and float(full["color"].sum()) > 0
        ok = ok and min(paths) > 0 and sum(paths) == 176 * 104 * c["spp"]
        q.put(bool(ok))
    finally:
        for r in rs:
            r.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_gpus_render_frame_multi():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_multi_worker, args=(q,))
    p.start()
    assert q.get(timeout=300) is True
    p.join(timeout=60)
    assert p.exitcode == 0


def test_single_rank_communicator_and_sharded_render_on_one_gpu():
    """world = 1 through the same entry points (runs on the 1-GPU box): comm_unique_id / comm_init_rank / render_frame_sharded /
    film_gather are exercised end to end and the film equals the plain render."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import small_config
    from rayn_b200 import _lib as L
    from rayn_b200 import configs
    from rayn_b200.dist import c_shard_tiles, shard_tiles
    from rayn_b200.film import Renderer, make_frame_desc, tile_grid
    for w, h, world in ((176, 104, 3), (100, 40, 2), (7680, 4320, 8)):
        for rank in range(world):
            assert c_shard_tiles(w, h, (16, 16), rank, world) == shard_tiles(*tile_grid(w, h, 16, 16), rank, world, "diagonal")
    c, inp = small_config(3, (80, 48), 1, 2)
    tr = configs.frame_time_range(1)
    r = Renderer(0)
    try:
        r.upload_scene(c["world"], c["camera"])
        full = r.render_host(inp, (16, 16), c["integrator"], tr)
        lib = L.lib()
        ident = (C.c_uint8 * L.COMM_ID_BYTES)()
        L.check(lib.rayn_b200_comm_unique_id(ident))
        L.check(lib.rayn_b200_comm_init_rank(r.ctx, ident, 0, 1), r.ctx)
        rk, wd = C.c_int32(-1), C.c_int32(-1)
        L.check(lib.rayn_b200_comm_info(r.ctx, C.byref(rk), C.byref(wd)))
        assert (rk.value, wd.value) == (0, 1)
        npx = 80 * 48
        host = {k: np.zeros(n * npx, np.float32) for k, n in (("color", 3), ("alpha", 1), ("background", 3), ("normal", 3))}
        hp = L.RaynFilmPlanes(host["color"].ctypes.data, host["alpha"].ctypes.data, host["background"].ctypes.data, host["normal"].ctypes.data, L.MEM_HOST)
        hd = make_frame_desc(80, 48, (16, 16), c["samples"], c["integrator"], 1, tr, tuple(a.ctypes.data for a in inp.arrays()), L.MEM_HOST, 0, 1,
                             (inp.sets_1d, inp.sets_2d))
        L.check(lib.rayn_b200_render_frame_sharded(r.ctx, C.byref(hd), C.byref(hp)), r.ctx)
        for k in full:
            assert np.array_equal(host[k].view(np.uint32), full[k].view(np.uint32)), k
        L.check(lib.rayn_b200_comm_destroy(r.ctx), r.ctx)
        with pytest.raises(L.RaynError):
            L.check(lib.rayn_b200_render_frame_sharded(r.ctx, C.byref(hd), C.byref(hp)), r.ctx)
    finally:
        r.close()
