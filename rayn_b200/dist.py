"""Multi-GPU: film tiles shard across ranks, NCCL only for the final film gather.

The reference's only parallelism is one rayon task per 16x16 tile over shared read-only state
(film.rs:640-649); tiles never communicate (film.rs:439-627).  So one process per GPU renders
the tiles with `tile_index % world == rank` (interleaved: fractal scenes concentrate work in
the image centre, contiguous slabs would load-imbalance) and the film is assembled with ONE
collective at the end of the frame: an all-gather of dense per-rank tile slabs.  Tiles are
disjoint, so the gather moves bytes and never reduces -> the N-GPU film is bit-identical to
the 1-GPU film.

The GPU path lives behind the C ABI (`rayn_b200_comm_*`, `rayn_b200_render_frame_sharded`, `rayn_b200_film_gather`): the
context owns the NCCL communicator.  The numpy functions below are the host-logic specification of the same gather
(used by the gloo CPU tests); `torch.distributed` is plumbing only (it ships the 128-byte communicator id).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .film import make_frame_desc, tile_grid

CHANNEL_FLOATS = (3, 1, 3, 3)  # color, alpha, background, normal (film.rs:103-120)


def shard_tiles(ntx, nty, rank, world, mode="diagonal"):
    """Ascending tile indices (film.rs:401-425 order: tile_x * n_tiles_y + tile_y) owned by `rank`.

    "diagonal": (tile_x + tile_y) % world == rank — neighbouring tiles in BOTH directions go to
    different ranks, which balances scenes whose cost is concentrated in the image centre.
    "index": tile_index % world == rank (what RaynFrameDesc.tile_offset/tile_stride express); with a
    tile-column height that is a multiple of `world` this hands each rank whole tile rows."""
    if mode == "index":
        return list(range(rank, ntx * nty, world))
    if mode == "diagonal":
        return [tx * nty + ty for tx in range(ntx) for ty in range(nty) if (tx + ty) % world == rank]
    raise ValueError(mode)


def max_slab_floats(width, height, tile_size, world, mode="diagonal"):
    ntx, nty = tile_grid(width, height, *tile_size)
    return max(len(shard_tiles(ntx, nty, r, world, mode)) for r in range(world)) * 10 * tile_size[0] * tile_size[1]


# ---- pack / unpack on plain arrays (host logic; used by the gloo CPU tests and as the spec
#      the CUDA kernel k_film_pack implements) ----------------------------------------------------
def pack_tiles_numpy(planes, width, height, tile_size, tiles):
    tw, th = tile_size
    ntx, nty = tile_grid(width, height, tw, th)
    slab = np.zeros((len(tiles), 10, th, tw), np.float32)
    chans = [planes["color"].reshape(height, width, 3), planes["alpha"].reshape(height, width, 1),
             planes["background"].reshape(height, width, 3), planes["normal"].reshape(height, width, 3)]
    full = np.concatenate(chans, axis=2)  # [H, W, 10]
    for k, idx in enumerate(tiles):
        x0, y0 = (idx // nty) * tw, (idx % nty) * th
        x1, y1 = min(x0 + tw, width), min(y0 + th, height)
        if x0 >= width or y0 >= height:
            continue
        slab[k, :, : y1 - y0, : x1 - x0] = np.moveaxis(full[y0:y1, x0:x1, :], 2, 0)
    return slab.reshape(-1)


def unpack_tiles_numpy(slab, planes, width, height, tile_size, tiles):
    tw, th = tile_size
    ntx, nty = tile_grid(width, height, tw, th)
    slab = np.asarray(slab, np.float32)[: len(tiles) * 10 * tw * th].reshape(len(tiles), 10, th, tw)
    views = [planes["color"].reshape(height, width, 3), planes["alpha"].reshape(height, width, 1),
             planes["background"].reshape(height, width, 3), planes["normal"].reshape(height, width, 3)]
    for k, idx in enumerate(tiles):
        x0, y0 = (idx // nty) * tw, (idx % nty) * th
        x1, y1 = min(x0 + tw, width), min(y0 + th, height)
        if x0 >= width or y0 >= height:
            continue
        c = 0
        for v in views:
            nc = v.shape[2]
            v[y0:y1, x0:x1, :] = np.moveaxis(slab[k, c:c + nc, : y1 - y0, : x1 - x0], 0, 2)
            c += nc


def gather_film_arrays(planes, width, height, tile_size, rank, world, all_gather, mode="diagonal"):
    """Backend-agnostic gather: `all_gather(vec) -> list of world vecs` (equal length)."""
    ntx, nty = tile_grid(width, height, *tile_size)
    n = max_slab_floats(width, height, tile_size, world, mode)
    mine = pack_tiles_numpy(planes, width, height, tile_size, shard_tiles(ntx, nty, rank, world, mode))
    padded = np.zeros(n, np.float32)
    padded[: mine.size] = mine
    for r, slab in enumerate(all_gather(padded)):
        if r != rank:
            unpack_tiles_numpy(slab, planes, width, height, tile_size, shard_tiles(ntx, nty, r, world, mode))
    return planes


def c_shard_tiles(width, height, tile_size, rank, world):
    """The shard the C ABI uses (rayn_b200_shard_tiles): must equal shard_tiles(..., "diagonal")."""
    lib = L.lib()
    n = lib.rayn_b200_shard_tiles(width, height, tile_size[0], tile_size[1], rank, world, None, 0)
    if n < 0:
        raise ValueError("rayn_b200_shard_tiles: bad arguments")
    arr = (C.c_int32 * max(n, 1))()
    lib.rayn_b200_shard_tiles(width, height, tile_size[0], tile_size[1], rank, world, arr, n)
    return list(arr[:n])


def init_comm(renderer, rank, world, group=None):
    """Create the context-owned NCCL communicator (rayn_b200_comm_init_rank).  The 128-byte unique id travels from rank 0
    to the others through torch.distributed — plumbing only; any transport works (the C++ host could use a file or MPI)."""
    import torch
    import torch.distributed as dist
    lib = L.lib()
    buf = (C.c_uint8 * L.COMM_ID_BYTES)()
    if rank == 0:
        L.check(lib.rayn_b200_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8, device=torch.device("cuda", renderer.device))
    dist.broadcast(t, src=0, group=group)
    ident = (C.c_uint8 * L.COMM_ID_BYTES)(*t.cpu().tolist())
    L.check(lib.rayn_b200_comm_init_rank(renderer.ctx, ident, rank, world), renderer.ctx)


class DistFilm:
    """Device-resident film of one rank.  Multi-GPU goes through the C ABI: the context owns the NCCL communicator and
    `rayn_b200_render_frame_sharded` renders this rank's tiles and all-gathers the film on the render stream
    (pack -> ncclAllGather -> one unpack kernel, no host synchronisation in between).  torch is imported lazily and is
    used only for device memory and to ship the communicator id."""

    def __init__(self, renderer, width, height, tile_size, rank=0, world=1, group=None):
        import torch
        self.torch = torch
        self.r, self.w, self.h, self.tile = renderer, width, height, tuple(tile_size)
        self.rank, self.world = rank, world
        dev = torch.device("cuda", renderer.device)
        npx = width * height
        self.store = torch.zeros(10 * npx, dtype=torch.float32, device=dev)
        self.planes_t = {"color": self.store[: 3 * npx], "alpha": self.store[3 * npx: 4 * npx],
                         "background": self.store[4 * npx: 7 * npx], "normal": self.store[7 * npx:]}
        self.planes = L.RaynFilmPlanes(self.planes_t["color"].data_ptr(), self.planes_t["alpha"].data_ptr(),
                                       self.planes_t["background"].data_ptr(), self.planes_t["normal"].data_ptr(), L.MEM_DEVICE)
        self.tile_list = c_shard_tiles(width, height, self.tile, rank, world) if world > 1 else None
        if world > 1:
            init_comm(renderer, rank, world, group)

    def render(self, frame_desc):
        """Render this rank's tiles into the device film (frame_desc carries this rank's tile list)."""
        self.r.render(frame_desc, self.planes)

    def gather(self):
        """All ranks end with the complete film: rayn_b200_film_gather (asynchronous) + sync."""
        if self.world == 1:
            return
        lib = L.lib()
        L.check(lib.rayn_b200_film_gather(self.r.ctx, self.w, self.h, self.tile[0], self.tile[1], C.byref(self.planes)), self.r.ctx)
        L.check(lib.rayn_b200_sync(self.r.ctx), self.r.ctx)

    def render_gathered(self, frame_desc, planes=None):
        """One call: shard render + gather (rayn_b200_render_frame_sharded); 1 GPU: plain render_frame."""
        planes = planes if planes is not None else self.planes
        lib = L.lib()
        if self.world == 1:
            L.check(lib.rayn_b200_render_frame(self.r.ctx, C.byref(frame_desc), C.byref(planes)), self.r.ctx)
        else:
            L.check(lib.rayn_b200_render_frame_sharded(self.r.ctx, C.byref(frame_desc), C.byref(planes)), self.r.ctx)

    def to_host(self):
        return {k: v.cpu().numpy() for k, v in self.planes_t.items()}


def device_frame_desc(inputs_dev, width, height, tile_size, samples, integrator, frame, time_range, sets, tile_list=None):
    ptrs = tuple(t.data_ptr() for t in inputs_dev)
    return make_frame_desc(width, height, tile_size, samples, integrator, frame, time_range, ptrs, L.MEM_DEVICE, 0, 1, sets, tile_list)
