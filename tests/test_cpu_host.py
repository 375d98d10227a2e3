"""Host logic: tile grid, sampler tables, scramble, filter table, scene flattening, film pack/unpack."""
import ctypes as C

import numpy as np
import pytest

from rayn_b200 import _lib as L
from rayn_b200 import configs
from rayn_b200.dist import gather_film_arrays, max_slab_floats, pack_tiles_numpy, shard_tiles, unpack_tiles_numpy
from rayn_b200.film import Film, FrameInputs, tile_grid
from rayn_b200.scene import Dielectric, PathTracingIntegrator, PinholeCamera, Srgb, Vec3


def test_tile_grid_follows_reference_formula():
    # film.rs:399-404 (res + res % tile) / tile, including the partial-tile quirk
    assert tile_grid(256, 256, 16, 16) == (16, 16)
    assert tile_grid(1920, 1080, 16, 16) == (120, 68)
    assert tile_grid(7680, 4320, 16, 16) == (480, 270)
    assert tile_grid(100, 40, 16, 16) == (6, 3)   # 100 % 16 = 4 < 8: last 4 columns are never rendered
    assert tile_grid(104, 40, 16, 16) == (7, 3)   # 104 % 16 = 8: partial tile kept, clipped


def test_sample_set_counts_match_integrator():
    i = PathTracingIntegrator(8, 2)
    assert 1 + i.requested_1d_sample_sets() == 46 and 2 + i.requested_2d_sample_sets() == 254
    i = PathTracingIntegrator(2, 2)
    assert 1 + i.requested_1d_sample_sets() == 16 and 2 + i.requested_2d_sample_sets() == 86


def test_rd_tables_are_a_rotated_additive_recurrence():
    inp = FrameInputs(16, 16, 8, PathTracingIntegrator(1, 2))
    n = inp.spp
    t1 = inp.samples_1d.reshape(inp.sets_1d, n)
    assert (t1 >= 0).all() and (t1 < 1).all()
    alpha = 0.6180339887498949
    d = np.mod(np.diff(t1[0].astype(np.float64)), 1.0)
    assert np.allclose(d, alpha, atol=2e-7)
    assert not np.allclose(t1[0], t1[1])
    t2 = inp.samples_2d.reshape(inp.sets_2d, n, 2)
    dx = np.mod(np.diff(t2[0, :, 0].astype(np.float64)), 1.0)
    dy = np.mod(np.diff(t2[0, :, 1].astype(np.float64)), 1.0)
    assert np.allclose(dx, 0.7548776662466927, atol=2e-7) and np.allclose(dy, 0.5698402909980532, atol=2e-7)
    assert abs(t1.mean() - 0.5) < 0.05


def test_scramble_and_fis_tables():
    a = FrameInputs(20, 12, 1, PathTracingIntegrator(1, 2))
    b = FrameInputs(20, 12, 1, PathTracingIntegrator(1, 2))
    assert np.array_equal(a.scramble, b.scramble)
    assert (a.scramble >= 0).all() and (a.scramble < 1).all()
    assert np.array_equal(a.scramble * 2 ** 24, np.round(a.scramble * 2 ** 24))  # 24-bit mantissa draws
    assert len(np.unique(a.scramble)) > 230
    # seed is x + y*width: pixel (x, y+1) of a width-20 film == pixel (x+20, y) of a wider one
    wide = FrameInputs(40, 12, 1, PathTracingIntegrator(1, 2))
    assert a.scramble[0 + 1 * 20] == wide.scramble[20 + 0 * 40]
    f = a.fis
    assert f[0] == 0 and f[-1] == pytest.approx(1.5) and (np.diff(f) >= 0).all()
    assert 0.2 < f[256] < 0.6  # Blackman-Harris r=1.5: median of the half-filter mass


def test_setup_scene_matches_setup_rs():
    cam, world = configs.setup()
    desc, keep = world.flatten(cam)
    assert desc.n_hitables == 7 and desc.n_materials == 4 and desc.n_lights == 5
    kinds = [desc.hitables[i].kind for i in range(7)]
    assert kinds == [L.HITABLE_SPHERE, L.HITABLE_MANDELBOX] + [L.HITABLE_SPHERE] * 5
    box = desc.hitables[1]
    assert box.iterations == 12 and box.box_l == 1.0 and box.scale == np.float32(-2.1)
    assert box.min_rad_sq == np.float32(0.01) * np.float32(0.01) and box.fixed_rad_sq == np.float32(1.9) * np.float32(1.9)
    assert desc.hitables[0].radius == 100.0 and desc.materials[0].kind == L.MATERIAL_SKY
    assert desc.materials[1].roughness == pytest.approx(8.68)  # new_remap(0.6), material.rs:167-174
    assert desc.volume.has_scattering == 1 and desc.volume.coeff_extinction == np.float32(0.035)
    assert desc.lights[4].rad == 0.25 and list(desc.lights[4].pos) == [0, 0, 0]
    assert desc.hitables[6].radius == np.float32(0.24)
    assert desc.camera.kind == L.CAMERA_PINHOLE
    assert desc.camera.half_size[1] == pytest.approx(np.tan(np.pi / 6), rel=1e-6)
    assert desc.camera.half_pixel_size == pytest.approx(np.tan(np.pi / 6) / 720, rel=1e-6)
    assert list(desc.camera.origin) == [np.float32(-0.45) * np.float32(2.25), np.float32(0.2) * np.float32(2.25), 4.5]
    assert desc.consts.max_marches == 256 and desc.consts.max_vis_marches == 100


def test_baseline_configs():
    for n, (w, h, spp, mb) in {1: (256, 256, 4, 2), 2: (1024, 1024, 128, 4), 3: (1920, 1080, 512, 8), 4: (2048, 2048, 256, 4),
                               5: (7680, 4320, 1024, 8)}.items():
        c = configs.BASELINE_CONFIGS[n]
        assert c["res"] == (w, h) and 4 * c["samples"] == spp and c["max_bounces"] == mb
    c4 = configs.baseline_config(4, res=(32, 32), samples=1)
    d, _ = c4["world"].flatten(c4["camera"])
    assert d.camera.kind == L.CAMERA_THINLENS and d.volume.has_scattering and d.hitables[1].kind == L.HITABLE_MANDELBULB


def test_film_rejects_duplicate_channels():
    with pytest.raises(ValueError):
        Film(["color", "color"], (16, 16))
    with pytest.raises(ValueError):
        Film(["depth"], (16, 16))


def test_pack_unpack_roundtrip_with_clipped_tiles():
    w, h, tile = 104, 40, (16, 16)
    rng = np.random.default_rng(0)
    full = {"color": rng.random(3 * w * h, dtype=np.float32), "alpha": rng.random(w * h, dtype=np.float32),
            "background": rng.random(3 * w * h, dtype=np.float32), "normal": rng.random(3 * w * h, dtype=np.float32)}
    ntx, nty = tile_grid(w, h, *tile)
    for mode in ("diagonal", "index"):
        out = {k: np.zeros_like(v) for k, v in full.items()}
        shards = [shard_tiles(ntx, nty, r, 3, mode) for r in range(3)]
        assert sorted(sum(shards, [])) == list(range(ntx * nty))  # a partition of the tile grid
        assert all(s == sorted(s) for s in shards)
        for r in range(3):
            slab = pack_tiles_numpy(full, w, h, tile, shards[r])
            assert slab.size <= max_slab_floats(w, h, tile, 3, mode)
            unpack_tiles_numpy(slab, out, w, h, tile, shards[r])
        for k in full:
            assert np.array_equal(out[k], full[k])


def test_diagonal_interleave_balances_rows_and_columns():
    ntx, nty = tile_grid(1024, 1024, 16, 16)
    for world in (2, 4, 8):
        sizes = [len(shard_tiles(ntx, nty, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        centre_counts = []
        for r in range(world):
            t = np.array(shard_tiles(ntx, nty, r, world))
            rows, cols = t % nty, t // nty
            centre_counts.append(int(((np.abs(rows - nty / 2) < 8) & (np.abs(cols - ntx / 2) < 8)).sum()))  # where the fractal is
        assert sum(centre_counts) == 15 * 15 and max(centre_counts) - min(centre_counts) <= 2


def test_gather_with_a_fake_collective():
    w, h, tile, world = 64, 48, (16, 16), 4
    rng = np.random.default_rng(1)
    full = {"color": rng.random(3 * w * h, dtype=np.float32), "alpha": rng.random(w * h, dtype=np.float32),
            "background": rng.random(3 * w * h, dtype=np.float32), "normal": rng.random(3 * w * h, dtype=np.float32)}
    ntx, nty = tile_grid(w, h, *tile)
    shards = [shard_tiles(ntx, nty, r, world) for r in range(world)]
    parts = []
    for r in range(world):
        mine = {k: np.zeros_like(v) for k, v in full.items()}
        unpack_tiles_numpy(pack_tiles_numpy(full, w, h, tile, shards[r]), mine, w, h, tile, shards[r])
        parts.append(mine)
    n = max_slab_floats(w, h, tile, world)
    slabs = []
    for r in range(world):
        s = np.zeros(n, np.float32)
        p = pack_tiles_numpy(parts[r], w, h, tile, shards[r])
        s[:p.size] = p
        slabs.append(s)
    for r in range(world):
        got = gather_film_arrays(parts[r], w, h, tile, r, world, lambda v: slabs)
        for k in full:
            assert np.array_equal(got[k], full[k])


def test_cpp_host_flattens_the_same_scene_as_the_python_host(tmp_path):
    """rayn_b200/host (C++ stand-in for setup.rs/main.rs) and rayn_b200/scene.py must hand the C ABI
    byte-identical descriptors for every BASELINE config (no GPU needed: --dump-scene stops before rendering)."""
    import os
    import subprocess
    from rayn_b200 import build
    build.build()
    exe = os.path.join(os.path.dirname(build.OUT), "rayn_host")
    for n in (1, 2, 3, 4, 5):
        out = tmp_path / f"scene{n}.bin"
        subprocess.run([exe, "--config", str(n), "--res", "160", "90", "--dump-scene", str(out)], check=True)
        c = configs.baseline_config(n, res=(160, 90))
        d, keep = c["world"].flatten(c["camera"])
        want = b"".join(bytes(d.hitables[i]) for i in range(d.n_hitables)) + b"".join(bytes(d.materials[i]) for i in range(d.n_materials)) + \
            b"".join(bytes(d.lights[i]) for i in range(d.n_lights)) + bytes(d.camera) + bytes(d.volume)
        assert out.read_bytes() == want, f"config {n}"


def test_c_abi_shard_equals_the_python_specification():
    """rayn_b200_shard_tiles (what render_frame_sharded / the NCCL gather use) is pure host arithmetic: callable without a GPU."""
    from rayn_b200.dist import c_shard_tiles
    for w, h, world in ((176, 104, 3), (100, 40, 2), (1920, 1080, 8), (7680, 4320, 8), (16, 16, 4)):
        ntx, nty = tile_grid(w, h, 16, 16)
        seen = []
        for rank in range(world):
            t = c_shard_tiles(w, h, (16, 16), rank, world)
            assert t == shard_tiles(ntx, nty, rank, world, "diagonal")
            seen += t
        assert sorted(seen) == list(range(ntx * nty))
