"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Integer/index work (queues, packet membership) and float work are both demanded BIT-exact:
the two implementations share only IEEE-exact primitives (rayn_b200/csrc/detmath.h), so any
difference is an algorithmic divergence.  The north star's 1e-4 relative tolerance is implied.
"""
import ctypes as C

import numpy as np
import pytest

from rayn_b200 import _lib as L
from rayn_b200 import configs
from rayn_b200.film import FrameInputs, Renderer

from helpers import CH, assert_bit_equal, random_rays, rel_err_stats, small_config

pytestmark = pytest.mark.gpu
TR = configs.frame_time_range(1)


def _sdf(kind):
    cam, world = configs.setup((64, 64), volume=False, fractal=kind)
    desc, keep = world.flatten(cam)
    return desc, keep, desc.hitables[1]


# ---- T0: deterministic math, host vs device -------------------------------------------------
@pytest.mark.parametrize("op,lo,hi", [(0, -90.0, 5.0), (1, 1e-6, 50.0), (3, -7.0, 7.0), (4, -7.0, 7.0), (5, -1.55, 1.55), (7, -2.0, 2.0)])
def test_detmath_unary_bit_equal(renderer, oracle, op, lo, hi):
    rng = np.random.default_rng(op)
    a = rng.uniform(lo, hi, 200_000).astype(np.float32)
    a[:8] = [0.0, -0.0, 1.0, np.nan, np.inf, -np.inf, 1e-38, hi]
    assert_bit_equal(renderer.kat_detmath(op, a), oracle.kat_detmath(op, a), f"detmath op {op}")


@pytest.mark.parametrize("op", [2, 6])
def test_detmath_binary_bit_equal(renderer, oracle, op):
    rng = np.random.default_rng(10 + op)
    if op == 2:
        a = rng.uniform(0.0, 1.5, 200_000).astype(np.float32)
        b = rng.uniform(0.05, 12.0, 200_000).astype(np.float32)
    else:
        a = rng.uniform(-200.0, 200.0, 200_000).astype(np.float32)
        b = rng.uniform(0.0, 50.0, 200_000).astype(np.float32)
    a[:4] = [0.0, 1.0, np.nan, 0.5]
    b[:4] = [2.0, 0.0, 1.0, np.nan]
    assert_bit_equal(renderer.kat_detmath(op, a, b), oracle.kat_detmath(op, a, b), f"detmath op {op}")


# ---- K2 stage tests: SDF dist / sphere-march / occlusion / closest hit ------------------------
@pytest.mark.parametrize("kind", ["mandelbox", "mandelbulb"])
def test_sdf_dist_bit_equal(renderer, oracle, kind):
    desc, keep, h = _sdf(kind)
    rng = np.random.default_rng(1)
    p = rng.uniform(-2.5, 2.5, size=(100_000, 3)).astype(np.float32)
    p[0] = 0.0
    p[1] = [0.0, 0.0, 1.0]
    p[2] = np.nan
    assert_bit_equal(renderer.kat_sdf_dist(h, p), oracle.kat_sdf_dist(h, p), f"{kind} dist")


@pytest.mark.parametrize("kind,variants", [("mandelbox", (-1, 0, 1, 2, 4, 5)), ("mandelbulb", (-1, 3))])
def test_packed_two_point_estimator_bit_equal(renderer, oracle, kind, variants):
    """rt_sdf2.cuh: the f32x2 (FFMA2/FMUL2/FADD2) two-point estimators the march kernels run, every specialisation,
    against the oracle's 4-lane SSE estimator; odd point count exercises the half-filled last pair."""
    desc, keep, h = _sdf(kind)
    rng = np.random.default_rng(5)
    p = rng.uniform(-2.5, 2.5, size=(100_001, 3)).astype(np.float32)
    p[:20_000] *= 0.2          # inside the sphere-fold radius: the division matters there
    p[20_000:20_100] *= 40.0   # far field
    p[0] = 0.0
    p[1] = [0.0, 0.0, 1.0]
    p[2] = np.nan
    p[3] = [np.inf, 0.0, 0.0]
    p[4] = [1e-30, -1e-30, 0.0]
    ref = oracle.kat_sdf_dist(h, p)
    for v in variants:
        assert_bit_equal(renderer.kat_sdf_dist2(h, p, v), ref, f"{kind} packed dist variant {v}")
    if kind == "mandelbox":  # non-default fold constants and iteration counts through the run-time specialisations
        import copy
        for iters, l, mn, fx, sc in [(7, 1.0, 0.25, 1.0, 2.0), (15, 0.8, 1e-4, 3.61, -1.5), (3, 1.5, 0.5, 0.4, -2.1)]:
            h2 = copy.copy(h)
            h2.iterations, h2.box_l, h2.min_rad_sq, h2.fixed_rad_sq, h2.scale = iters, l, mn, fx, sc
            ref2 = oracle.kat_sdf_dist(h2, p)
            for v in (-1, 0, 2) + ((5,) if mn <= fx else ()):  # 5 = three-operation division, needs a non-empty divisor interval
                assert_bit_equal(renderer.kat_sdf_dist2(h2, p, v), ref2, f"mandelbox {iters, l, mn, fx, sc} variant {v}")


def test_three_operation_division_is_selected_only_after_its_exhaustive_check(oracle):
    """upload_scene divides by every float of [min_rad_sq, fixed_rad_sq] on the device and selects the three-operation
    sphere-fold division (variants 4 / 5) only if all quotients equal IEEE division; RAYN_FLAG_NO_DIV3 keeps the
    five-operation form.  Both render the oracle's film."""
    c, inp = small_config(3, (48, 48), 2, 4)
    o, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    sdf_index = [i for i, h in enumerate(c["world"].hitables.items) if hasattr(h, "sdf")][0]
    # (third case: RAYN_FLAG_NO_FOLD_ALL keeps the closest-hit fold in insertion order instead of marching the Mandelbox last)
    for flags, want in ((0, (4,)), (L.FLAG_NO_DIV3, (1,)), (L.FLAG_NO_FOLD_ALL, (4,))):
        r = Renderer(0, flags=flags)
        try:
            r.upload_scene(c["world"], c["camera"])
            assert r.sdf_variant(sdf_index) in want, (flags, r.sdf_variant(sdf_index))
            assert r.sdf_variant(0) == -1 and r.sdf_variant(99) == -2
            g = r.render_host(inp, (16, 16), c["integrator"], TR)
        finally:
            r.close()
        for ch in CH:
            assert_bit_equal(g[ch], o[ch], f"flags {flags} {ch}")


def test_fastdiv_equals_ieee_division(renderer):
    """The Newton division of the Mandelbox sphere fold (rt_sdf2.cuh::fastdiv2, no FCHK slow path) against IEEE `/`,
    EXHAUSTIVELY over every float the clamped divisor can take for the reference's constants (setup.rs:84:
    min 0.01^2, fixed 1.9^2), and over every divisor in [2^-10, 2^10) for a few other numerators."""
    mn, fx = np.float32(0.01) * np.float32(0.01), np.float32(1.9) * np.float32(1.9)
    lo, hi = int(mn.view(np.uint32)), int(fx.view(np.uint32))
    assert renderer.kat_fastdiv(float(fx), lo, hi - lo + 1) == 0
    lo, hi = int(np.float32(2.0 ** -10).view(np.uint32)), int(np.float32(2.0 ** 10).view(np.uint32))
    for num in (1.0, 3.0, 0.7, 1e3, 1.1754944e-3):
        assert renderer.kat_fastdiv(num, lo, hi - lo) == 0, num


@pytest.mark.parametrize("kind", ["mandelbox", "mandelbulb"])
@pytest.mark.parametrize("thr", [(0.000563, 0), (0.0002, 0), (0.0006, 0), (0.001, 1)])
def test_sphere_march_bit_equal(renderer, oracle, kind, thr):
    desc, keep, h = _sdf(kind)
    o, d = random_rays(20_000, seed=7)
    tmax = np.full(len(o), 200.0, np.float32)
    tmax[::7] = 4.0  # some rays stopped early by a closer earlier hitable
    o[5] = np.nan
    g = renderer.kat_sdf_hit(h, desc.consts, o, d, tmax, thr[0], thr[1])
    r = oracle.kat_sdf_hit(h, desc.consts, o, d, tmax, thr[0], thr[1])
    assert_bit_equal(g, r, f"{kind} sphere-march t")
    assert np.isfinite(r).sum() > 0.9 * len(r)


@pytest.mark.parametrize("kind", ["mandelbox", "mandelbulb"])
def test_occluded_bit_equal(renderer, oracle, kind):
    cam, world = configs.setup((64, 64), volume=False, fractal=kind)
    desc, keep = world.flatten(cam)
    renderer.upload_scene_desc(desc)
    rng = np.random.default_rng(3)
    s = rng.uniform(-2.0, 2.0, size=(20_000, 3)).astype(np.float32)
    e = rng.uniform(-1.5, 1.5, size=(20_000, 3)).astype(np.float32)
    g = renderer.kat_occluded(s, e)
    assert (g >= 0).all(), "early-out occlusion disagrees with the reference product form"
    r = oracle.kat_occluded(desc, s, e)
    assert_bit_equal(g, r, f"{kind} occluded")
    assert 0.02 < r.mean() < 0.98


@pytest.mark.parametrize("kind", ["mandelbox", "mandelbulb"])
@pytest.mark.parametrize("depth", [0, 1, 3])
def test_closest_hit_bit_equal(renderer, oracle, kind, depth):
    cam, world = configs.setup((64, 64), volume=False, fractal=kind)
    desc, keep = world.flatten(cam)
    renderer.upload_scene_desc(desc)
    o, d = random_rays(20_000, seed=11 + depth, spread=1.5)
    gt, gobj = renderer.kat_closest_hit(depth, o, d)
    rt, robj = oracle.kat_closest_hit(desc, depth, o, d)
    assert (gobj == robj).all()
    assert_bit_equal(gt, rt, "closest-hit t")
    assert len(np.unique(robj)) >= 3


# ---- T2: packet order ------------------------------------------------------------
@pytest.mark.parametrize("n,res,samples,mb", [(1, (20, 12), 2, 2), (3, (24, 24), 2, 3), (4, (16, 16), 1, 2)])
def test_packet_order_identical(oracle, n, res, samples, mb):
    c, inp = small_config(n, res, samples, mb)
    r = Renderer(0)
    try:
        r.upload_scene(c["world"], c["camera"])
        r.enable_queue_log(True)
        r.render_host(inp, (16, 16), c["integrator"], TR)
        glog = r.read_queue_log()
    finally:
        r.close()
    _, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, n_threads=1, queue_log=True)

    def parse(log):
        out, i = {}, 0
        while i < len(log):
            depth, tile, ns = log[i:i + 3]
            out[(int(depth), int(tile))] = log[i + 3:i + 3 + ns].copy()
            i += 3 + ns
        return out
    g, o = parse(glog), parse(info["queue_log"])
    g = {k: v for k, v in g.items() if len(v)}
    o = {k: v for k, v in o.items() if len(v)}
    assert set(g) == set(o)
    for k in o:
        assert np.array_equal(g[k], o[k]), f"shading queue differs at depth/tile {k}"
    assert any((v < 0).any() for v in o.values()), "test scene produced no padded packet"


# ---- T3: image parity -------------------------------------------------------------------------
CASES = [
    (1, (64, 64), 1, 2),      # config 1 geometry at 64x64
    (1, (100, 40), 2, 3),     # partial tiles: 100 % 16 = 4 -> reference drops the last column of tiles (F8)
    (3, (64, 64), 2, 4),      # Mandelbox + NEE, roulette active at depth 3
    (2, (64, 64), 2, 4),      # Mandelbulb
    (4, (48, 48), 1, 3),      # Mandelbulb + volume + thin lens
]


@pytest.mark.parametrize("n,res,samples,mb", CASES)
def test_image_bit_exact(renderer, oracle, n, res, samples, mb):
    c, inp = small_config(n, res, samples, mb)
    renderer.upload_scene(c["world"], c["camera"])
    g = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    st = renderer.stats()
    from rayn_b200.film import tile_grid
    ntx, nty = tile_grid(res[0], res[1], 16, 16)
    assert st.paths == min(ntx * 16, res[0]) * min(nty * 16, res[1]) * 4 * samples and st.launches > 0
    assert st.extend_rays == info["extend_rays"]
    assert st.shade_lanes == info["shade_lanes"]
    # oracle counts dist() per 4-lane packet, the GPU per lane
    assert info["sdf_evals_extend"] <= st.sdf_evals_extend <= 4 * info["sdf_evals_extend"]
    stats = rel_err_stats(g["color"] + g["background"], o["color"] + o["background"])
    print(f"cfg{n} {res} rel-err {stats}")
    for ch in CH:
        assert_bit_equal(g[ch], o[ch], f"cfg{n} {res} {ch}")
    assert stats["max"] <= 1e-4  # the north star's stated tolerance (implied by bit equality)
    assert float(o["color"].sum() + o["background"].sum()) > 0


def test_orthographic_camera_bit_exact(renderer, oracle):
    from rayn_b200 import OrthographicCamera, Vec3
    c, inp = small_config(3, (48, 32), 1, 2)
    world = c["world"]
    h = world.cameras.add_camera(OrthographicCamera((48, 32), 11.0 / 4.0, Vec3(9.5, -3.5, 9.5), Vec3(0.0, 0.8, 0.0), Vec3(0.0, 1.0, 0.0)))
    renderer.upload_scene(world, h)
    g = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    o, _ = oracle.render(world, h, inp, (16, 16), c["integrator"], TR)
    for ch in CH:
        assert_bit_equal(g[ch], o[ch], f"ortho {ch}")


# ---- size-independent properties at larger sizes ------------------------------------------------
def test_pass_size_and_sharding_do_not_change_the_film(oracle):
    """Tiles are independent (film.rs:439-627): film must be bit-identical whatever the pass size,
    and the union of tile-sharded renders must equal the unsharded film (multi-GPU contract T4)."""
    c, inp = small_config(3, (160, 96), 2, 3)
    full = None
    for cap in (0, 16 * 16 * 8 * 3):  # default pass vs 3 tiles per pass
        r = Renderer(0, max_paths_per_pass=cap)
        try:
            r.upload_scene(c["world"], c["camera"])
            f = r.render_host(inp, (16, 16), c["integrator"], TR)
            if cap:
                assert r.stats().passes > 1
        finally:
            r.close()
        if full is None:
            full = f
        else:
            for ch in CH:
                assert_bit_equal(f[ch], full[ch], f"pass-size {ch}")
    r = Renderer(0)
    try:
        r.upload_scene(c["world"], c["camera"])
        acc = {ch: np.zeros_like(full[ch]) for ch in CH}
        from rayn_b200.dist import shard_tiles
        from rayn_b200.film import tile_grid
        ntx, nty = tile_grid(160, 96, 16, 16)
        for rank in range(3):
            if rank == 0:  # offset/stride form of the ABI ...
                part = r.render_host(inp, (16, 16), c["integrator"], TR, tile_offset=rank, tile_stride=3)
                r.render_host(inp, (16, 16), c["integrator"], TR, tile_list=shard_tiles(ntx, nty, rank, 3, "index"))
            else:          # ... and the explicit tile-list form give the same shard
                part = r.render_host(inp, (16, 16), c["integrator"], TR, tile_list=shard_tiles(ntx, nty, rank, 3, "index"))
            for ch in CH:
                assert not (np.logical_and(acc[ch] != 0, part[ch] != 0)).any()
                acc[ch] += part[ch]
        for ch in CH:
            assert_bit_equal(acc[ch], full[ch], f"sharded {ch}")
    finally:
        r.close()
    # and a tile subset of the oracle agrees with the same tiles of the full GPU film
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, subsample_k=7)
    mask = o["alpha"] + o["background"].reshape(-1, 3).sum(1) + o["color"].reshape(-1, 3).sum(1) != 0
    assert mask.sum() > 0
    assert_bit_equal(full["color"].reshape(-1, 3)[mask], o["color"].reshape(-1, 3)[mask], "subset color")


def test_run_to_run_determinism_full_tile_count(renderer):
    c, inp = small_config(2, (256, 256), 4, 4)
    renderer.upload_scene(c["world"], c["camera"])
    a = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    b = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    for ch in CH:
        assert_bit_equal(a[ch], b[ch], f"determinism {ch}")
    assert np.isfinite(a["color"]).all()
    assert abs(float(a["alpha"].max()) - 1.0) < 1e-6 or a["alpha"].max() <= 1.0


# ---- error behaviour of the boundary ----------------------------------------------------------
def test_error_codes():
    r = Renderer(0)
    try:
        c, inp = small_config(1, (32, 32), 1, 1)
        with pytest.raises(L.RaynError) as e:
            r.render_host(inp, (16, 16), c["integrator"], TR)
        assert e.value.code == L.RAYN_ERR_NO_SCENE
        r.upload_scene(c["world"], c["camera"])
        integ = configs.PathTracingIntegrator(1, 3)
        with pytest.raises(L.RaynError) as e:
            r.render_host(inp, (16, 16), integ, TR)
        assert e.value.code == L.RAYN_ERR_UNSUPPORTED
        deep = configs.PathTracingIntegrator(5, 2)  # needs more sample sets than the tables hold
        with pytest.raises(L.RaynError) as e:
            r.render_host(inp, (16, 16), deep, TR)
        assert e.value.code == L.RAYN_ERR_INVALID_ARG
        bad = c["world"].flatten(c["camera"])[0]
        bad.n_hitables = 0
        with pytest.raises(L.RaynError) as e:
            r.upload_scene_desc(bad)
        assert e.value.code == L.RAYN_ERR_INVALID_ARG
    finally:
        r.close()


# ---- GPU vs the committed golden fixtures (tests/golden/*.npz, generated by make_golden.py) ----
def test_gpu_matches_committed_golden(renderer):
    import os
    from test_cpu_oracle import GOLD, GOLD_SUFFIX, GOLDEN_CASES
    for name, (n, res, samples, mb) in sorted(GOLDEN_CASES.items()):
        c, inp = small_config(n, res, samples, mb)
        renderer.upload_scene(c["world"], c["camera"])
        g = renderer.render_host(inp, (16, 16), c["integrator"], TR)
        gold = np.load(os.path.join(GOLD, name + GOLD_SUFFIX + ".npz"))
        for ch in CH:
            assert_bit_equal(g[ch], gold[ch], f"golden {name} {ch}")


def _run_suite_variant(env_extra, args):
    import os
    import subprocess
    import sys
    env = dict(os.environ, **env_extra)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + args, cwd=root, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1500:]
    return r.stdout


@pytest.mark.skipif(L.MULADD_FUSED or L.LEGACY, reason="already inside a variant run")
def test_fused_mul_add_variant_passes_the_parity_suite():
    """oracle/README.md A6: `wide` f32x4::mul_add is unfused in a stock `cargo run --release` build of rayn (the default
    here) and fused with `-C target-feature=+fma`.  Both variants of the kernels (librayn_b200_fma.so) and of the oracle
    (librayn_oracle_fma.so) exist; this runs the stage, packet-order, image and golden tests again in the fused variant."""
    out = _run_suite_variant({"RAYN_MULADD_FUSED": "1"}, ["tests/test_gpu_parity.py", "-k", "not full_size and not variant and not legacy and not cpp_host"])
    assert " passed" in out


@pytest.mark.skipif(L.MULADD_FUSED or L.LEGACY, reason="already inside a variant run")
def test_legacy_one_thread_per_ray_kernels_agree():
    """TEST build librayn_b200_legacy.so (-DRAYN_LEGACY_KERNELS): the round-1 v0 kernels (one thread per ray, shadow marches
    fused into shading, scalar arithmetic, no queues) are a structurally independent second implementation; with
    RAYN_FLAG_SIMPLE_MARCH they must give the oracle's bits too."""
    out = _run_suite_variant({"RAYN_B200_LEGACY": "1"}, ["tests/test_gpu_parity.py", "-k", "legacy_family_inner or image_bit_exact"])
    assert " passed" in out


@pytest.mark.skipif(not L.LEGACY, reason="needs RAYN_B200_LEGACY=1 (run by test_legacy_one_thread_per_ray_kernels_agree)")
def test_legacy_family_inner(oracle):
    for n, res, samples, mb in [(3, (48, 48), 2, 4), (4, (32, 32), 1, 2), (2, (48, 32), 2, 3)]:
        c, inp = small_config(n, res, samples, mb)
        o, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
        r = Renderer(0, flags=L.FLAG_SIMPLE_MARCH)
        try:
            r.upload_scene(c["world"], c["camera"])
            g = r.render_host(inp, (16, 16), c["integrator"], TR)
        finally:
            r.close()
        for ch in CH:
            assert_bit_equal(g[ch], o[ch], f"legacy kernels cfg{n} {ch}")


def test_simple_march_flag_is_rejected_by_the_product_library():
    if L.LEGACY:
        pytest.skip("legacy build")
    with pytest.raises(L.RaynError) as e:
        Renderer(0, flags=L.FLAG_SIMPLE_MARCH)
    assert e.value.code == L.RAYN_ERR_UNSUPPORTED


def test_cpp_host_renders_the_same_film_as_the_python_host(renderer, tmp_path):
    """End to end through the C++ stand-in for main.rs (rayn_b200/host/main.cpp)."""
    import os
    import subprocess
    from rayn_b200 import build
    exe = os.path.join(os.path.dirname(build.OUT), "rayn_host")
    out = tmp_path / "planes.bin"
    r = subprocess.run([exe, "--config", "3", "--res", "48", "32", "--samples", "2", "--bounces", "3", "--dump", str(out)],
                       capture_output=True, text=True, check=True)
    assert "Done in" in r.stdout
    c, inp = small_config(3, (48, 32), 2, 3)
    renderer.upload_scene(c["world"], c["camera"])
    g = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    raw = np.fromfile(out, np.float32)
    npx = 48 * 32
    got = {"color": raw[:3 * npx], "alpha": raw[3 * npx:4 * npx], "background": raw[4 * npx:7 * npx], "normal": raw[7 * npx:]}
    for ch in CH:
        assert_bit_equal(got[ch], g[ch], f"C++ host {ch}")


# ---- BASELINE full sizes: sampled bit-exact parity + properties ---------------------------------
@pytest.mark.parametrize("n,k", [(2, 409), (3, 1361)])
def test_full_size_config_tile_sample_matches_oracle(oracle, n, k):
    """The GPU renders BASELINE config n at its FULL size (cfg2: 1024x1024x128spp, cfg3: 1920x1080x512spp);
    the oracle renders every k-th tile of the same frame (the CPU cannot do the whole frame in test time).
    Tiles are independent (film.rs:439-627), so those tiles must agree bit for bit."""
    c = configs.baseline_config(n)
    w, h = c["res"]
    inp = FrameInputs(w, h, c["samples"], c["integrator"])
    r = Renderer(0)
    try:
        r.upload_scene(c["world"], c["camera"])
        g = r.render_host(inp, (16, 16), c["integrator"], TR)
        st = r.stats()
    finally:
        r.close()
    assert st.paths == w * h * c["spp"]
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, subsample_k=k)
    assert 4 <= info["tiles"] <= 12
    from rayn_b200.film import tile_grid
    ntx, nty = tile_grid(w, h, 16, 16)
    mask = np.zeros((h, w), bool)
    for idx in range(0, ntx * nty, k):
        x0, y0 = (idx // nty) * 16, (idx % nty) * 16
        mask[y0:y0 + 16, x0:x0 + 16] = True
    m = mask.reshape(-1)
    assert m.sum() >= 4 * 128
    for ch in CH:
        a = g[ch].reshape(w * h, -1)[m]
        b = o[ch].reshape(w * h, -1)[m]
        assert_bit_equal(a, b, f"full-size cfg{n} {ch}")
    # properties of the whole frame
    assert np.isfinite(g["color"]).all() and (g["color"] >= 0).all() and (g["background"] >= 0).all()
    assert g["alpha"].min() >= 0 and g["alpha"].max() <= 1.0
    covered = (g["alpha"] > 0) | (g["background"].reshape(-1, 3).sum(1) > 0)
    assert covered.mean() > 0.999  # every camera ray ends on the sky sphere or the fractal


SAMPLED_TILES = ((0.5, 0.5), (0.4, 0.55), (0.62, 0.45), (0.05, 0.9), (0.33, 0.37), (0.7, 0.62))


@pytest.mark.parametrize("n", [4, 5])
def test_full_size_multi_gpu_configs_sampled_tiles_match_oracle(oracle, n):
    """BASELINE configs 4 (2048x2048, 256 spp, volume + thin lens) and 5 (7680x4320, 1024 spp, 8 bounces) at their FULL
    resolution and spp: GPU and oracle both render the same six 16x16 tiles (picked across the fractal and the sky) of the
    full-size frame through `tile_list`; tiles are independent (film.rs:439-627), so they must agree bit for bit.  Full
    frames of these sizes are what the multi-GPU bench renders; here one GPU and a few CPU seconds suffice."""
    c = configs.baseline_config(n)
    w, h = c["res"]
    inp = FrameInputs(w, h, c["samples"], c["integrator"])
    from rayn_b200.film import tile_grid
    ntx, nty = tile_grid(w, h, 16, 16)
    tiles = sorted({int(fx * ntx) * nty + int(fy * nty) for fx, fy in SAMPLED_TILES})
    r = Renderer(0)
    try:
        r.upload_scene(c["world"], c["camera"])
        g = r.render_host(inp, (16, 16), c["integrator"], TR, tile_list=tiles)
        st = r.stats()
    finally:
        r.close()
    assert st.paths == len(tiles) * 256 * c["spp"]
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, tile_list=tiles)
    assert info["tiles"] == len(tiles)
    assert st.extend_rays == info["extend_rays"] and st.shade_lanes == info["shade_lanes"]
    for ch in CH:
        assert_bit_equal(g[ch], o[ch], f"full-size cfg{n} sampled tiles {ch}")
    lit = g["color"].reshape(-1, 3).sum(1) + g["background"].reshape(-1, 3).sum(1)
    assert (lit > 0).sum() >= 0.9 * len(tiles) * 256 and g["alpha"].max() > 0  # the sample covers fractal and sky


def test_null_planes_are_skipped_like_absent_film_channels(renderer):
    """Film<N> may hold any subset of channels (film.rs:175-203); add_sample ignores absent ones (:167-172)."""
    c, inp = small_config(3, (48, 32), 1, 2)
    renderer.upload_scene(c["world"], c["camera"])
    full = renderer.render_host(inp, (16, 16), c["integrator"], TR)
    w, h = 48, 32
    color = np.zeros(3 * w * h, np.float32)
    alpha = np.zeros(w * h, np.float32)
    p = L.RaynFilmPlanes(color.ctypes.data, alpha.ctypes.data, None, None, L.MEM_HOST)
    from rayn_b200.film import make_frame_desc
    f = make_frame_desc(w, h, (16, 16), inp.samples, c["integrator"], 1, TR, tuple(a.ctypes.data for a in inp.arrays()), L.MEM_HOST, 0, 1,
                        (inp.sets_1d, inp.sets_2d))
    renderer.render(f, p)
    assert_bit_equal(color, full["color"], "color only")
    assert_bit_equal(alpha, full["alpha"], "alpha only")
    with pytest.raises(L.RaynError) as e:
        renderer.render(f, L.RaynFilmPlanes(None, None, None, None, L.MEM_HOST))
    assert e.value.code == L.RAYN_ERR_INVALID_ARG


def test_statistics_are_consistent(renderer, oracle):
    """evals / iterations / lanes reported by RaynStats feed bench.py's flop figures: cross-check them."""
    c, inp = small_config(2, (64, 64), 2, 3)
    renderer.upload_scene(c["world"], c["camera"])
    renderer.render_host(inp, (16, 16), c["integrator"], TR)
    st = renderer.stats()
    assert st.sdf_evals_extend > 0 and st.sdf_evals_shadow > 0 and st.sdf_evals_normals > 0
    assert st.sdf_evals_normals % 4 == 0 and st.sdf_evals_normals <= 4 * st.shade_lanes
    # the authored Mandelbulb leaves its loop between 0 and 8 iterations: the counted iterations must be below the cap
    assert 0 < st.bulb_iters_extend < 8 * st.sdf_evals_extend
    assert 0 < st.bulb_iters_shadow < 8 * st.sdf_evals_shadow


# ---- K4 stage tests: light sampling and BSDFs, host vs device -------------------------
def test_light_and_bsdf_stages_bit_equal(renderer, oracle):
    from rayn_b200.scene import Dielectric, Lambertian, SphereLight, Srgb, Vec3
    from test_cpu_closed_form import _hemisphere_inputs
    rng = np.random.default_rng(21)
    n = 50_000
    light = SphereLight(Vec3(1.2, -1.2, 1.2), 0.15, Srgb(1.5, 4.5, 3.0)).flatten()
    p = rng.uniform(-2.5, 2.5, size=(n, 3)).astype(np.float32)
    s0, s1 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    s0[:3] = [0.0, 1.0, 0.5]
    p[3] = light.pos[:]  # degenerate: shading point at the light centre
    gpt, gpdf = renderer.kat_light_sample(light, s0, s1, p)
    opt, opdf = oracle.kat_light_sample(light, s0, s1, p)
    assert_bit_equal(gpt, opt, "light sample point")
    assert_bit_equal(gpdf, opdf, "light sample pdf")
    o = rng.uniform(-3, 3, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    tm = rng.uniform(0.1, 50, n).astype(np.float32)
    gt, gp = renderer.kat_light_sample_volume(light, s0, o, d, tm)
    ot, op = oracle.kat_light_sample_volume(light, s0, o, d, tm)
    assert_bit_equal(gt, ot, "equi-angular t")
    assert_bit_equal(gp, op, "equi-angular pdf")
    nrm, wo, s1d, u4 = _hemisphere_inputs(n, 5)
    u4[0] = [0.5, 0.5, 0.0, 0.0]  # concentric map (0,0) guard, math.rs:206-207
    for mat in (Lambertian(Srgb(0.6, 0.4, 0.2)).flatten(), Dielectric.new_remap(Srgb(0.2, 0.2, 0.2), 0.6).flatten(),
                Dielectric(Srgb(0.9, 0.9, 0.9), 300.0).flatten()):
        g = renderer.kat_bsdf(mat, nrm, wo, s1d, u4)
        r = oracle.kat_bsdf(mat, nrm, wo, s1d, u4)
        for a, b, what in zip(g, r, ("wi", "f", "pdf", "f(wo,wi,n)")):
            assert_bit_equal(a, b, f"bsdf kind {mat.kind} {what}")
