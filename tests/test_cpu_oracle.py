"""The oracle against mathematics and against its own committed golden renders.

PARITY UNPINNED: the reference ships no tests, vectors or fixtures for this path,
so closed-form checks are the only external truth; the golden files pin the oracle against
regressions and let the GPU tests compare against committed numbers."""
import os

import numpy as np
import pytest

from rayn_b200 import _lib as L
from rayn_b200 import configs

from helpers import CH, assert_bit_equal, random_rays, small_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD_SUFFIX = "_fma" if L.MULADD_FUSED else ""  # fixtures exist for both `wide` mul_add variants (oracle/README.md A6)
TR = configs.frame_time_range(1)


def _scene(kind="mandelbox"):
    cam, world = configs.setup((64, 64), volume=False, fractal=kind)
    desc, keep = world.flatten(cam)
    return desc, keep


def mandelbox_ref(p, iters=12, l=1.0, minr2=np.float32(0.01) ** 2, fixr2=np.float32(1.9) ** 2, scale=-2.1):
    """Independent float64 statement of the Mandelbox DE (sdf.rs:125-188)."""
    p = p.astype(np.float64)
    off = p.copy()
    dr = np.ones(len(p))
    for _ in range(iters):
        p = np.clip(p, -l, l) * 2.0 - p
        r2 = (p * p).sum(1)
        m = np.maximum(1.0, float(fixr2) / np.maximum(float(minr2), r2))
        p = p * m[:, None]
        dr = dr * m
        p = p * float(np.float32(scale)) + off
        dr = -dr * float(np.float32(scale)) + 1.0
    return np.sqrt((p * p).sum(1)) / np.abs(dr)


def mandelbulb_ref(p, iters=8, bailout=2.0):
    """Textbook trigonometric power-8 Mandelbulb DE in float64 (z axis as pole)."""
    c = p.astype(np.float64)
    w = c.copy()
    dr = np.ones(len(p))
    alive = np.ones(len(p), bool)
    for _ in range(iters):
        r = np.sqrt((w * w).sum(1))
        alive &= ~(r * r > bailout * bailout)
        th = np.arccos(np.clip(w[:, 2] / np.maximum(r, 1e-300), -1, 1))
        ph = np.arctan2(w[:, 1], w[:, 0])
        ndr = 8.0 * r ** 7 * dr + 1.0
        r8 = r ** 8
        nw = np.stack([r8 * np.sin(8 * th) * np.cos(8 * ph), r8 * np.sin(8 * th) * np.sin(8 * ph), r8 * np.cos(8 * th)], 1) + c
        w = np.where(alive[:, None], nw, w)
        dr = np.where(alive, ndr, dr)
    r = np.sqrt((w * w).sum(1))
    return 0.5 * np.log(r) * r / dr


def test_mandelbox_dist_matches_float64_statement(oracle):
    desc, keep = _scene("mandelbox")
    p = np.random.default_rng(0).uniform(-3, 3, size=(20000, 3)).astype(np.float32)
    got = oracle.kat_sdf_dist(desc.hitables[1], p).astype(np.float64)
    ref = mandelbox_ref(p)
    assert np.median(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9)) < 1e-6
    assert np.quantile(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9), 0.99) < 1e-3  # chaotic folds amplify f32 rounding


def test_mandelbulb_algebraic_form_equals_trigonometric_form(oracle):
    """The authored trig-free power-8 step (Chebyshev polynomials) is the White/Nylander triplex power."""
    desc, keep = _scene("mandelbulb")
    p = np.random.default_rng(1).uniform(-1.3, 1.3, size=(20000, 3)).astype(np.float32)
    got = oracle.kat_sdf_dist(desc.hitables[1], p).astype(np.float64)
    ref = mandelbulb_ref(p)
    ok = np.isfinite(ref) & (np.abs(ref) > 1e-4)
    rel = np.abs(got[ok] - ref[ok]) / np.abs(ref[ok])
    assert ok.sum() > 10000
    assert np.median(rel) < 1e-5
    assert np.quantile(rel, 0.9) < 1e-2  # iterated 8th powers amplify f32 rounding near the set


def test_sphere_hit_closed_form(oracle):
    c = configs.baseline_config(1, res=(32, 32))
    desc, keep = c["world"].flatten(c["camera"])
    o, d = random_rays(5000, 5, origin_radius=4.0, spread=0.8)
    t, obj = oracle.kat_closest_hit(desc, 0, o, d)
    od = o.astype(np.float64), d.astype(np.float64)
    b = (od[0] * od[1]).sum(1)
    disc1 = b * b - ((od[0] ** 2).sum(1) - 1.0)
    hit1 = disc1 > 0
    t1 = -b - np.sqrt(np.maximum(disc1, 0))
    assert (obj[hit1 & (t1 > 1e-3)] == 1).all()
    assert np.allclose(t[hit1 & (t1 > 1e-3)], t1[hit1 & (t1 > 1e-3)], rtol=1e-4)
    sky = ~hit1
    t_sky = -b + np.sqrt(b * b - ((od[0] ** 2).sum(1) - 100.0 ** 2))
    assert (obj[sky] == 0).all() and np.allclose(t[sky], t_sky[sky], rtol=1e-4)


def test_sphere_march_lands_on_the_surface(oracle):
    desc, keep = _scene("mandelbox")
    o, d = random_rays(4000, 9)
    h = desc.hitables[1]
    t = oracle.kat_sdf_hit(h, desc.consts, o, d, np.full(len(o), 200.0, np.float32), 0.000563, 0)
    hit = np.isfinite(t) & (t < 200.0)
    assert hit.mean() > 0.5
    dist = oracle.kat_sdf_dist(h, o[hit] + d[hit] * t[hit, None])
    assert np.quantile(np.abs(dist), 0.95) < 1e-2  # stopped within the hit threshold band of the surface


def test_occlusion_semantics(oracle):
    desc, keep = _scene("mandelbox")
    s = np.array([[10, 10, 10], [5, 0, 0], [0.0, 0.0, 5.0]], np.float32)
    e = np.array([[12, 12, 10], [-5, 0, 0], [0.0, 0.0, -5.0]], np.float32)
    v = oracle.kat_occluded(desc, s, e)
    assert v[0] == 1.0          # free space
    assert v[1] == 0.0 and v[2] == 0.0  # straight through the fractal / the central emitter


def test_reference_tile_count_quirk(oracle):
    """film.rs:399-404: (res + res % tile) / tile drops the last partial tile when 0 < res % 16 < 8."""
    c, inp = small_config(1, (100, 40), 1, 1)
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    img = (o["color"] + o["background"]).reshape(40, 100, 3)
    assert info["tiles"] == 6 * 3
    assert img[:, 96:, :].sum() == 0 and img[:, :96, :].min() >= 0 and img[:40, :96].sum() > 0
    assert (img[:, :96].sum(axis=2) > 0).all()


def test_threads_and_tile_subsets_do_not_change_pixels(oracle):
    c, inp = small_config(3, (48, 48), 1, 2)
    a, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, n_threads=1)
    b, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, n_threads=4)
    for ch in CH:
        assert_bit_equal(a[ch], b[ch], ch)
    s, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR, subsample_k=4)
    assert info["tiles"] == 3
    m = s["alpha"] + s["color"].reshape(-1, 3).sum(1) + s["background"].reshape(-1, 3).sum(1) != 0
    assert_bit_equal(s["color"].reshape(-1, 3)[m], a["color"].reshape(-1, 3)[m], "subset")


def test_film_value_ranges(oracle):
    c, inp = small_config(3, (48, 48), 2, 3)
    o, info = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    assert np.isfinite(o["color"]).all() and (o["color"] >= 0).all() and (o["background"] >= 0).all()
    assert 0 < o["alpha"].max() <= 1.0 and o["alpha"].min() >= 0
    n = o["normal"].reshape(-1, 3)
    assert (np.linalg.norm(n, axis=1) <= 1.0 + 1e-5).all()
    assert info["extend_rays"] >= 48 * 48 * 8 and info["shadow_rays"] > 0


GOLDEN_CASES = {"cfg1_64x64_4spp_2b": (1, (64, 64), 1, 2), "cfg3_32x32_8spp_3b": (3, (32, 32), 2, 3),
                "cfg2_32x32_8spp_3b": (2, (32, 32), 2, 3), "cfg4_32x32_4spp_2b": (4, (32, 32), 1, 2)}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_reproduces_committed_golden(oracle, name):
    n, res, samples, mb = GOLDEN_CASES[name]
    c, inp = small_config(n, res, samples, mb)
    o, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    g = np.load(os.path.join(GOLD, name + GOLD_SUFFIX + ".npz"))
    for ch in CH:
        assert_bit_equal(o[ch], g[ch], f"{name} {ch}")


def test_film_postprocess_formulas(oracle):
    """save_to pixel arithmetic (film.rs:205-377) against an independent numpy statement (1 LSB slack for pow)."""
    rng = np.random.default_rng(0)
    w, h = 13, 9
    pl = {"color": rng.uniform(-0.2, 1.5, 3 * w * h).astype(np.float32), "alpha": rng.uniform(-0.1, 1.1, w * h).astype(np.float32),
          "background": rng.uniform(0, 0.5, 3 * w * h).astype(np.float32), "normal": rng.uniform(-1, 1, 3 * w * h).astype(np.float32)}
    col, bg = pl["color"].reshape(h, w, 3).astype(np.float64), pl["background"].reshape(h, w, 3).astype(np.float64)
    g = lambda x: np.clip(x, 0, 1) ** (1 / 2.2) * 255
    want = {L.POST_COLOR_PLUS_BACKGROUND: g(col + bg), L.POST_COLOR_ALPHA: g(col), L.POST_BACKGROUND: g(bg),
            L.POST_WORLD_NORMAL: np.clip((pl["normal"].reshape(h, w, 3) * 0.5 + 0.5) * 255, 0, 255),
            L.POST_ALPHA: np.clip(pl["alpha"].reshape(h, w, 1) * 255, 0, 255)}
    for mode, ref in want.items():
        got = oracle.film_postprocess(mode, w, h, pl).astype(int)
        assert got.shape[:2] == (h, w)
        assert np.abs(got[..., :ref.shape[2]] - np.floor(ref[::-1]).astype(int)).max() <= 1, mode  # y flipped, film.rs:236
    rgba = oracle.film_postprocess(L.POST_COLOR_ALPHA, w, h, pl)
    assert np.array_equal(rgba[..., 3:], oracle.film_postprocess(L.POST_ALPHA, w, h, pl))
    # COLOR_ONLY does not saturate (film.rs:275-293): values > 1 clip at 255, negatives become NaN.powf -> 255 after .min(255)
    only = oracle.film_postprocess(L.POST_COLOR_ONLY, w, h, pl)
    neg = (pl["color"].reshape(h, w, 3)[::-1] < 0)
    assert (only[neg] == 255).all()


def test_sphere_first_fold_equals_the_insertion_order_fold(oracle):
    """The argument behind `fold_all` (rt_kernels.cuh, k_extend_march): for a [spheres] Mandelbox [spheres] scene, testing ALL
    spheres first and marching the SDF last against the nearest sphere - accepting its t when it is smaller, or equal with the
    sphere later in insertion order - gives the reference's insertion-order fold (hitable.rs:177-198) bit for bit.  Checked here
    with the ORACLE's own closest-hit fold, its sphere-only fold and its sphere-march, on rays chosen so that every case of the
    proof occurs: the SDF wins, an earlier sphere wins, a LATER sphere cuts the march short, nothing is hit."""
    from rayn_b200.scene import HitableStore, World
    cam, world = configs.setup((64, 64), volume=False, fractal="mandelbox")
    desc, keep = world.flatten(cam)
    items = world.hitables.items
    hk = [i for i, h in enumerate(items) if hasattr(h, "sdf")][0]
    assert 0 < hk < len(items) - 1, "the setup.rs scene has spheres on both sides of the Mandelbox"
    spheres = HitableStore()
    sphere_index = []  # index in the full scene of every sphere of the sphere-only scene
    for i, h in enumerate(items):
        if i != hk:
            spheres.push(h)
            sphere_index.append(i)
    sphere_index = np.asarray(sphere_index)
    world_s = World(spheres, world.lights, world.materials, world.cameras, world.volume_params, world.consts)
    desc_s, keep_s = world_s.flatten(cam)
    rng = np.random.default_rng(11)
    o1, d1 = random_rays(6000, 3)                                   # from outside towards the fractal
    o2 = rng.uniform(-1.6, 1.6, size=(6000, 3)).astype(np.float32)  # from inside the scene towards the emitters
    emit = np.array([[1.2, 1.2, 1.2], [1.2, -1.2, 1.2], [-1.2, -1.2, 1.2], [-1.2, 1.2, 1.2], [0.0, 0.0, 0.0]])
    tgt = emit[rng.integers(0, 5, 6000)] + rng.normal(scale=0.12, size=(6000, 3))
    d2 = tgt - o2
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    for depth, thr_scale in ((1, 0.0001 * 2.0 * 1.0), (3, 0.0001 * 2.0 * 3.0)):  # film.rs:549 hit_threshold_at for depth > 0
        t_ref, id_ref = oracle.kat_closest_hit(desc, depth, o, d)
        c, ids = oracle.kat_closest_hit(desc_s, depth, o, d)         # fold over the spheres alone (same strict-min rule)
        owner = np.where(ids >= 0, sphere_index[np.maximum(ids, 0)], -1)
        t_sdf = oracle.kat_sdf_hit(desc.hitables[hk], desc.consts, o, d, c, np.float32(thr_scale), 0)
        with np.errstate(invalid="ignore"):
            take = (t_sdf < c) | ((t_sdf == c) & (owner > hk))
        t_new = np.where(take, t_sdf, c).astype(np.float32)
        id_new = np.where(take, hk, owner)
        assert_bit_equal(t_new, t_ref, f"sphere-first fold t, depth {depth}")
        assert (id_new == id_ref).all()
        later = (id_ref > hk).sum()
        assert (id_ref == hk).sum() > 500 and later > 200 and (id_ref == 0).sum() > 200, "every case of the proof must occur"
        # the interesting case: a later sphere ends the march before the SDF's own stop
        t_full = oracle.kat_sdf_hit(desc.hitables[hk], desc.consts, o, d, np.full(len(o), 200.0, np.float32), np.float32(thr_scale), 0)
        assert ((t_full != t_sdf) & (id_ref > hk)).sum() > 50
