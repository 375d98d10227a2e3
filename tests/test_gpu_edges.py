"""Edge cases of the path (the reference has no tests of its own; these are the ragged / empty /
extreme shapes its code can be driven into), each bit-exact vs the oracle."""
import numpy as np
import pytest

from rayn_b200 import _lib as L
from rayn_b200 import (BoxFold, CameraStore, Dielectric, Emissive, HitableStore, Lambertian, MandelBox, Mandelbulb, MaterialStore,
                       PathTracingIntegrator, PinholeCamera, Sky, Sphere, SphereFold, SphereLight, Srgb, TracedSDF, Vec3, VolumeParams,
                       World, configs)
from rayn_b200.film import FrameInputs, Renderer

from helpers import CH, assert_bit_equal, small_config

pytestmark = pytest.mark.gpu
TR = configs.frame_time_range(1)


def _both(renderer, oracle, world, cam, res, samples, mb, tile=(16, 16)):
    integ = PathTracingIntegrator(mb, 2)
    inp = FrameInputs(res[0], res[1], samples, integ)
    renderer.upload_scene(world, cam)
    g = renderer.render_host(inp, tile, integ, TR)
    o, info = oracle.render(world, cam, inp, tile, integ, TR)
    for ch in CH:
        assert_bit_equal(g[ch], o[ch], ch)
    return g, info


def _world(hitables, lights, materials, res, volume=VolumeParams(None, None)):
    cams = CameraStore()
    cam = cams.add_camera(PinholeCamera(res, 60.0, Vec3(-0.45, 0.2, 2.0) * 2.25, Vec3(0.0, 0.0, 0.0), Vec3(0.0, 1.0, 0.0)))
    return cam, World(hitables, lights, materials, cams, volume)


def test_image_smaller_than_a_tile_and_minimal_spp(renderer, oracle):
    c = configs.baseline_config(3, res=(5, 3), samples=1, max_bounces=2)
    g, info = _both(renderer, oracle, c["world"], c["camera"], (5, 3), 1, 2)
    assert info["tiles"] == 0 or g["alpha"].size == 15  # 5 % 16 = 5 < 8 -> the reference renders ZERO tiles (film.rs:399-404)
    c = configs.baseline_config(3, res=(9, 12), samples=1, max_bounces=2)  # 9 % 16 = 9 >= 8 -> one clipped tile
    g, info = _both(renderer, oracle, c["world"], c["camera"], (9, 12), 1, 2)
    assert info["tiles"] == 1 and float(g["color"].sum() + g["background"].sum()) > 0


@pytest.mark.parametrize("tile", [(8, 8), (32, 16), (16, 4)])
def test_other_tile_sizes(renderer, oracle, tile):
    c = configs.baseline_config(3, res=(64, 48), samples=1, max_bounces=2)
    _both(renderer, oracle, c["world"], c["camera"], (64, 48), 1, 2, tile)


def test_zero_bounces_and_no_lights(renderer, oracle):
    c = configs.baseline_config(3, res=(32, 32), samples=1, max_bounces=0)  # every path ends at depth 0 (integrator.rs:178)
    _both(renderer, oracle, c["world"], c["camera"], (32, 32), 1, 0)
    materials, hitables = MaterialStore(), HitableStore()
    sky = materials.add_material(Sky(Srgb(0.3, 0.4, 0.6), Srgb(0.2, 0.3, 0.6) * 0.05))
    grey = materials.add_material(Lambertian(Srgb(0.5, 0.4, 0.3)))
    hitables.push(Sphere(Vec3(0, 0, 0), 100.0, sky))
    hitables.push(Sphere(Vec3(0, 0, 0), 1.0, grey))
    cam, world = _world(hitables, [], materials, (32, 32))  # world.lights.len() == 0: NEE skipped (integrator.rs:73)
    g, _ = _both(renderer, oracle, world, cam, (32, 32), 2, 3)
    assert float(g["color"].sum()) > 0  # sky light arrives through Lambertian bounces


def test_sky_only_scene_all_paths_end_at_depth_zero(renderer, oracle):
    materials, hitables = MaterialStore(), HitableStore()
    sky = materials.add_material(Sky(Srgb(0.3, 0.4, 0.6), Srgb(0.2, 0.3, 0.6) * 0.05))
    hitables.push(Sphere(Vec3(0, 0, 0), 100.0, sky))
    cam, world = _world(hitables, [SphereLight(Vec3(1, 1, 1), 0.1, Srgb(1, 1, 1))], materials, (48, 32))
    g, _ = _both(renderer, oracle, world, cam, (48, 32), 2, 3)
    assert g["color"].sum() == 0 and g["alpha"].sum() == 0 and (g["background"] > 0).all()


def test_rays_that_hit_nothing_are_dropped_but_counted(renderer, oracle):
    """No sky sphere: camera rays that miss everything vanish, yet the divisor stays spp."""
    materials, hitables = MaterialStore(), HitableStore()
    em = materials.add_material(Emissive.new_splat(Srgb(2.0, 1.0, 0.5)))
    hitables.push(Sphere(Vec3(0, 0, 0), 1.0, em))
    cam, world = _world(hitables, [], materials, (32, 32))
    g, _ = _both(renderer, oracle, world, cam, (32, 32), 2, 2)
    img = g["background"].reshape(32, 32, 3)
    assert img[0, 0].sum() == 0 and img[16, 16, 0] > 0
    edge = img[:, :, 0][(img[:, :, 0] > 0) & (img[:, :, 0] < 2.0)]
    assert edge.size > 0  # partially covered pixels are darker: lost samples still divide


def test_two_sdfs_and_volume(renderer, oracle):
    """Two TracedSDF hitables (Mandelbox then Mandelbulb) around analytic spheres: exercises the per-SDF march
    kernels in fold order (hitable.rs:177-198) and shadow segments per SDF hitable."""
    materials, hitables, lights = MaterialStore(), HitableStore(), []
    sky = materials.add_material(Sky(Srgb(0.3, 0.4, 0.6), Srgb(0.2, 0.3, 0.6) * 0.05))
    grey = materials.add_material(Dielectric.new_remap(Srgb(0.2, 0.2, 0.2), 0.6))
    red = materials.add_material(Lambertian(Srgb(0.6, 0.2, 0.2)))
    hitables.push(TracedSDF(Mandelbulb(6, 8, 2.0), red))
    hitables.push(Sphere(Vec3(0, 0, 0), 100.0, sky))
    hitables.push(TracedSDF(MandelBox(8, BoxFold(1.0), SphereFold(0.5, 1.0), -2.0), grey))
    hitables.push(Sphere(Vec3(0.0, 1.6, 0.0), 0.3, red))
    lights.append(SphereLight(Vec3(2.5, 2.5, 2.5), 0.2, Srgb(1, 1, 1) * 60.0))
    lights.append(SphereLight(Vec3(-2.5, 1.0, 2.5), 0.2, Srgb(0.5, 0.7, 1.0) * 60.0))
    cam, world = _world(hitables, lights, materials, (40, 40), VolumeParams(0.25, 0.035))
    g, info = _both(renderer, oracle, world, cam, (40, 40), 1, 3)
    assert info["sdf_evals_extend"] > 0 and float(g["color"].sum()) > 0


def test_scattering_without_extinction_and_vice_versa(renderer, oracle):
    c = configs.baseline_config(4, res=(32, 32), samples=1, max_bounces=2)
    for vol in (VolumeParams(0.25, None), VolumeParams(None, 0.035)):  # Option<f32> pairs are independent (volume.rs:2-5)
        c["world"].volume_params = vol
        _both(renderer, oracle, c["world"], c["camera"], (32, 32), 1, 2)


def test_limits_are_reported_not_crashed():
    r = Renderer(0)
    try:
        c, inp = small_config(1, (16, 16), 1, 1)
        r.upload_scene(c["world"], c["camera"])
        big = PathTracingIntegrator(1, 2)
        huge = FrameInputs(16, 16, 4200, big)  # 16*16*16800 spp > 2^22 slot keys (and > the 16384 spp the film resolve stages)
        with pytest.raises(L.RaynError) as e:
            r.render_host(huge, (16, 16), big, TR)
        assert e.value.code == L.RAYN_ERR_UNSUPPORTED
        with pytest.raises(L.RaynError) as e:
            r.render_host(inp, (16, 16), big, TR, tile_list=[3, 1])  # not ascending
        assert e.value.code == L.RAYN_ERR_INVALID_ARG
        d, _ = c["world"].flatten(c["camera"])
        d.hitables[1].kind = L.HITABLE_MANDELBULB
        d.hitables[1].bulb_power = 5
        with pytest.raises(L.RaynError) as e:
            r.upload_scene_desc(d)
        assert e.value.code == L.RAYN_ERR_UNSUPPORTED
    finally:
        r.close()


def test_small_frames_replay_a_cuda_graph_and_stay_bit_exact(oracle):
    """Launch-bound frames (config 1 geometry) are captured into a CUDA graph once and replayed; the film must not depend on
    whether the kernels were launched directly (RAYN_FLAG_NO_GRAPH), captured, or replayed, and a changed scene or frame must
    invalidate the captured graph."""
    c, inp = small_config(1, (64, 64), 1, 2)
    o, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], TR)
    direct = Renderer(0, flags=L.FLAG_NO_GRAPH)
    r = Renderer(0)
    try:
        direct.upload_scene(c["world"], c["camera"])
        d = direct.render_host(inp, (16, 16), c["integrator"], TR)
        assert direct.stats().reserved_ == 0
        r.upload_scene(c["world"], c["camera"])
        # device-resident inputs and planes, so that every pointer baked into the graph stays the same between frames
        import torch
        dev = torch.device("cuda", 0)
        ins = [torch.from_numpy(a).to(dev) for a in inp.arrays()]
        from rayn_b200.dist import device_frame_desc
        npx = 64 * 64
        store = torch.zeros(10 * npx, dtype=torch.float32, device=dev)
        planes = L.RaynFilmPlanes(store[:3 * npx].data_ptr(), store[3 * npx:4 * npx].data_ptr(), store[4 * npx:7 * npx].data_ptr(), store[7 * npx:].data_ptr(), L.MEM_DEVICE)
        fd = device_frame_desc(ins, 64, 64, (16, 16), c["samples"], c["integrator"], 1, TR, (inp.sets_1d, inp.sets_2d))
        launches = []
        for i in range(3):
            store.zero_()
            torch.cuda.synchronize()
            r.render(fd, planes)
            st = r.stats()
            launches.append(st.launches)
            assert st.reserved_ == 1, f"frame {i} was not served by the captured graph"
            got = store.cpu().numpy()
            for ch, (a, b) in zip(CH, ((0, 3), (3, 4), (4, 7), (7, 10))):
                assert_bit_equal(got[a * npx:b * npx], o[ch], f"graph frame {i} {ch}")
                assert_bit_equal(got[a * npx:b * npx], d[ch], f"graph vs direct {i} {ch}")
        assert launches[0] == launches[1] == launches[2] > 0
        # another scene through the same context: the key changes, the graph is rebuilt, results follow the new scene
        c3, inp3 = small_config(3, (64, 64), 1, 2)
        r.upload_scene(c3["world"], c3["camera"])
        g3 = r.render_host(inp3, (16, 16), c3["integrator"], TR)
        o3, _ = oracle.render(c3["world"], c3["camera"], inp3, (16, 16), c3["integrator"], TR)
        for ch in CH:
            assert_bit_equal(g3[ch], o3[ch], f"after scene change {ch}")
    finally:
        direct.close()
        r.close()


def test_very_high_spp_tile_resolves_bit_exact(renderer, oracle):
    """4400 spp per pixel (more than the 4096 spp a weak-scaled 8-GPU config 3 frame uses): 1.1 M paths per tile, the film
    resolve orders 8192 keys per pixel in shared memory with 2 warps per CTA.  Scene with two objects so slots interleave."""
    c, _ = small_config(1, (16, 16), 1, 2)
    integ = PathTracingIntegrator(2, 2)
    inp = FrameInputs(16, 16, 1100, integ)
    renderer.upload_scene(c["world"], c["camera"])
    g = renderer.render_host(inp, (16, 16), integ, TR)
    o, _ = oracle.render(c["world"], c["camera"], inp, (16, 16), integ, TR)
    for ch in CH:
        assert_bit_equal(g[ch], o[ch], f"4400 spp {ch}")


def test_film_postprocess_bit_exact_and_save_to(renderer, oracle, tmp_path):
    """Film::save_to's pixel arithmetic on the device vs the oracle (u8 exact), then the
    Python Film.save_to writes the reference's file set."""
    rng = np.random.default_rng(5)
    w, h = 97, 41
    pl = {"color": rng.uniform(-0.2, 1.5, 3 * w * h).astype(np.float32), "alpha": rng.uniform(-0.1, 1.1, w * h).astype(np.float32),
          "background": rng.uniform(0, 0.5, 3 * w * h).astype(np.float32), "normal": rng.uniform(-1, 1, 3 * w * h).astype(np.float32)}
    pl["color"][:6] = [np.nan, np.inf, -np.inf, 0.0, 1.0, 1e-30]
    pl["alpha"][:3] = [np.nan, 2.0, -1.0]
    for mode in range(6):
        g = renderer.postprocess(mode, w, h, pl)
        o = oracle.film_postprocess(mode, w, h, pl)
        assert np.array_equal(g, o), f"post-process mode {mode}"
    from PIL import Image
    from rayn_b200 import BlackmanHarrisFilter, Film
    c = configs.baseline_config(3, res=(48, 32), samples=1, max_bounces=2)
    film = Film(["color", "alpha", "background", "normal"], (48, 32))
    film.render_frame_into(c["world"], c["camera"], c["integrator"], BlackmanHarrisFilter(1.5), (16, 16), 1, TR, 1)
    files = film.save_to(["alpha", "normal", "color"], str(tmp_path), "4_spp", False)  # main.rs:86-96
    assert [f.split("/")[-1] for f in files] == ["4_spp_alpha.png", "4_spp_normal.png", "4_spp_color.png"]
    img = np.asarray(Image.open(files[2]))
    flat = {k: np.ascontiguousarray(v, np.float32).reshape(-1) for k, v in film.channels.items()}
    assert np.array_equal(img, oracle.film_postprocess(L.POST_COLOR_PLUS_BACKGROUND, 48, 32, flat))
    with pytest.raises(ValueError):
        Film(["alpha"], (16, 16)).save_to(["color"], str(tmp_path), "x")


def test_device_side_frame_inputs_equal_host_builders(renderer):
    """R_d tables and the SmallRng scramble plane generated in HBM are bit-identical to the host builders."""
    import torch
    integ = PathTracingIntegrator(3, 2)
    inp = FrameInputs(301, 77, 5, integ, frame=7)
    dev = torch.device("cuda", 0)
    s1 = torch.empty(inp.samples_1d.size, dtype=torch.float32, device=dev)
    s2 = torch.empty(inp.samples_2d.size, dtype=torch.float32, device=dev)
    sc = torch.empty(inp.scramble.size, dtype=torch.float32, device=dev)
    L.check(L.lib().rayn_b200_device_frame_inputs(renderer.ctx, 301, 77, inp.spp, inp.sets_1d, inp.sets_2d, 7, s1.data_ptr(), s2.data_ptr(),
                                                  sc.data_ptr()), renderer.ctx)
    assert_bit_equal(s1.cpu().numpy(), inp.samples_1d, "1-D tables")
    assert_bit_equal(s2.cpu().numpy(), inp.samples_2d, "2-D tables")
    assert_bit_equal(sc.cpu().numpy(), inp.scramble, "scramble")


def test_time_varying_sphere_and_camera(renderer, oracle):
    """linear-in-time sphere centres and camera parameters (real motion blur over the shutter
    [1/24, 2/24], main.rs:47-49).  The reference evaluates a closure-backed parameter at LANE 0's time of the 4-lane
    packet (animation.rs:62-67); GPU and oracle must agree bit for bit, and the motion must be visible."""
    from rayn_b200 import Linear, ThinLensCamera
    materials, hitables, lights = MaterialStore(), HitableStore(), []
    sky = materials.add_material(Sky(Srgb(0.3, 0.4, 0.6), Srgb(0.2, 0.3, 0.6) * 0.05))
    grey = materials.add_material(Dielectric.new_remap(Srgb(0.5, 0.5, 0.5), 0.6))
    em = materials.add_material(Emissive.new_splat(Srgb(3.0, 2.0, 1.0)))
    hitables.push(Sphere(Vec3(0, 0, 0), 100.0, sky))
    hitables.push(TracedSDF(MandelBox(6, BoxFold(1.0), SphereFold(0.5, 1.0), -2.0), grey))
    hitables.push(Sphere(Linear(Vec3(-3.0, 1.9, 0.0), Vec3(40.0, 0.0, 0.0)), 0.4, grey))   # crosses ~1.7 units during the shutter
    hitables.push(Sphere(Linear(Vec3(0.0, -2.2, 1.0), Vec3(0.0, 6.0, 0.0)), 0.3, em))
    lights.append(SphereLight(Vec3(2.5, 2.5, 2.5), 0.2, Srgb(1, 1, 1) * 60.0))
    cams = CameraStore()
    res = (48, 40)
    moving_cam = cams.add_camera(ThinLensCamera(res, 60.0, Linear(0.02, 0.2), Linear(Vec3(-1.0, 0.45, 4.5), Vec3(2.0, 0.0, 0.0)),
                                                Vec3(0, 0, 0), Linear(Vec3(0, 1, 0), Vec3(0.5, 0, 0)), Linear(Vec3(0, 0, 0), Vec3(0, 1, 0))))
    still_cam = cams.add_camera(PinholeCamera(res, 60.0, Vec3(-1.0, 0.45, 4.5), Vec3(0, 0, 0), Vec3(0, 1, 0)))
    world = World(hitables, lights, materials, cams, VolumeParams(None, None))
    g_move, _ = _both(renderer, oracle, world, moving_cam, res, 2, 3)
    g_still_cam, _ = _both(renderer, oracle, world, still_cam, res, 2, 3)
    # same scene with the spheres frozen at t = 0 renders differently: the motion is really applied
    hitables.items[2] = Sphere(Vec3(-3.0, 1.9, 0.0), 0.4, grey)
    hitables.items[3] = Sphere(Vec3(0.0, -2.2, 1.0), 0.3, em)
    g_frozen, _ = _both(renderer, oracle, world, still_cam, res, 2, 3)
    assert not np.array_equal(g_still_cam["color"], g_frozen["color"])
    assert not np.array_equal(g_move["color"], g_still_cam["color"])
