"""T1 closed-form checks: the oracle's sampling routines against mathematics -
the only external truth available, since the reference ships no tests."""
import numpy as np

from rayn_b200 import _lib as L
from rayn_b200.scene import Dielectric, Lambertian, SphereLight, Srgb, Vec3


def test_sphere_light_cone_sampling(oracle):
    """light.rs:38-72: samples lie on the light's sphere, on the side facing the shading point, inside the tangent cone,
    uniformly in solid angle; pdf = 1 / (2 pi (1 - cos theta_max))."""
    light = SphereLight(Vec3(1.0, 2.0, -0.5), 0.4, Srgb(1, 1, 1)).flatten()
    rng = np.random.default_rng(0)
    n = 40000
    p = np.tile(np.array([[-2.0, 0.5, 1.5]], np.float32), (n, 1))
    s0, s1 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    pt, pdf = oracle.kat_light_sample(light, s0, s1, p)
    pos = np.array(light.pos[:], np.float64)
    assert np.allclose(np.linalg.norm(pt - pos, axis=1), 0.4, atol=2e-6)
    to_l = pos - p[0]
    dist = np.linalg.norm(to_l)
    cos_max = np.sqrt(1 - (0.4 / dist) ** 2)
    w = pt - p[0]
    w = w / np.linalg.norm(w, axis=1, keepdims=True)
    cos_t = w @ (to_l / dist)
    assert (cos_t >= cos_max - 1e-5).all()                       # inside the cone
    assert ((pt - pos) @ to_l < 1e-6).all()                      # on the near side of the sphere
    assert np.allclose(pdf, 1.0 / (2 * np.pi * (1 - cos_max)), rtol=2e-5)
    assert abs(cos_t.mean() - (1 + cos_max) / 2) < 3e-4          # uniform in solid angle <=> cos theta uniform on [cos_max, 1]
    axis = to_l / dist
    e1 = np.cross(axis, [0, 0, 1.0]); e1 /= np.linalg.norm(e1)
    e2 = np.cross(axis, e1)
    phi_moment = np.abs(np.mean(np.exp(1j * np.arctan2(w @ e2, w @ e1))))
    assert phi_moment < 0.02                                     # azimuth uniform around the cone axis


def test_equi_angular_sampling_pdf_integrates_to_one(oracle):
    """light.rs:75-102 (Kulla & Fajardo): t in [0, t_max], pdf(t) = D / ((theta_b - theta_a)(D^2 + (t - delta)^2)) and its integral is 1."""
    light = SphereLight(Vec3(0.3, 1.1, 0.2), 0.15, Srgb(1, 1, 1)).flatten()
    n = 20001
    u = np.linspace(0.0, 1.0, n).astype(np.float32)
    o = np.tile(np.array([[-1.0, 0.2, 3.0]], np.float32), (n, 1))
    d = np.array([0.25, 0.1, -1.0]); d = (d / np.linalg.norm(d)).astype(np.float32)
    tmax = 6.0
    t, pdf = oracle.kat_light_sample_volume(light, u, o, np.tile(d, (n, 1)), np.full(n, tmax, np.float32))
    assert t.min() >= -1e-4 and t.max() <= tmax + 1e-4 and (np.diff(t) >= -1e-5).all()   # monotone map of [0,1] onto [0,t_max]
    pos = np.array(light.pos[:], np.float64)
    delta = (pos - o[0]) @ d.astype(np.float64)
    D = np.linalg.norm(o[0] + delta * d - pos)
    th_a, th_b = np.arctan2(-delta, D), np.arctan2(tmax - delta, D)
    want = D / ((th_b - th_a) * (D * D + (t.astype(np.float64) - delta) ** 2))
    assert np.allclose(pdf, want, rtol=2e-4)
    integral = np.trapezoid(pdf.astype(np.float64), t.astype(np.float64)) if hasattr(np, "trapezoid") else np.trapz(pdf.astype(np.float64), t.astype(np.float64))
    assert abs(integral - 1.0) < 2e-3


def _hemisphere_inputs(n, seed):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    wo = rng.normal(size=(n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    wo = np.where(((wo * nrm).sum(1) < 0)[:, None], -wo, wo)  # outgoing direction on the normal's side
    return nrm.astype(np.float32), wo.astype(np.float32), rng.random(n, dtype=np.float32), rng.random((n, 4), dtype=np.float32)


def test_lambertian_scatter_is_cosine_weighted(oracle):
    """material.rs:117-142: wi unit length in the normal's hemisphere, pdf = cos/pi, f = albedo/pi  =>  f cos / pdf = albedo (white furnace)."""
    m = Lambertian(Srgb(0.6, 0.4, 0.2)).flatten()
    nrm, wo, s1d, u4 = _hemisphere_inputs(20000, 1)
    wi, f, pdf, fe = oracle.kat_bsdf(m, nrm, wo, s1d, u4)
    cos = (wi * nrm).sum(1)
    assert np.allclose(np.linalg.norm(wi, axis=1), 1.0, atol=3e-6) and (cos > -1e-6).all()
    ok = pdf > 1e-4
    assert np.allclose(pdf[ok], cos[ok] / np.pi, rtol=3e-4, atol=2e-6)
    assert np.allclose((f[ok] * (cos[ok] / pdf[ok])[:, None]), np.array([0.6, 0.4, 0.2]), rtol=1e-3)
    assert np.allclose(fe, np.array([0.6, 0.4, 0.2]) / np.pi, rtol=1e-6)
    assert abs(cos.mean() - 2.0 / 3.0) < 5e-3   # E[cos] under a cosine-weighted density


def test_dielectric_lobe_selection_and_reference_quirk(oracle):
    """material.rs:207-256: the lobe is chosen by `sample < fresnel(|n.wo|)`, pdf = F pdf_s + (1-F) pdf_d.
    SURVEY F8: `wo.reflected(norm)` of an OUTGOING vector points below the surface, so most sampled specular directions are
    below the horizon and their f is zeroed (material.rs:241-242) - reproduced, not fixed."""
    m = Dielectric.new_remap(Srgb(0.2, 0.2, 0.2), 0.6).flatten()
    nrm, wo, s1d, u4 = _hemisphere_inputs(20000, 2)
    wi, f, pdf, fe = oracle.kat_bsdf(m, nrm, wo, s1d, u4)
    cosv = np.abs((nrm * wo).sum(1)).astype(np.float64)
    fres = 0.04 + 0.96 * (1 - cosv) ** 5
    spec = s1d < fres - 1e-6
    diff = s1d > fres + 1e-6
    assert 0.02 < spec.mean() < 0.5 and diff.mean() > 0.5
    assert np.allclose(f[diff], 0.2 / np.pi, rtol=1e-5)                       # diffuse lobe: albedo / pi
    below = (wi[spec] * nrm[spec]).sum(1) < 0
    assert below.mean() > 0.5 and (f[spec][below] == 0).all()                 # the quirk: most specular samples land below the horizon, f = 0
    assert (f[spec][~below] > 0).all()                                        # the rest of the (wide, exponent 8.68) lobe keeps its Phong value
    assert (pdf > 0).all() and np.isfinite(pdf).all()
    assert np.isfinite(fe).all() and (fe >= 0).all()


def test_packet_coupled_light_selection_is_unbiased(oracle):
    """rayn picks the 4 NEE lights of a shading packet from the 4 lanes' samples and applies each to all lanes
    with weight L/4 (integrator.rs:76-93).  A per-lane choice with the same weight is also unbiased, so the two
    high-spp means must agree within Monte-Carlo error; a wrong correction factor would shift one of them."""
    from rayn_b200 import (CameraStore, HitableStore, Lambertian, MaterialStore, PathTracingIntegrator, PinholeCamera, Sky, Sphere, SphereLight,
                           VolumeParams, World, configs)
    from rayn_b200.film import FrameInputs
    materials, hitables = MaterialStore(), HitableStore()
    sky = materials.add_material(Sky(Srgb(0, 0, 0), Srgb(0, 0, 0)))
    white = materials.add_material(Lambertian(Srgb(0.7, 0.7, 0.7)))
    hitables.push(Sphere(Vec3(0, 0, 0), 100.0, sky))
    hitables.push(Sphere(Vec3(0, 0, 0), 1.0, white))
    lights = [SphereLight(Vec3(2.0, 2.0, 2.0), 0.2, Srgb(30, 5, 5)), SphereLight(Vec3(-2.5, 0.5, 2.0), 0.3, Srgb(5, 30, 5)),
              SphereLight(Vec3(0.0, -2.5, 2.5), 0.25, Srgb(5, 5, 30))]
    cams = CameraStore()
    cam = cams.add_camera(PinholeCamera((16, 16), 30.0, Vec3(0.0, 0.0, 5.0), Vec3(0, 0, 0), Vec3(0, 1, 0)))
    world = World(hitables, lights, materials, cams, VolumeParams(None, None))
    integ = PathTracingIntegrator(0, 2)  # direct lighting only: Color = NEE at depth 0
    inp = FrameInputs(16, 16, 512, integ)  # 2048 spp
    tr = configs.frame_time_range(1)
    a, _ = oracle.render(world, cam, inp, (16, 16), integ, tr)
    oracle.set_decoupled_lights(True)
    try:
        b, _ = oracle.render(world, cam, inp, (16, 16), integ, tr)
    finally:
        oracle.set_decoupled_lights(False)
    ca, cb = a["color"].reshape(-1, 3), b["color"].reshape(-1, 3)
    lit = ca.sum(1) > 0.05 * ca.sum(1).max()
    assert lit.sum() > 60
    assert not np.array_equal(ca, cb)                                         # genuinely different estimators
    assert np.allclose(ca[lit].mean(0), cb[lit].mean(0), rtol=0.01)           # same mean per colour channel (each light has its own colour)
    per_pixel = np.abs(ca[lit] - cb[lit]).sum(1) / ca[lit].sum(1)
    assert np.median(per_pixel) < 0.05
    again, _ = oracle.render(world, cam, inp, (16, 16), integ, tr)            # switch is off again: bit-identical to the first render
    assert np.array_equal(again["color"], a["color"])
