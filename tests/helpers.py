import numpy as np

from rayn_b200 import configs
from rayn_b200.film import FrameInputs

CH = ("color", "alpha", "background", "normal")


def small_config(n, res, samples, max_bounces):
    c = configs.baseline_config(n, res=res, samples=samples, max_bounces=max_bounces)
    inp = FrameInputs(res[0], res[1], c["samples"], c["integrator"])
    return c, inp


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, what
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        idx = np.flatnonzero(~same.ravel())
        raise AssertionError(f"{what}: {idx.size}/{a.size} values differ bitwise; first at {idx[:5]}: "
                             f"{a.ravel()[idx[:5]]} vs {b.ravel()[idx[:5]]}")


def rel_err_stats(gpu, ref):
    """per-pixel relative error of displayed radiance (color + background), SURVEY T3."""
    g = np.asarray(gpu, np.float64).reshape(-1, 3)
    r = np.asarray(ref, np.float64).reshape(-1, 3)
    denom = np.maximum(np.abs(r).max(axis=1), 1e-6)
    e = np.abs(g - r).max(axis=1) / denom
    return dict(max=float(e.max()), p999=float(np.quantile(e, 0.999)), rmse=float(np.sqrt(np.mean((g - r) ** 2))))


def random_rays(n, seed, origin_radius=4.5, spread=0.5):
    """Rays from a shell around the fractal aimed near the origin."""
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * origin_radius
    target = rng.uniform(-spread, spread, size=(n, 3)) * 2.0
    d = target - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)
