"""Accuracy of the deterministic math primitives (rayn_b200/csrc/detmath.h) against mpmath:
they stand in for libm calls of the reference, so they must be good-libm accurate (<= 1 ulp)."""
import mpmath as mp
import numpy as np
import pytest

mp.mp.dps = 40


def ulp_err(got, exact):
    got = np.asarray(got, np.float32)
    out = []
    for g, e in zip(got.tolist(), exact):
        e = mp.mpf(e)
        if e == 0:
            out.append(0.0 if g == 0 else float("inf"))
            continue
        ef = np.float32(float(e))
        ulp = float(np.spacing(np.abs(ef))) if np.isfinite(ef) and ef != 0 else 1e-45
        out.append(float(abs(mp.mpf(g) - e) / ulp))
    return np.array(out)


CASES = [
    (0, lambda r: r.uniform(-80, 5, 3000), lambda x: mp.e ** mp.mpf(x)),
    (1, lambda r: r.uniform(1e-6, 60, 3000), lambda x: mp.log(mp.mpf(x))),
    (3, lambda r: r.uniform(-7, 7, 3000), lambda x: mp.sin(mp.mpf(x))),
    (4, lambda r: r.uniform(-7, 7, 3000), lambda x: mp.cos(mp.mpf(x))),
    (5, lambda r: r.uniform(-1.5, 1.5, 3000), lambda x: mp.tan(mp.mpf(x))),
]


@pytest.mark.parametrize("op,gen,f", CASES)
def test_unary_accuracy(oracle, op, gen, f):
    x = gen(np.random.default_rng(op)).astype(np.float32)
    got = oracle.kat_detmath(op, x)
    err = ulp_err(got, [f(float(v)) for v in x])
    assert err.max() <= 1.0, f"op {op}: max {err.max()} ulp"


def test_pow_accuracy(oracle):
    r = np.random.default_rng(2)
    x = r.uniform(1e-4, 1.0, 3000).astype(np.float32)
    y = r.uniform(0.05, 12.0, 3000).astype(np.float32)
    got = oracle.kat_detmath(2, x, y)
    err = ulp_err(got, [mp.mpf(float(a)) ** mp.mpf(float(b)) for a, b in zip(x, y)])
    assert err.max() <= 1.0


def test_atan2_accuracy(oracle):
    r = np.random.default_rng(6)
    y = r.uniform(-200, 200, 3000).astype(np.float32)
    x = r.uniform(0, 50, 3000).astype(np.float32)
    got = oracle.kat_detmath(6, y, x)
    err = ulp_err(got, [mp.atan2(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(y, x)])
    assert err.max() <= 1.0


def test_special_values(oracle):
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    e = oracle.kat_detmath(0, np.array([0, -inf, inf, nan, -200, 100], np.float32))
    assert e[0] == 1 and e[1] == 0 and e[2] == inf and np.isnan(e[3]) and e[4] == 0 and e[5] == inf
    p = oracle.kat_detmath(2, np.array([0, 0, 1, 2, -1, nan, 0.5], np.float32), np.array([2, 0, 7, 0, 0.5, 1, 1], np.float32))
    assert p[0] == 0 and p[1] == 1 and p[2] == 1 and p[3] == 1 and np.isnan(p[4]) and np.isnan(p[5]) and p[6] == 0.5
    assert oracle.kat_detmath(7, np.array([2.0], np.float32))[0] == 32.0  # powi(5)
    s = oracle.kat_detmath(3, np.array([0.0, nan], np.float32))
    assert s[0] == 0 and np.isnan(s[1])
    a = oracle.kat_detmath(6, np.array([0.0, 1.0, -1.0], np.float32), np.array([0.0, 0.0, 0.0], np.float32))
    assert a[0] == 0 and abs(a[1] - np.pi / 2) < 1e-6 and abs(a[2] + np.pi / 2) < 1e-6
