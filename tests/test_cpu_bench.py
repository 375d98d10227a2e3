"""bench.py plumbing that can be checked without a GPU: workload construction, names, peaks fallback, clock-log parsing."""
import argparse
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _args(**kw):
    d = dict(config=2, samples=None, res=None, scaling="weak", flags=0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_default_workload_is_baseline_config_2_and_weak_scaling_multiplies_spp():
    b = _bench()
    c1 = b.build_workload(_args(), 1)
    assert c1["res"] == (1024, 1024) and c1["spp"] == 128 and c1["max_bounces"] == 4
    assert b.workload_name(c1, _args(), 1) == "cfg2 Mandelbulb(authored) 1024x1024 128spp 4b"
    c8 = b.build_workload(_args(), 8)
    assert c8["spp"] == 1024 and "spp x8" in b.workload_name(c8, _args(), 8)
    s8 = b.build_workload(_args(scaling="strong"), 8)
    assert s8["spp"] == 128 and "spp fixed" in b.workload_name(s8, _args(scaling="strong"), 8)
    c3 = b.build_workload(_args(config=3), 1)
    assert c3["res"] == (1920, 1080) and c3["spp"] == 512 and "mandelbox" in b.workload_name(c3, _args(config=3), 1)


def test_peaks_reads_measured_file_or_falls_back(tmp_path, monkeypatch):
    b = _bench()
    v, src = b.peaks()
    assert v > 1000 and ("measured" in src or "fallback" in src)
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    v, src = b.peaks()
    assert v == 6650.0 and "fallback" in src


def test_clock_log_parsing(tmp_path):
    b = _bench()
    s = b.ClockSampler.__new__(b.ClockSampler)
    f = open(tmp_path / "c.csv", "w+")
    f.write("0, 1965, 1965, 812.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active\n"
            "0, 1950, 1965, 990.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active\n"
            "garbage line\n"
            "0, 1305, 1965, 1001.0, 0x0, Not Active, Active, Not Active, Not Active\n")
    s.f = f

    class P:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

        def kill(self):
            pass
    s.p = P()
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
