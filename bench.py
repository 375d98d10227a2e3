#!/usr/bin/env python
"""bench.py — Msamples/s (pixels x spp) of the wavefront render path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of rayn's path (oracle)

A "step" is one full `render_frame_into` of the workload.  At N=1 the workload is BASELINE
config 2 (Mandelbulb 1024x1024, 128 spp, 4 bounces).  At N>1 film tiles are sharded
`tile % N == rank` (no data-path collective; one NCCL all-gather of the film at the end of
each step, inside the timed region) and spp grows with N so per-GPU work stays fixed
("weak"); `--scaling strong` keeps the config fixed instead.

`value` is measured with the sampler tables, scramble plane and film resident in HBM;
`e2e` goes through the public host-buffer API (H2D of the inputs and D2H of the film planes
inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_EXTEND = 40.0   # K2 reads float4 o+time, float4 d+tmax (32 B), writes t+key (8 B)
ALG_BYTES_SHADE = 184.0   # shade: read ray 68 + hit 8, write ray 68 or film <= 40
MANDELBULB_FLOP_PER_ITER = 75.0  # authored formula, counted in DESIGN.md
MANDELBOX_FLOP_PER_ITER = 25.0
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # non-tensor FP32: SMs x lanes x FMA x max clock


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--res", type=int, nargs=2, default=None, help="override resolution (debug)")
    ap.add_argument("--samples", type=int, default=None, help="override SAMPLES (spp/4) (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-paths", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0, help="RAYN_FLAG_* kernel-family selection (2 = v0 simple, 4 = v2 block pools)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def build_workload(args, world):
    from rayn_b200 import configs
    samples = args.samples
    base = configs.BASELINE_CONFIGS[args.config]
    if samples is None:
        samples = base["samples"]
    if args.scaling == "weak" and world > 1:
        samples *= world
    c = configs.baseline_config(args.config, res=args.res, samples=samples)
    c["time_range"] = configs.frame_time_range(1)
    return c


def workload_name(c, args, world):
    w, h = c["res"]
    s = f"{c['name'].split('-')[0]} {'Mandelbulb(authored)' if 'mandelbulb' in c['name'] else c['name'].split('-')[1]} {w}x{h} {c['spp']}spp {c['max_bounces']}b"
    if world > 1:
        s += f" tiles (tx+ty)%{world} ({args.scaling}: spp {'x' + str(world) if args.scaling == 'weak' else 'fixed'})"
    return s


def cpu_sample(c, inputs, target_seconds, threads=0):
    """Time the CPU oracle on every k-th tile, k chosen so the sample costs ~target_seconds."""
    from oracle import binding as ob
    from rayn_b200.film import tile_grid
    w, h = c["res"]
    n_tiles = int(np.prod(tile_grid(w, h, 16, 16)))
    ncores = os.cpu_count() or 1
    if threads == 0:
        threads = ncores  # explicit: torchrun exports OMP_NUM_THREADS=1, which would silently serialise the CPU baseline
    # probe: a spread of ~2*ncores tiles, to size the real sample
    k = max(1, n_tiles // max(2 * ncores, 8))
    t = time.perf_counter()
    _, info = ob.render(c["world"], c["camera"], inputs, (16, 16), c["integrator"], c["time_range"], n_threads=threads, subsample_k=k)
    dt = time.perf_counter() - t
    per_tile = dt / max(info["tiles"], 1)
    want = int(min(n_tiles, max(info["tiles"], target_seconds / max(per_tile, 1e-9))))
    k2 = max(1, n_tiles // max(want, 1))
    if k2 < k:
        t = time.perf_counter()
        _, info = ob.render(c["world"], c["camera"], inputs, (16, 16), c["integrator"], c["time_range"], n_threads=threads, subsample_k=k2)
        dt = time.perf_counter() - t
        k = k2
    samples_done = info["tiles"] * 256 * c["spp"]
    return dict(value=samples_done / dt / 1e6, unit="Msamples/s", cores=threads, kind="port",
                sample=f"every {k}-th 16x16 tile of the workload ({info['tiles']} of {n_tiles} tiles, {samples_done / 1e6:.2f} Msamples, {dt:.1f} s), "
                       f"OpenMP over tiles like rayon; C++ SSE-packet restatement of rayn's path (rayn itself cannot be built here)"), dt


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of rayn's render path on the host cores."""
    if rank != 0:
        return
    from rayn_b200.film import FrameInputs
    from oracle import binding as ob
    ob.build()
    c = build_workload(args, world)
    w, h = c["res"]
    inputs = FrameInputs(w, h, c["samples"], c["integrator"])
    per_step = max(3.0, min(args.cpu_seconds, 100.0 / max(args.steps + args.warmup, 1)))
    for _ in range(min(args.warmup, 1)):
        cpu_sample(c, inputs, min(per_step, 3.0))
    vals, secs, last = [], 0.0, None
    for _ in range(args.steps):
        last, dt = cpu_sample(c, inputs, per_step)
        vals.append(last["value"])
        secs += dt
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": "Msamples/sec (pixels x spp)", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / max(args.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(c, args, world), "bounded_sample": True},
            "cpu_baseline": last, "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from rayn_b200 import _lib as L
    from rayn_b200.dist import DistFilm, device_frame_desc
    from rayn_b200.film import FrameInputs, Renderer, make_frame_desc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rayn_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    c = build_workload(args, world)
    w, h = c["res"]
    tile = (16, 16)
    inputs = FrameInputs(w, h, c["samples"], c["integrator"])
    sets = (inputs.sets_1d, inputs.sets_2d)
    total_samples = None  # filled from stats (covered pixels x spp)

    # ---- resident inputs + film ---------------------------------------------------------------
    inputs_dev = [torch.from_numpy(a).to(dev) for a in inputs.arrays()]
    r = Renderer(local_rank, max_paths_per_pass=args.max_paths, flags=args.flags)
    r.upload_scene(c["world"], c["camera"])
    film = DistFilm(r, w, h, tile, rank, world)
    fdesc = device_frame_desc(inputs_dev, w, h, tile, c["samples"], c["integrator"], 1, c["time_range"], sets, film.tile_list)

    def step_resident():
        film.render(fdesc)
        film.gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    clocks = ClockSampler(local_rank) if rank == 0 else None  # started before warm-up (nvidia-smi needs ~0.3 s to emit its first line); every sample is under load
    for _ in range(args.warmup):
        step_resident()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    lib_ms = 0.0
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_resident()
        st = r.stats()
        launches += st.launches + (0 if world == 1 else world)  # + pack/unpack kernels of the gather
        lib_ms += st.total_ms
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ev_ms = e0.elapsed_time(e1)
    barrier()
    clock_info = clocks.stop() if clocks else None
    st = r.stats()
    my_paths = st.paths
    total_samples = allsum(float(my_paths))
    ms_per_step = allmax(max(ev_ms, lib_ms)) / args.steps
    value = total_samples / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel breakdown + roofline (separate TIMING context so events do not perturb `value`) ----
    rt = Renderer(local_rank, max_paths_per_pass=args.max_paths, flags=L.FLAG_TIMING | args.flags)
    rt.upload_scene(c["world"], c["camera"])
    filmt = DistFilm(rt, w, h, tile, rank, world)
    filmt.render(fdesc)
    filmt.render(fdesc)
    ts = rt.stats()
    kms = {L.KERNEL_NAMES[i]: float(ts.kernel_ms[i]) for i in range(len(L.KERNEL_NAMES)) if ts.kernel_launches[i]}
    klaunch = {L.KERNEL_NAMES[i]: int(ts.kernel_launches[i]) for i in range(len(L.KERNEL_NAMES)) if ts.kernel_launches[i]}
    ksum = sum(kms.values())
    hbm_peak, peak_src = peaks()
    is_bulb = "mandelbulb" in c["name"]
    sdf = [hh for hh in c["world"].hitables.items if hasattr(hh, "sdf")]
    iters = sdf[0].sdf.iterations if sdf else 0
    flop_eval = (MANDELBULB_FLOP_PER_ITER if is_bulb else MANDELBOX_FLOP_PER_ITER) * iters + 10
    ALG_BYTES_SHADOW = 40.0  # read seg_a + seg_b + owner (36 B), clear one visibility bit (4 B)
    ext_s = kms.get("extend", 0.0) * 1e-3
    shadow_s = kms.get("shadow", 0.0) * 1e-3
    fused_shade = "shadow" not in kms  # v0/v2 kernel families fuse the shadow march into shade
    shade_s = (kms.get("shade_pre", 0.0) + kms.get("shade_post", 0.0)) * 1e-3
    per_kernel = {}
    for k in kms:
        per_kernel[k] = {"ms": kms[k], "share": kms[k] / max(ksum, 1e-9), "launches": klaunch[k]}
    per_kernel.setdefault("extend", {}).update(
        rays=int(ts.extend_rays), sdf_evals=int(ts.sdf_evals_extend),
        hbm_gbs_algorithmic=ts.extend_rays * ALG_BYTES_EXTEND / max(ext_s, 1e-12) / 1e9,
        fp32_tflops_algorithmic=ts.sdf_evals_extend * flop_eval / max(ext_s, 1e-12) / 1e12)
    sh_key = "shade_pre" if fused_shade else "shadow"
    sh_s = shade_s if fused_shade else shadow_s
    per_kernel.setdefault(sh_key, {}).update(
        shadow_rays=int(ts.shadow_rays), sdf_evals=int(ts.sdf_evals_shadow),
        hbm_gbs_algorithmic=(ts.shade_lanes * ALG_BYTES_SHADE if fused_shade else ts.shadow_rays * ALG_BYTES_SHADOW) / max(sh_s, 1e-12) / 1e9,
        fp32_tflops_algorithmic=ts.sdf_evals_shadow * flop_eval / max(sh_s, 1e-12) / 1e12)
    if not fused_shade:
        per_kernel["shade_pre"]["lanes"] = int(ts.shade_lanes)
        per_kernel["shade_pre"]["hbm_gbs_algorithmic"] = ts.shade_lanes * ALG_BYTES_SHADE / max(shade_s, 1e-12) / 1e9
    # the genuinely HBM-bound streaming kernels, for context next to the issue-bound march kernels
    ALG_BYTES_RAYGEN = 88.0   # writes o_time, d_t, rad, thr, nrm0 (5 x float4) + term + q_live per path
    ALG_BYTES_RESOLVE = 36.0  # reads rad + nrm0 (2 x float4) + term per path (+ 40 B per pixel out)
    if "raygen" in per_kernel:
        per_kernel["raygen"]["hbm_gbs_algorithmic"] = ts.paths * ALG_BYTES_RAYGEN / max(kms["raygen"] * 1e-3, 1e-12) / 1e9
        per_kernel["raygen"]["hbm_frac"] = per_kernel["raygen"]["hbm_gbs_algorithmic"] / hbm_peak
    if "resolve" in per_kernel:
        per_kernel["resolve"]["hbm_gbs_algorithmic"] = (ts.paths * ALG_BYTES_RESOLVE + w * h * 40.0) / max(kms["resolve"] * 1e-3, 1e-12) / 1e9
        per_kernel["resolve"]["hbm_frac"] = per_kernel["resolve"]["hbm_gbs_algorithmic"] / hbm_peak
    dom = max(kms, key=kms.get) if kms else "extend"
    dom_name = {"extend": "k_extend_march (closest-hit sphere-march)", "shadow": "k_shadow (occlusion sphere-march)",
                "shade_pre": "k_shade_pre", "shade_post": "k_shade_post"}.get(dom, dom)
    achieved = per_kernel[dom].get("hbm_gbs_algorithmic", 0.0)
    roofline = {"kernel": dom_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src, "traffic": None,
                "note": "march kernels are FP32-issue/divergence bound, not HBM bound (SURVEY F7): achieved = algorithmic bytes / kernel time; "
                        "fp32 fraction of the non-tensor peak reported alongside",
                "fp32_peak_tflops_nominal": FP32_PEAK_TFLOPS,
                "fp32_frac_extend": per_kernel["extend"].get("fp32_tflops_algorithmic", 0.0) / FP32_PEAK_TFLOPS,
                "fp32_frac_shadow": per_kernel[sh_key].get("fp32_tflops_algorithmic", 0.0) / FP32_PEAK_TFLOPS,
                "launches_of_kernel_per_step": klaunch.get(dom, 0)}
    # measured DRAM traffic of the dominant kernel from the committed ncu capture (per launch, like `achieved`)
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.config == 2 and world == 1 and args.flags == 0:
        tj = json.load(open(tpath))
        key = {"extend": "k_extend_march", "shadow": "k_shadow"}.get(dom)
        if key in tj:
            roofline["traffic"] = tj[key]["dram_bytes_per_launch"][0]
            roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the depth-0 launch of pass 2 (50.3 M rays) from "
                                        + tj[key]["source"] + "; algorithmic bytes of that launch: "
                                        + str(tj[key].get("algorithmic_bytes_per_launch", [None])[0]))
    rt.close()
    del filmt

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region -----------------
    e2e = None
    if not args.no_e2e:
        pin = [torch.from_numpy(a).pin_memory() for a in inputs.arrays()]
        npx = w * h
        out_pin = torch.zeros(10 * npx, dtype=torch.float32).pin_memory()
        hp = L.RaynFilmPlanes(out_pin.data_ptr(), out_pin[3 * npx:].data_ptr(), out_pin[4 * npx:].data_ptr(), out_pin[7 * npx:].data_ptr(), L.MEM_HOST)
        hdesc = make_frame_desc(w, h, tile, c["samples"], c["integrator"], 1, c["time_range"], tuple(t.data_ptr() for t in pin), L.MEM_HOST,
                                0, 1, sets, film.tile_list)
        h2d = sum(t.numel() * 4 for t in pin)
        d2h = out_pin.numel() * 4

        def step_e2e():
            if world == 1:
                r.render(hdesc, hp)
                return float(out_pin[:3].sum())
            for src, dst in zip(pin, inputs_dev):
                dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            film.render(fdesc)
            film.gather()
            if rank == 0:
                out_pin.copy_(film.store, non_blocking=False)
            return 0.0

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = allmax(dt)
        e2e = {"value": total_samples / (dt / args.steps) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "timing": "host wall clock around the public host-buffer call, max over ranks"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_sample(c, inputs, args.cpu_seconds)

    if rank == 0:
        line = {"metric": "Msamples/sec (pixels x spp)", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(c, args, world), "tile": "16x16", "samples_total": int(total_samples),
                           "l2": "working set (path state >= 3 GB/pass) far exceeds the 126 MB L2; no explicit flush",
                           "parallelism": f"dp{world}: 16x16 film tiles, (tx+ty)%{world} interleave, NCCL all-gather of the film" if world > 1 else "1 GPU"},
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "kernels": per_kernel, "wall_ms_per_step": wall_ms / args.steps, "passes_per_step": int(st.passes)}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
