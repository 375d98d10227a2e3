"""Film + the render entry point: the host-side mirror of reference src/film.rs.

`Film.render_frame_into(world, camera, integrator, filter, tile_size, frame, time_range,
samples)` has the reference's signature (film.rs:382-395) and is the user-facing call; it
builds the host-owned sampler state exactly where the reference does (film.rs:429-434,
460-461), flattens the World and calls `rayn_b200_render_frame` through the C ABI.

`Renderer` is the thin handle on the C context for callers that keep buffers resident on the
device (bench, multi-GPU driver).  There is no CPU path: without the CUDA library or a GPU
these raise.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .scene import BlackmanHarrisFilter, PathTracingIntegrator

CHANNELS = ("color", "alpha", "background", "normal")  # ChannelKind, film.rs:103-120


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class FrameInputs:
    """Host-owned sampler state of one frame: `Samples::new_rd` tables (film.rs:434,
    sampler.rs:18-37), per-pixel SmallRng scramble (film.rs:460-461) and the
    FilterImportanceSampler table (film.rs:429).  Built by the pure-CPU helpers of the C ABI."""

    def __init__(self, width, height, samples, integrator, filt=None, frame=1):
        lib = L.host_lib()  # pure CPU: building frame inputs must not need (or map) the CUDA library
        filt = filt or BlackmanHarrisFilter(1.5)
        self.width, self.height, self.samples, self.spp, self.frame = width, height, samples, 4 * samples, frame
        self.sets_1d = 1 + integrator.requested_1d_sample_sets()  # film.rs:431
        self.sets_2d = 2 + integrator.requested_2d_sample_sets()  # film.rs:432
        self.samples_1d = np.empty(self.spp * self.sets_1d, np.float32)
        self.samples_2d = np.empty(2 * self.spp * self.sets_2d, np.float32)
        self.scramble = np.empty(width * height, np.float32)
        self.fis = np.empty(L.RAYN_FIS_TABLE_SIZE, np.float32)
        L.check(lib.rayn_b200_host_rd_tables(self.spp, self.sets_1d, self.sets_2d, frame, _fptr(self.samples_1d), _fptr(self.samples_2d)))
        L.check(lib.rayn_b200_host_scramble(width, height, _fptr(self.scramble)))
        L.check(lib.rayn_b200_host_fis_blackman_harris(filt.radius, _fptr(self.fis)))

    def arrays(self):
        return self.samples_1d, self.samples_2d, self.scramble, self.fis


def make_frame_desc(width, height, tile_size, samples, integrator, frame, time_range, ptrs, space, tile_offset=0,
                    tile_stride=1, sets=None, tile_list=None):
    f = L.RaynFrameDesc()
    f.width, f.height = width, height
    f.tile_w, f.tile_h = tile_size
    f.samples = samples
    f.max_bounces = integrator.max_bounces
    f.volume_marches = integrator.volume_marches
    f.frame = frame
    f.t0, f.t1 = float(np.float32(time_range[0])), float(np.float32(time_range[1]))
    f.sets_1d, f.sets_2d = sets if sets else (1 + integrator.requested_1d_sample_sets(), 2 + integrator.requested_2d_sample_sets())
    f.samples_1d, f.samples_2d, f.scramble, f.fis_inverse_cdf = ptrs
    f.input_space = space
    f.tile_offset, f.tile_stride = tile_offset, tile_stride
    if tile_list is not None:  # explicit shard (ascending tile indices); the array must outlive the render call
        arr = (C.c_int32 * len(tile_list))(*tile_list)
        f.tile_list, f.n_tile_list = C.cast(arr, C.POINTER(C.c_int32)), len(tile_list)
        f._keep_tile_list = arr
    return f


def tile_grid(width, height, tile_w, tile_h):
    nx, ny = C.c_int32(), C.c_int32()
    L.check(L.host_lib().rayn_b200_host_tile_grid(width, height, tile_w, tile_h, C.byref(nx), C.byref(ny)))
    return nx.value, ny.value


class Renderer:
    """Owns one RaynContext (one GPU)."""

    def __init__(self, device=0, max_paths_per_pass=0, flags=0):
        self._lib = L.lib()
        cfg = L.RaynConfig(device, max_paths_per_pass, flags)
        self._ctx = C.c_void_p()
        L.check(self._lib.rayn_b200_create(C.byref(cfg), C.byref(self._ctx)))
        self._keep = None
        self.device = device

    def close(self):
        if self._ctx:
            self._lib.rayn_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ctx(self):
        return self._ctx

    def upload_scene(self, world, camera):
        desc, keep = world.flatten(camera)
        self._keep = (desc, keep)
        L.check(self._lib.rayn_b200_upload_scene(self._ctx, C.byref(desc)), self._ctx)

    def upload_scene_desc(self, desc):
        L.check(self._lib.rayn_b200_upload_scene(self._ctx, C.byref(desc)), self._ctx)

    def render(self, frame_desc, planes):
        L.check(self._lib.rayn_b200_render_frame(self._ctx, C.byref(frame_desc), C.byref(planes)), self._ctx)

    def stats(self):
        s = L.RaynStats()
        L.check(self._lib.rayn_b200_get_stats(self._ctx, C.byref(s)), self._ctx)
        return s

    def render_host(self, inputs, tile_size, integrator, time_range, tile_offset=0, tile_stride=1, tile_list=None):
        """Host buffers in, host planes out (H2D + D2H inside the call).  Returns dict of numpy planes."""
        w, h = inputs.width, inputs.height
        planes = {"color": np.zeros(3 * w * h, np.float32), "alpha": np.zeros(w * h, np.float32),
                  "background": np.zeros(3 * w * h, np.float32), "normal": np.zeros(3 * w * h, np.float32)}
        p = L.RaynFilmPlanes(planes["color"].ctypes.data, planes["alpha"].ctypes.data, planes["background"].ctypes.data,
                             planes["normal"].ctypes.data, L.MEM_HOST)
        ptrs = tuple(a.ctypes.data for a in inputs.arrays())
        f = make_frame_desc(w, h, tile_size, inputs.samples, integrator, inputs.frame, time_range, ptrs, L.MEM_HOST, tile_offset,
                            tile_stride, (inputs.sets_1d, inputs.sets_2d), tile_list)
        self.render(f, p)
        return planes

    def postprocess(self, mode, width, height, planes):
        """Film::save_to pixel arithmetic on the device (film.rs:205-377): numpy planes in, uint8 [H, W, bpp] out (rows top to bottom)."""
        def ptr(k):
            return planes[k].ctypes.data if planes.get(k) is not None else None
        p = L.RaynFilmPlanes(ptr("color"), ptr("alpha"), ptr("background"), ptr("normal"), L.MEM_HOST)
        out = np.zeros((height, width, L.POST_BYTES[mode]), np.uint8)
        L.check(self._lib.rayn_b200_film_postprocess(self._ctx, mode, width, height, C.byref(p), out.ctypes.data, L.MEM_HOST), self._ctx)
        return out

    # ---- known-answer entry points (tests) ----
    def kat_detmath(self, op, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b if b is not None else a, np.float32)
        out = np.empty_like(a)
        L.check(self._lib.rayn_b200_kat_detmath(self._ctx, op, a.size, _fptr(a), _fptr(b), _fptr(out)), self._ctx)
        return out

    def kat_sdf_dist(self, hitable, points):
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.empty(len(p), np.float32)
        L.check(self._lib.rayn_b200_kat_sdf_dist(self._ctx, C.byref(hitable), len(p), _fptr(p), _fptr(out)), self._ctx)
        return out

    def kat_sdf_dist2(self, hitable, points, variant=-1):
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.empty(len(p), np.float32)
        L.check(self._lib.rayn_b200_kat_sdf_dist2(self._ctx, C.byref(hitable), variant, len(p), _fptr(p), _fptr(out)), self._ctx)
        return out

    def kat_fastdiv(self, num, first_bits, n):
        bad = C.c_int64(-1)
        L.check(self._lib.rayn_b200_kat_fastdiv(self._ctx, num, first_bits, n, C.byref(bad)), self._ctx)
        return bad.value

    def kat_sdf_hit(self, hitable, consts, origins, dirs, t_max, thr_scale, thr_const=0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        tm = np.ascontiguousarray(t_max, np.float32)
        out = np.empty(len(o), np.float32)
        L.check(self._lib.rayn_b200_kat_sdf_hit(self._ctx, C.byref(hitable), C.byref(consts), len(o), _fptr(o), _fptr(d), _fptr(tm),
                                                thr_scale, thr_const, _fptr(out)), self._ctx)
        return out

    def kat_occluded(self, start, end):
        s = np.ascontiguousarray(start, np.float32).reshape(-1, 3)
        e = np.ascontiguousarray(end, np.float32).reshape(-1, 3)
        out = np.empty(len(s), np.float32)
        L.check(self._lib.rayn_b200_kat_occluded(self._ctx, len(s), _fptr(s), _fptr(e), _fptr(out)), self._ctx)
        return out

    def kat_closest_hit(self, depth, origins, dirs):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        t = np.empty(len(o), np.float32)
        obj = np.empty(len(o), np.int32)
        L.check(self._lib.rayn_b200_kat_closest_hit(self._ctx, depth, len(o), _fptr(o), _fptr(d), _fptr(t),
                                                    obj.ctypes.data_as(C.POINTER(C.c_int32))), self._ctx)
        return t, obj

    def kat_light_sample(self, light, s0, s1, p):
        s0, s1 = np.ascontiguousarray(s0, np.float32), np.ascontiguousarray(s1, np.float32)
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        pt, pdf = np.empty_like(p), np.empty(len(p), np.float32)
        L.check(self._lib.rayn_b200_kat_light_sample(self._ctx, C.byref(light), len(p), _fptr(s0), _fptr(s1), _fptr(p), _fptr(pt), _fptr(pdf)), self._ctx)
        return pt, pdf

    def kat_light_sample_volume(self, light, sample, o, d, t_max):
        sample, t_max = np.ascontiguousarray(sample, np.float32), np.ascontiguousarray(t_max, np.float32)
        o, d = np.ascontiguousarray(o, np.float32).reshape(-1, 3), np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        t, pdf = np.empty(len(o), np.float32), np.empty(len(o), np.float32)
        L.check(self._lib.rayn_b200_kat_light_sample_volume(self._ctx, C.byref(light), len(o), _fptr(sample), _fptr(o), _fptr(d), _fptr(t_max),
                                                            _fptr(t), _fptr(pdf)), self._ctx)
        return t, pdf

    def kat_bsdf(self, mat, normals, wo, s1d, u4):
        n3, w3 = np.ascontiguousarray(normals, np.float32).reshape(-1, 3), np.ascontiguousarray(wo, np.float32).reshape(-1, 3)
        s1d, u4 = np.ascontiguousarray(s1d, np.float32), np.ascontiguousarray(u4, np.float32).reshape(-1, 4)
        wi, f, fe, pdf = np.empty_like(n3), np.empty_like(n3), np.empty_like(n3), np.empty(len(n3), np.float32)
        L.check(self._lib.rayn_b200_kat_bsdf(self._ctx, C.byref(mat), len(n3), _fptr(n3), _fptr(w3), _fptr(s1d), _fptr(u4), _fptr(wi), _fptr(f),
                                             _fptr(pdf), _fptr(fe)), self._ctx)
        return wi, f, pdf, fe

    def sdf_variant(self, hitable_index):
        """march-kernel specialisation upload_scene selected for a hitable (include/rayn_b200.h: rayn_b200_debug_sdf_variant)"""
        return int(self._lib.rayn_b200_debug_sdf_variant(self._ctx, hitable_index))

    def enable_queue_log(self, on=True):
        L.check(self._lib.rayn_b200_debug_enable_queue_log(self._ctx, 1 if on else 0), self._ctx)

    def read_queue_log(self):
        n = self._lib.rayn_b200_debug_read_queue_log(self._ctx, None, 0)
        out = np.empty(max(n, 1), np.int32)
        self._lib.rayn_b200_debug_read_queue_log(self._ctx, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]


class Film:
    """film.rs:175-203.  Channel planes are numpy arrays, row-major, y up, already / spp."""

    def __init__(self, channels, res, device=0):
        if len(set(channels)) != len(channels):
            raise ValueError("Attempted to create multiple channels of one kind")  # film.rs:187-189
        for c in channels:
            if c not in CHANNELS:
                raise ValueError(f"unknown channel {c}")
        self.channel_kinds = tuple(channels)
        self.res = (int(res[0]), int(res[1]))
        self.channels = {}
        self.progressive_epoch = 0
        self._renderer = None
        self._device = device
        self.last_stats = None

    def render_frame_into(self, world, camera, integrator, filt, tile_size, frame, time_range, samples):
        """Drop-in for film.rs:382-395.  time_range = (start, end)."""
        if self._renderer is None:
            self._renderer = Renderer(self._device)
        w, h = self.res
        inputs = FrameInputs(w, h, samples, integrator, filt, frame)
        self._renderer.upload_scene(world, camera)
        planes = self._renderer.render_host(inputs, tile_size, integrator, time_range)
        self.last_stats = self._renderer.stats()
        for k in self.channel_kinds:
            self.channels[k] = planes[k].reshape((h, w, 3) if k != "alpha" else (h, w))
        self.progressive_epoch += 1  # film.rs:657

    def save_to(self, write_channels, output_folder, base_name, transparent_background=False):
        """film.rs:205-377.  Same channel semantics and file names as the reference; the pixel arithmetic runs on the
        device (`rayn_b200_film_postprocess`), the PNG encoding stays host I/O (PIL)."""
        import os
        from PIL import Image
        os.makedirs(output_folder, exist_ok=True)
        w, h = self.res
        flat = {k: np.ascontiguousarray(v, np.float32).reshape(-1) for k, v in self.channels.items()}
        written = []
        for kind in write_channels:
            if kind == "color":
                if transparent_background and "color" in flat and "alpha" in flat:
                    mode, pil = L.POST_COLOR_ALPHA, "RGBA"
                elif not transparent_background and "color" in flat and "background" in flat:
                    mode, pil = L.POST_COLOR_PLUS_BACKGROUND, "RGB"
                elif not transparent_background and "color" in flat:
                    mode, pil = L.POST_COLOR_ONLY, "RGB"
                else:
                    raise ValueError("Attempted to write Color channel with insufficient channels")  # film.rs:294-298
            elif kind in ("background", "normal", "alpha"):
                if kind not in flat:
                    raise ValueError(f"Attempted to write {kind} channel but it didn't exist")
                mode, pil = {"background": (L.POST_BACKGROUND, "RGB"), "normal": (L.POST_WORLD_NORMAL, "RGB"), "alpha": (L.POST_ALPHA, "L")}[kind]
            else:
                raise ValueError(kind)
            px = self._renderer.postprocess(mode, w, h, flat)
            path = os.path.join(output_folder, f"{base_name}_{kind}.png")
            Image.fromarray(px[:, :, 0] if pil == "L" else px, pil).save(path)
            written.append(path)
        return written

    def tonemapped_rgb8(self):
        """The display formula of save_to (film.rs:253-267): (color + background).saturated().gamma(2.2), y flipped."""
        col = self.channels["color"] + self.channels.get("background", 0.0)
        rgb = np.clip(col, 0.0, 1.0) ** (1.0 / 2.2)
        return (np.clip(rgb * 255.0, 0, 255).astype(np.uint8))[::-1]
