// api.cu — C-ABI implementation (include/rayn_b200.h): context, scene upload, the tile-pass
// scheduler that drives the wavefront kernels, film gather helpers and the known-answer
// entry points.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rt_kernels.cuh"

using namespace rt;

static thread_local std::string g_last_error;

struct TimedLaunch {
  int kernel;
  cudaEvent_t a, b;
};

struct RaynContext {
  int device = 0;
  int flags = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  bool has_scene = false;
  DevScene scene;
  int64_t cap_paths = 0;  // requested paths per pass
  // pass buffers
  int64_t alloc_paths = 0, alloc_q = 0, alloc_seg = 0;
  int alloc_lc_ns = 0;
  int alloc_tiles = 0;
  PassBufs pb;
  int* d_tile_ids = nullptr;
  int* d_batch_prefix = nullptr;  // [alloc_tiles + 1]
  int* d_work_ctr = nullptr;      // [4] global work counters of the persistent kernels
  int n_sm = 148;
  // staging for host-space inputs / outputs
  float *d_s1 = nullptr, *d_s2 = nullptr, *d_scr = nullptr, *d_fis = nullptr;
  size_t cap_s1 = 0, cap_s2 = 0, cap_scr = 0;
  float* d_planes = nullptr;
  size_t cap_planes = 0;
  int* d_pack_ids = nullptr;
  size_t cap_pack_ids = 0;
  unsigned char* d_post = nullptr;
  size_t cap_post = 0;
  RaynStats stats;
  bool qlog_enabled = false;
  std::vector<int32_t> qlog;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<TimedLaunch> timed;
  size_t timed_used = 0;
};

static int32_t fail(RaynContext* ctx, int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
  return code;
}
#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? RAYN_ERR_OOM : RAYN_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                       \
  } while (0)

template <class T>
static cudaError_t regrow(T** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(need, 1) * sizeof(T));
  if (e == cudaSuccess) *cap = need;
  return e;
}

static void free_pass(RaynContext* c) {
  PassBufs& p = c->pb;
  cudaFree(p.o_time), cudaFree(p.d_t), cudaFree(p.rad), cudaFree(p.thr), cudaFree(p.nrm0), cudaFree(p.term);
  cudaFree(p.q_live), cudaFree(p.q_key), cudaFree(p.q_shade), cudaFree(p.n_live), cudaFree(p.n_slots), cudaFree(p.bin_start);
  cudaFree(c->d_tile_ids);
  cudaFree(c->d_batch_prefix);
  c->d_batch_prefix = nullptr;
  cudaFree(p.nrm), cudaFree(p.vis), cudaFree(p.seg_a), cudaFree(p.seg_b), cudaFree(p.seg_owner), cudaFree(p.lc_c), cudaFree(p.lc_t);
  unsigned long long* counters = p.counters;
  memset(&p, 0, sizeof p);
  p.counters = counters;
  c->d_tile_ids = nullptr;
  c->alloc_paths = c->alloc_q = c->alloc_seg = 0;
  c->alloc_lc_ns = 0;
  c->alloc_tiles = 0;
}

static int32_t ensure_pass(RaynContext* ctx, int n_tiles, int R, int QS, int seg_per_path, int lc_ns) {
  const int64_t need_paths = (int64_t)n_tiles * R, need_q = (int64_t)n_tiles * QS;
  const int64_t need_seg = need_paths * seg_per_path;
  if (need_paths <= ctx->alloc_paths && need_q <= ctx->alloc_q && n_tiles <= ctx->alloc_tiles && need_seg <= ctx->alloc_seg && lc_ns <= ctx->alloc_lc_ns)
    return RAYN_OK;
  free_pass(ctx);
  PassBufs& p = ctx->pb;
  CU(cudaMalloc(&p.o_time, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.d_t, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.rad, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.thr, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.nrm0, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.term, need_paths * sizeof(uint32_t)));
  CU(cudaMalloc(&p.q_live, need_paths * sizeof(int)));
  CU(cudaMalloc(&p.q_key, need_paths * sizeof(int)));
  CU(cudaMalloc(&p.q_shade, need_q * sizeof(int)));
  CU(cudaMalloc(&p.n_live, n_tiles * sizeof(int)));
  CU(cudaMalloc(&p.n_slots, n_tiles * sizeof(int)));
  CU(cudaMalloc(&p.bin_start, (size_t)n_tiles * (RAYN_MAX_HITABLES + 1) * sizeof(int)));
  CU(cudaMalloc(&ctx->d_tile_ids, n_tiles * sizeof(int)));
  CU(cudaMalloc(&ctx->d_batch_prefix, ((size_t)n_tiles + 1) * sizeof(int)));
  CU(cudaMalloc(&p.nrm, need_paths * sizeof(float4)));
  CU(cudaMalloc(&p.vis, need_paths * sizeof(uint32_t)));
  if (need_seg > 0) {
    CU(cudaMalloc(&p.seg_a, need_seg * sizeof(float4)));
    CU(cudaMalloc(&p.seg_b, need_seg * sizeof(float4)));
    CU(cudaMalloc(&p.seg_owner, need_seg * sizeof(int)));
  }
  p.seg_cap = need_seg;
  ctx->alloc_seg = need_seg;
  if (lc_ns > 0) {
    CU(cudaMalloc(&p.lc_c, need_paths * lc_ns * sizeof(float4)));
    if (lc_ns > 4) CU(cudaMalloc(&p.lc_t, need_paths * 8 * sizeof(float)));
  }
  ctx->alloc_lc_ns = lc_ns;
  ctx->alloc_paths = need_paths;
  ctx->alloc_q = need_q;
  ctx->alloc_tiles = n_tiles;
  return RAYN_OK;
}

static void timed_begin(RaynContext* ctx, int kernel) {
  if (!(ctx->flags & RAYN_FLAG_TIMING)) return;
  if (ctx->timed_used == ctx->timed.size()) {
    TimedLaunch t;
    t.kernel = kernel;
    cudaEventCreate(&t.a);
    cudaEventCreate(&t.b);
    ctx->timed.push_back(t);
  }
  ctx->timed[ctx->timed_used].kernel = kernel;
  cudaEventRecord(ctx->timed[ctx->timed_used].a, ctx->stream);
}
static void timed_end(RaynContext* ctx, int kernel) {
  ctx->stats.launches++;
  ctx->stats.kernel_launches[kernel]++;
  if (!(ctx->flags & RAYN_FLAG_TIMING)) return;
  cudaEventRecord(ctx->timed[ctx->timed_used].b, ctx->stream);
  ctx->timed_used++;
}

struct DevTmp {
  std::vector<void*> ptrs;
  ~DevTmp() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <class T>
  T* up(const T* h, size_t n, cudaError_t* e) {
    T* d = nullptr;
    if (*e != cudaSuccess) return nullptr;
    *e = cudaMalloc((void**)&d, std::max<size_t>(n, 1) * sizeof(T));
    if (*e != cudaSuccess) return nullptr;
    ptrs.push_back(d);
    if (h) *e = cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice);
    return d;
  }
};

extern "C" {

int32_t rayn_b200_abi_version(void) { return RAYN_B200_ABI_VERSION; }

const char* rayn_b200_last_error(const RaynContext* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int32_t rayn_b200_create(const RaynConfig* cfg, RaynContext** out_ctx) {
  RaynContext* ctx = nullptr;
  if (!out_ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "out_ctx is NULL");
  *out_ctx = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, RAYN_ERR_NO_DEVICE, "no CUDA device: the rayn_b200 render path has no CPU fallback");
  }
  const int dev = cfg ? cfg->device : 0;
  if (dev < 0 || dev >= ndev) return fail(nullptr, RAYN_ERR_INVALID_ARG, "device %d out of range (have %d)", dev, ndev);
  CU(cudaSetDevice(dev));
  ctx = new RaynContext();
  ctx->device = dev;
  ctx->flags = cfg ? cfg->flags : 0;
  ctx->cap_paths = (cfg && cfg->max_paths_per_pass > 0) ? cfg->max_paths_per_pass : (int64_t)96 << 20;  // ~25 GB of path state + shadow queue; fewer passes = fewer kernel tails (measured +3.5 %)
  memset(&ctx->pb, 0, sizeof ctx->pb);
  memset(&ctx->stats, 0, sizeof ctx->stats);
  memset(&ctx->scene, 0, sizeof ctx->scene);
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->pb.counters, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fis, RAYN_FIS_TABLE_SIZE * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_work_ctr, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  if (e != cudaSuccess) {
    fail(nullptr, RAYN_ERR_CUDA, "context setup: %s", cudaGetErrorString(e));
    delete ctx;
    return RAYN_ERR_CUDA;
  }
  *out_ctx = ctx;
  return RAYN_OK;
}

void rayn_b200_destroy(RaynContext* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_pass(ctx);
  cudaFree(ctx->pb.counters);
  cudaFree(ctx->d_work_ctr);
  cudaFree(ctx->d_pack_ids);
  cudaFree(ctx->d_post);
  cudaFree(ctx->d_s1), cudaFree(ctx->d_s2), cudaFree(ctx->d_scr), cudaFree(ctx->d_fis), cudaFree(ctx->d_planes);
  for (auto& t : ctx->timed) cudaEventDestroy(t.a), cudaEventDestroy(t.b);
  cudaEventDestroy(ctx->ev0), cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int32_t validate_scene(RaynContext* ctx, const RaynSceneDesc* s) {
  if (!s) return fail(ctx, RAYN_ERR_INVALID_ARG, "scene is NULL");
  if (s->n_hitables < 1 || s->n_hitables > RAYN_MAX_HITABLES)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "n_hitables %d not in [1,%d]", s->n_hitables, RAYN_MAX_HITABLES);
  if (s->n_materials < 1 || s->n_materials > RAYN_MAX_MATERIALS)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "n_materials %d not in [1,%d]", s->n_materials, RAYN_MAX_MATERIALS);
  if (s->n_lights < 0 || s->n_lights > RAYN_MAX_LIGHTS)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "n_lights %d not in [0,%d]", s->n_lights, RAYN_MAX_LIGHTS);
  if (!s->hitables || !s->materials || (s->n_lights && !s->lights)) return fail(ctx, RAYN_ERR_INVALID_ARG, "NULL scene array");
  for (int i = 0; i < s->n_hitables; ++i) {
    const RaynHitable& h = s->hitables[i];
    if (h.kind < 0 || h.kind > RAYN_HITABLE_MANDELBULB) return fail(ctx, RAYN_ERR_INVALID_ARG, "hitable %d: bad kind %d", i, h.kind);
    if (h.material < 0 || h.material >= s->n_materials)
      return fail(ctx, RAYN_ERR_INVALID_ARG, "hitable %d: material %d out of range", i, h.material);
    if (h.kind != RAYN_HITABLE_SPHERE && (h.iterations < 0 || h.iterations > 1024))
      return fail(ctx, RAYN_ERR_INVALID_ARG, "hitable %d: iterations %d", i, h.iterations);
    if (h.kind == RAYN_HITABLE_MANDELBULB && h.bulb_power != 8)
      return fail(ctx, RAYN_ERR_UNSUPPORTED, "hitable %d: Mandelbulb power %d (only 8 is built)", i, h.bulb_power);
  }
  for (int i = 0; i < s->n_materials; ++i)
    if (s->materials[i].kind < 0 || s->materials[i].kind > RAYN_MATERIAL_EMISSIVE)
      return fail(ctx, RAYN_ERR_INVALID_ARG, "material %d: bad kind %d", i, s->materials[i].kind);
  if (s->camera.kind < 0 || s->camera.kind > RAYN_CAMERA_ORTHOGRAPHIC) return fail(ctx, RAYN_ERR_INVALID_ARG, "bad camera kind");
  if (s->consts.max_marches < 1 || s->consts.max_vis_marches < 1) return fail(ctx, RAYN_ERR_INVALID_ARG, "march limits must be >= 1");
  return RAYN_OK;
}

int32_t rayn_b200_upload_scene(RaynContext* ctx, const RaynSceneDesc* s) {
  if (!ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "ctx is NULL");
  int32_t rc = validate_scene(ctx, s);
  if (rc) return rc;
  DevScene& d = ctx->scene;
  memset(&d, 0, sizeof d);
  d.n_hit = s->n_hitables;
  d.n_mat = s->n_materials;
  d.n_lights = s->n_lights;
  memcpy(d.hit, s->hitables, sizeof(RaynHitable) * s->n_hitables);
  memcpy(d.mat, s->materials, sizeof(RaynMaterial) * s->n_materials);
  if (s->n_lights) memcpy(d.light, s->lights, sizeof(RaynLight) * s->n_lights);
  d.cam = s->camera;
  d.vol = s->volume;
  d.rc = s->consts;
  ctx->has_scene = true;
  return RAYN_OK;
}

int32_t rayn_b200_get_stats(const RaynContext* ctx, RaynStats* out) {
  if (!ctx || !out) return RAYN_ERR_INVALID_ARG;
  *out = ctx->stats;
  return RAYN_OK;
}

int32_t rayn_b200_debug_enable_queue_log(RaynContext* ctx, int32_t enable) {
  if (!ctx) return RAYN_ERR_INVALID_ARG;
  ctx->qlog_enabled = enable != 0;
  ctx->qlog.clear();
  return RAYN_OK;
}
int64_t rayn_b200_debug_read_queue_log(RaynContext* ctx, int32_t* out, int64_t cap) {
  if (!ctx) return -1;
  const int64_t n = (int64_t)ctx->qlog.size();
  if (out && cap > 0) memcpy(out, ctx->qlog.data(), sizeof(int32_t) * (size_t)std::min(n, cap));
  return n;
}

int32_t rayn_b200_render_frame(RaynContext* ctx, const RaynFrameDesc* f, const RaynFilmPlanes* out) {
  if (!ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "ctx is NULL");
  if (!ctx->has_scene) return fail(ctx, RAYN_ERR_NO_SCENE, "render_frame before upload_scene");
  if (!f || !out) return fail(ctx, RAYN_ERR_INVALID_ARG, "frame/out is NULL");
  if (f->width <= 0 || f->height <= 0 || f->tile_w <= 0 || f->tile_h <= 0 || f->samples <= 0 || f->max_bounces < 0)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "bad frame geometry");
  if (f->volume_marches != 2)
    return fail(ctx, RAYN_ERR_UNSUPPORTED, "volume_marches = %d: the reference hard-wires samples_1d[3],[4] for vm = 2", f->volume_marches);
  const int spp = 4 * f->samples, vm = f->volume_marches, mb = f->max_bounces;
  const int need1 = 1 + (mb + 1) * (3 + vm), need2 = 2 + (mb + 1) * (12 + 8 * vm) / 2;
  if (f->sets_1d < need1 || f->sets_2d < need2)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "sample tables too small: have %d/%d sets, path needs %d/%d", f->sets_1d, f->sets_2d, need1, need2);
  if (!f->samples_1d || !f->samples_2d || !f->scramble || !f->fis_inverse_cdf) return fail(ctx, RAYN_ERR_INVALID_ARG, "NULL input table");
  if (!out->color || !out->alpha || !out->background || !out->normal) return fail(ctx, RAYN_ERR_INVALID_ARG, "NULL film plane");
  const int stride = f->tile_stride > 0 ? f->tile_stride : 1;
  if (f->tile_offset < 0 || f->tile_offset >= stride) return fail(ctx, RAYN_ERR_INVALID_ARG, "tile_offset %d not in [0,%d)", f->tile_offset, stride);
  if (mb > 1022) return fail(ctx, RAYN_ERR_UNSUPPORTED, "max_bounces > 1022");
  const int64_t R64 = (int64_t)f->tile_w * f->tile_h * spp;
  const int n_hit = ctx->scene.n_hit;
  if (R64 + 4 * n_hit >= (1 << 20))
    return fail(ctx, RAYN_ERR_UNSUPPORTED, "tile_w*tile_h*spp = %lld exceeds the 2^20 slot key space", (long long)R64);
  const int R = (int)R64, QS = R + 4 * n_hit;
  CU(cudaSetDevice(ctx->device));

  DevFrame fr;
  fr.W = f->width, fr.H = f->height, fr.tile_w = f->tile_w, fr.tile_h = f->tile_h;
  fr.samples = f->samples, fr.spp = spp, fr.max_bounces = mb, fr.vm = vm;
  fr.ntx = (f->width + f->width % f->tile_w) / f->tile_w;      // film.rs:399-404
  fr.nty = (f->height + f->height % f->tile_h) / f->tile_h;
  fr.sets_1d = f->sets_1d, fr.sets_2d = f->sets_2d;
  fr.t0 = f->t0, fr.t1 = f->t1;

  std::vector<int> my_tiles;
  if (f->tile_list) {
    if (f->n_tile_list < 0) return fail(ctx, RAYN_ERR_INVALID_ARG, "n_tile_list < 0");
    for (int i = 0; i < f->n_tile_list; ++i) {
      const int idx = f->tile_list[i];
      if (idx < 0 || idx >= fr.ntx * fr.nty || (i && idx <= f->tile_list[i - 1]))
        return fail(ctx, RAYN_ERR_INVALID_ARG, "tile_list must be ascending tile indices in [0,%d)", fr.ntx * fr.nty);
      my_tiles.push_back(idx);
    }
  } else {
    for (int idx = f->tile_offset; idx < fr.ntx * fr.nty; idx += stride) my_tiles.push_back(idx);
  }

  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->timed_used = 0;
  ctx->qlog.clear();
  cudaStream_t st = ctx->stream;

  // Device-space pointers may have been produced on another stream (e.g. torch's): fence.
  if (f->input_space == RAYN_MEM_DEVICE || out->space == RAYN_MEM_DEVICE) CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(ctx->ev0, st));

  const size_t n1 = (size_t)spp * f->sets_1d, n2 = (size_t)2 * spp * f->sets_2d, npx = (size_t)f->width * f->height;
  if (f->input_space == RAYN_MEM_HOST) {
    CU(regrow(&ctx->d_s1, &ctx->cap_s1, n1));
    CU(regrow(&ctx->d_s2, &ctx->cap_s2, n2));
    CU(regrow(&ctx->d_scr, &ctx->cap_scr, npx));
    CU(cudaMemcpyAsync(ctx->d_s1, f->samples_1d, n1 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_s2, f->samples_2d, n2 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_scr, f->scramble, npx * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_fis, f->fis_inverse_cdf, RAYN_FIS_TABLE_SIZE * 4, cudaMemcpyHostToDevice, st));
    fr.s1 = ctx->d_s1, fr.s2 = ctx->d_s2, fr.scramble = ctx->d_scr, fr.fis = ctx->d_fis;
  } else {
    fr.s1 = f->samples_1d, fr.s2 = f->samples_2d, fr.scramble = f->scramble, fr.fis = f->fis_inverse_cdf;
  }
  float *p_color, *p_alpha, *p_bg, *p_normal;
  if (out->space == RAYN_MEM_HOST) {
    CU(regrow(&ctx->d_planes, &ctx->cap_planes, npx * 10));
    CU(cudaMemsetAsync(ctx->d_planes, 0, npx * 10 * 4, st));
    p_color = ctx->d_planes, p_alpha = p_color + 3 * npx, p_bg = p_alpha + npx, p_normal = p_bg + 3 * npx;
  } else {
    p_color = out->color, p_alpha = out->alpha, p_bg = out->background, p_normal = out->normal;
  }

  int tiles_per_pass = (int)std::max<int64_t>(1, ctx->cap_paths / R);
  tiles_per_pass = std::min(tiles_per_pass, 65535);
  tiles_per_pass = std::min<int>(tiles_per_pass, std::max<size_t>(my_tiles.size(), 1));
  int n_sdf = 0;
  for (int i = 0; i < n_hit; ++i) n_sdf += ctx->scene.hit[i].kind != RAYN_HITABLE_SPHERE;
  // kernel family: v3 (default) pass-wide persistent march kernels; v2 per-block pools; v0 one thread per ray
  const bool simple = (ctx->flags & RAYN_FLAG_SIMPLE_MARCH) != 0;
  const bool block_pool = !simple && (ctx->flags & RAYN_FLAG_BLOCK_POOL) != 0 && n_sdf <= SH_MAX_SDF;
  const bool v3 = !simple && !block_pool;
  bool motion = false;  // time-varying sphere centres need the packet's lane-0 time: only the default kernel family plumbs it
  for (int i = 0; i < n_hit; ++i)
    motion |= ctx->scene.hit[i].kind == RAYN_HITABLE_SPHERE && (ctx->scene.hit[i].center_velocity[0] != 0.0f || ctx->scene.hit[i].center_velocity[1] != 0.0f ||
                                                                 ctx->scene.hit[i].center_velocity[2] != 0.0f);
  if (motion && !v3) return fail(ctx, RAYN_ERR_UNSUPPORTED, "time-varying sphere centres are only supported by the default kernel family (no RAYN_FLAG_SIMPLE_MARCH / BLOCK_POOL)");
  const bool no_flat = (ctx->flags & RAYN_FLAG_FLATTEN) == 0;
  // persistent kernels: exactly as many CTAs as can be resident (one wave), so every CTA pulls work until the pass is drained
  int occ_ext = 8, occ_ext_flat = 8, occ_shd = 8, occ_shd_flat = 8;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ext, k_extend_march<false>, EXT_T, 0));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ext_flat, k_extend_march<true>, EXT_T, 0));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_shd, k_shadow<false>, SHD_T, 0));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_shd_flat, k_shadow<true>, SHD_T, 0));
  bool any_bulb = false;
  for (int i = 0; i < n_hit; ++i) any_bulb |= ctx->scene.hit[i].kind == RAYN_HITABLE_MANDELBULB;
  const bool volume_on = ctx->scene.vol.has_scattering != 0 && ctx->scene.n_lights > 0;
  const int seg_per_path = v3 ? (volume_on ? 4 * (1 + vm) : 4) * n_sdf : 0;  // worst case shadow segments per path per depth
  if (v3 && volume_on) tiles_per_pass = std::max(1, tiles_per_pass / 3);
  const int lc_ns = v3 ? (volume_on ? 4 * (1 + vm) : 4) : 0;  // stored light contributions per path per depth
  int32_t rc = ensure_pass(ctx, tiles_per_pass, R, QS, seg_per_path, lc_ns);
  if (rc) return rc;
  PassBufs pb = ctx->pb;
  pb.R = R, pb.QS = QS, pb.tile_ids = ctx->d_tile_ids;
  pb.lc_ns = lc_ns;
  pb.seg_count = ctx->d_work_ctr + 2;
  CU(cudaMemsetAsync(pb.counters, 0, 8 * sizeof(unsigned long long), st));
  int np = 2;
  while (np < spp) np <<= 1;
  CU(cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resolve_smem_bytes(np)));
  if (block_pool) CU(cudaFuncSetAttribute(k_shade2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem_bytes(n_sdf)));

  std::vector<int> h_nslots, h_slots;
  for (size_t first = 0; first < my_tiles.size(); first += tiles_per_pass) {
    const int nt = (int)std::min<size_t>(tiles_per_pass, my_tiles.size() - first);
    pb.n_tiles = nt;
    CU(cudaMemcpyAsync(ctx->d_tile_ids, my_tiles.data() + first, nt * sizeof(int), cudaMemcpyHostToDevice, st));
    ctx->stats.passes++;
    const dim3 g_paths((R + 255) / 256, nt), g_ext((R + 127) / 128, nt), g_shade((QS + 127) / 128, nt);
    timed_begin(ctx, RAYN_K_RAYGEN);
    k_raygen<<<g_paths, 256, 0, st>>>(ctx->scene, fr, pb);
    timed_end(ctx, RAYN_K_RAYGEN);
    for (int depth = 0; depth <= mb; ++depth) {
      const Thr thr = make_thr(ctx->scene.cam, depth);
      timed_begin(ctx, RAYN_K_EXTEND);
      if (simple) {
        k_extend<<<g_ext, 128, 0, st>>>(ctx->scene, pb, thr);
      } else if (block_pool) {
        k_extend2<<<dim3((R + EXT_CHUNK - 1) / EXT_CHUNK, nt), EXT_T, 0, st>>>(ctx->scene, pb, thr);
      } else {
        k_scan_live<<<1, SCAN_T, 0, st>>>(pb, ctx->d_batch_prefix, ctx->d_work_ctr);
        ctx->stats.launches++;
        ctx->stats.kernel_launches[RAYN_K_MISC]++;
        // fold order of hitable.rs:177-198: runs of spheres as coherent kernels, each SDF as a persistent march
        int k = 0, first_kernel = 1, n_march = 0;
        while (k < n_hit || first_kernel) {
          int e = k;
          while (e < n_hit && ctx->scene.hit[e].kind == RAYN_HITABLE_SPHERE) ++e;
          if (e > k || first_kernel) {
            k_extend_spheres<<<g_paths, 256, 0, st>>>(ctx->scene, pb, k, e, first_kernel, motion ? 1 : 0);
            ctx->stats.launches++;
            first_kernel = 0;
          }
          if (e < n_hit) {
            if (n_march++ > 0) CU(cudaMemsetAsync(ctx->d_work_ctr, 0, sizeof(int), st));
            if (ctx->scene.hit[e].kind == RAYN_HITABLE_MANDELBULB && !no_flat)
              k_extend_march<true><<<ctx->n_sm * occ_ext_flat, EXT_T, 0, st>>>(ctx->scene, pb, thr, e, ctx->d_batch_prefix, ctx->d_work_ctr);
            else
              k_extend_march<false><<<ctx->n_sm * occ_ext, EXT_T, 0, st>>>(ctx->scene, pb, thr, e, ctx->d_batch_prefix, ctx->d_work_ctr);
            ctx->stats.launches++;
            ++e;
          }
          k = e;
        }
        ctx->stats.launches--;  // timed_end below counts one launch of this group
      }
      timed_end(ctx, RAYN_K_EXTEND);
      timed_begin(ctx, RAYN_K_BIN);
      k_bin<<<nt, BIN_T, 0, st>>>(pb, n_hit);
      timed_end(ctx, RAYN_K_BIN);
      if (ctx->qlog_enabled) {
        h_nslots.resize(nt);
        h_slots.resize((size_t)nt * QS);
        CU(cudaMemcpyAsync(h_nslots.data(), pb.n_slots, nt * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_slots.data(), pb.q_shade, (size_t)nt * QS * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int t = 0; t < nt; ++t) {
          ctx->qlog.push_back(depth);
          ctx->qlog.push_back(my_tiles[first + t]);
          ctx->qlog.push_back(h_nslots[t]);
          for (int s = 0; s < h_nslots[t]; ++s) ctx->qlog.push_back(h_slots[(size_t)t * QS + s]);
        }
      }
      if (v3) {
        timed_begin(ctx, RAYN_K_SHADE_PRE);
        k_shade_pre<<<g_shade, 128, 0, st>>>(ctx->scene, fr, pb, depth, thr);
        timed_end(ctx, RAYN_K_SHADE_PRE);
        if (n_sdf > 0 && ctx->scene.n_lights > 0) {
          timed_begin(ctx, RAYN_K_SHADOW);
          if (any_bulb && !no_flat)
            k_shadow<true><<<ctx->n_sm * occ_shd_flat, SHD_T, 0, st>>>(ctx->scene, pb, ctx->d_work_ctr + 1);
          else
            k_shadow<false><<<ctx->n_sm * occ_shd, SHD_T, 0, st>>>(ctx->scene, pb, ctx->d_work_ctr + 1);
          timed_end(ctx, RAYN_K_SHADOW);
        }
        timed_begin(ctx, RAYN_K_SHADE_POST);
        k_shade_post<<<g_shade, 128, 0, st>>>(ctx->scene, fr, pb, depth);
        timed_end(ctx, RAYN_K_SHADE_POST);
      } else {
        timed_begin(ctx, RAYN_K_SHADE_PRE);
        if (simple)
          k_shade<<<g_shade, 128, 0, st>>>(ctx->scene, fr, pb, depth, thr);
        else
          k_shade2<<<dim3((QS + SH_T - 1) / SH_T, nt), SH_T, shade_smem_bytes(n_sdf), st>>>(ctx->scene, fr, pb, depth, thr, SH_POOL * std::max(n_sdf, 1));
        timed_end(ctx, RAYN_K_SHADE_PRE);
      }
      if (depth < mb) {
        timed_begin(ctx, RAYN_K_COMPACT);
        k_compact<<<nt, CMP_T, 0, st>>>(pb);
        timed_end(ctx, RAYN_K_COMPACT);
      }
    }
    timed_begin(ctx, RAYN_K_RESOLVE);
    k_resolve<<<dim3(f->tile_w * f->tile_h, nt), RES_T, resolve_smem_bytes(np), st>>>(fr, pb, p_color, p_alpha, p_bg, p_normal, np);
    timed_end(ctx, RAYN_K_RESOLVE);
    CU(cudaGetLastError());
  }
  if (out->space == RAYN_MEM_HOST) {
    CU(cudaMemcpyAsync(out->color, p_color, npx * 3 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->alpha, p_alpha, npx * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->background, p_bg, npx * 3 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->normal, p_normal, npx * 3 * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaEventRecord(ctx->ev1, st));
  unsigned long long h_counters[8];
  CU(cudaMemcpyAsync(h_counters, pb.counters, sizeof h_counters, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&ctx->stats.total_ms, ctx->ev0, ctx->ev1));
  ctx->stats.extend_rays = (int64_t)h_counters[CNT_EXTEND_RAYS];
  ctx->stats.shade_lanes = (int64_t)h_counters[CNT_SHADE_LANES];
  ctx->stats.shadow_rays = (int64_t)h_counters[CNT_SHADOW_RAYS];
  ctx->stats.sdf_evals_extend = (int64_t)h_counters[CNT_EVALS_EXTEND];
  ctx->stats.sdf_evals_shadow = (int64_t)h_counters[CNT_EVALS_SHADOW];
  {
    int64_t paths = 0;
    for (int idx : my_tiles) {
      const int tx = idx / fr.nty, ty = idx % fr.nty;
      const int tw = std::min(tx * f->tile_w + f->tile_w, f->width) - tx * f->tile_w;
      const int th = std::min(ty * f->tile_h + f->tile_h, f->height) - ty * f->tile_h;
      paths += (int64_t)tw * th * spp;
    }
    ctx->stats.paths = paths;
  }
  for (size_t i = 0; i < ctx->timed_used; ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->timed[i].a, ctx->timed[i].b);
    ctx->stats.kernel_ms[ctx->timed[i].kernel] += ms;
  }
  return RAYN_OK;
}

// ---- multi-GPU film gather helpers ------------------------------------------------------------------
int64_t rayn_b200_film_slab_floats(int32_t tw, int32_t th, int32_t n_tiles) {
  if (tw <= 0 || th <= 0 || n_tiles < 0) return -1;
  return (int64_t)n_tiles * 10 * tw * th;
}
static int32_t pack_unpack(RaynContext* ctx, int W, int H, int tw, int th, const int32_t* tile_list, int n, const RaynFilmPlanes* pl,
                           float* slab, int unpack) {
  if (!ctx || !pl || !slab || n < 0 || (n && !tile_list) || W <= 0 || H <= 0 || tw <= 0 || th <= 0)
    return fail(ctx, RAYN_ERR_INVALID_ARG, "film pack/unpack: bad argument");
  if (n == 0) return RAYN_OK;
  CU(cudaSetDevice(ctx->device));
  const int ntx = (W + W % tw) / tw, nty = (H + H % th) / th;
  for (int i = 0; i < n; ++i)
    if (tile_list[i] < 0 || tile_list[i] >= ntx * nty) return fail(ctx, RAYN_ERR_INVALID_ARG, "film pack/unpack: tile %d out of range", tile_list[i]);
  CU(regrow(&ctx->d_pack_ids, &ctx->cap_pack_ids, (size_t)n));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyAsync(ctx->d_pack_ids, tile_list, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_film_pack<<<n, 256, 0, ctx->stream>>>(W, H, tw, th, nty, ctx->d_pack_ids, pl->color, pl->alpha, pl->background, pl->normal, slab, unpack,
                                          pl->color, pl->alpha, pl->background, pl->normal);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return RAYN_OK;
}
int32_t rayn_b200_film_pack_tiles(RaynContext* ctx, int32_t W, int32_t H, int32_t tw, int32_t th, const int32_t* tile_list, int32_t n_tiles,
                                  const RaynFilmPlanes* planes_dev, float* slab_dev) {
  return pack_unpack(ctx, W, H, tw, th, tile_list, n_tiles, planes_dev, slab_dev, 0);
}
int32_t rayn_b200_film_unpack_tiles(RaynContext* ctx, int32_t W, int32_t H, int32_t tw, int32_t th, const int32_t* tile_list, int32_t n_tiles,
                                    const float* slab_dev, const RaynFilmPlanes* planes_dev) {
  return pack_unpack(ctx, W, H, tw, th, tile_list, n_tiles, planes_dev, const_cast<float*>(slab_dev), 1);
}

// ---- film post-process (film.rs:205-377 arithmetic) -------------------------------------------------
int32_t rayn_b200_film_postprocess(RaynContext* ctx, int32_t mode, int32_t W, int32_t H, const RaynFilmPlanes* pl, uint8_t* out,
                                   int32_t out_space) {
  if (!ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "ctx is NULL");
  if (mode < 0 || mode > RAYN_POST_ALPHA || W <= 0 || H <= 0 || !pl || !out) return fail(ctx, RAYN_ERR_INVALID_ARG, "film_postprocess: bad argument");
  const bool need_color = mode <= RAYN_POST_COLOR_ONLY, need_bg = mode == RAYN_POST_COLOR_PLUS_BACKGROUND || mode == RAYN_POST_BACKGROUND;
  const bool need_alpha = mode == RAYN_POST_COLOR_ALPHA || mode == RAYN_POST_ALPHA, need_normal = mode == RAYN_POST_WORLD_NORMAL;
  if ((need_color && !pl->color) || (need_bg && !pl->background) || (need_alpha && !pl->alpha) || (need_normal && !pl->normal))
    return fail(ctx, RAYN_ERR_INVALID_ARG, "Attempted to write a channel with insufficient channels");  // film.rs:294-298
  CU(cudaSetDevice(ctx->device));
  const size_t npx = (size_t)W * H, nbytes = npx * post_bytes_per_pixel(mode);
  cudaStream_t st = ctx->stream;
  const float *c = pl->color, *a = pl->alpha, *b = pl->background, *n = pl->normal;
  CU(cudaDeviceSynchronize());
  if (pl->space == RAYN_MEM_HOST) {
    CU(regrow(&ctx->d_planes, &ctx->cap_planes, npx * 10));
    float* d = ctx->d_planes;
    if (need_color) CU(cudaMemcpyAsync(d, pl->color, npx * 12, cudaMemcpyHostToDevice, st));
    if (need_alpha) CU(cudaMemcpyAsync(d + 3 * npx, pl->alpha, npx * 4, cudaMemcpyHostToDevice, st));
    if (need_bg) CU(cudaMemcpyAsync(d + 4 * npx, pl->background, npx * 12, cudaMemcpyHostToDevice, st));
    if (need_normal) CU(cudaMemcpyAsync(d + 7 * npx, pl->normal, npx * 12, cudaMemcpyHostToDevice, st));
    c = d, a = d + 3 * npx, b = d + 4 * npx, n = d + 7 * npx;
  }
  unsigned char* dout = out;
  if (out_space == RAYN_MEM_HOST) {
    CU(regrow(&ctx->d_post, &ctx->cap_post, nbytes));
    dout = ctx->d_post;
  }
  k_postprocess<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(mode, W, H, c, a, b, n, dout);
  CU(cudaGetLastError());
  if (out_space == RAYN_MEM_HOST) CU(cudaMemcpyAsync(out, dout, nbytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return RAYN_OK;
}

int32_t rayn_b200_device_frame_inputs(RaynContext* ctx, int32_t W, int32_t H, int32_t spp, int32_t sets_1d, int32_t sets_2d, uint64_t offset,
                                      float* s1, float* s2, float* scramble) {
  if (!ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "ctx is NULL");
  if (spp <= 0 || sets_1d < 0 || sets_2d < 0 || (sets_1d && !s1) || (sets_2d && !s2) || (scramble && (W <= 0 || H <= 0)))
    return fail(ctx, RAYN_ERR_INVALID_ARG, "device_frame_inputs: bad argument");
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  const long long n = (long long)spp * (sets_1d + sets_2d);
  if (n > 0) k_gen_rd_tables<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(spp, sets_1d, sets_2d, offset, s1, s2);
  if (scramble) k_gen_scramble<<<(unsigned)(((long long)W * H + 255) / 256), 256, 0, ctx->stream>>>(W, H, scramble);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return RAYN_OK;
}

// ---- known-answer entry points ----------------------------------------------------------------------
#define KAT_PROLOGUE                                                     \
  if (!ctx) return fail(nullptr, RAYN_ERR_INVALID_ARG, "ctx is NULL");   \
  if (n < 0) return fail(ctx, RAYN_ERR_INVALID_ARG, "n < 0");            \
  CU(cudaSetDevice(ctx->device));                                        \
  if (n == 0) return RAYN_OK;                                            \
  DevTmp tmp;                                                            \
  cudaError_t e = cudaSuccess;                                           \
  const unsigned blocks = (unsigned)((n + 127) / 128);
#define KAT_EPILOGUE(dst, src, count, T)                                              \
  CU(e);                                                                              \
  CU(cudaGetLastError());                                                             \
  CU(cudaStreamSynchronize(ctx->stream));                                             \
  CU(cudaMemcpy(dst, src, (size_t)(count) * sizeof(T), cudaMemcpyDeviceToHost));

int32_t rayn_b200_kat_detmath(RaynContext* ctx, int32_t op, int64_t n, const float* a, const float* b, float* out) {
  KAT_PROLOGUE
  if (op < 0 || op > 7 || !a || !out) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_detmath: bad argument");
  float* da = tmp.up(a, n, &e);
  float* db = tmp.up(b ? b : a, n, &e);
  float* dout = tmp.up<float>(nullptr, n, &e);
  CU(e);
  k_kat_detmath<<<blocks, 128, 0, ctx->stream>>>(op, n, da, db, dout);
  KAT_EPILOGUE(out, dout, n, float)
  return RAYN_OK;
}
int32_t rayn_b200_kat_sdf_dist(RaynContext* ctx, const RaynHitable* sdf, int64_t n, const float* points3, float* out) {
  KAT_PROLOGUE
  if (!sdf || !points3 || !out) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_sdf_dist: NULL");
  float* dp = tmp.up(points3, 3 * n, &e);
  float* dout = tmp.up<float>(nullptr, n, &e);
  CU(e);
  k_kat_sdf_dist<<<blocks, 128, 0, ctx->stream>>>(*sdf, n, dp, dout);
  KAT_EPILOGUE(out, dout, n, float)
  return RAYN_OK;
}
int32_t rayn_b200_kat_sdf_hit(RaynContext* ctx, const RaynHitable* sdf, const RaynRenderConsts* consts, int64_t n, const float* origins3,
                              const float* dirs3, const float* t_max, float thr_scale, int32_t thr_const, float* out_t) {
  KAT_PROLOGUE
  if (!sdf || !consts || !origins3 || !dirs3 || !t_max || !out_t) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_sdf_hit: NULL");
  float* dor = tmp.up(origins3, 3 * n, &e);
  float* ddi = tmp.up(dirs3, 3 * n, &e);
  float* dtm = tmp.up(t_max, n, &e);
  float* dout = tmp.up<float>(nullptr, n, &e);
  CU(e);
  Thr thr;
  thr.scale = thr_scale;
  thr.is_const = thr_const;
  k_kat_sdf_hit<<<blocks, 128, 0, ctx->stream>>>(*sdf, *consts, n, dor, ddi, dtm, thr, dout);
  KAT_EPILOGUE(out_t, dout, n, float)
  return RAYN_OK;
}
int32_t rayn_b200_kat_occluded(RaynContext* ctx, int64_t n, const float* start3, const float* end3, float* out) {
  KAT_PROLOGUE
  if (!ctx->has_scene) return fail(ctx, RAYN_ERR_NO_SCENE, "kat_occluded before upload_scene");
  if (!start3 || !end3 || !out) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_occluded: NULL");
  float* ds = tmp.up(start3, 3 * n, &e);
  float* de = tmp.up(end3, 3 * n, &e);
  float* dout = tmp.up<float>(nullptr, n, &e);
  CU(e);
  k_kat_occluded<<<blocks, 128, 0, ctx->stream>>>(ctx->scene, n, ds, de, dout);
  KAT_EPILOGUE(out, dout, n, float)
  return RAYN_OK;
}
int32_t rayn_b200_kat_closest_hit(RaynContext* ctx, int32_t depth, int64_t n, const float* origins3, const float* dirs3, float* out_t,
                                  int32_t* out_obj) {
  KAT_PROLOGUE
  if (!ctx->has_scene) return fail(ctx, RAYN_ERR_NO_SCENE, "kat_closest_hit before upload_scene");
  if (!origins3 || !dirs3 || !out_t || !out_obj) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_closest_hit: NULL");
  float* dor = tmp.up(origins3, 3 * n, &e);
  float* ddi = tmp.up(dirs3, 3 * n, &e);
  float* dt = tmp.up<float>(nullptr, n, &e);
  int* dobj = tmp.up<int>(nullptr, n, &e);
  CU(e);
  k_kat_closest_hit<<<blocks, 128, 0, ctx->stream>>>(ctx->scene, make_thr(ctx->scene.cam, depth), n, dor, ddi, dt, dobj);
  KAT_EPILOGUE(out_t, dt, n, float)
  CU(cudaMemcpy(out_obj, dobj, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return RAYN_OK;
}

int32_t rayn_b200_kat_light_sample(RaynContext* ctx, const RaynLight* light, int64_t n, const float* s0, const float* s1, const float* points3,
                                   float* out_point3, float* out_pdf) {
  KAT_PROLOGUE
  if (!light || !s0 || !s1 || !points3 || !out_point3 || !out_pdf) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_light_sample: NULL");
  float *d0 = tmp.up(s0, n, &e), *d1 = tmp.up(s1, n, &e), *dp = tmp.up(points3, 3 * n, &e);
  float *dpt = tmp.up<float>(nullptr, 3 * n, &e), *dpdf = tmp.up<float>(nullptr, n, &e);
  CU(e);
  k_kat_light_sample<<<blocks, 128, 0, ctx->stream>>>(*light, n, d0, d1, dp, dpt, dpdf);
  KAT_EPILOGUE(out_point3, dpt, 3 * n, float)
  CU(cudaMemcpy(out_pdf, dpdf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return RAYN_OK;
}
int32_t rayn_b200_kat_light_sample_volume(RaynContext* ctx, const RaynLight* light, int64_t n, const float* sample, const float* origins3,
                                          const float* dirs3, const float* t_max, float* out_t, float* out_pdf) {
  KAT_PROLOGUE
  if (!light || !sample || !origins3 || !dirs3 || !t_max || !out_t || !out_pdf) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_light_sample_volume: NULL");
  float *ds = tmp.up(sample, n, &e), *dor = tmp.up(origins3, 3 * n, &e), *ddi = tmp.up(dirs3, 3 * n, &e), *dtm = tmp.up(t_max, n, &e);
  float *dt = tmp.up<float>(nullptr, n, &e), *dpdf = tmp.up<float>(nullptr, n, &e);
  CU(e);
  k_kat_light_sample_volume<<<blocks, 128, 0, ctx->stream>>>(*light, n, ds, dor, ddi, dtm, dt, dpdf);
  KAT_EPILOGUE(out_t, dt, n, float)
  CU(cudaMemcpy(out_pdf, dpdf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return RAYN_OK;
}
int32_t rayn_b200_kat_bsdf(RaynContext* ctx, const RaynMaterial* mat, int64_t n, const float* normals3, const float* wo3, const float* s1d,
                           const float* u4, float* out_wi3, float* out_f3, float* out_pdf, float* out_feval3) {
  KAT_PROLOGUE
  if (!mat || !normals3 || !wo3 || !s1d || !u4 || !out_wi3 || !out_f3 || !out_pdf || !out_feval3) return fail(ctx, RAYN_ERR_INVALID_ARG, "kat_bsdf: NULL");
  float *dn = tmp.up(normals3, 3 * n, &e), *dw = tmp.up(wo3, 3 * n, &e), *ds = tmp.up(s1d, n, &e), *du = tmp.up(u4, 4 * n, &e);
  float *dwi = tmp.up<float>(nullptr, 3 * n, &e), *df = tmp.up<float>(nullptr, 3 * n, &e), *dpdf = tmp.up<float>(nullptr, n, &e),
        *dfe = tmp.up<float>(nullptr, 3 * n, &e);
  CU(e);
  k_kat_bsdf<<<blocks, 128, 0, ctx->stream>>>(*mat, n, dn, dw, ds, du, dwi, df, dpdf, dfe);
  KAT_EPILOGUE(out_wi3, dwi, 3 * n, float)
  CU(cudaMemcpy(out_f3, df, (size_t)3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out_pdf, dpdf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out_feval3, dfe, (size_t)3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  return RAYN_OK;
}

}  // extern "C"
