// B200 (sm_100a) differentiable Gaussian rasterizer: per-Gaussian kernels + C ABI (include/gsb200.h).
//
// Pipeline (one view), every kernel hand-written for sm_100a, no library sort/scan, no host round trip:
//   k_setup_cam        camera/pose constants -> one CamConst in HBM (read by every block)
//   k_preprocess       fused pose transform + activations + EWA projection + SH->RGB, 128-bit
//                      coalesced loads staged through shared memory; writes packed splat records and
//                      counts instances per Gaussian and per tile (lossless alpha < 1/255 tile cull)
//   k_tile_scan        (gs_bin.cu) exclusive scan of the per-tile counts -> segment offsets, R on device
//   k_scatter          (gs_bin.cu) (depth | id) keys appended to the tiles' segments
//   k_tile_sort        (gs_bin.cu) per-tile shared-memory sort + gather into contiguous per-tile slabs
//   k_blend_fwd2       (gs_blend.cu) one CTA per 16x16 tile, TMA-staged slab chunks, packed f32x2 math
//   k_blend_bwd2       (gs_blend.cu) back-to-front replay, 9 gradients per (warp, Gaussian) reduced in-warp
//   k_preprocess_bwd   analytic backward to the model's own tensors + in-kernel pose-gradient
//                      reduction; k_pose_finalize chains it to dL/dP[7]
//
// Reference behaviour: SURVEY.md Appendix A; call site /root/reference/gaussian_renderer/__init__.py:60-135.
#include <string.h>

#include "gs_internal.cuh"
#include "gs_tiles.cuh"
#include "gs_tma.cuh"

using namespace gsb;

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void gsb_set_error(const char* s) { snprintf(g_err, sizeof(g_err), "%s", s); }
void gsb_set_errorf(const char* file, int line, const char* what, const char* detail) {
  snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", file, line, what, detail);
}
extern "C" GSB_API const char* gsb_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// optional per-kernel timing (CUDA events on the launching stream) and launch counting.
// The record list is guarded by a mutex: the backward runs on the autograd engine's thread.
// ------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::atomic<unsigned long long> g_launches{0};
struct ProfRec { cudaEvent_t a, b; int id; };
static ProfRec g_prof[8192];
static int g_prof_n = 0, g_prof_cap = 0;
void gsb_count_launch(int n) { g_launches += (unsigned long long)n; }
int gsb_prof_begin(int id, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_n >= 8192) return -1;
  if (g_prof_n >= g_prof_cap) {
    cudaEventCreate(&g_prof[g_prof_n].a);
    cudaEventCreate(&g_prof[g_prof_n].b);
    g_prof_cap = g_prof_n + 1;
  }
  g_prof[g_prof_n].id = id;
  cudaEventRecord(g_prof[g_prof_n].a, st);
  return g_prof_n++;
}
void gsb_prof_end(int slot, cudaStream_t st) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < g_prof_n) cudaEventRecord(g_prof[slot].b, st);
}
extern "C" GSB_API void gsb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  g_prof_n = 0;
}
// ms_sum[id] += elapsed, count[id] += 1 for every recorded interval; resets the record list.
extern "C" GSB_API int gsb_profile_collect(double* ms_sum, int64_t* count, int n_ids) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof[i].b) != cudaSuccess) return GSB_ERR_CUDA;
    if (cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) != cudaSuccess) return GSB_ERR_CUDA;
    if (g_prof[i].id >= 0 && g_prof[i].id < n_ids) { ms_sum[g_prof[i].id] += ms; count[g_prof[i].id] += 1; }
  }
  g_prof_n = 0;
  return GSB_OK;
}
extern "C" GSB_API uint64_t gsb_launch_count(void) { return g_launches.load(); }
static std::atomic<int> g_blend_version{2};
static std::atomic<int> g_stage_bulk{1};   // 1: slabs staged with cp.async.bulk (TMA) + mbarrier, 0: cooperative loads
int gsb_option_blend_version() { return g_blend_version.load(std::memory_order_relaxed); }
int gsb_option_stage_bulk() { return g_stage_bulk.load(std::memory_order_relaxed); }
// option "blend_version": 1 = one pixel per lane (8 warps / tile; cross-check), 2 = two pixels per lane + packed f32x2 (default)
extern "C" GSB_API int gsb_set_option(const char* name, int value) {
  if (name && strcmp(name, "blend_version") == 0 && value >= 1 && value <= 2) { g_blend_version = value; return GSB_OK; }
  if (name && strcmp(name, "stage_bulk") == 0 && (value == 0 || value == 1)) { g_stage_bulk = value; return GSB_OK; }
  gsb_set_error("gsb_set_option: unknown option or bad value");
  return GSB_ERR_INVALID;
}
extern "C" GSB_API int gsb_abi_version(void) { return 2; }

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kRowPad = 49;      // shared-memory SH row stride (48 + 1, conflict-free)

extern "C" GSB_API size_t gsb_geom_bytes(int32_t P) { return geom_view(nullptr, P).total; }
extern "C" GSB_API size_t gsb_binning_bytes(int64_t R, int32_t W, int32_t H) { return bin_view(nullptr, R, W, H).total; }
extern "C" GSB_API size_t gsb_image_bytes(int32_t W, int32_t H) { return img_view(nullptr, W, H).total; }

// ------------------------------------------------------------------------------------------
// kernel parameter blocks
// ------------------------------------------------------------------------------------------
struct InPtrs {
  int P;
  const float* means;
  const float* scales;
  const float* rots;
  const float* opac;
  const float* sh_dc;
  const float* sh_rest;
  const float* colors;
  const float* cov3D;
  int sh_packed;
  int vec_ok;       // all base pointers 16-byte aligned
  int exact_cull;
};

// ------------------------------------------------------------------------------------------
// k_setup_cam
// ------------------------------------------------------------------------------------------
struct CamArgs {
  const float* V; const float* Pm; const float* campos; const float* pose;
  int W, H; float tanfovx, tanfovy, scale_mod; int D, M, raw_params;
};
__device__ __forceinline__ void fill_cam(CamConst& c, const CamArgs& a) {
  const float* V = a.V; const float* Pm = a.Pm; const float* campos = a.campos; const float* pose = a.pose;
  const int W = a.W, H = a.H, D = a.D, M = a.M, raw_params = a.raw_params;
  const float tanfovx = a.tanfovx, tanfovy = a.tanfovy, scale_mod = a.scale_mod;
  for (int i = 0; i < 16; ++i) { c.V[i] = V[i]; c.Pm[i] = Pm[i]; }
  for (int i = 0; i < 3; ++i) c.campos[i] = campos[i];
  c.tanfovx = tanfovx; c.tanfovy = tanfovy;
  c.fx = W / (2.0f * tanfovx); c.fy = H / (2.0f * tanfovy);
  c.scale_mod = scale_mod; c.W = W; c.H = H;
  c.gx = (W + kBlock - 1) / kBlock; c.gy = (H + kBlock - 1) / kBlock;
  c.D = D; c.M = M; c.raw_params = raw_params; c.pose_on = 0;
  for (int i = 0; i < 9; ++i) c.Rc[i] = (i % 4 == 0) ? 1.f : 0.f;
  c.tc[0] = c.tc[1] = c.tc[2] = 0.f;
  c.qc[0] = 1.f; c.qc[1] = c.qc[2] = c.qc[3] = 0.f;
  if (pose) pose_to_const(pose, c);
}
__global__ void k_setup_cam(CamConst* out, CamArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  CamConst c;
  fill_cam(c, a);
  *out = c;
}

// ------------------------------------------------------------------------------------------
// staging helpers: copy `n` contiguous floats global -> shared (or back) with 128-bit accesses
// ------------------------------------------------------------------------------------------
// dst index of source element k is (k / row) * stride + col0 + (k % row)
__device__ __forceinline__ void stage_in(float* sm, const float* __restrict__ src, int n, int row,
                                         int stride, int col0, bool vec) {
  if (vec) {
    int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int k = threadIdx.x; k < n4; k += blockDim.x) {
      float4 v = __ldg(s4 + k);
      int e = 4 * k;
      int r = e / row, c = e - r * row;
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sm[r * stride + col0 + c] = vv[u];
        if (++c == row) { c = 0; ++r; }
      }
    }
    for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) {
      int r = e / row, c = e - r * row;
      sm[r * stride + col0 + c] = __ldg(src + e);
    }
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      int r = e / row, c = e - r * row;
      sm[r * stride + col0 + c] = __ldg(src + e);
    }
  }
}

__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* sm, int n, int row,
                                          int stride, int col0, bool vec) {
  if (vec) {
    int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int k = threadIdx.x; k < n4; k += blockDim.x) {
      int e = 4 * k;
      int r = e / row, c = e - r * row;
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vv[u] = sm[r * stride + col0 + c];
        if (++c == row) { c = 0; ++r; }
      }
      d4[k] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) {
      int r = e / row, c = e - r * row;
      dst[e] = sm[r * stride + col0 + c];
    }
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      int r = e / row, c = e - r * row;
      dst[e] = sm[r * stride + col0 + c];
    }
  }
}

// Shared-memory layout of the two per-Gaussian kernels (dynamic, floats, N = kPT Gaussians per CTA):
//   cam | xyz[3N] | scale[3N] | quat[4N] | opacity[N] | sh[49N]
// Every array is copied CONTIGUOUSLY (128-bit global loads -> 128-bit shared stores, conflict-free);
// each thread then walks its own row with an odd word stride (3, 45; the quaternion is one LDS.128), so
// the strided accesses are conflict-free too.  Split SH layout: dc at sh[3t], rest at sh[3N + 45t].
// Packed [P,M,3] layout (generic B2 boundary): rows padded to kRowPad = 49 words.
constexpr int kSmXyz = 0, kSmSc = 3 * kPT, kSmQ = 6 * kPT, kSmOp = 10 * kPT, kSmSh = 11 * kPT;
constexpr int kSmFloats = 11 * kPT + kRowPad * kPT;
constexpr size_t kPrepSmem = sizeof(CamConst) + 16 + (size_t)kSmFloats * 4;

__device__ __forceinline__ void copy_in(float* sm, const float* __restrict__ src, int n, bool vec) {
  if (vec) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(sm);
    for (int k = threadIdx.x; k < n4; k += blockDim.x) d4[k] = __ldg(s4 + k);
    for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) sm[e] = __ldg(src + e);
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) sm[e] = __ldg(src + e);
  }
}
__device__ __forceinline__ void copy_out(float* __restrict__ dst, const float* sm, int n, bool vec) {
  if (vec) {
    const int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(sm);
    for (int k = threadIdx.x; k < n4; k += blockDim.x) d4[k] = s4[k];
    for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) dst[e] = sm[e];
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = sm[e];
  }
}

struct ShRows { int dc_stride, rest_off, rest_stride; };   // thread t: dc at sh[dc_stride*t], rest at sh[rest_off + rest_stride*t]
__device__ __forceinline__ ShRows sh_rows(int sh_packed, int M) {
  ShRows r;
  if (sh_packed) { r.dc_stride = kRowPad; r.rest_off = 3; r.rest_stride = kRowPad; }
  else { r.dc_stride = 3; r.rest_off = 3 * kPT; r.rest_stride = 3 * (M - 1); }
  return r;
}

// Generic staging (ragged last CTA, unaligned pointers, packed SH layout, precomputed inputs); full aligned CTAs take
// the bulk-copy path below.
__device__ __forceinline__ void load_block_inputs(const InPtrs& in, int first, int nv, bool use_sh, int D, int M,
                                                  float* sm) {
  const bool vec = in.vec_ok != 0;
  copy_in(sm + kSmXyz, in.means + (size_t)3 * first, 3 * nv, vec);
  if (in.scales) copy_in(sm + kSmSc, in.scales + (size_t)3 * first, 3 * nv, vec);
  if (in.rots) copy_in(sm + kSmQ, in.rots + (size_t)4 * first, 4 * nv, vec);
  copy_in(sm + kSmOp, in.opac + first, nv, vec);
  if (use_sh) {
    if (in.sh_packed) {
      stage_in(sm + kSmSh, in.sh_dc + (size_t)3 * M * first, 3 * M * nv, 3 * M, kRowPad, 0, vec);
    } else {
      copy_in(sm + kSmSh, in.sh_dc + (size_t)3 * first, 3 * nv, vec);
      if (D > 0 && M > 1)
        copy_in(sm + kSmSh + 3 * kPT, in.sh_rest + (size_t)3 * (M - 1) * first, 3 * (M - 1) * nv, vec);
    }
  }
}

// The same full-CTA fast path with the copies done by the TMA unit: one thread arms an mbarrier with the byte count and
// issues one bulk copy per input array (cp.async.bulk.shared.global, SASS UBLKCP); nothing passes through registers or
// allocates L1 lines (with register staging, the 180 KB a resident set of CTAs keeps in flight is bounded by the L1
// carve-out: giving shared memory the whole unified storage made both per-Gaussian kernels 8-13 % slower).
__device__ __forceinline__ bool block_inputs_bulk_ok(const InPtrs& in, int nv, bool use_sh, int M) {
  return in.vec_ok != 0 && nv == kPT && in.scales && in.rots && (!use_sh || (!in.sh_packed && M == 16));
}
__device__ __forceinline__ void bulk_load_block_inputs(const InPtrs& in, int first, bool use_sh, int D, float* sm,
                                                       uint64_t* bar) {
  constexpr uint32_t B3 = 3 * kPT * 4, B4 = 4 * kPT * 4, B1 = kPT * 4, BR = 45 * kPT * 4;
  const bool rest = use_sh && D > 0;
  mbar_expect_tx(bar, 2 * B3 + B4 + B1 + (use_sh ? B3 : 0u) + (rest ? BR : 0u));
  bulk_g2s(sm + kSmXyz, in.means + (size_t)3 * first, B3, bar);
  bulk_g2s(sm + kSmSc, in.scales + (size_t)3 * first, B3, bar);
  bulk_g2s(sm + kSmQ, in.rots + (size_t)4 * first, B4, bar);
  bulk_g2s(sm + kSmOp, in.opac + first, B1, bar);
  if (use_sh) bulk_g2s(sm + kSmSh, in.sh_dc + (size_t)3 * first, B3, bar);
  if (rest) bulk_g2s(sm + kSmSh + 3 * kPT, in.sh_rest + (size_t)45 * first, BR, bar);
}
// all threads call; returns when the CTA's inputs are in shared memory
__device__ __forceinline__ void stage_block_inputs(const InPtrs& in, int first, int nv, bool use_sh, int D, int M, float* sm,
                                                   uint64_t* bar) {
  if (block_inputs_bulk_ok(in, nv, use_sh, M)) {
    if (threadIdx.x == 0) bulk_load_block_inputs(in, first, use_sh, D, sm, bar);
    mbar_wait(bar, 0u);
  } else {
    load_block_inputs(in, first, nv, use_sh, D, M, sm);
    __syncthreads();
  }
}

__device__ __forceinline__ void read_gauss(const float* sm, int t, bool has_sr, GaussIn& g) {
  g.m[0] = sm[kSmXyz + 3 * t]; g.m[1] = sm[kSmXyz + 3 * t + 1]; g.m[2] = sm[kSmXyz + 3 * t + 2];
  if (has_sr) {
    g.sc[0] = sm[kSmSc + 3 * t]; g.sc[1] = sm[kSmSc + 3 * t + 1]; g.sc[2] = sm[kSmSc + 3 * t + 2];
    const float4 q = *reinterpret_cast<const float4*>(sm + kSmQ + 4 * t);
    g.q[0] = q.x; g.q[1] = q.y; g.q[2] = q.z; g.q[3] = q.w;
  } else {
    g.sc[0] = g.sc[1] = g.sc[2] = 1.f;
    g.q[0] = 1.f; g.q[1] = g.q[2] = g.q[3] = 0.f;
  }
  g.op = sm[kSmOp + t];
}

// ------------------------------------------------------------------------------------------
// k_preprocess
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPT)
k_preprocess(InPtrs in, GeomView gv, int* __restrict__ radii) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CamConst* cam = reinterpret_cast<CamConst*>(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw + ((sizeof(CamConst) + 15) / 16) * 16);
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(gv.cam);
    uint32_t* d = reinterpret_cast<uint32_t*>(cam);
    for (int k = threadIdx.x; k < (int)(sizeof(CamConst) / 4); k += blockDim.x) d[k] = s[k];
  }
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  int first = blockIdx.x * kPT;
  int nv = min(kPT, in.P - first);
  bool use_sh = in.colors == nullptr;
  __syncthreads();
  stage_block_inputs(in, first, nv, use_sh, cam->D, cam->M, sm, &s_bar);
  const int t = threadIdx.x;
  const bool active = t < nv;
  const int i = first + (active ? t : 0);
  GaussIn g;
  read_gauss(sm, active ? t : 0, in.scales != nullptr, g);
  Proj p;
  project_geometry(*cam, g, in.cov3D ? in.cov3D + (size_t)6 * i : nullptr, p);
  if (!active) p.visible = 0;
  float qthr = -1.f;
  uint32_t ntiles = 0, kmask = 0u;
  bool coop = false, big = false;
  SplatRect sr_;
  sr_.x = sr_.y = sr_.A = sr_.B = sr_.C = 0.f; sr_.qthr = -1.f;
  sr_.rx0 = sr_.rx1 = sr_.ry0 = sr_.ry1 = 0;
  const TileSink sink{gv.tcount, nullptr, nullptr, nullptr, 0u};       // count mode
  const bool cull = in.exact_cull != 0;
  if (p.visible) {
    if (use_sh) {
      const ShRows sr = sh_rows(in.sh_packed, cam->M);
      project_color(*cam, sm + kSmSh + sr.dc_stride * t, sm + kSmSh + sr.rest_off + sr.rest_stride * t, p);
    } else {
      p.rgb[0] = in.colors[3 * i]; p.rgb[1] = in.colors[3 * i + 1]; p.rgb[2] = in.colors[3 * i + 2];
    }
    qthr = cull_threshold(p.opacity);
    // The tile coverage is decided on the record exactly as it is stored (conic and threshold in the log2 domain), so
    // that k_scatter -- which re-walks large rects from the stored record -- sees bit-identical inputs.
    sr_.x = p.x; sr_.y = p.y;
    sr_.A = 0.5f * kLog2e * p.A; sr_.B = 0.5f * kLog2e * p.B; sr_.C = 0.5f * kLog2e * p.C;
    sr_.qthr = qthr < 0.f ? -1.f : 0.5f * kLog2e * qthr;
    // Lossless tightening of the tile rect: alpha >= 1/255 needs q(d) <= qthr, an ellipse whose bounding box has the
    // half-extents sqrt(qthr C / det), sqrt(qthr A / det) -- for anisotropic or low-opacity Gaussians far inside the
    // reference's square 3-sigma rect (which stays what `radii` reports).  Tiles outside it would fail the per-tile test
    // anyway; not walking them keeps the count/emit loops proportional to what is kept, also for Gaussians that grow
    // large during training.  (+1 px margin; the tile range is rounded outwards.)
    if (cull && qthr >= 0.f) {
      const float det = sr_.A * sr_.C - sr_.B * sr_.B;
      if (det > 0.f) {
        const float hx = sqrtf(sr_.qthr * sr_.C / det) + 1.0f, hy = sqrtf(sr_.qthr * sr_.A / det) + 1.0f;
        const float big = 1.0e8f;
        const int tx0 = (int)floorf(fmaxf(-big, fminf(big, (p.x - hx) * (1.0f / kBlock))));
        const int tx1 = (int)floorf(fmaxf(-big, fminf(big, (p.x + hx) * (1.0f / kBlock)))) + 1;
        const int ty0 = (int)floorf(fmaxf(-big, fminf(big, (p.y - hy) * (1.0f / kBlock))));
        const int ty1 = (int)floorf(fmaxf(-big, fminf(big, (p.y + hy) * (1.0f / kBlock)))) + 1;
        p.rx0 = max(p.rx0, tx0); p.rx1 = min(p.rx1, tx1);
        p.ry0 = max(p.ry0, ty0); p.ry1 = min(p.ry1, ty1);
        if (p.rx1 <= p.rx0 || p.ry1 <= p.ry0) { p.rx0 = p.rx1 = p.ry0 = p.ry1 = 0; }
      }
    }
    sr_.rx0 = p.rx0; sr_.rx1 = p.rx1; sr_.ry0 = p.ry0; sr_.ry1 = p.ry1;
    const int area = (p.rx1 - p.rx0) * (p.ry1 - p.ry0);
    if (cull && qthr < 0.f) ntiles = 0;
    else if (area > kBigRect) big = true;        // counted later by k_big_rects (which also fills gv.tiles[i])
    else if (area > kCoopTiles) coop = true;
    else {
      // per-tile counts: plain fire-and-forget RED.ADD per kept tile (measured faster here than warp-aggregating
      // them: nothing waits for the result, unlike the emit pass in k_scatter)
      kmask = rect_keep_mask_count(sr_, cam->W, cam->H, cam->gx, cull, gv.tcount);
      ntiles = (uint32_t)__popc(kmask);
    }
  } else {
    p.x = p.y = p.A = p.B = p.C = 0.f; p.rgb[0] = p.rgb[1] = p.rgb[2] = 0.f;
  }
  {
    const uint32_t c = visit_tiles_coop(coop, sr_, cam->W, cam->H, cam->gx, cull, sink, 0ull);
    if (coop) ntiles = c;
  }
  // ---- outputs.  The 48-byte splat record (rows x,y,A',B' | C',o,qthr',id | r,g,b,radius; conic pre-scaled into the log2
  // domain: power*log2(e) = A' dx^2 + B' dx dy + C' dy^2) is what the blend kernels gather by id, so it is stored as an
  // array of structures -- transposed through shared memory so that the CTA's 128 x 48 B leave as one contiguous block
  // of 128-bit stores.  The binning record (depth, tile rect, keep mask) is a plain coalesced float4 array.
  __syncthreads();                                   // every thread is done with the staged inputs: reuse `sm`
  float4* sm4 = reinterpret_cast<float4*>(sm);
  if (active) {
    sm4[3 * t + 0] = make_float4(p.x, p.y, -0.5f * kLog2e * p.A, -kLog2e * p.B);
    sm4[3 * t + 1] = make_float4(-0.5f * kLog2e * p.C, p.opacity, 0.5f * kLog2e * qthr, __uint_as_float((uint32_t)i));
    sm4[3 * t + 2] = make_float4(p.rgb[0], p.rgb[1], p.rgb[2], (float)p.radius);
    gv.brec[i] = make_float4(p.depth, __uint_as_float((uint32_t)p.rx0 | ((uint32_t)p.rx1 << 16)),
                             __uint_as_float((uint32_t)p.ry0 | ((uint32_t)p.ry1 << 16)), __uint_as_float(kmask));
    gv.tiles[i] = ntiles;
    if (big) gv.q_big[atomicAdd(gv.aux, 1u)] = (uint32_t)i;     // at most once per Gaussian: the queue holds P entries
    gv.clamped[i] = (uint8_t)p.clamped;
    radii[i] = p.radius;
  }
  tma_store_fence();                                 // the records leave as ONE bulk store (shared -> global, 48 nv bytes)
  __syncthreads();
  if (threadIdx.x == 0) {
    bulk_s2g(gv.rec + 3 * (size_t)first, sm4, (uint32_t)(48 * nv));
    bulk_commit();
    bulk_wait_read();
  }
}

// ------------------------------------------------------------------------------------------
// k_preprocess_bwd
// ------------------------------------------------------------------------------------------
struct OutPtrs {
  float* dmeans; float* dmeans2D; float* dscales; float* drots; float* dopac;
  float* dsh_dc; float* dsh_rest; float* dcolors; float* dcov3D;
};

#ifndef GSB_PBWD_MINB
#define GSB_PBWD_MINB 4
#endif
__global__ void __launch_bounds__(kPT, GSB_PBWD_MINB)
k_preprocess_bwd(InPtrs in, GeomView gv, const int pose_only_layout, OutPtrs out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CamConst* cam = reinterpret_cast<CamConst*>(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw + ((sizeof(CamConst) + 15) / 16) * 16);
  __shared__ float s_pose[kPT / 32][16];
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(gv.cam);
    uint32_t* d = reinterpret_cast<uint32_t*>(cam);
    for (int k = threadIdx.x; k < (int)(sizeof(CamConst) / 4); k += blockDim.x) d[k] = s[k];
  }
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  const int first = blockIdx.x * kPT;
  const int nv = min(kPT, in.P - first);
  const bool use_sh = in.colors == nullptr;
  const bool vec = in.vec_ok != 0;
  __syncthreads();
  const int D = cam->D, M = cam->M;
  stage_block_inputs(in, first, nv, use_sh, D, M, sm, &s_bar);
  const int t = threadIdx.x;
  const int i = first + t;
  float pa[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) pa[k] = 0.f;
  GaussGrad gg;
  gg.dm[0] = gg.dm[1] = gg.dm[2] = 0.f; gg.dsc[0] = gg.dsc[1] = gg.dsc[2] = 0.f;
  gg.dq[0] = gg.dq[1] = gg.dq[2] = gg.dq[3] = 0.f; gg.dop = 0.f;
  gg.dmeans2D[0] = gg.dmeans2D[1] = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) gg.dcov3D[k] = 0.f;
  gg.dcolor[0] = gg.dcolor[1] = gg.dcolor[2] = 0.f;
  // Every thread touches only its own shared-memory rows from here on (no barrier needed until the
  // cooperative stores): the SH gradient is written IN PLACE over the staged SH row.
  const ShRows sr = sh_rows(in.sh_packed, M);
  float* row_dc = sm + kSmSh + sr.dc_stride * t;
  float* row_rest = sm + kSmSh + sr.rest_off + sr.rest_stride * t;
  bool has = false;
  if (t < nv) {
    float4 d0 = gv.dacc[3 * (size_t)i], d1 = gv.dacc[3 * (size_t)i + 1], d2 = gv.dacc[3 * (size_t)i + 2];
    float ds[9] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x};
    if (pose_only_layout) { ds[8] = ds[5]; ds[5] = 0.f; }     // tracking mode: dacc[5] carries dL/db, no dL/dopacity
    bool any = false;
#pragma unroll
    for (int k = 0; k < 9; ++k) any |= (ds[k] != 0.f);
    if (any && gv.tiles[i] > 0) {
      GaussIn g;
      read_gauss(sm, t, in.scales != nullptr, g);
      Proj p;
      project_geometry(*cam, g, in.cov3D ? in.cov3D + (size_t)6 * i : nullptr, p);
      if (p.visible) {
        p.clamped = gv.clamped[i];
        // The blend backward accumulates RAW pixel moments of w = dL/dalpha * (opacity * G) per Gaussian
        //   ds = [S w dx, S w dy, S w dx^2, S w dx dy, S w dy^2, S w, dL/dr, dL/dg, dL/db];
        // the per-Gaussian constants are applied once here instead of once per (warp, Gaussian) in the blend kernel:
        //   dL/dx = -(A Sx + B Sy), dL/dy = -(C Sy + B Sx), dL/dA = -Sxx/2, dL/dB = -Sxy, dL/dC = -Syy/2,
        //   dL/dopacity = Sw / opacity.
        {
          const float sx = ds[0], sy = ds[1];
          ds[0] = -(p.A * sx + p.B * sy);
          ds[1] = -(p.C * sy + p.B * sx);
          ds[2] *= -0.5f;
          ds[3] = -ds[3];
          ds[4] *= -0.5f;
          ds[5] = p.opacity > 0.f ? ds[5] / p.opacity : 0.f;     // contributing pairs have opacity >= 1/255
        }
        project_bwd(*cam, g, p, row_rest, use_sh, in.cov3D != nullptr, ds, gg, row_dc, row_rest, pa);
        has = use_sh;
      }
    }
    // inactive coefficients (and everything of a Gaussian that did not contribute) get zero gradient
    if (!has) { row_dc[0] = 0.f; row_dc[1] = 0.f; row_dc[2] = 0.f; }
    for (int k = has ? 3 * ((D + 1) * (D + 1) - 1) : 0; k < 3 * (M - 1); ++k) row_rest[k] = 0.f;
    sm[kSmXyz + 3 * t] = gg.dm[0]; sm[kSmXyz + 3 * t + 1] = gg.dm[1]; sm[kSmXyz + 3 * t + 2] = gg.dm[2];
    sm[kSmSc + 3 * t] = gg.dsc[0]; sm[kSmSc + 3 * t + 1] = gg.dsc[1]; sm[kSmSc + 3 * t + 2] = gg.dsc[2];
    *reinterpret_cast<float4*>(sm + kSmQ + 4 * t) = make_float4(gg.dq[0], gg.dq[1], gg.dq[2], gg.dq[3]);
    sm[kSmOp + t] = gg.dop;
    if (out.dmeans2D) {
      out.dmeans2D[3 * (size_t)i] = gg.dmeans2D[0];
      out.dmeans2D[3 * (size_t)i + 1] = gg.dmeans2D[1];
      out.dmeans2D[3 * (size_t)i + 2] = 0.f;
    }
    if (out.dcolors) {
      out.dcolors[3 * (size_t)i] = gg.dcolor[0]; out.dcolors[3 * (size_t)i + 1] = gg.dcolor[1];
      out.dcolors[3 * (size_t)i + 2] = gg.dcolor[2];
    }
    if (out.dcov3D) {
#pragma unroll
      for (int k = 0; k < 6; ++k) out.dcov3D[6 * (size_t)i + k] = gg.dcov3D[k];
    }
  }
  // Full CTA, aligned, split SH layout: every gradient array of the CTA leaves as one bulk store (shared -> global)
  // issued by one thread; the generic path below covers ragged / unaligned / packed cases.
  const bool bulk_out = vec && nv == kPT && !in.sh_packed && M == 16;
  if (bulk_out) tma_store_fence();
  __syncthreads();
  if (bulk_out) {
    if (threadIdx.x == 0) {
      constexpr uint32_t B3 = 3 * kPT * 4, B4 = 4 * kPT * 4, B1 = kPT * 4, BR = 45 * kPT * 4;
      if (out.dmeans) bulk_s2g(out.dmeans + (size_t)3 * first, sm + kSmXyz, B3);
      if (out.dscales) bulk_s2g(out.dscales + (size_t)3 * first, sm + kSmSc, B3);
      if (out.drots) bulk_s2g(out.drots + (size_t)4 * first, sm + kSmQ, B4);
      if (out.dopac) bulk_s2g(out.dopac + first, sm + kSmOp, B1);
      if (use_sh && out.dsh_dc) {
        bulk_s2g(out.dsh_dc + (size_t)3 * first, sm + kSmSh, B3);
        if (out.dsh_rest) bulk_s2g(out.dsh_rest + (size_t)45 * first, sm + kSmSh + 3 * kPT, BR);
      }
      bulk_commit();
    }
  } else {
  if (out.dmeans) copy_out(out.dmeans + (size_t)3 * first, sm + kSmXyz, 3 * nv, vec);
  if (out.dscales) copy_out(out.dscales + (size_t)3 * first, sm + kSmSc, 3 * nv, vec);
  if (out.drots) copy_out(out.drots + (size_t)4 * first, sm + kSmQ, 4 * nv, vec);
  if (out.dopac) copy_out(out.dopac + first, sm + kSmOp, nv, vec);
  if (use_sh && out.dsh_dc) {
    if (in.sh_packed) {
      stage_out(out.dsh_dc + (size_t)3 * M * first, sm + kSmSh, 3 * M * nv, 3 * M, kRowPad, 0, vec);
    } else {
      copy_out(out.dsh_dc + (size_t)3 * first, sm + kSmSh, 3 * nv, vec);
      if (M > 1 && out.dsh_rest)
        copy_out(out.dsh_rest + (size_t)3 * (M - 1) * first, sm + kSmSh + 3 * kPT, 3 * (M - 1) * nv, vec);
    }
  }
  }
  // pose-gradient partials: warp shuffle, then across the 8 warps
  if (cam->pose_on) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float v = pa[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_pose[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kPT / 32; ++w) v += s_pose[w][threadIdx.x];
      gv.pose_part[(size_t)blockIdx.x * 16 + threadIdx.x] = v;
    }
  }
  if (bulk_out && threadIdx.x == 0) bulk_wait_read();   // shared memory must outlive the bulk stores reading it
}

// Pose-gradient partials [nblocks][16] -> 16 column sums (one CTA per column, deterministic order) ...
__global__ void __launch_bounds__(256) k_pose_reduce(const float* __restrict__ part, int nblocks, float* __restrict__ acc16) {
  __shared__ float s[256];
  const int c = blockIdx.x;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
  int b = threadIdx.x;
  for (; b + 768 < nblocks; b += 1024) {
    v0 += part[(size_t)b * 16 + c];
    v1 += part[(size_t)(b + 256) * 16 + c];
    v2 += part[(size_t)(b + 512) * 16 + c];
    v3 += part[(size_t)(b + 768) * 16 + c];
  }
  for (; b < nblocks; b += 256) v0 += part[(size_t)b * 16 + c];
  s[threadIdx.x] = (v0 + v1) + (v2 + v3);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) acc16[c] = s[0];
}
// ... and the chain to dL/dP[7] (utils/pose_utils.py quad2rotation + normalisation backward).
__global__ void k_pose_finalize(const float* __restrict__ acc16, const float* pose, float* dpose) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float acc[16];
  for (int c = 0; c < 16; ++c) acc[c] = acc16[c];
  float dp[7];
  pose_grad_finalize(pose, acc, dp);
  for (int c = 0; c < 7; ++c) dpose[c] = dp[c];
}

__global__ void k_mark_visible(int P, const float* means, const float* V, uint8_t* present) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float z = means[3 * i] * V[2] + means[3 * i + 1] * V[6] + means[3 * i + 2] * V[10] + V[14];
  present[i] = z > kNear ? 1 : 0;
}
// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static int make_inptrs(const GsbCamera* cam, const GsbGaussians* g, InPtrs& in) {
  GSB_REQUIRE(cam && g, "null camera / gaussians");
  GSB_REQUIRE(g->P >= 0, "P < 0");
  GSB_REQUIRE(cam->width > 0 && cam->height > 0, "image size");
  GSB_REQUIRE((long long)((cam->width + kBlock - 1) / kBlock) * ((cam->height + kBlock - 1) / kBlock) < kMaxTiles,
              "image too large (more than 65535 tiles)");
  GSB_REQUIRE(cam->sh_degree >= 0 && cam->sh_degree <= 3, "sh_degree must be 0..3");
  GSB_REQUIRE(cam->sh_coeffs >= 1 && cam->sh_coeffs <= 16, "sh_coeffs must be 1..16");
  GSB_REQUIRE((cam->sh_degree + 1) * (cam->sh_degree + 1) <= cam->sh_coeffs || g->colors_precomp,
              "active SH degree exceeds stored coefficients");
  GSB_REQUIRE(cam->bg && cam->viewmatrix && cam->projmatrix && cam->campos, "null camera tensor");
  if (g->P > 0) {   // an empty cloud may come with null tensors
    GSB_REQUIRE(g->means3D && g->opacities, "means3D / opacities are required");
    GSB_REQUIRE((g->sh_dc != nullptr) != (g->colors_precomp != nullptr),
                "provide exactly one of SHs / precomputed colours");
    GSB_REQUIRE((g->scales != nullptr && g->rotations != nullptr) != (g->cov3D_precomp != nullptr),
                "provide exactly one of scale+rotation / precomputed 3D covariance");
    if (g->sh_dc && !g->sh_packed && cam->sh_coeffs > 1 && cam->sh_degree > 0)
      GSB_REQUIRE(g->sh_rest != nullptr, "sh_rest missing");
  }
  in.P = g->P; in.means = g->means3D; in.scales = g->scales; in.rots = g->rotations; in.opac = g->opacities;
  in.sh_dc = g->sh_dc; in.sh_rest = g->sh_rest; in.colors = g->colors_precomp; in.cov3D = g->cov3D_precomp;
  in.sh_packed = g->sh_packed; in.exact_cull = cam->exact_cull;
  uintptr_t a = (uintptr_t)g->means3D | (uintptr_t)g->scales | (uintptr_t)g->rotations |
                (uintptr_t)g->opacities | (uintptr_t)g->sh_dc | (uintptr_t)g->sh_rest;
  in.vec_ok = (a & 15) == 0;
  return GSB_OK;
}

// cudaFuncSetAttribute is per device: remember which devices have been configured
static int ensure_attrs() {
  static bool done[64] = {};
  int dev = 0;
  GSB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    GSB_CUDA(cudaFuncSetAttribute(k_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmem));
    GSB_CUDA(cudaFuncSetAttribute(k_preprocess_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmem));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return GSB_OK;
}

extern "C" GSB_API uint32_t* gsb_status_device(void* geom, int32_t P) { return geom ? geom_view(geom, P).status : nullptr; }

static CamArgs cam_args(const GsbCamera* cam, const GsbGaussians* g) {
  CamArgs a;
  a.V = cam->viewmatrix; a.Pm = cam->projmatrix; a.campos = cam->campos; a.pose = g->pose;
  a.W = cam->width; a.H = cam->height; a.tanfovx = cam->tanfovx; a.tanfovy = cam->tanfovy;
  a.scale_mod = cam->scale_modifier; a.D = cam->sh_degree; a.M = cam->sh_coeffs; a.raw_params = g->raw_params;
  return a;
}

extern "C" GSB_API int gsb_preprocess(const GsbCamera* cam, const GsbGaussians* g, void* geom, size_t geom_bytes,
                              int32_t* radii, uint32_t* status_host, gsb_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  InPtrs in;
  int rc = make_inptrs(cam, g, in);
  if (rc) return rc;
  GSB_REQUIRE(geom && (radii || g->P == 0), "null buffer");
  const int P = g->P;
  GeomView gv = geom_view(geom, P);
  if (gv.total > geom_bytes) { gsb_set_error("geom buffer too small"); return GSB_ERR_CAPACITY; }
  rc = ensure_attrs();
  if (rc) return rc;
  const int ntiles = ((cam->width + kBlock - 1) / kBlock) * ((cam->height + kBlock - 1) / kBlock);
  gsb_count_launch(1);
  k_setup_cam<<<1, 32, 0, st>>>(gv.cam, cam_args(cam, g));
  // aux (the per-forward counters) sits immediately before tcount: one memset clears both
  GSB_CUDA(cudaMemsetAsync(gv.aux, 0, (size_t)((char*)gv.tcount - (char*)gv.aux) + (size_t)ntiles * 4, st));
  if (P > 0) {
    const int nb = (P + kPT - 1) / kPT;
    ProfScope ps(GSB_K_PREPROCESS, st, 2);
    k_preprocess<<<nb, kPT, kPrepSmem, st>>>(in, gv, radii);
    rc = gsb_launch_big_rects(P, gv, nullptr, cam->width, cam->height, cam->exact_cull, 0u, st);   // count mode
    if (rc) return rc;
  }
  rc = gsb_launch_tile_scan(gv, ntiles, st);
  if (rc) return rc;
  if (status_host) GSB_CUDA(cudaMemcpyAsync(status_host, gv.status, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaGetLastError());
  return GSB_OK;
}

extern "C" GSB_API int gsb_render(const GsbCamera* cam, int32_t P, void* geom, void* binning, size_t binning_bytes,
                          int64_t R, void* image, float* out_color, uint32_t* status_host, gsb_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  GSB_REQUIRE(cam && geom && binning && image && out_color, "null buffer");
  GSB_REQUIRE(R >= 0 && R < (int64_t)0x7fffffff, "R out of range");
  const int W = cam->width, H = cam->height;
  GeomView gv = geom_view(geom, P);
  BinView bv = bin_view(binning, R, W, H);
  if (bv.total > binning_bytes) { gsb_set_error("binning buffer too small"); return GSB_ERR_CAPACITY; }
  ImgView iv = img_view(image, W, H);
  int rc = gsb_launch_binning(P, gv, bv, W, H, cam->exact_cull, (uint32_t)R, st);
  if (rc) return rc;
  rc = gsb_launch_blend_fwd(gv, bv, iv, cam->bg, W, H, out_color, st);
  if (rc) return rc;
  if (status_host) GSB_CUDA(cudaMemcpyAsync(status_host, gv.status, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaGetLastError());
  return GSB_OK;
}

extern "C" GSB_API int gsb_backward(const GsbCamera* cam, const GsbGaussians* g, void* geom, void* binning, int64_t R,
                            void* image, const float* dL_dout, const GsbGrads* grads, gsb_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  InPtrs in;
  int rc = make_inptrs(cam, g, in);
  if (rc) return rc;
  GSB_REQUIRE(geom && binning && image && dL_dout && grads, "null buffer");
  if (g->pose) GSB_REQUIRE(grads->dL_dpose != nullptr, "dL_dpose required when pose is fused");
  const int P = g->P, W = cam->width, H = cam->height;
  GeomView gv = geom_view(geom, P);
  BinView bv = bin_view(binning, R, W, H);
  ImgView iv = img_view(image, W, H);
  rc = ensure_attrs();
  if (rc) return rc;
  if (P == 0) {
    if (grads->dL_dpose) GSB_CUDA(cudaMemsetAsync(grads->dL_dpose, 0, 7 * 4, st));
    return GSB_OK;
  }
  GSB_CUDA(cudaMemsetAsync(gv.dacc, 0, (size_t)P * 48, st));
  // tracking mode (render.py:99-170): only dL/dpose is wanted -> the blend backward reduces 8 values instead of 9
  const bool pose_only = g->pose && grads->dL_dpose && !grads->dL_dmeans3D && !grads->dL_dmeans2D && !grads->dL_dscales &&
                         !grads->dL_drotations && !grads->dL_dopacities && !grads->dL_dsh_dc && !grads->dL_dsh_rest &&
                         !grads->dL_dcolors && !grads->dL_dcov3D && gsb_option_blend_version() == 2 &&
                         gsb_option_stage_bulk() == 1;
  rc = gsb_launch_blend_bwd(gv, bv, iv, cam->bg, W, H, dL_dout, (float*)gv.dacc, pose_only, st);
  if (rc) return rc;
  OutPtrs out;
  out.dmeans = grads->dL_dmeans3D; out.dmeans2D = grads->dL_dmeans2D; out.dscales = grads->dL_dscales;
  out.drots = grads->dL_drotations; out.dopac = grads->dL_dopacities; out.dsh_dc = grads->dL_dsh_dc;
  out.dsh_rest = grads->dL_dsh_rest; out.dcolors = grads->dL_dcolors; out.dcov3D = grads->dL_dcov3D;
  uintptr_t a = (uintptr_t)out.dmeans | (uintptr_t)out.dscales | (uintptr_t)out.drots | (uintptr_t)out.dopac |
                (uintptr_t)out.dsh_dc | (uintptr_t)out.dsh_rest;
  if (a & 15) in.vec_ok = 0;
  const int nb = (P + kPT - 1) / kPT;
  { ProfScope ps(GSB_K_PREPROCESS_BWD, st, g->pose ? 3 : 1);
    k_preprocess_bwd<<<nb, kPT, kPrepSmem, st>>>(in, gv, pose_only ? 1 : 0, out);
    if (g->pose) {
      k_pose_reduce<<<16, 256, 0, st>>>(gv.pose_part, nb, gv.pose_acc);
      k_pose_finalize<<<1, 32, 0, st>>>(gv.pose_acc, g->pose, grads->dL_dpose);
    } }
  GSB_CUDA(cudaGetLastError());
  return GSB_OK;
}

extern "C" GSB_API int gsb_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, gsb_stream_t stream_) {
  (void)projmatrix;
  GSB_REQUIRE(P >= 0 && means3D && viewmatrix && present, "null buffer");
  if (P > 0)
    k_mark_visible<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(P, means3D, viewmatrix, present);
  GSB_CUDA(cudaGetLastError());
  return GSB_OK;
}
