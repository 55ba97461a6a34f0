// B200 (sm_100a) differentiable Gaussian rasterizer: kernels + C ABI (include/gsb200.h).
//
// Pipeline (one view):
//   k_setup_cam        camera/pose constants -> one CamConst in HBM (read by every block)
//   k_preprocess       fused pose transform + activations + EWA projection + SH->RGB, 128-bit
//                      coalesced loads staged through shared memory; writes packed splat records
//   cub sort           depth keys (32 bit) -> depth order                     [library: cub]
//   cub scan           tile counts in depth order -> offsets, R
//   k_duplicate        (tile id, gaussian id) instances emitted in depth order, lossless culling
//   cub sort           stable sort on the tile bits only (13 bits at 1080p)   [library: cub]
//   k_ranges_gather    per-tile ranges + gather of the splat records into contiguous per-tile
//                      slabs (3 x float4 per instance)
//   k_blend_fwd        one CTA per 16x16 tile, one warp per 8x4 sub-tile; slab chunks staged in
//                      shared memory; per-warp ballot-compacted sub-tile culling
//   k_blend_bwd        back-to-front replay; 9 gradients per (warp, Gaussian) reduced with a
//                      transposing butterfly (14 shuffles) then 9 RED.ADD.F32
//   k_preprocess_bwd   analytic backward to the model's own tensors + in-kernel pose-gradient
//                      reduction; k_pose_finalize chains it to dL/dP[7]
//
// Reference behaviour: SURVEY.md Appendix A; call site /root/reference/gaussian_renderer/__init__.py:60-135.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdio.h>
#include <string.h>

#include "../../include/gsb200.h"
#include "gs_math.cuh"

using namespace gsb;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
#define GSB_CUDA(x)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #x,              \
               cudaGetErrorString(e_));                                                   \
      return GSB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
#define GSB_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      snprintf(g_err, sizeof(g_err), "%s:%d invalid argument: %s", __FILE__, __LINE__, msg); \
      return GSB_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

extern "C" GSB_API const char* gsb_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// optional per-kernel timing (CUDA events on the launching stream) and launch counting
// ------------------------------------------------------------------------------------------
static bool g_prof_on = false;
static unsigned long long g_launches = 0;
struct ProfRec { cudaEvent_t a, b; int id; };
static ProfRec g_prof[8192];
static int g_prof_n = 0, g_prof_cap = 0;
void gsb_count_launch(int n) { g_launches += (unsigned long long)n; }
int gsb_prof_begin(int id, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= 8192) return -1;
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
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}
extern "C" GSB_API void gsb_profile_enable(int on) { g_prof_on = on != 0; g_prof_n = 0; }
// ms_sum[id] += elapsed, count[id] += 1 for every recorded interval; resets the record list.
extern "C" GSB_API int gsb_profile_collect(double* ms_sum, int64_t* count, int n_ids) {
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof[i].b) != cudaSuccess) return GSB_ERR_CUDA;
    if (cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) != cudaSuccess) return GSB_ERR_CUDA;
    if (g_prof[i].id >= 0 && g_prof[i].id < n_ids) { ms_sum[g_prof[i].id] += ms; count[g_prof[i].id] += 1; }
  }
  g_prof_n = 0;
  return GSB_OK;
}
extern "C" GSB_API uint64_t gsb_launch_count(void) { return g_launches; }
static int g_blend_version = 2;
static int g_stage_bulk = 1;   // 1: slabs staged with cp.async.bulk (TMA) + mbarrier, 0: cooperative loads
// option "blend_version": 1 = one pixel per lane (8 warps / tile), 2 = two pixels per lane + packed f32x2 (default),
// 3 = four pixels per lane (2 warps / tile; measured ~10 % slower than v2 on B200, kept as an experiment)
extern "C" GSB_API int gsb_set_option(const char* name, int value) {
  if (name && strcmp(name, "blend_version") == 0 && value >= 1 && value <= 3) { g_blend_version = value; return GSB_OK; }
  if (name && strcmp(name, "stage_bulk") == 0 && (value == 0 || value == 1)) { g_stage_bulk = value; return GSB_OK; }
  snprintf(g_err, sizeof(g_err), "gsb_set_option: unknown option or bad value");
  return GSB_ERR_INVALID;
}
struct ProfScope {
  int slot; cudaStream_t st;
  ProfScope(int id, cudaStream_t s, int launches = 1) : st(s) { gsb_count_launch(launches); slot = gsb_prof_begin(id, s); }
  ~ProfScope() { gsb_prof_end(slot, st); }
};
extern "C" GSB_API int gsb_abi_version(void) { return 1; }
void gsb_set_error(const char* s) { snprintf(g_err, sizeof(g_err), "%s", s); }

// ------------------------------------------------------------------------------------------
// buffer layouts
// ------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

constexpr int kThreads = 256;
constexpr int kPT = 128;         // threads (= Gaussians) per CTA in the per-Gaussian kernels
constexpr int kRowPad = 49;      // shared-memory SH row stride (48 + 1, conflict-free)
#ifndef GSB_CHUNK
#define GSB_CHUNK 256
#endif
#ifndef GSB_FWD_MINB
#define GSB_FWD_MINB 7
#endif
#ifndef GSB_BWD_MINB
#define GSB_BWD_MINB 6
#endif
constexpr int kChunk1 = 256;           // v1 blend kernels: one entry per thread
constexpr int kChunk = GSB_CHUNK;      // slab entries staged per step in the blend kernels

struct GeomView {
  CamConst* cam;
  float4* xyAB;        // x, y, conic A, conic B
  float4* Codq;        // conic C, opacity, depth, cull threshold
  float4* rgbr;        // r, g, b, radius
  uint2* rect;         // x: rx0 | rx1<<16   y: ry0 | ry1<<16
  uint32_t* tiles;     // tile instances per Gaussian (after culling)
  uint32_t* dkey;      // depth key (0xFFFFFFFF: not visible)
  uint32_t* iota;
  uint32_t* dkey_s;
  uint32_t* order;     // Gaussian ids in depth order
  uint32_t* offs;      // inclusive scan of tiles[order[j]]
  uint8_t* clamped;
  float4* dacc;        // [3P] backward accumulators
  float* pose_part;    // [nblocks*16]
  float* pose_acc;     // [16]
  uint32_t* nrend;     // [1]
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static size_t cub_bytes_geom(int P) {
  size_t a = 0, b = 0;
  uint32_t* k = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, a, k, k, k, k, P, 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, b, k, k, P);
  return (a > b ? a : b) + 1024;
}

static GeomView geom_view(void* base, int P) {
  GeomView v;
  size_t off = 0;
  char* p = (char*)base;
  size_t Pp = (size_t)(P > 0 ? P : 1);
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  v.cam = (CamConst*)take(sizeof(CamConst));
  v.xyAB = (float4*)take(Pp * 16);
  v.Codq = (float4*)take(Pp * 16);
  v.rgbr = (float4*)take(Pp * 16);
  v.rect = (uint2*)take(Pp * 8);
  v.tiles = (uint32_t*)take(Pp * 4);
  v.dkey = (uint32_t*)take(Pp * 4);
  v.iota = (uint32_t*)take(Pp * 4);
  v.dkey_s = (uint32_t*)take(Pp * 4);
  v.order = (uint32_t*)take(Pp * 4);
  v.offs = (uint32_t*)take(Pp * 4);
  v.clamped = (uint8_t*)take(Pp);
  v.dacc = (float4*)take(Pp * 48);
  size_t nb = (Pp + kPT - 1) / kPT;
  v.pose_part = (float*)take(nb * 16 * 4);
  v.pose_acc = (float*)take(16 * 4);
  v.nrend = (uint32_t*)take(4);
  v.cub_bytes = cub_bytes_geom((int)Pp);
  v.cub_tmp = take(v.cub_bytes);
  v.total = off;
  return v;
}

typedef uint16_t tkey_t;   // tile id: 16 bits cover 65535 tiles (a 4K frame has 32400)
struct BinView {
  tkey_t* keys;
  tkey_t* keys_s;
  uint32_t* vals;
  uint32_t* vals_s;
  float4* s0;          // x, y, A, B
  float4* s1;          // C, opacity, cull threshold, gaussian id (bits)
  float4* s2;          // r, g, b, -
  uint2* ranges;       // [tiles]
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static int tile_bits(int ntiles) {
  int b = 1;
  while ((1 << b) < ntiles + 1) ++b;
  return b;
}

static BinView bin_view(void* base, int64_t R, int W, int H) {
  BinView v;
  size_t off = 0;
  char* p = (char*)base;
  size_t Rp = (size_t)(R > 0 ? R : 1);
  int ntiles = ((W + kBlock - 1) / kBlock) * ((H + kBlock - 1) / kBlock);
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  v.keys = (tkey_t*)take(Rp * sizeof(tkey_t));
  v.keys_s = (tkey_t*)take(Rp * sizeof(tkey_t));
  v.vals = (uint32_t*)take(Rp * 4);
  v.vals_s = (uint32_t*)take(Rp * 4);
  v.s0 = (float4*)take(Rp * 16);
  v.s1 = (float4*)take(Rp * 16);
  v.s2 = (float4*)take(Rp * 16);
  v.ranges = (uint2*)take((size_t)ntiles * 8);
  size_t a = 0;
  tkey_t* k = nullptr;
  uint32_t* vv = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, a, k, k, vv, vv, (int)Rp, 0, tile_bits(ntiles));
  v.cub_bytes = a + 1024;
  v.cub_tmp = take(v.cub_bytes);
  v.total = off;
  return v;
}

struct ImgView {
  float* final_T;
  uint32_t* n_contrib;
  size_t total;
};
static ImgView img_view(void* base, int W, int H) {
  ImgView v;
  size_t hw = (size_t)W * H;
  char* p = (char*)base;
  v.final_T = (float*)p;
  v.n_contrib = (uint32_t*)(p ? p + align_up(hw * 4) : nullptr);
  v.total = 2 * align_up(hw * 4);
  return v;
}

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
__global__ void k_setup_cam(CamConst* out, const float* V, const float* Pm, const float* campos,
                            const float* pose, int W, int H, float tanfovx, float tanfovy,
                            float scale_mod, int D, int M, int raw_params) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  CamConst c;
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

// Full-CTA fast path (aligned pointers, split SH layout with 16 coefficients): all 128-bit global loads of
// every array are issued before the first shared store, so each thread has ~15 independent 16-byte loads in
// flight (memory-level parallelism instead of a load->store loop).
template <int NF4, int MAXIT>
__device__ __forceinline__ void ld_batch(float4 (&r)[MAXIT], const float* __restrict__ src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < MAXIT; ++i) {
    const int k = threadIdx.x + i * kPT;
    if (k < NF4) r[i] = __ldg(s4 + k);
  }
}
template <int NF4, int MAXIT>
__device__ __forceinline__ void st_batch(const float4 (&r)[MAXIT], float* sm) {
  float4* d4 = reinterpret_cast<float4*>(sm);
#pragma unroll
  for (int i = 0; i < MAXIT; ++i) {
    const int k = threadIdx.x + i * kPT;
    if (k < NF4) d4[k] = r[i];
  }
}

__device__ __forceinline__ void load_block_inputs_full(const InPtrs& in, int first, bool use_sh, int D, float* sm) {
  constexpr int N3 = 3 * kPT / 4, N4 = kPT, N1 = kPT / 4, NR = 45 * kPT / 4;
  float4 rx[1], rs[1], rq[1], ro[1], rd[1], rr[(NR + kPT - 1) / kPT];
  ld_batch<N3, 1>(rx, in.means + (size_t)3 * first);
  ld_batch<N3, 1>(rs, in.scales + (size_t)3 * first);
  ld_batch<N4, 1>(rq, in.rots + (size_t)4 * first);
  ld_batch<N1, 1>(ro, in.opac + first);
  if (use_sh) {
    ld_batch<N3, 1>(rd, in.sh_dc + (size_t)3 * first);
    if (D > 0) ld_batch<NR, (NR + kPT - 1) / kPT>(rr, in.sh_rest + (size_t)45 * first);
  }
  st_batch<N3, 1>(rx, sm + kSmXyz);
  st_batch<N3, 1>(rs, sm + kSmSc);
  st_batch<N4, 1>(rq, sm + kSmQ);
  st_batch<N1, 1>(ro, sm + kSmOp);
  if (use_sh) {
    st_batch<N3, 1>(rd, sm + kSmSh);
    if (D > 0) st_batch<NR, (NR + kPT - 1) / kPT>(rr, sm + kSmSh + 3 * kPT);
  }
}

__device__ __forceinline__ void load_block_inputs(const InPtrs& in, int first, int nv, bool use_sh, int D, int M,
                                                  float* sm) {
  const bool vec = in.vec_ok != 0;
  if (vec && nv == kPT && in.scales && in.rots && (!use_sh || (!in.sh_packed && M == 16))) {
    load_block_inputs_full(in, first, use_sh, D, sm);
    return;
  }
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

__device__ __forceinline__ uint32_t count_or_emit_tiles(const Proj& p, float qthr, int W, int H, int gx,
                                                        bool cull, tkey_t* keys, uint32_t* vals,
                                                        uint32_t id) {
  uint32_t n = 0;
  for (int ty = p.ry0; ty < p.ry1; ++ty)
    for (int tx = p.rx0; tx < p.rx1; ++tx) {
      bool keep = true;
      if (cull) {
        float x0 = (float)(tx * kBlock), y0 = (float)(ty * kBlock);
        float x1 = fminf(x0 + kBlock - 1, (float)(W - 1)), y1 = fminf(y0 + kBlock - 1, (float)(H - 1));
        keep = rect_may_contribute(p.x, p.y, p.A, p.B, p.C, qthr, x0, y0, x1, y1);
      }
      if (keep) {
        if (keys) { keys[n] = (tkey_t)(ty * gx + tx); vals[n] = id; }
        ++n;
      }
    }
  return n;
}

// Gaussians whose rect spans more than kCoopTiles tiles are handled by the WHOLE WARP (one tile per lane per
// step, ballot-ranked so the emission order stays row-major) instead of one thread looping over thousands of
// tiles -- the per-thread loop is a performance cliff once a few Gaussians grow large.  Must be called by all
// 32 lanes; `mine` says whether this lane's Gaussian wants the cooperative path.  Returns the lane's count.
constexpr int kCoopTiles = 24;
__device__ __forceinline__ uint32_t coop_count_or_emit(bool mine, const Proj& p, float qthr, int W, int H, int gx,
                                                       tkey_t* keys, uint32_t* vals, uint32_t id) {
  const int lane = threadIdx.x & 31;
  uint32_t result = 0;
  unsigned big = __ballot_sync(0xffffffffu, mine);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const float bx = __shfl_sync(0xffffffffu, p.x, src), by = __shfl_sync(0xffffffffu, p.y, src);
    const float bA = __shfl_sync(0xffffffffu, p.A, src), bB = __shfl_sync(0xffffffffu, p.B, src);
    const float bC = __shfl_sync(0xffffffffu, p.C, src), bq = __shfl_sync(0xffffffffu, qthr, src);
    const int x0t = __shfl_sync(0xffffffffu, p.rx0, src), x1t = __shfl_sync(0xffffffffu, p.rx1, src);
    const int y0t = __shfl_sync(0xffffffffu, p.ry0, src), y1t = __shfl_sync(0xffffffffu, p.ry1, src);
    const uint32_t bid = __shfl_sync(0xffffffffu, id, src);
    unsigned long long kp = (unsigned long long)keys, vp = (unsigned long long)vals;
    kp = __shfl_sync(0xffffffffu, kp, src);
    vp = __shfl_sync(0xffffffffu, vp, src);
    tkey_t* bkeys = (tkey_t*)kp;
    uint32_t* bvals = (uint32_t*)vp;
    const int w = x1t - x0t, n = w * (y1t - y0t);
    uint32_t run = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      bool keep = false;
      int tx = 0, ty = 0;
      if (i < n) {
        ty = y0t + i / w; tx = x0t + i - (i / w) * w;
        const float x0 = (float)(tx * kBlock), y0 = (float)(ty * kBlock);
        const float x1 = fminf(x0 + kBlock - 1, (float)(W - 1)), y1 = fminf(y0 + kBlock - 1, (float)(H - 1));
        keep = rect_may_contribute(bx, by, bA, bB, bC, bq, x0, y0, x1, y1);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep && bkeys) {
        const uint32_t pos = run + __popc(m & ((1u << lane) - 1u));
        bkeys[pos] = (tkey_t)(ty * gx + tx);
        bvals[pos] = bid;
      }
      run += __popc(m);
    }
    if (lane == src) result = run;
  }
  return result;
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
  int first = blockIdx.x * kPT;
  int nv = min(kPT, in.P - first);
  bool use_sh = in.colors == nullptr;
  __syncthreads();
  load_block_inputs(in, first, nv, use_sh, cam->D, cam->M, sm);
  __syncthreads();
  const int t = threadIdx.x;
  const bool active = t < nv;
  const int i = first + (active ? t : 0);
  GaussIn g;
  read_gauss(sm, active ? t : 0, in.scales != nullptr, g);
  Proj p;
  project_geometry(*cam, g, in.cov3D ? in.cov3D + (size_t)6 * i : nullptr, p);
  if (!active) p.visible = 0;
  float qthr = -1.f;
  uint32_t ntiles = 0;
  bool coop = false;
  if (p.visible) {
    if (use_sh) {
      const ShRows sr = sh_rows(in.sh_packed, cam->M);
      project_color(*cam, sm + kSmSh + sr.dc_stride * t, sm + kSmSh + sr.rest_off + sr.rest_stride * t, p);
    } else {
      p.rgb[0] = in.colors[3 * i]; p.rgb[1] = in.colors[3 * i + 1]; p.rgb[2] = in.colors[3 * i + 2];
    }
    qthr = cull_threshold(p.opacity);
    const int area = (p.rx1 - p.rx0) * (p.ry1 - p.ry0);
    if (in.exact_cull == 0) ntiles = (uint32_t)area;
    else if (qthr < 0.f) ntiles = 0;
    else if (area > kCoopTiles) coop = true;
    else ntiles = count_or_emit_tiles(p, qthr, cam->W, cam->H, cam->gx, true, nullptr, nullptr, 0);
  } else {
    p.x = p.y = p.A = p.B = p.C = 0.f; p.rgb[0] = p.rgb[1] = p.rgb[2] = 0.f;
  }
  {
    const uint32_t c = coop_count_or_emit(coop, p, qthr, cam->W, cam->H, cam->gx, nullptr, nullptr, 0);
    if (coop) ntiles = c;
  }
  if (!active) return;
  gv.xyAB[i] = make_float4(p.x, p.y, p.A, p.B);
  gv.Codq[i] = make_float4(p.C, p.opacity, p.depth, qthr);
  gv.rgbr[i] = make_float4(p.rgb[0], p.rgb[1], p.rgb[2], (float)p.radius);
  gv.rect[i] = make_uint2((uint32_t)p.rx0 | ((uint32_t)p.rx1 << 16), (uint32_t)p.ry0 | ((uint32_t)p.ry1 << 16));
  gv.tiles[i] = ntiles;
  gv.dkey[i] = p.visible ? __float_as_uint(p.depth) : 0xFFFFFFFFu;
  gv.iota[i] = (uint32_t)i;
  gv.clamped[i] = (uint8_t)p.clamped;
  radii[i] = p.radius;
}

struct TilesInOrder {
  const uint32_t* tiles;
  const uint32_t* order;
  __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& j) const { return tiles[order[j]]; }
};

__global__ void k_store_total(const uint32_t* offs, int P, uint32_t* nrend) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *nrend = P > 0 ? offs[P - 1] : 0u;
}

// ------------------------------------------------------------------------------------------
// k_duplicate: one thread per depth-ordered Gaussian
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_duplicate(int P, GeomView gv, BinView bv, int W, int H, int gx, int exact_cull, uint32_t R) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i = 0, n = 0, off = 0;
  if (j < P) {
    i = gv.order[j];
    n = gv.tiles[i];
    if (n) {
      off = gv.offs[j] - n;
      if (off + n > R) n = 0;   // defensive: never write past the binning buffer
    }
  }
  Proj p;
  p.x = p.y = p.A = p.B = p.C = 0.f;
  p.rx0 = p.rx1 = p.ry0 = p.ry1 = 0;
  float qthr = -1.f;
  bool coop = false;
  if (n) {
    const float4 a = gv.xyAB[i], b = gv.Codq[i];
    const uint2 rc = gv.rect[i];
    p.x = a.x; p.y = a.y; p.A = a.z; p.B = a.w; p.C = b.x;
    p.rx0 = rc.x & 0xFFFF; p.rx1 = rc.x >> 16; p.ry0 = rc.y & 0xFFFF; p.ry1 = rc.y >> 16;
    qthr = b.w;
    const int area = (p.rx1 - p.rx0) * (p.ry1 - p.ry0);
    coop = exact_cull != 0 && area > kCoopTiles;
    if (!coop) count_or_emit_tiles(p, qthr, W, H, gx, exact_cull != 0, bv.keys + off, bv.vals + off, i);
  }
  coop_count_or_emit(coop, p, qthr, W, H, gx, bv.keys + off, bv.vals + off, i);
}

// ------------------------------------------------------------------------------------------
// k_ranges_gather: tile ranges + contiguous per-tile slabs
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_ranges_gather(uint32_t R, GeomView gv, BinView bv) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R) return;
  uint32_t tile = bv.keys_s[e];
  uint32_t i = bv.vals_s[e];
  if (e == 0 || bv.keys_s[e - 1] != tile) bv.ranges[tile].x = e;
  if (e == R - 1 || bv.keys_s[e + 1] != tile) bv.ranges[tile].y = e + 1;
  float4 a = gv.xyAB[i], b = gv.Codq[i], c = gv.rgbr[i];
  // conic pre-scaled into the log2 domain: power*log2(e) = A' dx^2 + B' dx dy + C' dy^2
  bv.s0[e] = make_float4(a.x, a.y, -0.5f * kLog2e * a.z, -kLog2e * a.w);
  bv.s1[e] = make_float4(-0.5f * kLog2e * b.x, b.y, 0.5f * kLog2e * b.w, __uint_as_float(i));
  bv.s2[e] = make_float4(c.x, c.y, c.z, 0.f);
}

// ------------------------------------------------------------------------------------------
// blend
// ------------------------------------------------------------------------------------------
struct PairEval { float dx, dy, power, G, alpha; };   // power is in the log2 domain (power * log2 e)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ PairEval pair_eval(const float4& e0, const float4& e1, float fx, float fy) {
  PairEval r;
  r.dx = e0.x - fx;
  r.dy = e0.y - fy;
  r.power = e0.z * r.dx * r.dx + (e0.w * r.dx + e1.x * r.dy) * r.dy;
  r.G = ex2_approx(r.power);   // MUFU.EX2 directly (valid pairs have power >= -8)
  r.alpha = fminf(0.99f, e1.y * r.G);
  return r;
}

// cull test on a slab entry: q' = -(A'dx^2 + B'dxdy + C'dy^2) = q * log2(e)/2 against qthr' = qthr * log2(e)/2
__device__ __forceinline__ bool slab_may_contribute(const float4& e0, const float4& e1, float rx0, float ry0,
                                                    float rx1, float ry1) {
  return rect_may_contribute(e0.x, e0.y, -e0.z, -0.5f * e0.w, -e1.x, e1.z, rx0, ry0, rx1, ry1);
}

__global__ void __launch_bounds__(kThreads)
k_blend_fwd(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
            const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
            float* __restrict__ out_color, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
  __shared__ float4 sm0[kChunk1], sm1[kChunk1], sm2[kChunk1];
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + (warp & 1) * 8, sy0 = ty * kBlock + (warp >> 1) * 4;
  const int px = sx0 + (lane & 7), py = sy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float fx = (float)px, fy = (float)py;
  const float rx0 = (float)sx0, ry0 = (float)sy0;
  const float rx1 = (float)min(sx0 + 7, W - 1), ry1 = (float)min(sy0 + 3, H - 1);
  const uint2 rg = ranges[tile];
  const int n = (int)(rg.y - rg.x);
  float T = 1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f;
  uint32_t last = 0;
  bool done = !inside;
  bool wdone = !(sx0 < W && sy0 < H);
  for (int base = 0; base < n; base += kChunk1) {
    const int cnt = min(kChunk1, n - base);
    if ((int)threadIdx.x < cnt) {
      size_t e = (size_t)rg.x + base + threadIdx.x;
      sm0[threadIdx.x] = s0[e];
      sm1[threadIdx.x] = s1[e];
      sm2[threadIdx.x] = s2[e];
    }
    __syncthreads();
    if (!wdone) {
      for (int b = 0; b < cnt; b += 32) {
        const int j = b + lane;
        bool hit = false;
        if (j < cnt) {
          float4 e0 = sm0[j], e1 = sm1[j];
          hit = slab_may_contribute(e0, e1, rx0, ry0, rx1, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k];
          PairEval pe = pair_eval(e0, e1, fx, fy);
          bool valid = !done && pe.power <= 0.f && pe.alpha >= kAlphaMin;
          float testT = T * (1.f - pe.alpha);
          if (valid && testT < kTEps) { done = true; valid = false; }
          if (valid) {
            const float4 c = sm2[b + k];
            float w = pe.alpha * T;
            Cr += c.x * w; Cg += c.y * w; Cb += c.z * w;
            T = testT;
            last = (uint32_t)(base + b + k + 1);
          }
        }
        if (__all_sync(0xffffffffu, done)) { wdone = true; break; }
      }
    }
    if (__syncthreads_and(wdone)) break;
  }
  if (inside) {
    size_t pix = (size_t)py * W + px, hw = (size_t)W * H;
    final_T[pix] = T;
    n_contrib[pix] = last;
    out_color[pix] = Cr + T * bg[0];
    out_color[hw + pix] = Cg + T * bg[1];
    out_color[2 * hw + pix] = Cb + T * bg[2];
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Reduce 9 per-lane values over the warp with 14 shuffles.  On return, lane l with (l & 3) == 0
// holds the total of v[l >> 2] in v[0]; every lane holds the total of v[8] in v[8].
__device__ __forceinline__ void warp_reduce9(float* v, int lane) {
  const unsigned full = 0xffffffffu;
  float b[4], c[2], d;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float send = hi ? v[k] : v[k + 4];
      float keep = hi ? v[k + 4] : v[k];
      b[k] = keep + __shfl_xor_sync(full, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float send = hi ? b[k] : b[k + 2];
      float keep = hi ? b[k + 2] : b[k];
      c[k] = keep + __shfl_xor_sync(full, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    float send = hi ? c[0] : c[1];
    float keep = hi ? c[1] : c[0];
    d = keep + __shfl_xor_sync(full, send, 4);
  }
  d += __shfl_xor_sync(full, d, 2);
  d += __shfl_xor_sync(full, d, 1);
  v[0] = d;
  float e = v[8];
  e += __shfl_xor_sync(full, e, 16);
  e += __shfl_xor_sync(full, e, 8);
  e += __shfl_xor_sync(full, e, 4);
  e += __shfl_xor_sync(full, e, 2);
  e += __shfl_xor_sync(full, e, 1);
  v[8] = e;
}

__global__ void __launch_bounds__(kThreads)
k_blend_bwd(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
            const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
            const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
            const float* __restrict__ dL_dpix, float* __restrict__ dacc) {
  __shared__ float4 sm0[kChunk1], sm1[kChunk1], sm2[kChunk1];
  __shared__ int s_bmax;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + (warp & 1) * 8, sy0 = ty * kBlock + (warp >> 1) * 4;
  const int px = sx0 + (lane & 7), py = sy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float fx = (float)px, fy = (float)py;
  const float rx0 = (float)sx0, ry0 = (float)sy0;
  const float rx1 = (float)min(sx0 + 7, W - 1), ry1 = (float)min(sy0 + 3, H - 1);
  const uint2 rg = ranges[tile];
  const size_t pix = (size_t)py * W + px, hw = (size_t)W * H;
  const float T_final = inside ? final_T[pix] : 0.f;
  const int last_contrib = inside ? (int)n_contrib[pix] : 0;
  float dLr = 0.f, dLg = 0.f, dLb = 0.f;
  if (inside) { dLr = dL_dpix[pix]; dLg = dL_dpix[hw + pix]; dLb = dL_dpix[2 * hw + pix]; }
  const float bg_dot = bg[0] * dLr + bg[1] * dLg + bg[2] * dLb;
  int wmax = last_contrib;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (threadIdx.x == 0) s_bmax = 0;
  __syncthreads();
  if (lane == 0) atomicMax(&s_bmax, wmax);
  __syncthreads();
  const int bmax = s_bmax;
  float T = T_final;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, last_alpha = 0.f, last_r = 0.f, last_g = 0.f, last_b = 0.f;
  const int nchunks = (bmax + kChunk1 - 1) / kChunk1;
  for (int ch = nchunks - 1; ch >= 0; --ch) {
    const int base = ch * kChunk1;
    const int cnt = min(kChunk1, bmax - base);
    if ((int)threadIdx.x < cnt) {
      size_t e = (size_t)rg.x + base + threadIdx.x;
      sm0[threadIdx.x] = s0[e];
      sm1[threadIdx.x] = s1[e];
      sm2[threadIdx.x] = s2[e];
    }
    __syncthreads();
    if (base < wmax) {
      for (int b = (cnt - 1) & ~31; b >= 0; b -= 32) {
        if (base + b >= wmax) continue;
        const int j = b + lane;
        bool hit = false;
        if (j < cnt && base + j < wmax) {
          float4 e0 = sm0[j], e1 = sm1[j];
          hit = slab_may_contribute(e0, e1, rx0, ry0, rx1, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);
          mask &= ~(1u << k);
          const int pos = base + b + k;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k];
          PairEval pe = pair_eval(e0, e1, fx, fy);
          const bool valid = inside && pos < last_contrib && pe.power <= 0.f && pe.alpha >= kAlphaMin;
          if (!__any_sync(0xffffffffu, valid)) continue;
          float v[9];
#pragma unroll
          for (int u = 0; u < 9; ++u) v[u] = 0.f;
          if (valid) {
            const float4 c = sm2[b + k];
            const float inv1ma = rcp_approx(1.f - pe.alpha);      // 1-alpha in [0.01, 1]
            T = T * inv1ma;
            const float dchannel_dcolor = pe.alpha * T;
            acc_r = last_alpha * last_r + (1.f - last_alpha) * acc_r; last_r = c.x;
            acc_g = last_alpha * last_g + (1.f - last_alpha) * acc_g; last_g = c.y;
            acc_b = last_alpha * last_b + (1.f - last_alpha) * acc_b; last_b = c.z;
            float dL_dalpha = (c.x - acc_r) * dLr + (c.y - acc_g) * dLg + (c.z - acc_b) * dLb;
            dL_dalpha *= T;
            last_alpha = pe.alpha;
            dL_dalpha -= T_final * inv1ma * bg_dot;
            const float dL_dG = e1.y * dL_dalpha;
            const float gdx = pe.G * pe.dx, gdy = pe.G * pe.dy;
            // -A = 2 ln2 A', -B = ln2 B', -C = 2 ln2 C'
            v[0] = dL_dG * kLn2 * (2.f * gdx * e0.z + gdy * e0.w);
            v[1] = dL_dG * kLn2 * (2.f * gdy * e1.x + gdx * e0.w);
            v[2] = -0.5f * gdx * pe.dx * dL_dG;
            v[3] = -gdx * pe.dy * dL_dG;
            v[4] = -0.5f * gdy * pe.dy * dL_dG;
            v[5] = pe.G * dL_dalpha;
            v[6] = dchannel_dcolor * dLr;
            v[7] = dchannel_dcolor * dLg;
            v[8] = dchannel_dcolor * dLb;
          }
          warp_reduce9(v, lane);
          // lane 4i holds total i (i < 8): pull 4 consecutive totals into lanes 0 and 16 and issue two
          // 128-bit reductions (REDG.E.ADD.F32x4) + one scalar instead of nine scalar atomics
          const float a1 = __shfl_down_sync(0xffffffffu, v[0], 4);
          const float a2 = __shfl_down_sync(0xffffffffu, v[0], 8);
          const float a3 = __shfl_down_sync(0xffffffffu, v[0], 12);
          float* dst = dacc + (size_t)__float_as_uint(e1.w) * 12;
          if ((lane & 15) == 0) red_add_v4(dst + (lane >> 2), v[0], a1, a2, a3);
          if (lane == 1) atomicAdd(dst + 8, v[8]);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// blend v2: 4 warps per tile, each warp owns an 8x8 block = two 8x4 halves; every lane carries TWO
// pixels (x, y) and (x, y+4) that share dx, and the per-pair arithmetic is issued as packed
// FFMA2/FMUL2/FADD2 (Blackwell f32x2), so one instruction stream serves 64 (pixel, Gaussian) pairs.
// The sub-tile cull is evaluated per half and the loop runs over the union of the two masks.
// ------------------------------------------------------------------------------------------
constexpr int kThreads2 = 128;

// ------------------------------------------------------------------------------------------
// TMA (bulk async copy) staging of the per-tile slabs: one elected thread arms an mbarrier with the
// byte count and issues cp.async.bulk.shared.global for the three slab arrays of the NEXT chunk
// while the CTA blends the current one (double buffered).  SASS: UBLKCP + SYNCS.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct SlabStage {
  float4 s0[kChunk], s1[kChunk], s2[kChunk];
};

// Stage `cnt` slab entries starting at global entry `e0` into `dst`.  BULK: thread 0 issues three bulk copies
// that complete on `bar`; otherwise all threads copy cooperatively (caller synchronises).
template <bool BULK>
__device__ __forceinline__ void stage_slab(SlabStage* dst, const float4* __restrict__ s0, const float4* __restrict__ s1,
                                           const float4* __restrict__ s2, size_t e0, int cnt, uint64_t* bar,
                                           int nthreads) {
  if (BULK) {
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)cnt * 16u;
      mbar_expect_tx(bar, 3u * bytes);
      bulk_g2s(dst->s0, s0 + e0, bytes, bar);
      bulk_g2s(dst->s1, s1 + e0, bytes, bar);
      bulk_g2s(dst->s2, s2 + e0, bytes, bar);
    }
  } else {
    for (int k = threadIdx.x; k < cnt; k += nthreads) {
      dst->s0[k] = s0[e0 + k];
      dst->s1[k] = s1[e0 + k];
      dst->s2[k] = s2[e0 + k];
    }
  }
}



__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }

template <bool BULK>
__global__ void __launch_bounds__(kThreads2, GSB_FWD_MINB)
k_blend_fwd2(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
             const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
             float* __restrict__ out_color, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
  __shared__ __align__(128) SlabStage stg[BULK ? 2 : 1];
  __shared__ __align__(8) uint64_t bars[2];
  if (BULK) {
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
  }
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + (warp & 1) * 8, sy0 = ty * kBlock + (warp >> 1) * 8;
  const int px = sx0 + (lane & 7), pyA = sy0 + (lane >> 3), pyB = pyA + 4;
  const bool inA = px < W && pyA < H, inB = px < W && pyB < H;
  const float fx = (float)px;
  const float2 fy = f2((float)pyA, (float)pyB);
  const float rx0 = (float)sx0, rx1 = (float)min(sx0 + 7, W - 1);
  const float ryA0 = (float)sy0, ryA1 = (float)min(sy0 + 3, H - 1);
  const float ryB0 = (float)(sy0 + 4), ryB1 = (float)min(sy0 + 7, H - 1);
  const uint2 rg = ranges[tile];
  const int n = (int)(rg.y - rg.x);
  float2 T = f2(1.f, 1.f), Cr = f2(0.f, 0.f), Cg = f2(0.f, 0.f), Cb = f2(0.f, 0.f);
  uint32_t lastA = 0, lastB = 0;
  bool doneA = !inA, doneB = !inB;
  bool wdoneA = !(sx0 < W && sy0 < H), wdoneB = !(sx0 < W && sy0 + 4 < H);
  const int nch = (n + kChunk - 1) / kChunk;
  if (BULK && nch > 0) stage_slab<true>(&stg[0], s0, s1, s2, (size_t)rg.x, min(kChunk, n), &bars[0], kThreads2);
  int pending = -1;    // chunk whose bulk copy is in flight but has not been waited for
  for (int ci = 0; ci < nch; ++ci) {
    const int base = ci * kChunk;
    const int cnt = min(kChunk, n - base);
    const SlabStage* cur = &stg[BULK ? (ci & 1) : 0];
    if (BULK) {
      pending = -1;
      if (ci + 1 < nch) {     // prefetch the next chunk into the other stage (freed by the barrier below)
        stage_slab<true>(&stg[(ci + 1) & 1], s0, s1, s2, (size_t)rg.x + base + kChunk, min(kChunk, n - base - kChunk),
                         &bars[(ci + 1) & 1], kThreads2);
        pending = ci + 1;
      }
      mbar_wait(&bars[ci & 1], (uint32_t)((ci >> 1) & 1));
    } else {
      stage_slab<false>(&stg[0], s0, s1, s2, (size_t)rg.x + base, cnt, nullptr, kThreads2);
      __syncthreads();
    }
    const float4* sm0 = cur->s0;
    const float4* sm1 = cur->s1;
    const float4* sm2 = cur->s2;
    if (!(wdoneA && wdoneB)) {
      for (int b = 0; b < cnt; b += 32) {
        const int j = b + lane;
        bool hitA = false, hitB = false;
        if (j < cnt) {
          const float4 e0 = sm0[j], e1 = sm1[j];
          if (!wdoneA) hitA = slab_may_contribute(e0, e1, rx0, ryA0, rx1, ryA1);
          if (!wdoneB) hitB = slab_may_contribute(e0, e1, rx0, ryB0, rx1, ryB1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hitA || hitB);
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k], c = sm2[b + k];
          const float dx = e0.x - fx;
          const float2 dy = f2(e0.y - fy.x, e0.y - fy.y);
          const float c1 = e0.w * dx, c0 = e0.z * dx * dx;
          // power' = c0 + dy * (c1 + C' * dy)
          const float2 pw = __ffma2_rn(dy, __ffma2_rn(f2s(e1.x), dy, f2s(c1)), f2s(c0));
          const float2 G = f2(ex2_approx(pw.x), ex2_approx(pw.y));
          float2 al = __fmul2_rn(f2s(e1.y), G);
          al.x = fminf(0.99f, al.x); al.y = fminf(0.99f, al.y);
          bool vA = !doneA && pw.x <= 0.f && al.x >= kAlphaMin;
          bool vB = !doneB && pw.y <= 0.f && al.y >= kAlphaMin;
          const float2 tT = __fmul2_rn(T, __ffma2_rn(al, f2s(-1.f), f2s(1.f)));
          if (vA && tT.x < kTEps) { doneA = true; vA = false; }
          if (vB && tT.y < kTEps) { doneB = true; vB = false; }
          float2 w = __fmul2_rn(al, T);
          w.x = vA ? w.x : 0.f; w.y = vB ? w.y : 0.f;
          Cr = __ffma2_rn(f2s(c.x), w, Cr);
          Cg = __ffma2_rn(f2s(c.y), w, Cg);
          Cb = __ffma2_rn(f2s(c.z), w, Cb);
          const uint32_t pos = (uint32_t)(base + b + k + 1);
          T.x = vA ? tT.x : T.x; T.y = vB ? tT.y : T.y;
          lastA = vA ? pos : lastA; lastB = vB ? pos : lastB;
        }
        wdoneA = __all_sync(0xffffffffu, doneA);
        wdoneB = __all_sync(0xffffffffu, doneB);
        if (wdoneA && wdoneB) break;
      }
    }
    if (__syncthreads_and(wdoneA && wdoneB)) break;
    pending = -1;
  }
  if (BULK && pending >= 0) mbar_wait(&bars[pending & 1], (uint32_t)((pending >> 1) & 1));   // never exit with a copy in flight
  const size_t hw = (size_t)W * H;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  if (inA) {
    size_t pix = (size_t)pyA * W + px;
    final_T[pix] = T.x; n_contrib[pix] = lastA;
    out_color[pix] = Cr.x + T.x * b0; out_color[hw + pix] = Cg.x + T.x * b1; out_color[2 * hw + pix] = Cb.x + T.x * b2;
  }
  if (inB) {
    size_t pix = (size_t)pyB * W + px;
    final_T[pix] = T.y; n_contrib[pix] = lastB;
    out_color[pix] = Cr.y + T.y * b0; out_color[hw + pix] = Cg.y + T.y * b1; out_color[2 * hw + pix] = Cb.y + T.y * b2;
  }
}

template <bool BULK>
__global__ void __launch_bounds__(kThreads2, GSB_BWD_MINB)
k_blend_bwd2(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
             const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
             const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
             const float* __restrict__ dL_dpix, float* __restrict__ dacc) {
  __shared__ __align__(128) SlabStage stg[BULK ? 2 : 1];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ int s_bmax;
  if (BULK && threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + (warp & 1) * 8, sy0 = ty * kBlock + (warp >> 1) * 8;
  const int px = sx0 + (lane & 7), pyA = sy0 + (lane >> 3), pyB = pyA + 4;
  const bool inA = px < W && pyA < H, inB = px < W && pyB < H;
  const float fx = (float)px;
  const float2 fy = f2((float)pyA, (float)pyB);
  const float rx0 = (float)sx0, rx1 = (float)min(sx0 + 7, W - 1);
  const float ryA0 = (float)sy0, ryA1 = (float)min(sy0 + 3, H - 1);
  const float ryB0 = (float)(sy0 + 4), ryB1 = (float)min(sy0 + 7, H - 1);
  const uint2 rg = ranges[tile];
  const size_t hw = (size_t)W * H;
  const size_t pixA = (size_t)pyA * W + px, pixB = (size_t)pyB * W + px;
  const float2 T_final = f2(inA ? final_T[pixA] : 0.f, inB ? final_T[pixB] : 0.f);
  const int lcA = inA ? (int)n_contrib[pixA] : 0, lcB = inB ? (int)n_contrib[pixB] : 0;
  float2 dLr = f2(0.f, 0.f), dLg = f2(0.f, 0.f), dLb = f2(0.f, 0.f);
  if (inA) { dLr.x = dL_dpix[pixA]; dLg.x = dL_dpix[hw + pixA]; dLb.x = dL_dpix[2 * hw + pixA]; }
  if (inB) { dLr.y = dL_dpix[pixB]; dLg.y = dL_dpix[hw + pixB]; dLb.y = dL_dpix[2 * hw + pixB]; }
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  const float2 tf_bg = __fmul2_rn(T_final, f2(b0 * dLr.x + b1 * dLg.x + b2 * dLb.x, b0 * dLr.y + b1 * dLg.y + b2 * dLb.y));
  int wmaxA = lcA, wmaxB = lcB;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmaxA = max(wmaxA, __shfl_xor_sync(0xffffffffu, wmaxA, o));
    wmaxB = max(wmaxB, __shfl_xor_sync(0xffffffffu, wmaxB, o));
  }
  const int wmax = max(wmaxA, wmaxB);
  if (threadIdx.x == 0) s_bmax = 0;
  __syncthreads();
  if (lane == 0) atomicMax(&s_bmax, wmax);
  __syncthreads();
  const int bmax = s_bmax;
  float2 T = T_final;
  float2 acc_r = f2(0.f, 0.f), acc_g = f2(0.f, 0.f), acc_b = f2(0.f, 0.f);
  const int nchunks = (bmax + kChunk - 1) / kChunk;
  // chunks are visited back to front; it = 0 is the LAST chunk
  if (BULK && nchunks > 0)
    stage_slab<true>(&stg[0], s0, s1, s2, (size_t)rg.x + (size_t)(nchunks - 1) * kChunk,
                     min(kChunk, bmax - (nchunks - 1) * kChunk), &bars[0], kThreads2);
  for (int it = 0; it < nchunks; ++it) {
    const int ch = nchunks - 1 - it;
    const int base = ch * kChunk;
    const int cnt = min(kChunk, bmax - base);
    const SlabStage* cur = &stg[BULK ? (it & 1) : 0];
    if (BULK) {
      if (it + 1 < nchunks)
        stage_slab<true>(&stg[(it + 1) & 1], s0, s1, s2, (size_t)rg.x + base - kChunk, kChunk, &bars[(it + 1) & 1],
                         kThreads2);
      mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
    } else {
      stage_slab<false>(&stg[0], s0, s1, s2, (size_t)rg.x + base, cnt, nullptr, kThreads2);
      __syncthreads();
    }
    const float4* sm0 = cur->s0;
    const float4* sm1 = cur->s1;
    const float4* sm2 = cur->s2;
    if (base < wmax) {
      for (int b = (cnt - 1) & ~31; b >= 0; b -= 32) {
        if (base + b >= wmax) continue;
        const int j = b + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 e0 = sm0[j], e1 = sm1[j];
          if (base + j < wmaxA) hit = slab_may_contribute(e0, e1, rx0, ryA0, rx1, ryA1);
          if (!hit && base + j < wmaxB) hit = slab_may_contribute(e0, e1, rx0, ryB0, rx1, ryB1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);
          mask &= ~(1u << k);
          const int pos = base + b + k;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k];
          const float dx = e0.x - fx;
          const float2 dy = f2(e0.y - fy.x, e0.y - fy.y);
          const float c1 = e0.w * dx, c0 = e0.z * dx * dx;
          const float2 pw = __ffma2_rn(dy, __ffma2_rn(f2s(e1.x), dy, f2s(c1)), f2s(c0));
          const float2 G = f2(ex2_approx(pw.x), ex2_approx(pw.y));
          float2 al = __fmul2_rn(f2s(e1.y), G);
          al.x = fminf(0.99f, al.x); al.y = fminf(0.99f, al.y);
          const bool vA = inA && pos < lcA && pw.x <= 0.f && al.x >= kAlphaMin;
          const bool vB = inB && pos < lcB && pw.y <= 0.f && al.y >= kAlphaMin;
          if (!__any_sync(0xffffffffu, vA || vB)) continue;
          const float4 c = sm2[b + k];
          // masked alpha: an invalid pixel behaves as alpha = 0 (T, accumulator and gradients unchanged)
          const float2 am = f2(vA ? al.x : 0.f, vB ? al.y : 0.f);
          const float2 one_m = __ffma2_rn(am, f2s(-1.f), f2s(1.f));
          const float2 inv = f2(rcp_approx(one_m.x), rcp_approx(one_m.y));
          T = __fmul2_rn(T, inv);
          const float2 dcol = __fmul2_rn(am, T);                 // dchannel/dcolor = alpha * T
          // acc = colour composited from everything BEHIND this entry (the reference's accum_rec);
          // dL/dalpha = T * sum_ch (c - acc) dL_ch  -  T_final/(1-alpha) * bg.dL ; then acc += alpha (c - acc)
          const float2 d_r = __fadd2_rn(f2s(c.x), f2(-acc_r.x, -acc_r.y));
          const float2 d_g = __fadd2_rn(f2s(c.y), f2(-acc_g.x, -acc_g.y));
          const float2 d_b = __fadd2_rn(f2s(c.z), f2(-acc_b.x, -acc_b.y));
          float2 da = __ffma2_rn(d_b, dLb, __ffma2_rn(d_g, dLg, __fmul2_rn(d_r, dLr)));
          da = __fmul2_rn(da, T);
          da = __ffma2_rn(f2(-tf_bg.x, -tf_bg.y), inv, da);
          da.x = vA ? da.x : 0.f; da.y = vB ? da.y : 0.f;
          acc_r = __ffma2_rn(am, d_r, acc_r);
          acc_g = __ffma2_rn(am, d_g, acc_g);
          acc_b = __ffma2_rn(am, d_b, acc_b);
          const float2 dG = __fmul2_rn(f2s(e1.y), da);           // dL/dG = opacity * dL/dalpha
          const float2 gdG = __fmul2_rn(G, dG);                  // G * dL/dG
          const float2 gy = __fmul2_rn(gdG, dy);                 // G dL/dG dy
          const float gxs = (gdG.x + gdG.y) * dx;                // sum over the two pixels of G dL/dG dx
          const float gys = gy.x + gy.y;
          float v[9];
          v[0] = kLn2 * (2.f * gxs * e0.z + gys * e0.w);
          v[1] = kLn2 * (2.f * gys * e1.x + gxs * e0.w);
          v[2] = -0.5f * gxs * dx;
          v[3] = -dx * gys;
          v[4] = -0.5f * (gy.x * dy.x + gy.y * dy.y);
          v[5] = G.x * da.x + G.y * da.y;
          const float dcs_r = dcol.x * dLr.x + dcol.y * dLr.y;
          v[6] = dcs_r;
          v[7] = dcol.x * dLg.x + dcol.y * dLg.y;
          v[8] = dcol.x * dLb.x + dcol.y * dLb.y;
          warp_reduce9(v, lane);
          const float a1 = __shfl_down_sync(0xffffffffu, v[0], 4);
          const float a2 = __shfl_down_sync(0xffffffffu, v[0], 8);
          const float a3 = __shfl_down_sync(0xffffffffu, v[0], 12);
          float* dst = dacc + (size_t)__float_as_uint(e1.w) * 12;
          if ((lane & 15) == 0) red_add_v4(dst + (lane >> 2), v[0], a1, a2, a3);
          if (lane == 1) atomicAdd(dst + 8, v[8]);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// blend v3: 2 warps per tile, each warp owns an 8x16 block = four 8x4 quarters; every lane carries FOUR
// pixels (x, y + 4q) sharing dx, processed as two packed f32x2 pairs.  One instruction stream serves 128
// (pixel, Gaussian) pairs, so the per-Gaussian overheads (loop, slab reads, the 9-value warp reduction and
// the REDs in the backward) are amortised over twice as many pairs as in v2.  Culling stays per 8x4 quarter.
// ------------------------------------------------------------------------------------------
constexpr int kThreads3 = 64;
#ifndef GSB_FWD3_MINB
#define GSB_FWD3_MINB 10
#endif
#ifndef GSB_BWD3_MINB
#define GSB_BWD3_MINB 8
#endif

template <bool BULK>
__global__ void __launch_bounds__(kThreads3, GSB_FWD3_MINB)
k_blend_fwd3(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
             const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
             float* __restrict__ out_color, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
  __shared__ __align__(128) SlabStage stg[BULK ? 2 : 1];
  __shared__ __align__(8) uint64_t bars[2];
  if (BULK) {
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
  }
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + warp * 8, sy0 = ty * kBlock;
  const int px = sx0 + (lane & 7), py0 = sy0 + (lane >> 3);
  const float fx = (float)px;
  const float rx0 = (float)sx0, rx1 = (float)min(sx0 + 7, W - 1);
  float2 fy[2];
  bool in[4], done[4], wdone[4];
  float ry0[4], ry1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int py = py0 + 4 * q;
    in[q] = px < W && py < H;
    done[q] = !in[q];
    wdone[q] = !(sx0 < W && sy0 + 4 * q < H);
    ry0[q] = (float)(sy0 + 4 * q);
    ry1[q] = (float)min(sy0 + 4 * q + 3, H - 1);
  }
  fy[0] = f2((float)py0, (float)(py0 + 4));
  fy[1] = f2((float)(py0 + 8), (float)(py0 + 12));
  const uint2 rg = ranges[tile];
  const int n = (int)(rg.y - rg.x);
  float2 T[2], Cr[2], Cg[2], Cb[2];
  uint32_t last[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; ++h) { T[h] = f2(1.f, 1.f); Cr[h] = Cg[h] = Cb[h] = f2(0.f, 0.f); }
  const int nch = (n + kChunk - 1) / kChunk;
  if (BULK && nch > 0) stage_slab<true>(&stg[0], s0, s1, s2, (size_t)rg.x, min(kChunk, n), &bars[0], kThreads3);
  int pending = -1;
  for (int ci = 0; ci < nch; ++ci) {
    const int base = ci * kChunk;
    const int cnt = min(kChunk, n - base);
    const SlabStage* cur = &stg[BULK ? (ci & 1) : 0];
    if (BULK) {
      pending = -1;
      if (ci + 1 < nch) {
        stage_slab<true>(&stg[(ci + 1) & 1], s0, s1, s2, (size_t)rg.x + base + kChunk, min(kChunk, n - base - kChunk),
                         &bars[(ci + 1) & 1], kThreads3);
        pending = ci + 1;
      }
      mbar_wait(&bars[ci & 1], (uint32_t)((ci >> 1) & 1));
    } else {
      stage_slab<false>(&stg[0], s0, s1, s2, (size_t)rg.x + base, cnt, nullptr, kThreads3);
      __syncthreads();
    }
    const float4* sm0 = cur->s0;
    const float4* sm1 = cur->s1;
    const float4* sm2 = cur->s2;
    bool wall = wdone[0] && wdone[1] && wdone[2] && wdone[3];
    if (!wall) {
      for (int b = 0; b < cnt; b += 32) {
        const int j = b + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 e0 = sm0[j], e1 = sm1[j];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (!hit && !wdone[q]) hit = slab_may_contribute(e0, e1, rx0, ry0[q], rx1, ry1[q]);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k], c = sm2[b + k];
          const float dx = e0.x - fx;
          const float c1 = e0.w * dx, c0 = e0.z * dx * dx;
          const uint32_t pos = (uint32_t)(base + b + k + 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 dy = f2(e0.y - fy[h].x, e0.y - fy[h].y);
            const float2 pw = __ffma2_rn(dy, __ffma2_rn(f2s(e1.x), dy, f2s(c1)), f2s(c0));
            const float2 G = f2(ex2_approx(pw.x), ex2_approx(pw.y));
            float2 al = __fmul2_rn(f2s(e1.y), G);
            al.x = fminf(0.99f, al.x); al.y = fminf(0.99f, al.y);
            bool vA = !done[2 * h] && pw.x <= 0.f && al.x >= kAlphaMin;
            bool vB = !done[2 * h + 1] && pw.y <= 0.f && al.y >= kAlphaMin;
            const float2 tT = __fmul2_rn(T[h], __ffma2_rn(al, f2s(-1.f), f2s(1.f)));
            if (vA && tT.x < kTEps) { done[2 * h] = true; vA = false; }
            if (vB && tT.y < kTEps) { done[2 * h + 1] = true; vB = false; }
            float2 w = __fmul2_rn(al, T[h]);
            w.x = vA ? w.x : 0.f; w.y = vB ? w.y : 0.f;
            Cr[h] = __ffma2_rn(f2s(c.x), w, Cr[h]);
            Cg[h] = __ffma2_rn(f2s(c.y), w, Cg[h]);
            Cb[h] = __ffma2_rn(f2s(c.z), w, Cb[h]);
            T[h].x = vA ? tT.x : T[h].x; T[h].y = vB ? tT.y : T[h].y;
            last[2 * h] = vA ? pos : last[2 * h]; last[2 * h + 1] = vB ? pos : last[2 * h + 1];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) wdone[q] = __all_sync(0xffffffffu, done[q]);
        wall = wdone[0] && wdone[1] && wdone[2] && wdone[3];
        if (wall) break;
      }
    }
    if (__syncthreads_and(wall)) break;
    pending = -1;
  }
  if (BULK && pending >= 0) mbar_wait(&bars[pending & 1], (uint32_t)((pending >> 1) & 1));
  const size_t hw = (size_t)W * H;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (in[q]) {
      const int h = q >> 1;
      const float t = (q & 1) ? T[h].y : T[h].x;
      const float r = (q & 1) ? Cr[h].y : Cr[h].x, g = (q & 1) ? Cg[h].y : Cg[h].x, bb = (q & 1) ? Cb[h].y : Cb[h].x;
      const size_t pix = (size_t)(py0 + 4 * q) * W + px;
      final_T[pix] = t; n_contrib[pix] = last[q];
      out_color[pix] = r + t * b0; out_color[hw + pix] = g + t * b1; out_color[2 * hw + pix] = bb + t * b2;
    }
  }
}

template <bool BULK>
__global__ void __launch_bounds__(kThreads3, GSB_BWD3_MINB)
k_blend_bwd3(const uint2* __restrict__ ranges, const float4* __restrict__ s0, const float4* __restrict__ s1,
             const float4* __restrict__ s2, const float* __restrict__ bg, int W, int H, int gx,
             const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
             const float* __restrict__ dL_dpix, float* __restrict__ dacc) {
  __shared__ __align__(128) SlabStage stg[BULK ? 2 : 1];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ int s_bmax;
  if (BULK && threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sx0 = tx * kBlock + warp * 8, sy0 = ty * kBlock;
  const int px = sx0 + (lane & 7), py0 = sy0 + (lane >> 3);
  const float fx = (float)px;
  const float rx0 = (float)sx0, rx1 = (float)min(sx0 + 7, W - 1);
  const uint2 rg = ranges[tile];
  const size_t hw = (size_t)W * H;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  float2 fy[2], T[2], dLr[2], dLg[2], dLb[2], tf_bg[2], acc_r[2], acc_g[2], acc_b[2];
  int lc[4], wmaxq[4];
  float ry0[4], ry1[4];
  bool in[4];
  fy[0] = f2((float)py0, (float)(py0 + 4));
  fy[1] = f2((float)(py0 + 8), (float)(py0 + 12));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int py = py0 + 4 * q;
    in[q] = px < W && py < H;
    const size_t pix = (size_t)py * W + px;
    const float tfin = in[q] ? final_T[pix] : 0.f;
    lc[q] = in[q] ? (int)n_contrib[pix] : 0;
    const float r = in[q] ? dL_dpix[pix] : 0.f, g = in[q] ? dL_dpix[hw + pix] : 0.f, bb = in[q] ? dL_dpix[2 * hw + pix] : 0.f;
    const int h = q >> 1;
    if (q & 1) { T[h].y = tfin; dLr[h].y = r; dLg[h].y = g; dLb[h].y = bb; tf_bg[h].y = tfin * (b0 * r + b1 * g + b2 * bb); }
    else       { T[h].x = tfin; dLr[h].x = r; dLg[h].x = g; dLb[h].x = bb; tf_bg[h].x = tfin * (b0 * r + b1 * g + b2 * bb); }
    ry0[q] = (float)(sy0 + 4 * q);
    ry1[q] = (float)min(sy0 + 4 * q + 3, H - 1);
    int m = lc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    wmaxq[q] = m;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) acc_r[h] = acc_g[h] = acc_b[h] = f2(0.f, 0.f);
  const int wmax = max(max(wmaxq[0], wmaxq[1]), max(wmaxq[2], wmaxq[3]));
  if (threadIdx.x == 0) s_bmax = 0;
  __syncthreads();
  if (lane == 0) atomicMax(&s_bmax, wmax);
  __syncthreads();
  const int bmax = s_bmax;
  const int nchunks = (bmax + kChunk - 1) / kChunk;
  if (BULK && nchunks > 0)
    stage_slab<true>(&stg[0], s0, s1, s2, (size_t)rg.x + (size_t)(nchunks - 1) * kChunk,
                     min(kChunk, bmax - (nchunks - 1) * kChunk), &bars[0], kThreads3);
  for (int it = 0; it < nchunks; ++it) {
    const int ch = nchunks - 1 - it;
    const int base = ch * kChunk;
    const int cnt = min(kChunk, bmax - base);
    const SlabStage* cur = &stg[BULK ? (it & 1) : 0];
    if (BULK) {
      if (it + 1 < nchunks)
        stage_slab<true>(&stg[(it + 1) & 1], s0, s1, s2, (size_t)rg.x + base - kChunk, kChunk, &bars[(it + 1) & 1],
                         kThreads3);
      mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
    } else {
      stage_slab<false>(&stg[0], s0, s1, s2, (size_t)rg.x + base, cnt, nullptr, kThreads3);
      __syncthreads();
    }
    const float4* sm0 = cur->s0;
    const float4* sm1 = cur->s1;
    const float4* sm2 = cur->s2;
    if (base < wmax) {
      for (int b = (cnt - 1) & ~31; b >= 0; b -= 32) {
        if (base + b >= wmax) continue;
        const int j = b + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 e0 = sm0[j], e1 = sm1[j];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (!hit && base + j < wmaxq[q]) hit = slab_may_contribute(e0, e1, rx0, ry0[q], rx1, ry1[q]);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);
          mask &= ~(1u << k);
          const int pos = base + b + k;
          const float4 e0 = sm0[b + k], e1 = sm1[b + k];
          const float dx = e0.x - fx;
          const float c1 = e0.w * dx, c0 = e0.z * dx * dx;
          float2 dy[2], G[2], al[2];
          bool v[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            dy[h] = f2(e0.y - fy[h].x, e0.y - fy[h].y);
            const float2 pw = __ffma2_rn(dy[h], __ffma2_rn(f2s(e1.x), dy[h], f2s(c1)), f2s(c0));
            G[h] = f2(ex2_approx(pw.x), ex2_approx(pw.y));
            al[h] = __fmul2_rn(f2s(e1.y), G[h]);
            al[h].x = fminf(0.99f, al[h].x); al[h].y = fminf(0.99f, al[h].y);
            v[2 * h] = in[2 * h] && pos < lc[2 * h] && pw.x <= 0.f && al[h].x >= kAlphaMin;
            v[2 * h + 1] = in[2 * h + 1] && pos < lc[2 * h + 1] && pw.y <= 0.f && al[h].y >= kAlphaMin;
          }
          if (!__any_sync(0xffffffffu, v[0] || v[1] || v[2] || v[3])) continue;
          const float4 c = sm2[b + k];
          float s_u = 0.f, s_uy = 0.f, s_uyy = 0.f, s_do = 0.f, s_r = 0.f, s_g = 0.f, s_b = 0.f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 am = f2(v[2 * h] ? al[h].x : 0.f, v[2 * h + 1] ? al[h].y : 0.f);
            const float2 one_m = __ffma2_rn(am, f2s(-1.f), f2s(1.f));
            const float2 inv = f2(rcp_approx(one_m.x), rcp_approx(one_m.y));
            T[h] = __fmul2_rn(T[h], inv);
            const float2 dcol = __fmul2_rn(am, T[h]);
            const float2 d_r = __fadd2_rn(f2s(c.x), f2(-acc_r[h].x, -acc_r[h].y));
            const float2 d_g = __fadd2_rn(f2s(c.y), f2(-acc_g[h].x, -acc_g[h].y));
            const float2 d_b = __fadd2_rn(f2s(c.z), f2(-acc_b[h].x, -acc_b[h].y));
            float2 da = __ffma2_rn(d_b, dLb[h], __ffma2_rn(d_g, dLg[h], __fmul2_rn(d_r, dLr[h])));
            da = __fmul2_rn(da, T[h]);
            da = __ffma2_rn(f2(-tf_bg[h].x, -tf_bg[h].y), inv, da);
            da.x = v[2 * h] ? da.x : 0.f; da.y = v[2 * h + 1] ? da.y : 0.f;
            acc_r[h] = __ffma2_rn(am, d_r, acc_r[h]);
            acc_g[h] = __ffma2_rn(am, d_g, acc_g[h]);
            acc_b[h] = __ffma2_rn(am, d_b, acc_b[h]);
            const float2 gda = __fmul2_rn(G[h], da);                 // G * dL/dalpha
            const float2 u = __fmul2_rn(f2s(e1.y), gda);             // u = G * dL/dG = opacity * G * dL/dalpha
            const float2 uy = __fmul2_rn(u, dy[h]);
            const float2 uyy = __fmul2_rn(uy, dy[h]);
            const float2 cr = __fmul2_rn(dcol, dLr[h]), cg = __fmul2_rn(dcol, dLg[h]), cb = __fmul2_rn(dcol, dLb[h]);
            s_u += u.x + u.y; s_uy += uy.x + uy.y; s_uyy += uyy.x + uyy.y; s_do += gda.x + gda.y;
            s_r += cr.x + cr.y; s_g += cg.x + cg.y; s_b += cb.x + cb.y;
          }
          // raw moments (dx is common to the lane's four pixels); the conic / ln2 factors are applied after the
          // warp reduction by the two lanes that issue the REDs
          float vv[9];
          vv[0] = s_u * dx;            // S_x  = sum u dx
          vv[1] = s_uy;                // S_y  = sum u dy
          vv[2] = s_u * dx * dx;       // S_xx
          vv[3] = s_uy * dx;           // S_xy
          vv[4] = s_uyy;               // S_yy
          vv[5] = s_do;
          vv[6] = s_r; vv[7] = s_g; vv[8] = s_b;
          warp_reduce9(vv, lane);
          const float a1 = __shfl_down_sync(0xffffffffu, vv[0], 4);
          const float a2 = __shfl_down_sync(0xffffffffu, vv[0], 8);
          const float a3 = __shfl_down_sync(0xffffffffu, vv[0], 12);
          float* dst = dacc + (size_t)__float_as_uint(e1.w) * 12;
          if (lane == 0) {        // holds S_x, S_y, S_xx, S_xy
            const float Sx = vv[0], Sy = a1;
            red_add_v4(dst, kLn2 * (2.f * Sx * e0.z + Sy * e0.w), kLn2 * (2.f * Sy * e1.x + Sx * e0.w), -0.5f * a2, -a3);
          } else if (lane == 16) {  // holds S_yy, do, r, g
            red_add_v4(dst + 4, -0.5f * vv[0], a1, a2, a3);
          }
          if (lane == 1) atomicAdd(dst + 8, vv[8]);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// k_preprocess_bwd
// ------------------------------------------------------------------------------------------
struct OutPtrs {
  float* dmeans; float* dmeans2D; float* dscales; float* drots; float* dopac;
  float* dsh_dc; float* dsh_rest; float* dcolors; float* dcov3D;
};

__global__ void __launch_bounds__(kPT, 4)
k_preprocess_bwd(InPtrs in, GeomView gv, const int* __restrict__ radii_unused, OutPtrs out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CamConst* cam = reinterpret_cast<CamConst*>(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw + ((sizeof(CamConst) + 15) / 16) * 16);
  __shared__ float s_pose[kPT / 32][16];
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(gv.cam);
    uint32_t* d = reinterpret_cast<uint32_t*>(cam);
    for (int k = threadIdx.x; k < (int)(sizeof(CamConst) / 4); k += blockDim.x) d[k] = s[k];
  }
  const int first = blockIdx.x * kPT;
  const int nv = min(kPT, in.P - first);
  const bool use_sh = in.colors == nullptr;
  const bool vec = in.vec_ok != 0;
  __syncthreads();
  const int D = cam->D, M = cam->M;
  load_block_inputs(in, first, nv, use_sh, D, M, sm);
  __syncthreads();
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
  __syncthreads();
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
  GSB_REQUIRE((long long)((cam->width + kBlock - 1) / kBlock) * ((cam->height + kBlock - 1) / kBlock) <= 65535,
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

static bool g_attr_set = false;
static int ensure_attrs() {
  if (!g_attr_set) {
    GSB_CUDA(cudaFuncSetAttribute(k_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmem));
    GSB_CUDA(cudaFuncSetAttribute(k_preprocess_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmem));
    g_attr_set = true;
  }
  return GSB_OK;
}

extern "C" GSB_API int gsb_preprocess(const GsbCamera* cam, const GsbGaussians* g, void* geom, size_t geom_bytes,
                              int32_t* radii, uint32_t* num_rendered_host, gsb_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  InPtrs in;
  int rc = make_inptrs(cam, g, in);
  if (rc) return rc;
  GSB_REQUIRE(geom && (radii || g->P == 0) && num_rendered_host, "null buffer");
  const int P = g->P;
  GeomView gv = geom_view(geom, P);
  if (gv.total > geom_bytes) { gsb_set_error("geom buffer too small"); return GSB_ERR_CAPACITY; }
  rc = ensure_attrs();
  if (rc) return rc;
  gsb_count_launch(1);
  k_setup_cam<<<1, 32, 0, st>>>(gv.cam, cam->viewmatrix, cam->projmatrix, cam->campos, g->pose, cam->width,
                                cam->height, cam->tanfovx, cam->tanfovy, cam->scale_modifier, cam->sh_degree,
                                cam->sh_coeffs, g->raw_params);
  if (P == 0) {
    GSB_CUDA(cudaMemsetAsync(gv.nrend, 0, 4, st));
  } else {
    const int nb = (P + kPT - 1) / kPT;
    { ProfScope ps(GSB_K_PREPROCESS, st); k_preprocess<<<nb, kPT, kPrepSmem, st>>>(in, gv, radii); }
    size_t tb = gv.cub_bytes;
    { ProfScope ps(GSB_K_SORT_DEPTH, st, 0);
      GSB_CUDA(cub::DeviceRadixSort::SortPairs(gv.cub_tmp, tb, gv.dkey, gv.dkey_s, gv.iota, gv.order, P, 0, 32, st)); }
    cub::CountingInputIterator<uint32_t> cnt(0);
    TilesInOrder op{gv.tiles, gv.order};
    cub::TransformInputIterator<uint32_t, TilesInOrder, cub::CountingInputIterator<uint32_t>> it(cnt, op);
    tb = gv.cub_bytes;
    { ProfScope ps(GSB_K_SCAN, st, 1);
      GSB_CUDA(cub::DeviceScan::InclusiveSum(gv.cub_tmp, tb, it, gv.offs, P, st));
      k_store_total<<<1, 32, 0, st>>>(gv.offs, P, gv.nrend); }
  }
  GSB_CUDA(cudaMemcpyAsync(num_rendered_host, gv.nrend, 4, cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaGetLastError());
  return GSB_OK;
}

extern "C" GSB_API int gsb_render(const GsbCamera* cam, int32_t P, void* geom, void* binning, size_t binning_bytes,
                          int64_t R, void* image, float* out_color, gsb_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  GSB_REQUIRE(cam && geom && binning && image && out_color, "null buffer");
  GSB_REQUIRE(R >= 0 && R < (int64_t)0x7fffffff, "R out of range");
  const int W = cam->width, H = cam->height;
  const int gx = (W + kBlock - 1) / kBlock, gy = (H + kBlock - 1) / kBlock;
  GeomView gv = geom_view(geom, P);
  BinView bv = bin_view(binning, R, W, H);
  if (bv.total > binning_bytes) { gsb_set_error("binning buffer too small"); return GSB_ERR_CAPACITY; }
  ImgView iv = img_view(image, W, H);
  GSB_CUDA(cudaMemsetAsync(bv.ranges, 0, (size_t)gx * gy * 8, st));
  if (R > 0) {
    { ProfScope ps(GSB_K_DUPLICATE, st);
      k_duplicate<<<(P + kThreads - 1) / kThreads, kThreads, 0, st>>>(P, gv, bv, W, H, gx, cam->exact_cull, (uint32_t)R); }
    size_t tb = bv.cub_bytes;
    { ProfScope ps(GSB_K_SORT_TILE, st, 0);
      GSB_CUDA(cub::DeviceRadixSort::SortPairs(bv.cub_tmp, tb, bv.keys, bv.keys_s, bv.vals, bv.vals_s, (int)R, 0,
                                               tile_bits(gx * gy), st)); }
    { ProfScope ps(GSB_K_GATHER, st);
      k_ranges_gather<<<(unsigned)((R + kThreads - 1) / kThreads), kThreads, 0, st>>>((uint32_t)R, gv, bv); }
  }
  { ProfScope ps(GSB_K_BLEND_FWD, st);
    if (g_blend_version == 3 && g_stage_bulk)
      k_blend_fwd3<true><<<gx * gy, kThreads3, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, out_color,
                                                        iv.final_T, iv.n_contrib);
    else if (g_blend_version == 3)
      k_blend_fwd3<false><<<gx * gy, kThreads3, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, out_color,
                                                         iv.final_T, iv.n_contrib);
    else if (g_blend_version == 2 && g_stage_bulk)
      k_blend_fwd2<true><<<gx * gy, kThreads2, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, out_color,
                                                        iv.final_T, iv.n_contrib);
    else if (g_blend_version == 2)
      k_blend_fwd2<false><<<gx * gy, kThreads2, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, out_color,
                                                         iv.final_T, iv.n_contrib);
    else
      k_blend_fwd<<<gx * gy, kThreads, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, out_color,
                                                iv.final_T, iv.n_contrib); }
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
  const int gx = (W + kBlock - 1) / kBlock, gy = (H + kBlock - 1) / kBlock;
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
  if (R > 0) {
    ProfScope ps(GSB_K_BLEND_BWD, st);
    if (g_blend_version == 3 && g_stage_bulk)
      k_blend_bwd3<true><<<gx * gy, kThreads3, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, iv.final_T,
                                                        iv.n_contrib, dL_dout, (float*)gv.dacc);
    else if (g_blend_version == 3)
      k_blend_bwd3<false><<<gx * gy, kThreads3, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, iv.final_T,
                                                         iv.n_contrib, dL_dout, (float*)gv.dacc);
    else if (g_blend_version == 2 && g_stage_bulk)
      k_blend_bwd2<true><<<gx * gy, kThreads2, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, iv.final_T,
                                                        iv.n_contrib, dL_dout, (float*)gv.dacc);
    else if (g_blend_version == 2)
      k_blend_bwd2<false><<<gx * gy, kThreads2, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, iv.final_T,
                                                         iv.n_contrib, dL_dout, (float*)gv.dacc);
    else
      k_blend_bwd<<<gx * gy, kThreads, 0, st>>>(bv.ranges, bv.s0, bv.s1, bv.s2, cam->bg, W, H, gx, iv.final_T,
                                                iv.n_contrib, dL_dout, (float*)gv.dacc);
  }
  OutPtrs out;
  out.dmeans = grads->dL_dmeans3D; out.dmeans2D = grads->dL_dmeans2D; out.dscales = grads->dL_dscales;
  out.drots = grads->dL_drotations; out.dopac = grads->dL_dopacities; out.dsh_dc = grads->dL_dsh_dc;
  out.dsh_rest = grads->dL_dsh_rest; out.dcolors = grads->dL_dcolors; out.dcov3D = grads->dL_dcov3D;
  uintptr_t a = (uintptr_t)out.dmeans | (uintptr_t)out.dscales | (uintptr_t)out.drots | (uintptr_t)out.dopac |
                (uintptr_t)out.dsh_dc | (uintptr_t)out.dsh_rest;
  if (a & 15) in.vec_ok = 0;
  const int nb = (P + kPT - 1) / kPT;
  { ProfScope ps(GSB_K_PREPROCESS_BWD, st, g->pose ? 3 : 1);
    k_preprocess_bwd<<<nb, kPT, kPrepSmem, st>>>(in, gv, nullptr, out);
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
