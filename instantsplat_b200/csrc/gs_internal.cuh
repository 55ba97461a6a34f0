// Internal declarations shared by the translation units of libgsb200.so (not part of the C ABI):
// error handling, per-kernel timing hooks, and the layouts of the three caller-owned scratch buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsb200.h"
#include "gs_math.cuh"

// ------------------------------------------------------------------------------------------
// error handling (thread-local message, see gsb_last_error)
// ------------------------------------------------------------------------------------------
void gsb_set_error(const char* s);
void gsb_set_errorf(const char* file, int line, const char* what, const char* detail);
#define GSB_CUDA(x)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      gsb_set_errorf(__FILE__, __LINE__, #x, cudaGetErrorString(e_));                     \
      return GSB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
#define GSB_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      gsb_set_errorf(__FILE__, __LINE__, "invalid argument", msg);                        \
      return GSB_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

// ------------------------------------------------------------------------------------------
// optional per-kernel timing (CUDA events on the launching stream) and launch counting
// ------------------------------------------------------------------------------------------
void gsb_count_launch(int n);
int gsb_prof_begin(int id, cudaStream_t st);
void gsb_prof_end(int slot, cudaStream_t st);
struct ProfScope {
  int slot; cudaStream_t st;
  ProfScope(int id, cudaStream_t s, int launches = 1) : st(s) { gsb_count_launch(launches); slot = gsb_prof_begin(id, s); }
  ~ProfScope() { gsb_prof_end(slot, st); }
};
int gsb_option_blend_version();
int gsb_option_stage_bulk();

namespace gsb {

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

constexpr int kThreads = 256;
constexpr int kPT = 128;           // threads (= Gaussians) per CTA in the per-Gaussian kernels
constexpr int kMaxTiles = 65536;   // tile ids are 16 bit (a 4K frame has 32400 tiles)

// status words written by the forward (device) and copied to the caller's pinned host array
enum { kStR = 0, kStOverflow = 1, kStMaxList = 2, kStHugeTiles = 3, kStNLarge = 4, kStNHuge = 5, kStSeq = 6,
       kStWords = 8 };
// tile-list size classes of the per-tile sort (gs_bin.cu): <= kSmallList entries are sorted by the one-CTA-per-tile
// kernel; longer lists are queued by k_tile_scan and drained by two persistent kernels
constexpr uint32_t kSmallList = 2048, kLargeList = 8192;

// ---- geom: P-sized scratch + per-tile counters (caller-owned, gsb_geom_bytes) -------------
struct GeomView {
  CamConst* cam;
  float4* rec;         // [3P] one 48-byte splat record per Gaussian, array of structures (the blend kernels gather it by id):
                       //   [0] x, y, A', B'      [1] C', opacity, cull threshold', id (bits)      (conic in the log2 domain:
                       //   [2] r, g, b, radius                                                     power*log2e = A'dx^2+B'dxdy+C'dy^2)
  float4* brec;        // [P] binning record: depth, rect x (rx0 | rx1<<16), rect y (ry0 | ry1<<16), keep mask (bits)
  uint32_t* tiles;     // tile instances per Gaussian (after culling)
  uint8_t* clamped;
  float4* dacc;        // [3P] backward accumulators
  float* pose_part;    // [nblocks*16]
  float* pose_acc;     // [16]
  uint32_t* status;    // [kStWords]
  uint32_t* aux;       // [64] per-forward counters, zeroed together with tcount (they are adjacent): [0] = length of q_big
  uint32_t* tcount;    // [kMaxTiles] instances per tile (counted by k_preprocess / k_big_rects)
  uint32_t* tstart;    // [kMaxTiles] exclusive scan of tcount
  uint32_t* tcursor;   // [kMaxTiles] emission cursors (k_scatter)
  uint32_t* q_large;   // [kMaxTiles] tiles whose list has kSmallList < n <= kLargeList entries
  uint32_t* q_huge;    // [kMaxTiles] tiles with longer lists
  uint32_t* q_big;     // [P] Gaussians whose tile rect exceeds kBigRect tiles (walked by k_big_rects, flattened)
  size_t total;
};

static inline GeomView geom_view(void* base, int P) {
  GeomView v;
  size_t off = 0;
  char* p = (char*)base;
  size_t Pp = (size_t)(P > 0 ? P : 1);
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  v.cam = (CamConst*)take(sizeof(CamConst));
  v.rec = (float4*)take(Pp * 48);
  v.brec = (float4*)take(Pp * 16);
  v.tiles = (uint32_t*)take(Pp * 4);
  v.clamped = (uint8_t*)take(Pp);
  v.dacc = (float4*)take(Pp * 48);
  size_t nb = (Pp + kPT - 1) / kPT;
  v.pose_part = (float*)take(nb * 16 * 4);
  v.pose_acc = (float*)take(16 * 4);
  v.status = (uint32_t*)take(kStWords * 4);
  v.aux = (uint32_t*)take(64 * 4);                        // exactly 256 bytes: tcount follows immediately
  v.tcount = (uint32_t*)take((size_t)kMaxTiles * 4);
  v.tstart = (uint32_t*)take((size_t)kMaxTiles * 4);
  v.tcursor = (uint32_t*)take((size_t)kMaxTiles * 4);
  v.q_large = (uint32_t*)take((size_t)kMaxTiles * 4);
  v.q_huge = (uint32_t*)take((size_t)kMaxTiles * 4);
  v.q_big = (uint32_t*)take(Pp * 4);
  v.total = off;
  return v;
}

// ---- binning: capacity-sized scratch (caller-owned, gsb_binning_bytes) --------------------
struct BinView {
  unsigned long long* pairs;   // [cap] (depth bits << 32 | Gaussian id), unsorted per-tile segments
  uint32_t* ids;               // [cap + 8] Gaussian ids, per tile in (depth, id) order (what the blend kernels walk)
  uint2* ranges;               // [tiles]
  size_t total;
};

static inline BinView bin_view(void* base, int64_t cap, int W, int H) {
  BinView v;
  size_t off = 0;
  char* p = (char*)base;
  size_t Rp = (size_t)(cap > 0 ? cap : 1);
  int ntiles = ((W + kBlock - 1) / kBlock) * ((H + kBlock - 1) / kBlock);
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
  v.pairs = (unsigned long long*)take(Rp * 8);
  v.ids = (uint32_t*)take((Rp + 8) * 4);       // + 8: bulk copies of id chunks are rounded to 16 bytes
  v.ranges = (uint2*)take((size_t)ntiles * 8);
  v.total = off;
  return v;
}

struct ImgView {
  float* final_T;
  uint32_t* n_contrib;
  size_t total;
};
static inline ImgView img_view(void* base, int W, int H) {
  ImgView v;
  size_t hw = (size_t)W * H;
  char* p = (char*)base;
  v.final_T = (float*)p;
  v.n_contrib = (uint32_t*)(p ? p + align_up(hw * 4) : nullptr);
  v.total = 2 * align_up(hw * 4);
  return v;
}

}  // namespace gsb

// ---- launchers implemented in gs_bin.cu / gs_blend.cu -----------------------------------------
// tile scan: tstart = exclusive scan(tcount), status = {R, 0, max list, 0, long-list queue lengths, serial}
int gsb_launch_tile_scan(const gsb::GeomView& gv, int ntiles, cudaStream_t st);
int gsb_launch_big_rects(int P, const gsb::GeomView& gv, unsigned long long* pairs, int W, int H, int exact_cull,
                         uint32_t cap, cudaStream_t st);
// scatter + per-tile (depth, id) sort + slab gather + ranges; cap = instance capacity of the binning buffer
int gsb_launch_binning(int P, const gsb::GeomView& gv, const gsb::BinView& bv, int W, int H, int exact_cull,
                       uint32_t cap, cudaStream_t st);
int gsb_launch_blend_fwd(const gsb::GeomView& gv, const gsb::BinView& bv, const gsb::ImgView& iv, const float* bg, int W, int H,
                         float* out_color, cudaStream_t st);
// pose_only: 8-value accumulator layout (dacc[5] = dL/db, no dL/dopacity), see k_blend_bwd2
int gsb_launch_blend_bwd(const gsb::GeomView& gv, const gsb::BinView& bv, const gsb::ImgView& iv, const float* bg, int W, int H,
                         const float* dL_dout, float* dacc, bool pose_only, cudaStream_t st);
