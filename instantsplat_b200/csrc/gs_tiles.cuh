// Tile coverage of one projected Gaussian: the ONE definition of "which 16x16 tiles receive this splat",
// used twice -- by k_preprocess to COUNT instances per Gaussian and per tile, and by k_scatter to EMIT them.
// Coverage = the reference's tile rect minus the tiles that the lossless cull
// (gs_math.cuh rect_may_contribute) proves cannot reach alpha >= 1/255.
//
// Small rects (<= kCoopTiles tiles): k_preprocess stores the keep MASK, and k_scatter replays exactly those tiles.
// Large rects are re-walked by the whole warp in both passes; there the emit side does not trust the two
// evaluations to agree bit for bit: an instance is written only while the tile's cursor is below the counted
// length (and below the buffer capacity), and the consumer uses min(count, cursor) as the list length.  A
// borderline tile that flips between the two passes carries no visible contribution by construction of the cull
// margin, so dropping / missing it does not change the image.
#pragma once
#include "gs_math.cuh"

namespace gsb {

constexpr int kCoopTiles = 24;   // rects larger than this are walked by the whole warp
constexpr int kTripTiles = 128;  // tiles per cooperative trip (4 per lane)
constexpr int kBigRect = 128;    // rects larger than this leave the per-Gaussian kernels altogether: k_preprocess queues them
                                 // and k_big_rects walks them as (Gaussian, trip) work items spread over the whole grid

struct TileSink {
  // count mode: tcount != nullptr, pairs == nullptr
  // emit  mode: pairs  != nullptr
  uint32_t* tcount;
  const uint32_t* tstart;
  uint32_t* tcursor;
  unsigned long long* pairs;
  uint32_t cap;
};

__device__ __forceinline__ void sink_tile(const TileSink& s, uint32_t tile, unsigned long long key) {
  if (s.pairs) {
    const uint32_t pos = atomicAdd(s.tcursor + tile, 1u);
    if (pos < s.tcount[tile]) {
      const uint32_t dst = s.tstart[tile] + pos;
      if (dst < s.cap) s.pairs[dst] = key;
    }
  } else {
    atomicAdd(s.tcount + tile, 1u);
  }
}

struct SplatRect {
  float x, y, A, B, C, qthr;
  int rx0, ry0, rx1, ry1;
};

__device__ __forceinline__ bool tile_kept(const SplatRect& p, int tx, int ty, int W, int H, bool cull) {
  if (!cull) return true;
  const float x0 = (float)(tx * kBlock), y0 = (float)(ty * kBlock);
  const float x1 = fminf(x0 + kBlock - 1, (float)(W - 1)), y1 = fminf(y0 + kBlock - 1, (float)(H - 1));
  return rect_may_contribute(p.x, p.y, p.A, p.B, p.C, p.qthr, x0, y0, x1, y1);
}

// Small rects (area <= kCoopTiles <= 32): bit k of the keep mask = tile (ry0 + k / w, rx0 + k % w) survives the cull.
// Computed once by k_preprocess and stored, so the emit pass replays exactly the tiles that were counted.
// Same walk, also counting every kept tile into tcount (RED.ADD, no return value).
__device__ __forceinline__ uint32_t rect_keep_mask_count_tiles(const SplatRect& p, int W, int H, int gx, bool cull,
                                                               uint32_t* tcount) {
  uint32_t m = 0u, bit = 1u;
  for (int ty = p.ry0; ty < p.ry1; ++ty)
    for (int tx = p.rx0; tx < p.rx1; ++tx) {
      if (tile_kept(p, tx, ty, W, H, cull)) { m |= bit; atomicAdd(tcount + (uint32_t)(ty * gx + tx), 1u); }
      bit <<= 1;
    }
  return m;
}

// The same mask computed per tile ROW (gs_math.cuh row_keep_range: two square roots per row instead of one rectangle
// test per tile); degenerate conics and cull-off fall back to the per-tile walk.
__device__ __forceinline__ uint32_t rect_keep_mask_count(const SplatRect& p, int W, int H, int gx, bool cull,
                                                         uint32_t* tcount) {
  const RowCull rc = row_cull_setup(p.x, p.y, p.A, p.B, p.C, p.qthr);
  if (!cull || !rc.ok) return rect_keep_mask_count_tiles(p, W, H, gx, cull, tcount);
  const int w = p.rx1 - p.rx0;
  uint32_t m = 0u;
  for (int ty = p.ry0; ty < p.ry1; ++ty) {
    int ta, tb;
    if (!row_keep_range(rc, ty, H, p.rx0, p.rx1, ta, tb)) continue;
    const int n = tb - ta + 1;                            // <= 24: the caller's rect has at most kCoopTiles tiles
    m |= ((1u << n) - 1u) << ((ty - p.ry0) * w + (ta - p.rx0));
    for (int tx = ta; tx <= tb; ++tx) atomicAdd(tcount + (uint32_t)(ty * gx + tx), 1u);
  }
  return m;
}

// Warp-synchronous EMISSION of every lane's kept tiles (all 32 lanes must call).  Each round every lane offers its
// next TWO kept tiles; lanes offering the SAME tile in the same slot are merged with match.any so that one atomic per
// distinct tile is issued (neighbouring Gaussians overlap the same tiles, which otherwise serialises in the L2 atomic
// unit): the group leader reserves cnt slots with one atomicAdd and every member takes base + rank.  Both slots'
// atomics are issued before either result is consumed -- the kernel is a chain of atomic round trips per warp, two
// in flight halve it.
__device__ __forceinline__ void warp_sink_masks(uint32_t mask, int rx0, int ry0, int w, int gx, const TileSink& s,
                                                unsigned long long key) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const uint32_t inv_w = w > 0 ? (65536u + (uint32_t)w - 1u) / (uint32_t)w : 0u;   // k / w for k < 32, w <= 32
  const uint32_t below = (1u << lane) - 1u;
  while (__any_sync(full, mask != 0u)) {
    bool has[2];
    uint32_t tile[2], base[2], cnt[2];
    unsigned grp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      has[u] = mask != 0u;
      tile[u] = 0xFFFFFFFFu;
      if (has[u]) {
        const uint32_t k = (uint32_t)__ffs(mask) - 1u;
        mask &= mask - 1u;
        const uint32_t ky = (k * inv_w) >> 16;
        tile[u] = (uint32_t)(ry0 + (int)ky) * (uint32_t)gx + (uint32_t)(rx0 + (int)(k - ky * (uint32_t)w));
      }
      grp[u] = __match_any_sync(full, tile[u]);
      cnt[u] = (uint32_t)__popc(grp[u]);
    }
    uint32_t tc[2], ts[2];                               // loaded before the atomics (which the compiler must not reorder)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      tc[u] = has[u] ? s.tcount[tile[u]] : 0u;
      ts[u] = has[u] ? s.tstart[tile[u]] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      base[u] = 0u;
      if (has[u] && lane == __ffs(grp[u]) - 1) base[u] = atomicAdd(s.tcursor + tile[u], cnt[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      base[u] = __shfl_sync(full, base[u], __ffs(grp[u]) - 1);
      if (has[u]) {
        const uint32_t pos = base[u] + (uint32_t)__popc(grp[u] & below);
        if (pos < tc[u]) {
          const uint32_t dst = ts[u] + pos;
          if (dst < s.cap) s.pairs[dst] = key;
        }
      }
    }
  }
}

// One cooperative trip: tiles [base, base + kTripTiles) of the rect (row-major index i -> (ry0 + i / w, rx0 + i % w)),
// four per lane, every atomic issued before the first result is consumed.  All 32 lanes call with the same arguments;
// returns the number of kept tiles of the trip.
__device__ __forceinline__ uint32_t visit_trip(const SplatRect& b, int base, int n, int w, int W, int H, int gx, bool cull,
                                               const TileSink& s, unsigned long long bkey) {
  const int lane = threadIdx.x & 31;
  constexpr int kU = kTripTiles / 32;
  bool keep[kU];
  uint32_t tile[kU], pos[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int i = base + u * 32 + lane;
    keep[u] = false;
    tile[u] = 0u;
    if (i < n) {
      const int ty = b.ry0 + i / w, tx = b.rx0 + i - (i / w) * w;
      keep[u] = tile_kept(b, tx, ty, W, H, cull);
      tile[u] = (uint32_t)(ty * gx + tx);
    }
  }
  if (s.pairs) {
#pragma unroll
    for (int u = 0; u < kU; ++u) pos[u] = keep[u] ? atomicAdd(s.tcursor + tile[u], 1u) : 0xFFFFFFFFu;
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (keep[u] && pos[u] < s.tcount[tile[u]]) {
        const uint32_t dst = s.tstart[tile[u]] + pos[u];
        if (dst < s.cap) s.pairs[dst] = bkey;
      }
  } else {
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (keep[u]) atomicAdd(s.tcount + tile[u], 1u);
  }
  uint32_t run = 0;
#pragma unroll
  for (int u = 0; u < kU; ++u) run += __popc(__ballot_sync(0xffffffffu, keep[u]));
  return run;
}

// Gaussians whose rect spans more than kCoopTiles tiles are walked by the WHOLE WARP (one tile per lane per
// step) instead of one thread looping over up to thousands of tiles -- the per-thread loop is a performance
// cliff once a few Gaussians grow large.  (Rects beyond kBigRect do not come here: a Gaussian that has grown over the
// whole frame is thousands of tiles = a serial chain of trips for ONE warp and the tail of the whole kernel; those are
// queued and flattened over the grid by k_big_rects.)  Must be called by all 32 lanes; `mine` says whether this lane's
// Gaussian wants the cooperative path.  Returns the lane's kept-tile count.
__device__ __forceinline__ uint32_t visit_tiles_coop(bool mine, const SplatRect& p, int W, int H, int gx, bool cull,
                                                     const TileSink& s, unsigned long long key) {
  const int lane = threadIdx.x & 31;
  uint32_t result = 0;
  unsigned big = __ballot_sync(0xffffffffu, mine);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    SplatRect b;
    b.x = __shfl_sync(0xffffffffu, p.x, src); b.y = __shfl_sync(0xffffffffu, p.y, src);
    b.A = __shfl_sync(0xffffffffu, p.A, src); b.B = __shfl_sync(0xffffffffu, p.B, src);
    b.C = __shfl_sync(0xffffffffu, p.C, src); b.qthr = __shfl_sync(0xffffffffu, p.qthr, src);
    b.rx0 = __shfl_sync(0xffffffffu, p.rx0, src); b.rx1 = __shfl_sync(0xffffffffu, p.rx1, src);
    b.ry0 = __shfl_sync(0xffffffffu, p.ry0, src); b.ry1 = __shfl_sync(0xffffffffu, p.ry1, src);
    const unsigned long long bkey = __shfl_sync(0xffffffffu, key, src);
    const int w = b.rx1 - b.rx0, n = w * (b.ry1 - b.ry0);
    uint32_t run = 0;
    for (int base = 0; base < n; base += kTripTiles) run += visit_trip(b, base, n, w, W, H, gx, cull, s, bkey);
    if (lane == src) result = run;
  }
  return result;
}

}  // namespace gsb
