// Tile-ordered mirror of the resident cloud -- the copy the Lloyd iterations stream.
//
// The resident cloud keeps the reference's point order (np.where order: day-major, row-major,
// members/rafael/disparity/plugin.py:157), because labels, k-means++ draws and relocation ties
// are defined in that order.  In that order a group of 128 consecutive points is a 128 x 1
// pixel strip: long and thin, so a cluster boundary crosses many groups and each of them
// needs the per-point pass.  The mirror holds the same points re-ordered by the cells of an
// x-y grid with about 256 points per cell (the points of all days at those pixels): 128 consecutive
// points of the mirror are compact in x and y, far fewer groups touch a boundary, and most of
// them are settled from their cached summaries alone.  Sums are integers and the labels are
// recomputed in the reference's order at the end of a fit, so results do not depend on this
// order (nor on the arrival order of the points inside a cell, which the atomic cursors leave
// open).
//
// Built once per (cloud, frame): histogram of the points over the cells, exclusive scan,
// scatter -- two reads of the cloud and one write.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

constexpr int kMirrorMaxSeg = 1024;  // segment table held in shared memory

struct MirrorGrid {
  float x0, y0;          // lower corner of the cloud's bounding box
  float inv_cx, inv_cy;  // 1 / cell size
  int gx, gy;            // cells per dimension and segment
  int n_seg;
  const long long* seg_off;  // [n_seg + 1] point offsets of the segments (device)
  long long gpb;             // groups per band and segment
  long long n_virtual;       // n_bands * n_seg * gpb
};

// Cell of a point.  Cells are wide in x (a raster row contributes a run of consecutive points
// to a cell, which keeps the scattered writes sector-sized) and short in y.
__device__ __forceinline__ int mirror_cell(const MirrorGrid& g, int seg, float x, float y) {
  const int cx = min(g.gx - 1, max(0, (int)((x - g.x0) * g.inv_cx)));
  const int cy = min(g.gy - 1, max(0, (int)((y - g.y0) * g.inv_cy)));
  (void)seg;
  return cy * g.gx + cx;
}

// The cell of this lane's four points (-1 beyond n).  `seg` is the segment of the group's
// first point; a group that straddles a boundary walks on to the next segments.
__device__ __forceinline__ void mirror_cells4(const MirrorGrid& g, const long long* s_off, int seg, long long i0,
                                              long long n, const float (&xs)[4], const float (&ys)[4], int (&cell)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long i = i0 + e;
    while (seg < g.n_seg - 1 && i >= s_off[seg + 1]) ++seg;
    cell[e] = i < n ? mirror_cell(g, seg, xs[e], ys[e]) : -1;
  }
}

// Traversal order of the two passes below: band of rows outermost, segment (day) inside.  The
// points of a cell come from the same few raster rows of EVERY day; visiting those rows of all
// days together keeps the partial writes to the cell's range close in time, so they merge in
// L2 instead of being evicted half-filled.  Days are cut into n_bands runs of `gpb` groups;
// virtual index v = (band * n_seg + seg) * gpb + j  ->  group seg_first[seg] + band * gpb + j.
__device__ __forceinline__ long long mirror_group_of(const MirrorGrid& g, const long long* s_off, long long v,
                                                     int& seg) {
  const long long per_band = (long long)g.n_seg * g.gpb;
  const long long band = v / per_band;
  const long long r = v - band * per_band;
  seg = (int)(r / g.gpb);
  const long long j = r - (long long)seg * g.gpb;
  // groups that START inside the segment belong to it (a straddling group goes with its first point)
  const long long first = (s_off[seg] + kGroup - 1) / kGroup;
  const long long last = (s_off[seg + 1] + kGroup - 1) / kGroup;  // exclusive
  const long long grp = first + band * g.gpb + j;
  return grp < last ? grp : -1;
}

// counts[cell] += 1 for every point.  A lane's four consecutive points usually share a cell,
// and so do neighbouring lanes: equal cells are merged first in the lane, then across the
// warp (match.any), and one atomic is issued per distinct cell.
__global__ void __launch_bounds__(kThreads) mirror_count_kernel(const float* pts, long long n, MirrorGrid g,
                                                                unsigned int* counts) {
  __shared__ long long s_off[kMirrorMaxSeg + 1];
  for (int i = threadIdx.x; i <= g.n_seg; i += kThreads) s_off[i] = g.seg_off[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long v = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); v < g.n_virtual; v += stride) {
    int seg;
    const long long grp = mirror_group_of(g, s_off, v, seg);
    if (grp < 0) continue;  // warp-uniform
    const float* blk = pts + grp * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup);
    const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w};
    int cell[4];
    mirror_cells4(g, s_off, seg, grp * kGroup + lane * 4, n, xs, ys, cell);
    const bool same = cell[0] == cell[1] && cell[0] == cell[2] && cell[0] == cell[3];
    if (__all_sync(0xffffffffu, same)) {
      const unsigned int peers = __match_any_sync(0xffffffffu, cell[0]);
      if (cell[0] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&counts[cell[0]], 4u * (unsigned int)__popc(peers));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned int peers = __match_any_sync(0xffffffffu, cell[e]);
        if (cell[e] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&counts[cell[e]], (unsigned int)__popc(peers));
      }
    }
  }
}

// Exclusive scan of n 32-bit counts into 64-bit offsets in two launches: per-CTA totals of
// kScanTile counts, then every CTA adds up the totals before it (a few dozen values) and
// scans its own tile.
constexpr int kScanItems = 16;
constexpr int kScanTile = 1024 * kScanItems;

__global__ void __launch_bounds__(1024) mirror_tile_sums_kernel(const unsigned int* counts, long long n,
                                                                long long* tile_sums) {
  __shared__ long long s_warp[32];
  const long long base = (long long)blockIdx.x * kScanTile;
  long long mine = 0;
  for (int e = 0; e < kScanItems; ++e) {
    const long long i = base + (long long)e * 1024 + threadIdx.x;  // coalesced
    mine += i < n ? counts[i] : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < 32; ++w) t += s_warp[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) mirror_scan_kernel(const unsigned int* counts, long long n, long long* offsets,
                                                           const long long* tile_sums) {
  constexpr int kItems = kScanItems;
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < 32) {  // totals of the tiles before this one
    long long t = 0;
    for (int b = tid; b < (int)blockIdx.x; b += 32) t += tile_sums[b];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (tid == 0) s_carry = t;
  }
  __syncthreads();
  {
    const long long base = (long long)blockIdx.x * kScanTile;
    const long long i0 = base + (long long)tid * kItems;
    unsigned int v[kItems];
    long long mine = 0;
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
      v[e] = (i0 + e < n) ? counts[i0 + e] : 0u;
      mine += v[e];
    }
    long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
      long long ws = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      s_warp[lane] = ws;
    }
    __syncthreads();
    const long long carry = s_carry;
    long long run = carry + (w ? s_warp[w - 1] : 0) + (incl - mine);
#pragma unroll
    for (int e = 0; e < kItems; ++e) {
      if (i0 + e < n) offsets[i0 + e] = run;
      run += v[e];
    }
  }
}

// dst = offsets[cell] + (arrival rank inside the cell); copies x, y, z into the mirror.
// One point per lane and round (four rounds per group): the lanes that share a cell hold
// consecutive points of a raster row and get consecutive slots, so every store instruction
// writes whole runs instead of 4-byte fragments.
__global__ void __launch_bounds__(kThreads) mirror_scatter_kernel(const float* pts, long long n, MirrorGrid g,
                                                                  const long long* offsets, unsigned int* cursor,
                                                                  float* tpts) {
  __shared__ long long s_off[kMirrorMaxSeg + 1];
  for (int i = threadIdx.x; i <= g.n_seg; i += kThreads) s_off[i] = g.seg_off[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long v = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); v < g.n_virtual; v += stride) {
    int seg;
    const long long grp = mirror_group_of(g, s_off, v, seg);
    if (grp < 0) continue;  // warp-uniform
    const float* blk = pts + grp * kBlockFloats;
    float xs[4], ys[4], zs[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      xs[r] = __ldg(blk + r * 32 + lane);
      ys[r] = __ldg(blk + kGroup + r * 32 + lane);
      zs[r] = __ldg(blk + 2 * kGroup + r * 32 + lane);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long i = grp * kGroup + r * 32 + lane;
      const int cell = i < n ? mirror_cell(g, seg, xs[r], ys[r]) : -1;
      const unsigned int peers = __match_any_sync(0xffffffffu, cell);
      const int leader = __ffs(peers) - 1;
      unsigned int base = 0;
      if (cell >= 0 && lane == leader) base = atomicAdd(&cursor[cell], (unsigned int)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (cell >= 0) {
        float* q = tpts + pt_off(offsets[cell] + base + __popc(peers & ((1u << lane) - 1u)));
        q[0] = xs[r];
        q[kGroup] = ys[r];
        q[2 * kGroup] = zs[r];
      }
    }
  }
}

}  // namespace mdkm
