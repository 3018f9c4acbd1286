// Tile-ordered mirror of the resident cloud -- the copy the Lloyd iterations stream.
//
// The resident cloud keeps the reference's point order (np.where order: day-major, row-major,
// members/rafael/disparity/plugin.py:157), because labels, k-means++ draws and relocation ties
// are defined in that order.  In that order a group of 128 consecutive points is a 128 x 1
// pixel strip: long and thin, so a cluster boundary crosses many groups and each of them
// needs the per-point pass.  The mirror holds the same points re-ordered by cells of an x-y
// grid (about 256 points per cell, cells in row-major order): 128 consecutive points of the
// mirror are compact in x and y, far fewer groups touch a boundary, and most of them are
// settled from their cached summaries alone.  Sums are integers and labels are recomputed in
// the reference's order at the end of a fit, so results do not depend on this order (nor on
// the arrival order of the points inside a cell, which the atomic cursors leave open).
//
// Built once per (cloud, frame): histogram of the points over the cells, exclusive scan,
// scatter.  Two reads of the cloud and one write.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

struct MirrorGrid {
  float x0, y0;        // lower corner of the cloud's bounding box
  float inv_cx, inv_cy;  // 1 / cell size
  int gx, gy;          // cells per dimension
};

__device__ __forceinline__ int mirror_cell(const MirrorGrid& g, float x, float y) {
  const int cx = min(g.gx - 1, max(0, (int)((x - g.x0) * g.inv_cx)));
  const int cy = min(g.gy - 1, max(0, (int)((y - g.y0) * g.inv_cy)));
  return cy * g.gx + cx;
}

// counts[cell] += 1 for every point; consecutive points mostly share a cell, so each warp
// first merges equal cells (match.any) and issues one atomic per distinct cell.
__global__ void __launch_bounds__(kThreads) mirror_count_kernel(const float* pts, long long n, MirrorGrid g,
                                                                unsigned int* counts) {
  const int lane = threadIdx.x & 31;
  const long long n_groups = (n + kGroup - 1) / kGroup;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long grp = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); grp < n_groups; grp += stride) {
    const float* blk = pts + grp * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup);
    const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w};
    const long long i0 = grp * kGroup + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = i0 + e < n;
      const int cell = ok ? mirror_cell(g, xs[e], ys[e]) : -1;
      const unsigned int peers = __match_any_sync(0xffffffffu, cell);
      if (ok && lane == __ffs(peers) - 1) atomicAdd(&counts[cell], (unsigned int)__popc(peers));
    }
  }
}

// dst = offsets[cell] + (arrival rank inside the cell); copies x, y, z into the mirror.
__global__ void __launch_bounds__(kThreads) mirror_scatter_kernel(const float* pts, long long n, MirrorGrid g,
                                                                  const long long* offsets, unsigned int* cursor,
                                                                  float* tpts) {
  const int lane = threadIdx.x & 31;
  const long long n_groups = (n + kGroup - 1) / kGroup;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long grp = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); grp < n_groups; grp += stride) {
    const float* blk = pts + grp * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup), vz = ldg_stream_f4(blk + 2 * kGroup);
    const float xs[4] = {vx.x, vx.y, vx.z, vx.w}, ys[4] = {vy.x, vy.y, vy.z, vy.w}, zs[4] = {vz.x, vz.y, vz.z, vz.w};
    const long long i0 = grp * kGroup + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = i0 + e < n;
      const int cell = ok ? mirror_cell(g, xs[e], ys[e]) : -1;
      const unsigned int peers = __match_any_sync(0xffffffffu, cell);
      const int leader = __ffs(peers) - 1;
      unsigned int base = 0;
      if (ok && lane == leader) base = atomicAdd(&cursor[cell], (unsigned int)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (ok) {
        const long long dst = offsets[cell] + base + __popc(peers & ((1u << lane) - 1u));
        float* q = tpts + pt_off(dst);
        q[0] = xs[e];
        q[kGroup] = ys[e];
        q[2 * kGroup] = zs[e];
      }
    }
  }
}

}  // namespace mdkm
