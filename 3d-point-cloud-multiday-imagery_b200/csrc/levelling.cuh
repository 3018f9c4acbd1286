// Percentile ground-levelling of the resident cloud (members/rafael/disparity/plugin.py:181-192):
//   h_min, h_max = np.percentile(z, 2), np.percentile(z, 98)   (numpy "linear" method)
//   h_norm = clip((z - h_min) / (h_max - h_min + 1e-6), 0, 1)  -> the 'height' colour property
//   z = z - h_min                                               -> ground level at 0
// per SEGMENT of the cloud (one segment per day, as the reference does it per stereo pair).
//
// The four order statistics a segment needs (floor and ceil neighbours of the two virtual
// indices) are found exactly with an MSB-first radix select over the order-preserving 32-bit
// key of z: three histogram passes (11 + 11 + 10 bits), each reading only z (4 B/point).  The
// interpolation itself is four scalars per segment and is done by the host in FP64 exactly as
// numpy's _lerp does it.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "unproject.cuh"  // f2ord / ord2f

namespace mdkm {

constexpr int kSelTargets = 4;
constexpr int kSelBins = 2048;
constexpr int kSelCtasPerSeg = 64;

struct SelTarget {
  unsigned int prefix;      // key bits decided so far (right-aligned)
  unsigned int pad;
  long long rank;           // rank still to resolve inside the matching subset
};

struct SelParams {
  const float* pts;           // blocked cloud
  const long long* seg_off;   // [n_seg + 1] point offsets (device)
  unsigned int* hist;         // [n_seg][kSelTargets][kSelBins]
  SelTarget* targets;         // [n_seg][kSelTargets]
  int n_seg;
  int shift;                  // bin = (key >> shift) & (2^bits - 1)
  int bits;
  int prefix_shift;           // a point matches target t iff (key >> prefix_shift) == prefix; 32 = all
};

// grid (kSelCtasPerSeg, n_seg): CTA (b, s) histograms a strided share of segment s.
__global__ void __launch_bounds__(kThreads) select_hist_kernel(const SelParams p) {
  __shared__ unsigned int s_hist[kSelTargets][kSelBins];
  const int s = blockIdx.y;
  const int n_t = p.prefix_shift >= 32 ? 1 : kSelTargets;
  for (int i = threadIdx.x; i < n_t * kSelBins; i += kThreads) (&s_hist[0][0])[i] = 0u;
  unsigned int pre[kSelTargets];
#pragma unroll
  for (int t = 0; t < kSelTargets; ++t) pre[t] = p.targets[s * kSelTargets + t].prefix;
  __syncthreads();
  const long long lo = p.seg_off[s], hi = p.seg_off[s + 1];
  const unsigned int mask = (1u << p.bits) - 1u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long c_lo = lo / kGroup, c_hi = (hi + kGroup - 1) / kGroup;
  for (long long cell = c_lo + (long long)blockIdx.x * (kThreads / 32) + warp; cell < c_hi;
       cell += (long long)gridDim.x * (kThreads / 32)) {
    const float4 vz = ldg_stream_f4(p.pts + cell * kBlockFloats + 2 * kGroup + lane * 4);
    const float z[4] = {vz.x, vz.y, vz.z, vz.w};
    const long long i0 = cell * kGroup + lane * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = i0 + e;
      if (i < lo || i >= hi) continue;
      const unsigned int key = f2ord(z[e]);
      const unsigned int bin = (key >> p.shift) & mask;
      if (p.prefix_shift >= 32) {
        atomicAdd(&s_hist[0][bin], 1u);
      } else {
        const unsigned int top = key >> p.prefix_shift;
#pragma unroll
        for (int t = 0; t < kSelTargets; ++t)
          if (top == pre[t]) atomicAdd(&s_hist[t][bin], 1u);
      }
    }
  }
  __syncthreads();
  unsigned int* out = p.hist + (size_t)s * kSelTargets * kSelBins;
  for (int i = threadIdx.x; i < n_t * kSelBins; i += kThreads) {
    const unsigned int v = (&s_hist[0][0])[i];
    if (v) atomicAdd(&out[i], v);
  }
}

// grid (n_seg), 4 warps: warp t walks the histogram of target t to the bin holding its rank,
// appends the bin to the prefix and makes the rank relative to that bin.  Clears the histogram.
__global__ void __launch_bounds__(kSelTargets * 32) select_pick_kernel(const SelParams p) {
  const int s = blockIdx.x, t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool shared_hist = p.prefix_shift >= 32;  // first pass: one histogram for all targets
  unsigned int* hist = p.hist + ((size_t)s * kSelTargets + (shared_hist ? 0 : t)) * kSelBins;
  SelTarget tg = p.targets[s * kSelTargets + t];
  const int n_bins = 1 << p.bits;
  const int per = n_bins / 32;
  long long mine = 0;
  for (int j = 0; j < per; ++j) mine += hist[lane * per + j];
  long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned int owners = __ballot_sync(0xffffffffu, incl > tg.rank);
  const int owner = owners ? __ffs(owners) - 1 : 31;
  int bin = n_bins - 1;
  long long before = 0;
  if (lane == owner) {
    long long run = incl - mine;
    for (int j = 0; j < per; ++j) {
      const long long c = hist[lane * per + j];
      if (run + c > tg.rank) { bin = lane * per + j; before = run; break; }
      run += c;
      before = run;
    }
  }
  bin = __shfl_sync(0xffffffffu, bin, owner);
  before = __shfl_sync(0xffffffffu, before, owner);
  __syncthreads();  // every warp has read the (possibly shared) histogram
  if (lane == 0) {
    tg.prefix = (tg.prefix << p.bits) | (unsigned int)bin;
    tg.rank -= before;
    p.targets[s * kSelTargets + t] = tg;
  }
  if (!shared_hist || t == 0)
    for (int i = lane; i < n_bins; i += 32) hist[i] = 0u;
}

struct LevelParams {
  float* pts;                // blocked cloud, z updated in place
  const long long* seg_off;  // [n_seg + 1]
  const double* levels;      // [n_seg][2] = h_min, h_max - h_min + 1e-6
  float* height_norm;        // [n] or nullptr
  long long n;
  int n_seg;
};

// z -= h_min (plugin.py:191) and h_norm = clip((z - h_min) / div, 0, 1) (plugin.py:184-185),
// evaluated in FP64 like the reference and rounded once to FP32.
__global__ void __launch_bounds__(kThreads) level_apply_kernel(const LevelParams p) {
  extern __shared__ long long s_off[];
  for (int i = threadIdx.x; i <= p.n_seg; i += kThreads) s_off[i] = p.seg_off[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_cells = (p.n + kGroup - 1) / kGroup;
  for (long long cell = (long long)blockIdx.x * (kThreads / 32) + warp; cell < n_cells;
       cell += (long long)gridDim.x * (kThreads / 32)) {
    float* zp = p.pts + cell * kBlockFloats + 2 * kGroup + lane * 4;
    const float4 vz = *reinterpret_cast<const float4*>(zp);
    float z[4] = {vz.x, vz.y, vz.z, vz.w};
    float hn[4] = {0.f, 0.f, 0.f, 0.f};
    const long long i0 = cell * kGroup + lane * 4;
    // segment of the first point (upper bound search), later points walk forward
    int lo = 0, hi = p.n_seg;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid + 1] <= i0) lo = mid + 1; else hi = mid;
    }
    int s = lo;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long i = i0 + e;
      if (i >= p.n) continue;
      while (s < p.n_seg - 1 && i >= s_off[s + 1]) ++s;
      const double h_min = p.levels[2 * s], div = p.levels[2 * s + 1];
      const double zr = (double)z[e] - h_min;
      hn[e] = (float)fmin(fmax(zr / div, 0.0), 1.0);
      z[e] = (float)zr;
    }
    *reinterpret_cast<float4*>(zp) = make_float4(z[0], z[1], z[2], z[3]);
    if (p.height_norm) {
      if (i0 + 3 < p.n && ((reinterpret_cast<uintptr_t>(p.height_norm) & 15) == 0)) {
        *reinterpret_cast<float4*>(p.height_norm + i0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (i0 + e < p.n) p.height_norm[i0 + e] = hn[e];
      }
    }
  }
}

// Output offset of the first pixel of every day of the unprojected range: offsets of the
// chunk holding the day boundary plus the valid pixels of that chunk before it.  One warp per
// boundary; day_off[0] = 0 and day_off[n_days] = total are written by the host.
__global__ void day_offsets_kernel(const UnprojParams p, int n_days, long long* day_off) {
  const int d = blockIdx.x + 1;  // boundary between day d-1 and day d of the range
  if (d >= n_days) return;
  const int lane = threadIdx.x;
  const long long first_day_end = ((p.pix_begin / p.HW) + 1) * p.HW - p.pix_begin;  // local index
  const long long bl = first_day_end + (long long)(d - 1) * p.HW;
  const long long c = bl / kChunk;
  unsigned int cnt = 0;
  for (long long i = c * kChunk + lane; i < bl; i += 32) {
    float hv;
    const bool ok = load_height1(p, i, hv);
    cnt += ok ? 1u : 0u;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) day_off[d] = p.chunk_offsets[c] + (long long)cnt;
}

}  // namespace mdkm
