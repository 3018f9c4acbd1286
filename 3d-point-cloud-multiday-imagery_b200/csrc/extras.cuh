// Rare-path kernels: empty-cluster relocation (sklearn/cluster/_k_means_common.pyx:167-211).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

// ---------------------------------------------------------------------------------------
// Empty-cluster relocation.  One round per empty cluster, sequenced by the host while the
// Lloyd loop is paused: (1) max over points of the squared distance to the point's own OLD
// centre, (2) lowest global index attaining it, (3) the owner publishes the point, (4) every
// rank applies  sums[old] -= x ; sums[new] = x ; count[old] -= 1 ; count[new] = 1  to its
// (identical) global accumulators.  Distances in FP64; ties -> lowest index; rounds proceed
// in descending distance, empty clusters in ascending index (what np.argpartition(...)
// [:-n_empty-1:-1] yields for n_empty <= 2; for more its order is unspecified).
// scratch: [0] max dist bits  [1] arg index  [2..4] q  [5] old label  [6] "max>0 in round 0"
//          [8 .. 8+kMaxK) taken indices  [8+kMaxK .. 8+2kMaxK) empty-cluster list
// ---------------------------------------------------------------------------------------
struct RelocParams {
  const float* pts;  // blocked cloud
  long long n;
  const int* labels;   // int32, reference point order
  const unsigned char* table;
  unsigned long long* scratch;
  unsigned long long* acc;
  long long rank_offset;
  FrameF f;
  int k, kpad, round;
};

__device__ __forceinline__ int reloc_label(const RelocParams& p, long long i) { return p.labels[i]; }

__device__ __forceinline__ bool reloc_taken(const RelocParams& p, long long gi) {
  for (int t = 0; t < p.round; ++t)
    if (p.scratch[8 + t] == (unsigned long long)gi) return true;
  return false;
}

__device__ __forceinline__ double reloc_dist(const RelocParams& p, const double4* c64, long long i) {
  const double4 c = c64[reloc_label(p, i)];
  const float* q = p.pts + pt_off(i);
  const double dx = ((double)q[0] - (double)p.f.ox) - c.x;
  const double dy = ((double)q[kGroup] - (double)p.f.oy) - c.y;
  const double dz = ((double)q[2 * kGroup] - (double)p.f.oz) - c.z;
  return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(kThreads) reloc_maxdist_kernel(const RelocParams p) {
  const double4* c64 = reinterpret_cast<const double4*>(p.table + exact_offset(p.kpad));
  unsigned long long best = 0ull;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += (long long)gridDim.x * kThreads) {
    if (reloc_taken(p, p.rank_offset + i)) continue;
    const unsigned long long b = (unsigned long long)__double_as_longlong(reloc_dist(p, c64, i));
    best = b > best ? b : best;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_down_sync(0xffffffffu, best, o);
    best = t > best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(&p.scratch[0], best);
}

__global__ void __launch_bounds__(kThreads) reloc_argidx_kernel(const RelocParams p) {
  const double4* c64 = reinterpret_cast<const double4*>(p.table + exact_offset(p.kpad));
  const unsigned long long target = p.scratch[0];
  unsigned long long best = ~0ull;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += (long long)gridDim.x * kThreads) {
    const long long gi = p.rank_offset + i;
    if (reloc_taken(p, gi)) continue;
    const unsigned long long b = (unsigned long long)__double_as_longlong(reloc_dist(p, c64, i));
    if (b == target && (unsigned long long)gi < best) best = (unsigned long long)gi;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_down_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&p.scratch[1], best);
}

__global__ void reloc_payload_kernel(const RelocParams p) {
  if (threadIdx.x != 0) return;
  const long long gi = (long long)p.scratch[1];
  const long long i = gi - p.rank_offset;
  if (p.scratch[1] == ~0ull || i < 0 || i >= p.n) return;  // another rank owns the point
  const float* q = p.pts + pt_off(i);
  const float xc = q[0] - p.f.ox, yc = q[kGroup] - p.f.oy, zc = q[2 * kGroup] - p.f.oz;
  const int qx = (int)(__float_as_uint(fmaf(xc, p.f.sx, kMagic)) - kMagicBits);
  const int qy = (int)(__float_as_uint(fmaf(yc, p.f.sy, kMagic)) - kMagicBits);
  const int qz = (int)(__float_as_uint(fmaf(zc, p.f.sz, kMagic)) - kMagicBits);
  p.scratch[2] = (unsigned long long)(long long)qx;
  p.scratch[3] = (unsigned long long)(long long)qy;
  p.scratch[4] = (unsigned long long)(long long)qz;
  p.scratch[5] = (unsigned long long)reloc_label(p, i);
}

__global__ void reloc_apply_kernel(const RelocParams p, DevStatus* st) {
  if (threadIdx.x != 0) return;
  unsigned long long* sc = p.scratch;
  if (p.round == 0) {
    sc[6] = sc[0] != 0ull ? 1ull : 0ull;  // pyx:188-191: nothing to do when max distance is 0
    int ne = 0;
    for (int j = 0; j < p.k; ++j)
      if (p.acc[j * 4 + 3] == 0ull) sc[8 + kMaxK + ne++] = (unsigned long long)j;
  }
  sc[8 + p.round] = sc[1];  // taken (also when skipped: harmless)
  if (!sc[6] || sc[1] == ~0ull) return;
  const int new_id = (int)sc[8 + kMaxK + p.round];
  const int old_id = (int)sc[5];
  for (int d = 0; d < 3; ++d) {
    p.acc[old_id * 4 + d] -= sc[2 + d];  // pyx:205
    p.acc[new_id * 4 + d] = sc[2 + d];   // pyx:206
  }
  p.acc[new_id * 4 + 3] = 1ull;   // pyx:208
  p.acc[old_id * 4 + 3] -= 1ull;  // pyx:209
  st->n_relocated += 1ull;
}

}  // namespace mdkm
