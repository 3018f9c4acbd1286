// k-means++ seeding on the device (sklearn/cluster/_kmeans.py:180-278, dense, unit weights).
//
// The reference's only k-means call, KMeans(n_clusters, random_state=42, n_init=10) at
// members/jasraj/land_use_classification/core.py:227-228, seeds every restart this way.  The
// host keeps numpy's RandomState stream (the uniforms are uploaded once, in scikit-learn's draw
// order); distances, potentials, the cumulative-sum search and the greedy choice among the
// local trials all run here, with no host synchronisation between centres.
//
// State per point: closest[i] = squared distance to the nearest centre chosen so far, FP64,
// direct form sum((x - c)^2) on the original FP32 coordinates (differences are exact in FP64).
//
// Cumulative sum used by the search (np.searchsorted(np.cumsum(closest), r), side="left"):
// a fixed three-level order -- cell = 128 points (one block of the resident cloud), block =
// 32 cells, prefix over blocks -- so the result is deterministic and independent of the grid:
//   C[i] = P_blk[b-1] + (cell sums of block b before cell c, added in order)
//          + (closest of cell c up to i, added in order)
// It differs from numpy's strictly sequential cumsum only by FP64 rounding (relative 1e-16),
// i.e. a draw picks another index only if it lands within that distance of a bucket edge.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

constexpr int kKppCellsPerBlock = 32;                    // 4096 points per block
constexpr int kKppBlockPts = kKppCellsPerBlock * kGroup;
constexpr int kKppMaxTrials = 16;
// exchange scratch of the sharded variant (doubles): rank totals | candidates | rank point counts
constexpr int kXTot = 0;
constexpr int kXCand = kMaxRanks;
constexpr int kXCnt = kMaxRanks + 4 * kKppMaxTrials;
constexpr int kXWords = kXCnt + kMaxRanks;

struct KppState {
  double pot;                          // current potential = C[n-1]
  double cand_xyz[kKppMaxTrials][3];   // candidate centres of the current round
  long long cand_idx[kKppMaxTrials];
  double best_xyz[3];                  // centre committed by the last choose step
  long long best_idx;
  double pots[kKppMaxTrials];          // candidate potentials of the current round
  unsigned int ticket;
  unsigned int pad;
};

struct KppParams {
  const float* pts;        // blocked cloud
  long long n;
  double* closest;         // [cap] (whole cells)
  double* cell_sum;        // [n_cells]
  double* blk_sum;         // [n_blk]
  double* blk_prefix;      // [n_blk] inclusive
  double* partials;        // [grid][kKppMaxTrials]
  const double* rand_vals; // [(k-1) * n_trials] uniforms in [0,1)
  KppState* st;
  double* centers_out;     // [k*3] device
  long long* indices_out;  // [k] device
  int n_trials;
  int round;               // index of the centre being chosen (1..k-1); 0 = first centre
  long long first_index;   // centre 0 (local index; -1: another rank owns it, see st->best_xyz)
  // multi-rank (points sharded over ranks, one cumulative sum over all of them in rank order)
  int n_ranks, rank;
  long long rank_offset;   // global index of this rank's first point
  double* xbuf;            // exchange scratch, layout kXTot / kXCand / kXCnt
  int choose;              // pot kernel: 1 = pick the winner in-kernel, 0 = the host reduces first
};

__device__ __forceinline__ double kpp_dist(float x, float y, float z, double cx, double cy, double cz) {
  const double dx = (double)x - cx, dy = (double)y - cy, dz = (double)z - cz;
  return fma(dz, dz, fma(dy, dy, dx * dx));
}

// Commits a centre: closest = min(closest, d(., centre)) (first centre: closest = d), and the
// cell / block sums of the new closest.  One CTA per block of 32 cells, one warp per 4 cells.
// round 0 takes the centre from first_index, later rounds from st->best_*.
__global__ void __launch_bounds__(kThreads) kpp_commit_kernel(const KppParams p) {
  __shared__ double s_cell[kKppCellsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double cx, cy, cz;
  if (p.round == 0 && p.n_ranks <= 1) {
    const float* q = p.pts + pt_off(p.first_index);
    cx = (double)q[0]; cy = (double)q[kGroup]; cz = (double)q[2 * kGroup];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      p.centers_out[0] = cx; p.centers_out[1] = cy; p.centers_out[2] = cz;
      p.indices_out[0] = p.first_index;
    }
  } else {
    cx = p.st->best_xyz[0]; cy = p.st->best_xyz[1]; cz = p.st->best_xyz[2];
  }
  const long long n_cells = (p.n + kGroup - 1) / kGroup;
  const long long cell0 = (long long)blockIdx.x * kKppCellsPerBlock;
  for (int j = warp; j < kKppCellsPerBlock; j += kThreads / 32) {
    const long long cell = cell0 + j;
    double s = 0.0;
    if (cell < n_cells) {
      const float* blk = p.pts + cell * kBlockFloats + lane * 4;
      const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup), vz = ldg_stream_f4(blk + 2 * kGroup);
      const float ax[4] = {vx.x, vx.y, vx.z, vx.w}, ay[4] = {vy.x, vy.y, vy.z, vy.w}, az[4] = {vz.x, vz.y, vz.z, vz.w};
      const long long i0 = cell * kGroup + lane * 4;
      double2* cp = reinterpret_cast<double2*>(p.closest + i0);
      double c[4] = {0, 0, 0, 0};
      if (p.round != 0) {
        const double2 a = cp[0], b = cp[1];
        c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double d = kpp_dist(ax[e], ay[e], az[e], cx, cy, cz);
        if (p.round != 0) d = fmin(c[e], d);
        c[e] = (i0 + e < p.n) ? d : 0.0;  // tail of the last cell contributes nothing
      }
      cp[0] = make_double2(c[0], c[1]);
      cp[1] = make_double2(c[2], c[3]);
      s = ((c[0] + c[1]) + c[2]) + c[3];
      // fixed butterfly: every lane ends with the same, order-defined cell sum
      for (int o = 1; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) p.cell_sum[cell] = s;
    }
    if (lane == 0) s_cell[j] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int j = 0; j < kKppCellsPerBlock; ++j) t += s_cell[j];
    p.blk_sum[blockIdx.x] = t;
  }
}

// searchsorted(C, r) on this rank's cumulative sum, clipped to the last point; one warp.
// The candidate goes to the state block (single rank) or, as (global index, x, y, z), to
// `out` for the exchange between ranks.
__device__ __forceinline__ void kpp_locate(const KppParams& p, long long n_blk, double r, int trial, int lane,
                                           double* out) {
  // first block whose inclusive prefix reaches r
  long long lo = 0, hi = n_blk - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (p.blk_prefix[mid] >= r) hi = mid; else lo = mid + 1;
  }
  const long long b = lo;
  double base = b > 0 ? p.blk_prefix[b - 1] : 0.0;
  const long long n_cells = (p.n + kGroup - 1) / kGroup;
  const long long cell0 = b * kKppCellsPerBlock;
  const int cells_here = (int)min((long long)kKppCellsPerBlock, n_cells - cell0);
  const double my_cell = (lane < cells_here) ? p.cell_sum[cell0 + lane] : 0.0;
  int c = cells_here - 1;
  for (int j = 0; j < cells_here; ++j) {
    const double v = __shfl_sync(0xffffffffu, my_cell, j);
    if (base + v >= r) { c = j; break; }
    base += v;
  }
  const long long cell = cell0 + c;
  const long long i0 = cell * kGroup;
  const int pts_here = (int)min((long long)kGroup, p.n - i0);
  double v4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v4[e] = p.closest[i0 + lane * 4 + e];
  int found = pts_here - 1;
  double acc = base;
  for (int j = 0; j < pts_here; ++j) {
    const int e = j & 3;
    const double mine = e == 0 ? v4[0] : (e == 1 ? v4[1] : (e == 2 ? v4[2] : v4[3]));
    const double v = __shfl_sync(0xffffffffu, mine, j >> 2);
    acc += v;
    if (acc >= r) { found = j; break; }
  }
  long long idx = i0 + found;
  if (idx > p.n - 1) idx = p.n - 1;  // np.clip(candidate_ids, None, n - 1)
  if (lane == 0) {
    const float* q = p.pts + pt_off(idx);
    if (out) {
      out[0] = (double)(p.rank_offset + idx);
      out[1] = (double)q[0];
      out[2] = (double)q[kGroup];
      out[3] = (double)q[2 * kGroup];
    } else {
      p.st->cand_idx[trial] = idx;
      p.st->cand_xyz[trial][0] = (double)q[0];
      p.st->cand_xyz[trial][1] = (double)q[kGroup];
      p.st->cand_xyz[trial][2] = (double)q[2 * kGroup];
    }
  }
}

// Inclusive prefix over the block sums (single CTA, fixed order), the potential, and -- on a
// single rank -- the search of this round's n_trials draws:
// cand = searchsorted(C, rand * pot), clipped to n-1.
__global__ void __launch_bounds__(1024) kpp_search_kernel(const KppParams p, long long n_blk) {
  __shared__ double s_part[1024];
  __shared__ double s_pot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // each thread owns a contiguous run of blocks; runs are combined in thread order
  const long long per = (n_blk + 1023) / 1024;
  const long long b0 = (long long)tid * per, b1 = min(n_blk, b0 + per);
  double t = 0.0;
  for (long long b = b0; b < b1; ++b) t += p.blk_sum[b];
  s_part[tid] = t;
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int i = 0; i < 1024; ++i) {
      const double v = s_part[i];
      s_part[i] = run;  // exclusive
      run += v;
    }
    s_pot = run;
    p.st->pot = run;
    if (p.n_ranks > 1) p.xbuf[kXTot + p.rank] = run;  // this rank's slot; the others stay zero for the exchange
  }
  __syncthreads();
  double run = s_part[tid];
  for (long long b = b0; b < b1; ++b) {
    run += p.blk_sum[b];
    p.blk_prefix[b] = run;
  }
  __syncthreads();
  if (p.n_ranks > 1) return;  // the draws need every rank's total: kpp_search_sharded_kernel
  if (warp >= p.n_trials) return;
  const double r = p.rand_vals[(size_t)(p.round - 1) * p.n_trials + warp] * s_pot;  // one warp per trial
  kpp_locate(p, n_blk, r, warp, lane, nullptr);
}

// Sharded variant, once the ranks' totals are gathered in xbuf[0..n_ranks): the global
// cumulative sum is the concatenation of the ranks' sums in rank order; the rank whose range
// contains a draw locates it and publishes the candidate in xbuf[kXCand + 4*trial ..),
// everybody else leaves zeros there for the sum-exchange that follows.
__global__ void __launch_bounds__(kKppMaxTrials * 32) kpp_search_sharded_kernel(const KppParams p, long long n_blk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= p.n_trials) return;
  double pot = 0.0, before = 0.0;
  for (int q = 0; q < p.n_ranks; ++q) {
    if (q == p.rank) before = pot;
    pot += p.xbuf[q];
  }
  if (lane == 0 && warp == 0) p.st->pot = pot;
  const double r = p.rand_vals[(size_t)(p.round - 1) * p.n_trials + warp] * pot;
  // owner: the first rank with points whose inclusive prefix reaches r; overshoots (rounding)
  // go to the last rank that has points, like np.clip does on one rank
  int owner = -1, last_nonempty = -1;
  double acc = 0.0;
  for (int q = 0; q < p.n_ranks; ++q) {
    const double tq = p.xbuf[q];
    const bool has = p.xbuf[kXCnt + q] > 0.0;  // point counts, gathered once per call
    if (has) last_nonempty = q;
    acc += tq;
    if (owner < 0 && has && acc >= r) owner = q;
  }
  if (owner < 0) owner = last_nonempty;
  if (owner != p.rank) return;
  kpp_locate(p, n_blk, r - before, warp, lane, p.xbuf + kXCand + warp * 4);
}

// After the candidate exchange: unpack xbuf into the state block (all ranks, identical).
__global__ void kpp_unpack_candidates_kernel(const KppParams p) {
  const int t = threadIdx.x;
  if (t >= p.n_trials) return;
  const double* in = p.xbuf + kXCand + t * 4;
  p.st->cand_idx[t] = (long long)in[0];
  p.st->cand_xyz[t][0] = in[1];
  p.st->cand_xyz[t][1] = in[2];
  p.st->cand_xyz[t][2] = in[3];
}

// Sharded: np.argmin over the exchanged potentials, or (round 0) the exchanged first centre.
__global__ void kpp_choose_kernel(const KppParams p) {
  if (threadIdx.x != 0) return;
  int best = 0;
  if (p.round > 0) {
    for (int t = 1; t < p.n_trials; ++t)
      if (p.st->pots[t] < p.st->pots[best]) best = t;  // first minimum
  }
  for (int d = 0; d < 3; ++d) {
    p.st->best_xyz[d] = p.st->cand_xyz[best][d];
    p.centers_out[(size_t)p.round * 3 + d] = p.st->cand_xyz[best][d];
  }
  p.st->best_idx = p.st->cand_idx[best];
  p.indices_out[p.round] = p.st->cand_idx[best];
}

// Sharded round 0: the owner of the first centre publishes it as candidate 0.
__global__ void kpp_first_candidate_kernel(const KppParams p) {
  if (threadIdx.x != 0 || p.first_index < 0) return;
  const float* q = p.pts + pt_off(p.first_index);
  double* out = p.xbuf + kXCand;
  out[0] = (double)(p.rank_offset + p.first_index);
  out[1] = (double)q[0];
  out[2] = (double)q[kGroup];
  out[3] = (double)q[2 * kGroup];
}

// Potential of every candidate: sum_i min(closest[i], d(x_i, cand_t)); fixed-order reductions;
// the last CTA adds the per-CTA partials in CTA order, takes np.argmin (first minimum) and
// publishes the winner as centre `round`.
template <int T>
__global__ void __launch_bounds__(kThreads) kpp_pot_kernel(const KppParams p) {
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double cx[T], cy[T], cz[T], pot[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    cx[t] = p.st->cand_xyz[t][0]; cy[t] = p.st->cand_xyz[t][1]; cz[t] = p.st->cand_xyz[t][2];
    pot[t] = 0.0;
  }
  const long long n_cells = (p.n + kGroup - 1) / kGroup;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long cell = (long long)blockIdx.x * (kThreads / 32) + warp; cell < n_cells; cell += stride) {
    const float* blk = p.pts + cell * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup), vz = ldg_stream_f4(blk + 2 * kGroup);
    const float ax[4] = {vx.x, vx.y, vx.z, vx.w}, ay[4] = {vy.x, vy.y, vy.z, vy.w}, az[4] = {vz.x, vz.y, vz.z, vz.w};
    const long long i0 = cell * kGroup + lane * 4;
    const double2* cp = reinterpret_cast<const double2*>(p.closest + i0);
    const double2 a = cp[0], b = cp[1];
    const double c[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i0 + e < p.n) {
#pragma unroll
        for (int t = 0; t < T; ++t) pot[t] += fmin(c[e], kpp_dist(ax[e], ay[e], az[e], cx[t], cy[t], cz[t]));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    double v = pot[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) s += s_red[w];
      p.partials[(size_t)blockIdx.x * kKppMaxTrials + t] = s;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&p.st->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < T) {
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += p.partials[(size_t)b * kKppMaxTrials + threadIdx.x];
    p.st->pots[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0 && !p.choose) p.st->ticket = 0u;  // sharded: the ranks' potentials are summed first
  if (threadIdx.x == 0 && p.choose) {
    int best = 0;
    for (int t = 1; t < T; ++t)
      if (p.st->pots[t] < p.st->pots[best]) best = t;  // np.argmin: first minimum
    for (int d = 0; d < 3; ++d) {
      p.st->best_xyz[d] = p.st->cand_xyz[best][d];
      p.centers_out[(size_t)p.round * 3 + d] = p.st->cand_xyz[best][d];
    }
    p.st->best_idx = p.st->cand_idx[best];
    p.indices_out[p.round] = p.st->cand_idx[best];
    p.st->ticket = 0u;
  }
}

typedef void (*KppPotKernel)(const KppParams);
inline KppPotKernel kpp_pot_variant(int T) {
  switch (T) {
    case 1: return kpp_pot_kernel<1>;
    case 2: return kpp_pot_kernel<2>;
    case 3: return kpp_pot_kernel<3>;
    case 4: return kpp_pot_kernel<4>;
    case 5: return kpp_pot_kernel<5>;
    case 6: return kpp_pot_kernel<6>;
    case 7: return kpp_pot_kernel<7>;
    case 8: return kpp_pot_kernel<8>;
    case 9: return kpp_pot_kernel<9>;
    case 10: return kpp_pot_kernel<10>;
    case 11: return kpp_pot_kernel<11>;
    case 12: return kpp_pot_kernel<12>;
    case 13: return kpp_pot_kernel<13>;
    case 14: return kpp_pot_kernel<14>;
    case 15: return kpp_pot_kernel<15>;
    case 16: return kpp_pot_kernel<16>;
    default: return nullptr;
  }
}

}  // namespace mdkm
