// Shared device/host definitions for libmdkm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmdkm is written for sm_100a (B200); compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace mdkm {

constexpr int kThreads = 256;          // threads per CTA of the streaming kernels
constexpr int kMaxK = 2048;            // largest supported cluster count (shared-memory tables)
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: float -> integer by mantissa alignment
constexpr uint32_t kMagicBits = 0x4B400000u;
constexpr int kQuantBits = 22;         // |quantised coordinate| < 2^22

// Resident cloud layout ("blocked SoA"): consecutive blocks of kGroup = 128 points, each block
// x[128] y[128] z[128] (1536 contiguous bytes), so that one warp-group is ONE 1-D TMA bulk
// copy and every lane still reads 16 B-aligned float4s.  Capacity is whole blocks; the unused
// tail of the last block is zero.
constexpr int kGroup = 128;
constexpr int kBlockFloats = 3 * kGroup;
__host__ __device__ inline long long pt_off(long long i) { return (i >> 7) * kBlockFloats + (i & (kGroup - 1)); }

// Device-resident control block.  Written by the update / finalize kernels, read by every
// kernel of the Lloyd loop (early exit once `done`), mirrored to pinned host memory between
// batches of iterations only.
struct DevStatus {
  int done;       // 1: converged or max_iter reached
  int strict;     // 1: exit because no label changed (sklearn/_kmeans.py:721-726)
  int paused;     // 1: an empty cluster needs relocation (host-sequenced rare path)
  int iter;       // Lloyd iterations completed
  int max_iter;
  int n_empty;    // empty clusters found by the last update
  int k;
  int first;      // 1 until the first step has run (labels_old = -1 in sklearn)
  int pending;    // 1: the sums of E-step `last_seq` wait for their centroid update (applied by the next launch)
  int last_seq;   // launch index (StepParams::seq) of the last E-step that ran: its sums sit in acc[last_seq % 3],
                  // its table in table[last_seq & 1]
  unsigned long long n_changed;  // labels changed in the last step (global after allreduce)
  unsigned long long n_refined;  // point-iterations re-decided in float64 (this rank)
  unsigned long long n_relocated;
  double shift2;   // sum_j ||c_new - c_old||^2 of the last update
  double tol;      // scaled tolerance  mean(var(X)) * tol
  double inertia;  // written by the finalize pass
  float thresh;    // 2 * FP32 error bound of the fast distances for the current table
  unsigned int ticket;            // CTAs of the running step kernel that have flushed their sums
  unsigned long long epoch;       // fused steps completed since the communicator was created
  int xchg_timeout;               // 1: a kernel-side wait (peer sums, grid barrier) timed out (fatal)
  unsigned int grid_bar;          // arrival counter of the step kernel's grid barrier
  unsigned int upd_flag;          // (epoch << 2) | verdict of the deferred update CTA 0 applied for the grid (large tables)
  float next_thresh;              // its error threshold (published with upd_flag)
  unsigned long long work_sum;    // groups handed to the per-point pass, summed over the fused steps of a fit
  // diagnostics (MDKM_TIMING builds only): globaltimer stamps of the last fused step, ns
  unsigned long long t_start, t_first_done, t_last_done, t_update_done, t_classify_start, t_classify_done;
  unsigned long long t_a, t_b;  // stamps inside the deferred update
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Peer exchange of the K x 4 partial sums over NVLink (buffers shared through CUDA IPC, or plain
// peer pointers inside one process).  Every rank owns a buffer  packets[2][n_ranks][slot][2] of
// 8-byte (value half, epoch) packets; rank r writes its partials into [parity][r] of EVERY other
// rank's buffer (lloyd.cuh: peer_exchange_sums).
constexpr int kMaxRanks = 8;
struct PeerXchg {
  unsigned long long* data[kMaxRanks];   // data base of rank q's buffer (peer-mapped pointer)
  unsigned long long* flags[kMaxRanks];  // (unused by the packet protocol; kept for layout stability)
  int n_ranks, rank;
  int slot;                              // elements per slot (>= kpad*4 + 8)
  int pad;
};

// Frame of the resident cloud: x' = x - origin (exact for pixel grids), fixed-point scale.
struct Frame {
  double origin[3];
  double scale[3];     // power of two; q = rint((x - origin) * scale), |q| < 2^22
  double halfrange[3]; // max |x - origin| over ALL ranks (upper bound)
};

struct FrameF {
  float ox, oy, oz;    // origin as float (exactly representable)
  float sx, sy, sz;    // fixed-point scales (powers of two)
};

// One centroid table in global memory: [Kpad] float4 fast rows, then [Kpad] double4 exact rows.
//   fast row  = (-2c'x, -2c'y, -2c'z, ||c'||^2) rounded to FP32 ; padding rows = (0,0,0,+inf)
//   exact row = ( c'x,   c'y,   c'z,  ||c'||^2) in FP64, c' = c - origin
__host__ __device__ inline int pad_k(int k) { return (k + 7) & ~7; }
__host__ __device__ inline size_t exact_offset(int kpad) { return (size_t)kpad * 16; }

// Third section of the table, for k > kBucketMinK only: the centroids bucketed by an x-y grid
// of G x G cells over the frame (G = bucket_g(k)), so that the classification pass looks at
// the centroids near a group instead of all k:
//   BucketHdr | uint16 start[G*G + 1] | uint16 perm[Kpad]     (16-byte aligned, zero padded)
// perm lists the centroid indices bucket by bucket (any order inside a bucket).
constexpr int kBucketMinK = 33;
struct BucketHdr {
  float hx, hy;        // half-ranges of the frame in x and y
  float inv_x, inv_y;  // G / (2 hx), G / (2 hy)
};
__host__ __device__ inline int bucket_g(int k) {
  int g = 4;
  while (g < 32 && g * g < k) g *= 2;
  return g;
}
__host__ __device__ inline size_t bucket_offset(int kpad) { return (size_t)kpad * 48; }
__host__ __device__ inline size_t bucket_bytes(int k, int kpad) {
  if (k < kBucketMinK) return 0;
  const int g = bucket_g(k);
  return (sizeof(BucketHdr) + (size_t)(g * g + 1 + kpad) * 2 + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t table_bytes(int k, int kpad) { return (size_t)kpad * 48 + bucket_bytes(k, kpad); }
__device__ __forceinline__ int bucket_coord(float c, float h, float inv, int g) {
  return min(g - 1, max(0, (int)floorf((c + h) * inv)));
}

}  // namespace mdkm
