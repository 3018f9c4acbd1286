// Lloyd k-means kernels for d = 3 on sm_100a.
//
//   lloyd_step_kernel   K2+K3 fused: nearest-centroid assignment + per-cluster sum/count
//                       (replaces sklearn/cluster/_k_means_lloyd.pyx:168-218, one pass over
//                        the points; the reference reaches it through KMeans.fit at
//                        members/jasraj/land_use_classification/core.py:227-228)
//   lloyd_update_kernel K4: centroid update, centre shift, convergence, next centroid table
//                       (sklearn/cluster/_k_means_common.pyx:274-311, _kmeans.py:721-738)
//   lloyd_final_kernel  final E-step when the exit was not strict + inertia + int32 labels
//                       (sklearn/cluster/_kmeans.py:742-756, _k_means_common.pyx:94-124)
//
// Numerics (DESIGN.md "Exactness"):
//   * distances: FP32 CUDA cores, expanded form  ||c'||^2 - 2 x'.c'  (3 FFMA per pair) in a
//     frame whose origin makes pixel-grid coordinates exact.  A rigorous bound E on the
//     FP32 error of that expression is carried with every centroid table; a point whose best
//     and second-best FP32 distances are closer than 2E is re-decided in FP64 with the exact
//     centroids, so the label equals the FP64 argmin (lowest index on ties) up to FP64
//     rounding.
//   * sums: coordinates are rounded once to a 2^-22-of-range fixed-point grid and summed as
//     integers (warp REDUX -> shared int64 -> global int64).  Integer addition is
//     associative, so the sums are bit-identical for any grid size, run, or number of GPUs.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

struct StepParams {
  const float* x;
  const float* y;
  const float* z;
  long long n;
  void* labels;               // uint8 (k <= 256) or uint16, n rounded up to the tile
  const unsigned char* table; // centroid table (see common.cuh)
  unsigned long long* acc;    // [kpad*4] (qx,qy,qz,count) + [kpad*4 + 0] n_changed
  DevStatus* st;
  FrameF f;
  int k, kpad;
  int ignore_status;          // 1: test hook (run even when done/paused)
};

struct FinalParams {
  const float* x;
  const float* y;
  const float* z;
  long long n;
  const void* labels;         // stored labels of the last step
  int* labels_out;            // int32[n] or nullptr
  const unsigned char* table;
  double* partials;           // [gridDim.x] inertia partials
  unsigned int* ticket;
  DevStatus* st;
  FrameF f;
  int k, kpad;
  int force_assign;           // 1: always recompute labels (predict / test hook)
};

// exact centroid row through the read-only path (two 16 B loads)
__device__ __forceinline__ double4 ld_c64(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// ---------------------------------------------------------------------------------------
// Assignment of P points held in registers.  xc/yc/zc are centred FP32 coordinates,
// xo/yo/zo the original ones (for the FP64 refine).  Returns labels in lab[].
// ---------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void assign_points(const float (&xc)[P], const float (&yc)[P],
                                              const float (&zc)[P], const float (&xo)[P],
                                              const float (&yo)[P], const float (&zo)[P],
                                              const float4* __restrict__ s_c,
                                              const double4* __restrict__ c64, int k, int kpad,
                                              float thresh, const FrameF& f, int (&lab)[P],
                                              unsigned int& n_refined) {
  float best[P], second[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    best[p] = __int_as_float(0x7f800000);
    second[p] = __int_as_float(0x7f800000);
    lab[p] = 0;
  }
#pragma unroll 4
  for (int j = 0; j < kpad; ++j) {
    const float4 c = s_c[j];  // LDS.128, warp broadcast
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float d = fmaf(xc[p], c.x, fmaf(yc[p], c.y, fmaf(zc[p], c.z, c.w)));
      const bool lt = d < best[p];  // strict: lowest index wins ties (pyx:205-213)
      second[p] = fminf(second[p], fmaxf(d, best[p]));
      best[p] = fminf(best[p], d);
      lab[p] = lt ? j : lab[p];
    }
  }
  // FP64 refine of the points the FP32 pass cannot decide (rare; see file header).
#pragma unroll
  for (int p = 0; p < P; ++p) {
    if (!(second[p] - best[p] > thresh)) {
      const double X = (double)xo[p] - (double)f.ox;
      const double Y = (double)yo[p] - (double)f.oy;
      const double Z = (double)zo[p] - (double)f.oz;
      double bd = 1.0 / 0.0;
      int bi = 0;
      for (int j = 0; j < k; ++j) {
        const double4 c = ld_c64(&c64[j]);
        const double d = fma(-2.0, fma(X, c.x, fma(Y, c.y, Z * c.z)), c.w);
        if (d < bd) {
          bd = d;
          bi = j;
        }
      }
      lab[p] = bi;
      ++n_refined;
    }
  }
}

// Flush one per-thread run (label, biased sums, count) into the CTA's shared accumulators.
__device__ __forceinline__ void flush_run(unsigned long long* s_acc, int lab, unsigned int ax,
                                          unsigned int ay, unsigned int az, int an) {
  if (an > 0) {
    const unsigned int m = (unsigned int)an * kMagicBits;
    atomicAdd(&s_acc[lab * 4 + 0], (unsigned long long)(long long)(int)(ax - m));
    atomicAdd(&s_acc[lab * 4 + 1], (unsigned long long)(long long)(int)(ay - m));
    atomicAdd(&s_acc[lab * 4 + 2], (unsigned long long)(long long)(int)(az - m));
    atomicAdd(&s_acc[lab * 4 + 3], (unsigned long long)an);
  }
}

template <typename LabT>
struct LabPack;
template <>
struct LabPack<unsigned char> {
  using V = unsigned int;  // 4 labels
  static __device__ __forceinline__ V load(const unsigned char* p) {
    return *reinterpret_cast<const unsigned int*>(p);
  }
  static __device__ __forceinline__ void store(unsigned char* p, const int (&l)[4]) {
    *reinterpret_cast<unsigned int*>(p) =
        (unsigned)l[0] | ((unsigned)l[1] << 8) | ((unsigned)l[2] << 16) | ((unsigned)l[3] << 24);
  }
  static __device__ __forceinline__ int get(V v, int e) { return (v >> (8 * e)) & 0xff; }
};
template <>
struct LabPack<unsigned short> {
  using V = uint2;
  static __device__ __forceinline__ V load(const unsigned short* p) {
    return *reinterpret_cast<const uint2*>(p);
  }
  static __device__ __forceinline__ void store(unsigned short* p, const int (&l)[4]) {
    uint2 v;
    v.x = (unsigned)l[0] | ((unsigned)l[1] << 16);
    v.y = (unsigned)l[2] | ((unsigned)l[3] << 16);
    *reinterpret_cast<uint2*>(p) = v;
  }
  static __device__ __forceinline__ int get(V v, int e) {
    const unsigned w = (e < 2) ? v.x : v.y;
    return (w >> (16 * (e & 1))) & 0xffff;
  }
};

// ---------------------------------------------------------------------------------------
// K2 + K3: assignment and per-cluster sums in one pass.  G float4 groups per thread.
// Tile = kThreads * 4 * G consecutive points; group g of a tile is kThreads*4 consecutive
// points, thread t owns points [4t, 4t+4) of each group (coalesced 16 B per lane).
// ---------------------------------------------------------------------------------------
template <typename LabT, int G>
__global__ void __launch_bounds__(kThreads, 2) lloyd_step_kernel(const StepParams p) {
  constexpr int P = 4 * G;
  constexpr int TILE = kThreads * P;
  if (!p.ignore_status && (p.st->done | p.st->paused)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_c = reinterpret_cast<float4*>(smem_raw);
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)p.kpad * 16);
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ unsigned int s_changed;
  __shared__ unsigned int s_refined;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    s_changed = 0;
    s_refined = 0;
  }
  for (int i = tid; i < p.kpad * 4; i += kThreads) s_acc[i] = 0ull;
  __syncthreads();
  if (tid == 0) {
    // centroid fast rows: global -> shared through the TMA unit (1-D bulk copy)
    mbar_expect_tx(&s_bar, (uint32_t)p.kpad * 16u);
    tma_load_1d(s_c, p.table, (uint32_t)p.kpad * 16u, &s_bar);
  }
  const double4* c64 = reinterpret_cast<const double4*>(p.table + (size_t)p.kpad * 16);
  const float thresh = p.st->thresh;
  const bool first = p.st->first != 0;
  const FrameF f = p.f;
  mbar_wait(&s_bar, 0);

  // warp-level run accumulator (identical in every lane)
  long long wx = 0, wy = 0, wz = 0;
  unsigned int wn = 0;
  int wlab = -1;
  unsigned int n_chg = 0, n_ref = 0;
  LabT* labels = reinterpret_cast<LabT*>(p.labels);

  const long long n_tiles = (p.n + TILE - 1) / TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * (long long)TILE;
    float xo[P], yo[P], zo[P];
    typename LabPack<LabT>::V oldl[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long i0 = base + (long long)g * (kThreads * 4) + tid * 4;
      const float4 vx = ldg_stream_f4(p.x + i0);
      const float4 vy = ldg_stream_f4(p.y + i0);
      const float4 vz = ldg_stream_f4(p.z + i0);
      xo[4 * g + 0] = vx.x; xo[4 * g + 1] = vx.y; xo[4 * g + 2] = vx.z; xo[4 * g + 3] = vx.w;
      yo[4 * g + 0] = vy.x; yo[4 * g + 1] = vy.y; yo[4 * g + 2] = vy.z; yo[4 * g + 3] = vy.w;
      zo[4 * g + 0] = vz.x; zo[4 * g + 1] = vz.y; zo[4 * g + 2] = vz.z; zo[4 * g + 3] = vz.w;
      oldl[g] = LabPack<LabT>::load(labels + i0);
    }
    float xc[P], yc[P], zc[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      xc[q] = xo[q] - f.ox;
      yc[q] = yo[q] - f.oy;
      zc[q] = zo[q] - f.oz;
    }
    int lab[P];
    assign_points<P>(xc, yc, zc, xo, yo, zo, s_c, c64, p.k, p.kpad, thresh, f, lab, n_ref);

    const bool full_tile = (base + TILE <= p.n);
    // labels out + changed count
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long i0 = base + (long long)g * (kThreads * 4) + tid * 4;
      int l4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        l4[e] = lab[4 * g + e];
        const bool valid = full_tile || (i0 + e < p.n);
        n_chg += (valid && (first || l4[e] != LabPack<LabT>::get(oldl[g], e))) ? 1u : 0u;
      }
      LabPack<LabT>::store(labels + i0, l4);
    }

    // per-thread runs of equal labels -> warp aggregate -> shared accumulators
    int cur = lab[0];
    unsigned int ax = 0, ay = 0, az = 0;
    int an = 0;
    bool single = true;
#pragma unroll
    for (int q = 0; q < P; ++q) {
      const long long i = base + (long long)(q >> 2) * (kThreads * 4) + tid * 4 + (q & 3);
      if (full_tile || i < p.n) {
        if (lab[q] != cur) {
          flush_run(s_acc, cur, ax, ay, az, an);
          single = false;
          cur = lab[q];
          ax = ay = az = 0;
          an = 0;
        }
        // q = rint(x' * scale) by mantissa alignment (|q| < 2^22): bits(x'*s + 1.5*2^23) - bits(1.5*2^23)
        ax += __float_as_uint(fmaf(xc[q], f.sx, kMagic));
        ay += __float_as_uint(fmaf(yc[q], f.sy, kMagic));
        az += __float_as_uint(fmaf(zc[q], f.sz, kMagic));
        ++an;
      }
    }
    const int l0 = __shfl_sync(0xffffffffu, cur, 0);
    const bool uni = __all_sync(0xffffffffu, single && cur == l0 && an == P);
    if (uni) {
      const unsigned int m = (unsigned int)P * kMagicBits;
      const int sx = __reduce_add_sync(0xffffffffu, (int)(ax - m));
      const int sy = __reduce_add_sync(0xffffffffu, (int)(ay - m));
      const int sz = __reduce_add_sync(0xffffffffu, (int)(az - m));
      if (l0 != wlab) {
        if (lane == 0 && wn > 0) {
          atomicAdd(&s_acc[wlab * 4 + 0], (unsigned long long)wx);
          atomicAdd(&s_acc[wlab * 4 + 1], (unsigned long long)wy);
          atomicAdd(&s_acc[wlab * 4 + 2], (unsigned long long)wz);
          atomicAdd(&s_acc[wlab * 4 + 3], (unsigned long long)wn);
        }
        wx = wy = wz = 0;
        wn = 0;
        wlab = l0;
      }
      wx += sx;
      wy += sy;
      wz += sz;
      wn += 32u * P;
    } else {
      flush_run(s_acc, cur, ax, ay, az, an);
    }
  }
  if (lane == 0 && wn > 0) {
    atomicAdd(&s_acc[wlab * 4 + 0], (unsigned long long)wx);
    atomicAdd(&s_acc[wlab * 4 + 1], (unsigned long long)wy);
    atomicAdd(&s_acc[wlab * 4 + 2], (unsigned long long)wz);
    atomicAdd(&s_acc[wlab * 4 + 3], (unsigned long long)wn);
  }
  n_chg = __reduce_add_sync(0xffffffffu, n_chg);
  n_ref = __reduce_add_sync(0xffffffffu, n_ref);
  if (lane == 0) {
    if (n_chg) atomicAdd(&s_changed, n_chg);
    if (n_ref) atomicAdd(&s_refined, n_ref);
  }
  __syncthreads();
  // CTA partials -> global int64 accumulators (RED.ADD.64; order-independent, exact)
  for (int i = tid; i < p.kpad * 4; i += kThreads) {
    const unsigned long long v = s_acc[i];
    if (v) atomicAdd(&p.acc[i], v);
  }
  if (tid == 0) {
    if (s_changed) atomicAdd(&p.acc[p.kpad * 4 + 0], (unsigned long long)s_changed);
    if (s_refined) atomicAdd(&p.st->n_refined, (unsigned long long)s_refined);
  }
}

// ---------------------------------------------------------------------------------------
// K4: centroid update + convergence.  One CTA.
// ---------------------------------------------------------------------------------------
struct UpdateParams {
  unsigned long long* acc;  // [kpad*4 + 8], global sums (already allreduced)
  unsigned char* table;     // centroid table, updated in place
  DevStatus* st;
  Frame fr;
  double mean[3];           // data mean (only used for sklearn's empty-cluster copy quirk)
  int k, kpad;
  int allow_pause;          // 1: pause for relocation when a cluster is empty
  int ignore_status;
};

__device__ __forceinline__ double block_sum_fixed(double v, double* s_red) {
  // fixed-order reduction: shuffle tree inside the warp, warps combined in index order
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
  return t;
}

__device__ __forceinline__ double block_max(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, s_red[i]);
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) lloyd_update_kernel(const UpdateParams u) {
  DevStatus* st = u.st;
  if (!u.ignore_status && (st->done || (st->paused && u.allow_pause))) return;
  __shared__ double s_red[kThreads / 32];
  __shared__ int s_nempty;
  __shared__ unsigned long long s_maxcnt;  // (count << 20) | (kMaxK*... - j): argmax, first wins
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_nempty = 0;
    s_maxcnt = 0ull;
  }
  __syncthreads();
  // empty clusters and the heaviest cluster (np.argmax: first maximum)
  int my_empty = 0;
  unsigned long long my_max = 0ull;
  for (int j = tid; j < u.k; j += kThreads) {
    const unsigned long long cnt = u.acc[j * 4 + 3];
    if (cnt == 0ull) ++my_empty;
    const unsigned long long key = (cnt << 13) | (unsigned long long)(kMaxK * 2 - 1 - j);
    my_max = key > my_max ? key : my_max;
  }
  if (my_empty) atomicAdd(&s_nempty, my_empty);
  atomicMax(&s_maxcnt, my_max);
  __syncthreads();
  const int n_empty = s_nempty;
  if (n_empty > 0 && u.allow_pause) {
    if (tid == 0) {
      st->paused = 1;
      st->n_empty = n_empty;
    }
    return;  // sums are kept; the host sequences the relocation kernels and re-runs update
  }
  const int jmax = kMaxK * 2 - 1 - (int)(s_maxcnt & 0x1fffull);

  float4* fast = reinterpret_cast<float4*>(u.table);
  double4* exact = reinterpret_cast<double4*>(u.table + (size_t)u.kpad * 16);

  double shift2 = 0.0, m_cn = 0.0, m_cx = 0.0, m_cy = 0.0, m_cz = 0.0;
  for (int j = tid; j < u.kpad; j += kThreads) {
    if (j < u.k) {
      const double4 old = exact[j];
      const unsigned long long cnt = u.acc[j * 4 + 3];
      double cx, cy, cz;
      int src = j;
      bool raw = false;
      if (cnt == 0ull) {
        // sklearn/_k_means_common.pyx:289-293: copy of the heaviest cluster's row -- which is
        // still the un-averaged sum when that row comes later in the loop.
        src = jmax;
        raw = jmax > j;
      }
      const double sc = (double)u.acc[src * 4 + 3];
      const double qx = (double)(long long)u.acc[src * 4 + 0] / u.fr.scale[0];
      const double qy = (double)(long long)u.acc[src * 4 + 1] / u.fr.scale[1];
      const double qz = (double)(long long)u.acc[src * 4 + 2] / u.fr.scale[2];
      if (!raw) {
        const double alpha = 1.0 / sc;  // pyx:284-287: centers *= 1/weight
        cx = qx * alpha;
        cy = qy * alpha;
        cz = qz * alpha;
      } else {
        // raw sum in sklearn's mean-centred frame, expressed in ours
        cx = qx + sc * (u.fr.origin[0] - u.mean[0]) + (u.mean[0] - u.fr.origin[0]);
        cy = qy + sc * (u.fr.origin[1] - u.mean[1]) + (u.mean[1] - u.fr.origin[1]);
        cz = qz + sc * (u.fr.origin[2] - u.mean[2]) + (u.mean[2] - u.fr.origin[2]);
      }
      const double dx = cx - old.x, dy = cy - old.y, dz = cz - old.z;
      const double sh = sqrt(dx * dx + dy * dy + dz * dz);  // _center_shift, pyx:298-311
      shift2 += sh * sh;                                     // (center_shift**2).sum()
      const double cn = cx * cx + cy * cy + cz * cz;
      exact[j] = make_double4(cx, cy, cz, cn);
      fast[j] = make_float4((float)(-2.0 * cx), (float)(-2.0 * cy), (float)(-2.0 * cz), (float)cn);
      m_cn = fmax(m_cn, cn);
      m_cx = fmax(m_cx, fabs(cx));
      m_cy = fmax(m_cy, fabs(cy));
      m_cz = fmax(m_cz, fabs(cz));
    } else {
      fast[j] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
      exact[j] = make_double4(0.0, 0.0, 0.0, 1.0 / 0.0);
    }
  }
  shift2 = block_sum_fixed(shift2, s_red);
  m_cn = block_max(m_cn, s_red);
  m_cx = block_max(m_cx, s_red);
  m_cy = block_max(m_cy, s_red);
  m_cz = block_max(m_cz, s_red);
  const unsigned long long n_changed = u.acc[u.kpad * 4 + 0];
  __syncthreads();
  for (int i = tid; i < u.kpad * 4 + 8; i += kThreads) u.acc[i] = 0ull;
  if (tid == 0) {
    // FP32 error bound of the fast distances (DESIGN.md "Exactness"): u = 2^-24
    const double ue = 5.9604644775390625e-08;
    const double E = ue * (4.0 * m_cn + 10.0 * (u.fr.halfrange[0] * m_cx + u.fr.halfrange[1] * m_cy +
                                                 u.fr.halfrange[2] * m_cz));
    st->thresh = __double2float_ru(2.0 * E * 1.001 + 1e-37);
    st->shift2 = shift2;
    st->n_changed = n_changed;
    st->n_empty = n_empty;
    st->paused = 0;
    const int it = st->iter + 1;
    st->iter = it;
    if (!st->first && n_changed == 0ull) {  // _kmeans.py:721-726
      st->strict = 1;
      st->done = 1;
    } else if (shift2 <= st->tol) {         // _kmeans.py:729-738
      st->done = 1;
    }
    if (it >= st->max_iter) st->done = 1;
    st->first = 0;
  }
}

// Builds the centroid table from K x 3 float64 centroids in ORIGINAL coordinates.
struct InitTableParams {
  const double* centers;  // device, k*3
  unsigned char* table;
  DevStatus* st;
  Frame fr;
  int k, kpad;
};

__global__ void __launch_bounds__(kThreads, 1) init_table_kernel(const InitTableParams u) {
  __shared__ double s_red[kThreads / 32];
  float4* fast = reinterpret_cast<float4*>(u.table);
  double4* exact = reinterpret_cast<double4*>(u.table + (size_t)u.kpad * 16);
  double m_cn = 0.0, m_cx = 0.0, m_cy = 0.0, m_cz = 0.0;
  for (int j = threadIdx.x; j < u.kpad; j += kThreads) {
    if (j < u.k) {
      const double cx = u.centers[3 * j + 0] - u.fr.origin[0];
      const double cy = u.centers[3 * j + 1] - u.fr.origin[1];
      const double cz = u.centers[3 * j + 2] - u.fr.origin[2];
      const double cn = cx * cx + cy * cy + cz * cz;
      exact[j] = make_double4(cx, cy, cz, cn);
      fast[j] = make_float4((float)(-2.0 * cx), (float)(-2.0 * cy), (float)(-2.0 * cz), (float)cn);
      m_cn = fmax(m_cn, cn);
      m_cx = fmax(m_cx, fabs(cx));
      m_cy = fmax(m_cy, fabs(cy));
      m_cz = fmax(m_cz, fabs(cz));
    } else {
      fast[j] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
      exact[j] = make_double4(0.0, 0.0, 0.0, 1.0 / 0.0);
    }
  }
  m_cn = block_max(m_cn, s_red);
  m_cx = block_max(m_cx, s_red);
  m_cy = block_max(m_cy, s_red);
  m_cz = block_max(m_cz, s_red);
  if (threadIdx.x == 0) {
    const double ue = 5.9604644775390625e-08;
    const double E = ue * (4.0 * m_cn + 10.0 * (u.fr.halfrange[0] * m_cx + u.fr.halfrange[1] * m_cy +
                                                 u.fr.halfrange[2] * m_cz));
    u.st->thresh = __double2float_ru(2.0 * E * 1.001 + 1e-37);
  }
}

// Reads the table back as K x 3 float64 centroids in original coordinates.
__global__ void read_table_kernel(const unsigned char* table, int k, int kpad, Frame fr,
                                  double* centers_out) {
  const double4* exact = reinterpret_cast<const double4*>(table + (size_t)kpad * 16);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    const double4 c = exact[j];
    centers_out[3 * j + 0] = c.x + fr.origin[0];
    centers_out[3 * j + 1] = c.y + fr.origin[1];
    centers_out[3 * j + 2] = c.z + fr.origin[2];
  }
}

// Converts the int64 accumulators into float64 coordinate sums (original frame) + counts.
__global__ void read_sums_kernel(const unsigned long long* acc, int k, Frame fr, double* sums_out,
                                 long long* counts_out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    const long long cnt = (long long)acc[j * 4 + 3];
    for (int d = 0; d < 3; ++d)
      sums_out[3 * j + d] = (double)(long long)acc[j * 4 + d] / fr.scale[d] + (double)cnt * fr.origin[d];
    counts_out[j] = cnt;
  }
}

// ---------------------------------------------------------------------------------------
// Final pass: labels as int32 (recomputed with the final centroids unless the exit was
// strict) and inertia in FP64 (direct form, fixed-order reduction).
// ---------------------------------------------------------------------------------------
template <typename LabT>
__global__ void __launch_bounds__(kThreads, 2) lloyd_final_kernel(const FinalParams p) {
  constexpr int P = 4;
  constexpr int TILE = kThreads * P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_c = reinterpret_cast<float4*>(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ double s_red[kThreads / 32];
  __shared__ unsigned int s_refined;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    s_refined = 0;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, (uint32_t)p.kpad * 16u);
    tma_load_1d(s_c, p.table, (uint32_t)p.kpad * 16u, &s_bar);
  }
  const double4* c64 = reinterpret_cast<const double4*>(p.table + (size_t)p.kpad * 16);
  const float thresh = p.st->thresh;
  const bool reassign = p.force_assign || !p.st->strict;
  const FrameF f = p.f;
  const LabT* labels = reinterpret_cast<const LabT*>(p.labels);
  mbar_wait(&s_bar, 0);

  double inert = 0.0;
  unsigned int n_ref = 0;
  const long long n_tiles = (p.n + TILE - 1) / TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long i0 = tile * (long long)TILE + tid * 4;
    const float4 vx = ldg_stream_f4(p.x + i0);
    const float4 vy = ldg_stream_f4(p.y + i0);
    const float4 vz = ldg_stream_f4(p.z + i0);
    const float xo[P] = {vx.x, vx.y, vx.z, vx.w};
    const float yo[P] = {vy.x, vy.y, vy.z, vy.w};
    const float zo[P] = {vz.x, vz.y, vz.z, vz.w};
    int lab[P];
    if (reassign) {
      float xc[P], yc[P], zc[P];
#pragma unroll
      for (int q = 0; q < P; ++q) {
        xc[q] = xo[q] - f.ox;
        yc[q] = yo[q] - f.oy;
        zc[q] = zo[q] - f.oz;
      }
      assign_points<P>(xc, yc, zc, xo, yo, zo, s_c, c64, p.k, p.kpad, thresh, f, lab, n_ref);
    } else {
      const typename LabPack<LabT>::V v = LabPack<LabT>::load(labels + i0);
#pragma unroll
      for (int q = 0; q < P; ++q) lab[q] = LabPack<LabT>::get(v, q);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (i0 + q < p.n) {
        const double4 c = ld_c64(&c64[lab[q]]);
        const double dx = ((double)xo[q] - (double)f.ox) - c.x;
        const double dy = ((double)yo[q] - (double)f.oy) - c.y;
        const double dz = ((double)zo[q] - (double)f.oz) - c.z;
        inert += dx * dx + dy * dy + dz * dz;
      }
    }
    if (p.labels_out) {
      if (i0 + 3 < p.n && ((reinterpret_cast<uintptr_t>(p.labels_out) & 15) == 0)) {
        *reinterpret_cast<int4*>(p.labels_out + i0) = make_int4(lab[0], lab[1], lab[2], lab[3]);
      } else {
#pragma unroll
        for (int q = 0; q < P; ++q)
          if (i0 + q < p.n) p.labels_out[i0 + q] = lab[q];
      }
    }
  }
  const double bsum = block_sum_fixed(inert, s_red);
  n_ref = __reduce_add_sync(0xffffffffu, n_ref);
  if ((tid & 31) == 0 && n_ref) atomicAdd(&s_refined, n_ref);
  __syncthreads();
  if (tid == 0) {
    p.partials[blockIdx.x] = bsum;
    if (s_refined) atomicAdd(&p.st->n_refined, (unsigned long long)s_refined);
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    // last CTA: add the per-CTA partials in CTA-index order (deterministic)
    __threadfence();
    double v = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += kThreads) v += p.partials[i];
    const double tot = block_sum_fixed(v, s_red);
    if (tid == 0) {
      p.st->inertia = tot;
      *p.ticket = 0u;
    }
  }
}

}  // namespace mdkm
