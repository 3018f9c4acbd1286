// K1: height-map stack -> XYZ points with NaN / nodata / range masking fused in, stable
// (np.where-order) compaction, optional per-day plane detrend.
//
// Replaces members/rafael/disparity/plugin.py:148 (h = -disp/16), :151-152 (validity),
// :157-160 (np.where + stack), :161-171 (SVD plane fit, via 9 moments + a 3x3 symmetric
// eigen-solve) for every day of the stack, and concatenates the days (absent in the
// reference).
//
// ONE pass over the rasters (4 B read + 12 B written per pixel, the algorithmic minimum): a warp
// takes a tile of kTile = 1024 consecutive pixels (8 rounds of 32 lanes x 4 px, 16 B loads),
// ranks its valid pixels with ballots, parks them in shared memory, obtains the number of
// points of all earlier tiles by a decoupled look-back over per-tile status words (tiles are
// handed out by a ticket counter, so every predecessor of a running tile is itself running or
// done), and writes x, y, z as contiguous runs at offset + rank -- the output order is exactly
// np.where's.  Offsets at every kChunk = 4096 pixels are kept for the day boundaries.  (The
// look-back works on super-tiles of eight warp tiles -- one CTA -- see unproject_fused_kernel.)
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mdkm {

constexpr int kChunk = 4096;
constexpr int kTile = 1024;  // pixels per warp tile of the fused pass
constexpr int kFusedSmem = (kThreads / 32) * kTile * 6;  // dynamic shared memory of unproject_fused_kernel
// per-tile status word of the decoupled look-back: flag in the two top bits, count below
constexpr unsigned long long kStAggregate = 1ull << 62;  // value = valid pixels of this tile
constexpr unsigned long long kStPrefix = 2ull << 62;     // value = valid pixels up to and including this tile
constexpr unsigned long long kStValue = (1ull << 62) - 1ull;

struct UnprojParams {
  const void* hm;         // points at pixel pix_begin
  const uint8_t* mask;    // idem, or nullptr
  long long pix_begin;    // global flat index of the first pixel
  long long pix_count;
  long long HW;
  int W, H;
  int dtype;              // MDKM_HM_F32 / MDKM_HM_I16 / MDKM_HM_F32_GTIFF3
  int vec_ok;             // hm (and mask) aligned for 16 B / 4 B vector loads
  float scale;            // for I16
  float max_abs;
  long long* chunk_offsets;  // [n_chunks]: points produced before every kChunk-th pixel of the range
  float* pts;             // blocked cloud (common.cuh)
  const double* planes;   // [n_days][8]: centre xyz, normal xyz, pad -- or nullptr
  int day0;               // day index of planes[0]
  long long chunk_begin;  // this launch handles chunks [chunk_begin, chunk_end) of the range
  long long chunk_end;
  unsigned int* run_src;  // optional [pix_count / 8 + 1]: entry i = points produced by the pixels before
                          // local pixel 8 i (the raster mirror build reads runs of pixels from it, mirror.cuh)
  // fused pass: this launch handles tiles [tile_begin, tile_end) of the range
  long long tile_begin, tile_end;
  unsigned long long* status;  // [n_tiles / 8] look-back words (one per super-tile), zeroed once per mdkm_unproject
  unsigned int* ticket;        // zeroed before every launch; ticket[1] = fault flag (look-back timed out)
  long long* total_out;        // points of all tiles up to tile_end - 1 (running total of the range)
  unsigned int* minmax;        // optional [6]: ordered-uint min x,y,z / max x,y,z of the points written
                               // (pre-filled with 0xffffffff / 0 by the host, as minmax_kernel expects)
};

// One pixel: height and validity (plugin.py:151-152).  dtype 2 is the reference's own
// "5-out-F.tif" raster (disparity.py:213-224): three pixel-interleaved float32 bands, band 0
// the height (-disp/16), band 2 `final_defined`.
__device__ __forceinline__ bool load_height1(const UnprojParams& p, long long i, float& hv) {
  bool ok = true;
  if (p.dtype == 0) {
    hv = __ldg(reinterpret_cast<const float*>(p.hm) + i);
  } else if (p.dtype == 1) {
    hv = p.scale * (float)__ldg(reinterpret_cast<const short*>(p.hm) + i);
  } else {
    const float* s3 = reinterpret_cast<const float*>(p.hm) + 3 * i;
    hv = __ldg(s3);
    ok = __ldg(s3 + 2) != 0.f;
  }
  ok = ok && (fabsf(hv) <= p.max_abs);
  if (p.mask) ok = ok && (__ldg(p.mask + i) != 0);
  return ok;
}

// Loads 4 consecutive pixel heights starting at local index i (i % 4 == 0); invalid -> NaN.
__device__ __forceinline__ void load_heights4(const UnprojParams& p, long long i, float (&h)[4],
                                              unsigned int& valid_bits) {
  valid_bits = 0;
  const bool full = (i + 3 < p.pix_count);
  if (p.dtype == 0) {
    const float* src = reinterpret_cast<const float*>(p.hm);
    if (full && p.vec_ok) {
      const float4 v = ldg_stream_f4(src + i);
      h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = (i + e < p.pix_count) ? __ldg(src + i + e) : __int_as_float(0x7fc00000);
    }
  } else if (p.dtype == 1) {
    const short* src = reinterpret_cast<const short*>(p.hm);
    if (full && p.vec_ok) {
      const uint2 v = ldg_stream_u64(src + i);
      h[0] = p.scale * (float)(short)(v.x & 0xffff);
      h[1] = p.scale * (float)(short)(v.x >> 16);
      h[2] = p.scale * (float)(short)(v.y & 0xffff);
      h[3] = p.scale * (float)(short)(v.y >> 16);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = (i + e < p.pix_count) ? p.scale * (float)__ldg(src + i + e) : __int_as_float(0x7fc00000);
    }
  } else {
    // 4 pixels = 12 floats (h a d | h a d | h a d | h a d): three aligned 16 B loads
    const float* src = reinterpret_cast<const float*>(p.hm) + 3 * i;
    float d[4];
    if (full && p.vec_ok) {
      const float4 v0 = ldg_stream_f4(src), v1 = ldg_stream_f4(src + 4), v2 = ldg_stream_f4(src + 8);
      h[0] = v0.x; d[0] = v0.z; h[1] = v0.w; d[1] = v1.y; h[2] = v1.z; d[2] = v2.x; h[3] = v2.y; d[3] = v2.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool in = i + e < p.pix_count;
        h[e] = in ? __ldg(src + 3 * e) : __int_as_float(0x7fc00000);
        d[e] = in ? __ldg(src + 3 * e + 2) : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (d[e] == 0.f) h[e] = __int_as_float(0x7fc00000);  // not `final_defined` -> nodata
  }
  unsigned int m4 = 0x01010101u;
  if (p.mask) {
    if (full && p.vec_ok) {
      m4 = ldg_stream_u32(p.mask + i);
    } else {
      m4 = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i + e < p.pix_count && __ldg(p.mask + i + e)) m4 |= (1u << (8 * e));
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    // plugin.py:151-152: isfinite(h) & (|h| <= limit) & validity_mask
    const bool ok = (i + e < p.pix_count) && (fabsf(h[e]) <= p.max_abs) && ((m4 >> (8 * e)) & 0xff);
    valid_bits |= ok ? (1u << e) : 0u;  // NaN and inf fail the <= test
  }
}

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Points of all super-tiles before super-tile t (warp-collective).  Flag and value share one 64-bit
// word, so no fence is involved: a word is either not there yet, an aggregate or a prefix.  Every
// lane requests kLookWin status words at once, so ONE round trip covers the 32 * kLookWin
// predecessors of t -- about as many as there are tickets in flight without a prefix yet; walking
// back one window of 32 per round trip made the look-back the longest phase of the kernel.
constexpr int kLookWin = 8;

__device__ __forceinline__ unsigned long long tile_lookback(unsigned long long* status, long long t, unsigned int cnt,
                                                            int lane, unsigned int* fault) {
  if (lane == 0 && t > 0) st_status(status + t, kStAggregate | cnt);
  unsigned long long excl = 0ull;
  const long long t0 = clock64();
  for (long long base = t - 1;; base -= 32 * kLookWin) {
    unsigned long long acc;
    int state;  // 0: no prefix in these windows, 1: reached a prefix, 2: a word is missing, ask again
    do {
      // bounded wait (about 4 s): every predecessor holds an earlier ticket, so it is running or done
      // and this never triggers; if it ever did, a flagged failure beats a hung GPU
      if (clock64() - t0 > (8ll << 30)) {
        if (lane == 0) atomicExch(fault, 1u);
        return 0ull;
      }
      unsigned long long s[kLookWin];
#pragma unroll
      for (int i = 0; i < kLookWin; ++i) {
        const long long jj = base - (i * 32 + lane);
        s[i] = jj >= 0 ? ld_status(status + jj) : kStPrefix;  // before tile 0: a prefix of zero points
      }
      acc = 0ull;
      state = 0;
#pragma unroll
      for (int i = 0; i < kLookWin; ++i) {
        if (state == 0) {  // warp-uniform
          const unsigned int flag = (unsigned int)(s[i] >> 62);
          const unsigned int pre = __ballot_sync(0xffffffffu, flag == 2u);
          const unsigned int first_pre = pre ? (unsigned int)__ffs(pre) - 1u : 32u;
          // every tile between t and the nearest prefix must have published its aggregate
          const unsigned int need = first_pre < 32u ? ((2u << first_pre) - 1u) : 0xffffffffu;
          if (__ballot_sync(0xffffffffu, flag == 0u) & need) {
            state = 2;
          } else {
            acc += __reduce_add_sync(0xffffffffu, (unsigned int)lane < first_pre ? (unsigned int)(s[i] & kStValue) : 0u);
            if (first_pre < 32u) {
              acc += __shfl_sync(0xffffffffu, s[i], first_pre) & kStValue;
              state = 1;
            }
          }
        }
      }
    } while (state == 2);
    excl += acc;
    if (state == 1) break;
  }
  if (lane == 0) st_status(status + t, kStPrefix | (excl + cnt));
  return excl;
}

// Raster position (and detrended height) of the pixel `pix` of a tile whose first pixel sits at
// (day0, row0, col0): 32-bit arithmetic, no division for rasters at least a tile wide.
struct TilePos {
  unsigned int col0, row0, W, H;
  int day0, plane_day0;
  const double* planes;
};
__device__ __forceinline__ void tile_point(const TilePos& t, unsigned int pix, const float* wz, float& fx, float& fy,
                                           float& zz) {
  unsigned int col = t.col0 + pix, row = t.row0;
  int dcur = t.day0;
  if (t.W >= (unsigned int)kTile) {  // at most one row boundary inside the tile
    if (col >= t.W) { col -= t.W; ++row; }
  } else {
    const unsigned int q = col / t.W;
    col -= q * t.W;
    row += q;
  }
  while (row >= t.H) {  // a tile may run into the next day(s)
    row -= t.H;
    ++dcur;
  }
  zz = wz[pix];
  if (t.planes) {
    // plugin.py:171: height_rel = dot(P - center, normal)
    const double* pl = t.planes + (size_t)(dcur - t.plane_day0) * 8;
    zz = (float)(((double)col - pl[0]) * pl[3] + ((double)row - pl[1]) * pl[4] + ((double)zz - pl[2]) * pl[5]);
  }
  fx = (float)col;
  fy = (float)row;
}

// One CTA takes a super-tile of kThreads / 32 = 8 consecutive warp tiles per ticket: the warps rank
// their tiles independently, ONE look-back per super-tile (by warp 0) yields the CTA's offset, and
// the warps' own offsets follow from the eight counts in shared memory.  The look-back depth is
// bounded by the number of resident CTAs (a few hundred), not by the number of resident warps.
__global__ void __launch_bounds__(kThreads) unproject_fused_kernel(const UnprojParams p) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char s_stage[];  // kFusedSmem bytes: heights [8][1024] f32 (pixel order), valid-pixel lists [8][1024] u16
  __shared__ unsigned int s_cnt[kWarps];
  __shared__ unsigned long long s_base;
  __shared__ long long s_super;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wz = reinterpret_cast<float*>(s_stage) + warp * kTile;
  unsigned short* wix = reinterpret_cast<unsigned short*>(s_stage + kWarps * kTile * 4) + warp * kTile;
  // bounding box of the points this warp writes (the frame of the cloud needs it: no extra pass)
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  const long long super_begin = p.tile_begin / kWarps;  // (launch boundaries are whole super-tiles)
  while (true) {
    if (threadIdx.x == 0) s_super = super_begin + (long long)atomicAdd(p.ticket, 1u);
    __syncthreads();
    const long long sup = s_super;
    if (sup * kWarps >= p.tile_end) break;  // CTA-uniform
    const long long t = sup * kWarps + warp;
    const bool active = t < p.tile_end;     // warp-uniform
    const long long pix0 = t * kTile;  // local index of the tile's first pixel
    unsigned int cnt = 0;
    unsigned int head[kTile / 128];  // rank of this lane's first pixel of round r
    if (active) {
    // 1. the tile's pixels: 8 x 16 B per lane in flight; the heights are parked in shared memory
    // in PIXEL order at once (the registers are free again), only the validity bits stay
    unsigned int vbits = 0;  // 4 bits per round
#pragma unroll
    for (int r = 0; r < kTile / 128; ++r) {
      float hv[4];
      unsigned int vb;
      load_heights4(p, pix0 + r * 128 + lane * 4, hv, vb);
      *reinterpret_cast<float4*>(wz + r * 128 + lane * 4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      vbits |= vb << (4 * r);
    }
    // 2. rank of every valid pixel inside the tile, in pixel order: the list of valid pixels
#pragma unroll
    for (int r = 0; r < kTile / 128; ++r) {
      const unsigned int lt = (1u << lane) - 1u;
      const unsigned int vb = (vbits >> (4 * r)) & 0xfu;
      unsigned int before = 0, total = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned int m = __ballot_sync(0xffffffffu, (vb >> e) & 1u);
        before += __popc(m & lt);
        total += __popc(m);
      }
      head[r] = cnt + before;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((vb >> e) & 1u) wix[head[r] + __popc(vb & ((1u << e) - 1u))] = (unsigned short)(r * 128 + lane * 4 + e);
      cnt += total;
    }
    }  // active
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    // 4. points of all earlier super-tiles (one look-back per CTA), then of the earlier warps
    if (warp == 0) {
      const unsigned int agg = __reduce_add_sync(0xffffffffu, lane < kWarps ? s_cnt[lane] : 0u);
      const unsigned long long base = tile_lookback(p.status, sup, agg, lane, p.ticket + 1);
      if (lane == 0) s_base = base;
    }
    __syncthreads();
    if (!active) continue;
    unsigned long long excl = s_base;
    for (int w = 0; w < warp; ++w) excl += s_cnt[w];
    if (lane == 0) {
      if ((t & (kChunk / kTile - 1)) == 0) p.chunk_offsets[t / (kChunk / kTile)] = (long long)excl;
      if (t == p.tile_end - 1 && p.total_out) *p.total_out = (long long)(excl + cnt);
    }
    if (p.run_src && (lane & 1) == 0) {
#pragma unroll
      for (int r = 0; r < kTile / 128; ++r) {
        const long long o = pix0 + r * 128 + lane * 4;  // a multiple of 8
        if (o < p.pix_count) p.run_src[o >> 3] = (unsigned int)(excl + head[r]);
      }
    }
    // 5. contiguous runs of x, y, z.  Position of the tile's first pixel once per tile (64-bit);
    // per point only 32-bit arithmetic, and no division at all for rasters at least a tile wide
    const long long gp0 = p.pix_begin + pix0;
    const long long day0 = gp0 / p.HW;
    const long long rem0 = gp0 - day0 * p.HW;
    const unsigned int row0 = (unsigned int)(rem0 / p.W);
    const unsigned int col0 = (unsigned int)(rem0 - (long long)row0 * p.W);
    const unsigned int W = (unsigned int)p.W, H = (unsigned int)p.H;
    const TilePos tp{col0, row0, W, H, (int)day0, p.day0, p.planes};
    // outputs [excl, excl + cnt): a few single points up to the next multiple of four, then four
    // consecutive points per lane and round as three 16-byte stores (a block of the cloud holds
    // 128 points, so an aligned quad never straddles two blocks), then the rest
    const unsigned int lead = min(cnt, (4u - (unsigned int)(excl & 3ull)) & 3u);
    const unsigned int quads = (cnt - lead) >> 2;
    for (unsigned int j = lane; j < quads; j += 32) {
      const unsigned int i = lead + 4u * j;
      float4 vx, vy, vz;
      tile_point(tp, wix[i + 0], wz, vx.x, vy.x, vz.x);
      tile_point(tp, wix[i + 1], wz, vx.y, vy.y, vz.y);
      tile_point(tp, wix[i + 2], wz, vx.z, vy.z, vz.z);
      tile_point(tp, wix[i + 3], wz, vx.w, vy.w, vz.w);
      float* dst = p.pts + pt_off((long long)excl + i);
      *reinterpret_cast<float4*>(dst) = vx;
      *reinterpret_cast<float4*>(dst + kGroup) = vy;
      *reinterpret_cast<float4*>(dst + 2 * kGroup) = vz;
      mn[0] = fminf(fminf(mn[0], vx.x), fminf(fminf(vx.y, vx.z), vx.w)); mx[0] = fmaxf(fmaxf(mx[0], vx.x), fmaxf(fmaxf(vx.y, vx.z), vx.w));
      mn[1] = fminf(fminf(mn[1], vy.x), fminf(fminf(vy.y, vy.z), vy.w)); mx[1] = fmaxf(fmaxf(mx[1], vy.x), fmaxf(fmaxf(vy.y, vy.z), vy.w));
      mn[2] = fminf(fminf(mn[2], vz.x), fminf(fminf(vz.y, vz.z), vz.w)); mx[2] = fmaxf(fmaxf(mx[2], vz.x), fmaxf(fmaxf(vz.y, vz.z), vz.w));
    }
    {  // the (at most 3 + 3) points before and behind the quads: one lane each
      const unsigned int tail0 = lead + 4u * quads;
      const unsigned int n_single = lead + (cnt - tail0);
      if ((unsigned int)lane < n_single) {
        const unsigned int i = (unsigned int)lane < lead ? (unsigned int)lane : tail0 + ((unsigned int)lane - lead);
        float fx, fy, zz;
        tile_point(tp, wix[i], wz, fx, fy, zz);
        float* dst = p.pts + pt_off((long long)excl + i);
        dst[0] = fx;
        dst[kGroup] = fy;
        dst[2 * kGroup] = zz;
        mn[0] = fminf(mn[0], fx); mx[0] = fmaxf(mx[0], fx);
        mn[1] = fminf(mn[1], fy); mx[1] = fmaxf(mx[1], fy);
        mn[2] = fminf(mn[2], zz); mx[2] = fmaxf(mx[2], zz);
      }
    }
  }
  if (p.minmax) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const unsigned int a = __reduce_min_sync(0xffffffffu, f2ord(mn[d]));
      const unsigned int b = __reduce_max_sync(0xffffffffu, f2ord(mx[d]));
      if (lane == 0) {
        atomicMin(&p.minmax[d], a);
        atomicMax(&p.minmax[3 + d], b);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Per-day plane fit (plugin.py:161-171) from 9 moments.  Grid (kPlaneBlocks, n_days): CTA b of
// day d reduces a fixed slice of the day's pixels in a fixed order -> deterministic partials;
// plane_solve_kernel adds them in order and solves the 3x3 symmetric eigenproblem (Jacobi).
// ---------------------------------------------------------------------------------------
constexpr int kPlaneBlocks = 64;

__global__ void __launch_bounds__(kThreads) plane_moments_kernel(const UnprojParams p, int n_days,
                                                                 double* partials /*[d][b][10]*/) {
  __shared__ double s_red[kThreads / 32];
  const int d = blockIdx.y;
  const long long day_begin = (long long)(p.day0 + d) * p.HW - p.pix_begin;  // local index
  const long long quads = (p.HW + 3) / 4;
  const double px = 0.5 * (double)(p.W - 1), py = 0.5 * (double)(p.H - 1);
  double m[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = 0.0;
  for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < quads;
       q += (long long)gridDim.x * kThreads) {
    const long long li = q * 4;  // index inside the day
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long l = li + e;
      if (l >= p.HW) break;
      const long long i = day_begin + l;
      float hv;
      if (load_height1(p, i, hv)) {
        const int row = (int)(l / p.W);
        const int col = (int)(l - (long long)row * p.W);
        const double X = (double)col - px, Y = (double)row - py, Z = (double)hv;
        m[0] += 1.0; m[1] += X; m[2] += Y; m[3] += Z;
        m[4] += X * X; m[5] += X * Y; m[6] += X * Z; m[7] += Y * Y; m[8] += Y * Z; m[9] += Z * Z;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double v = m[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
      partials[((size_t)d * gridDim.x + blockIdx.x) * 10 + i] = t;
    }
  }
}

__global__ void plane_solve_kernel(const double* partials, int n_blocks, int W, int H,
                                   double* planes /*[d][8]*/) {
  const int d = blockIdx.x;
  if (threadIdx.x != 0) return;
  double m[10];
  for (int i = 0; i < 10; ++i) {
    double t = 0.0;
    for (int b = 0; b < n_blocks; ++b) t += partials[((size_t)d * n_blocks + b) * 10 + i];
    m[i] = t;
  }
  double* pl = planes + (size_t)d * 8;
  const double px = 0.5 * (double)(W - 1), py = 0.5 * (double)(H - 1);
  const double n = m[0];
  if (n < 3.0) {  // oracle skips the fit for < 3 points: identity plane z_rel = z
    pl[0] = px; pl[1] = py; pl[2] = 0.0; pl[3] = 0.0; pl[4] = 0.0; pl[5] = 1.0; pl[6] = n; pl[7] = 0.0;
    return;
  }
  const double cx = m[1] / n, cy = m[2] / n, cz = m[3] / n;
  // scatter matrix of the centred points
  double A[3][3];
  A[0][0] = m[4] - n * cx * cx; A[0][1] = m[5] - n * cx * cy; A[0][2] = m[6] - n * cx * cz;
  A[1][1] = m[7] - n * cy * cy; A[1][2] = m[8] - n * cy * cz; A[2][2] = m[9] - n * cz * cz;
  A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int pi = 0; pi < 2; ++pi)
      for (int qi = pi + 1; qi < 3; ++qi) {
        if (fabs(A[pi][qi]) < 1e-300) continue;
        const double theta = (A[qi][qi] - A[pi][pi]) / (2.0 * A[pi][qi]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][pi], akq = A[k][qi];
          A[k][pi] = c * akp - s * akq;
          A[k][qi] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[pi][k], aqk = A[qi][k];
          A[pi][k] = c * apk - s * aqk;
          A[qi][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][pi], vkq = V[k][qi];
          V[k][pi] = c * vkp - s * vkq;
          V[k][qi] = s * vkp + c * vkq;
        }
      }
  }
  int mi = 0;
  if (A[1][1] < A[mi][mi]) mi = 1;
  if (A[2][2] < A[mi][mi]) mi = 2;
  double nx = V[0][mi], ny = V[1][mi], nz = V[2][mi];
  const double nn = sqrt(nx * nx + ny * ny + nz * nz);
  nx /= nn; ny /= nn; nz /= nn;
  if (nz < 0) { nx = -nx; ny = -ny; nz = -nz; }  // plugin.py:167-168
  pl[0] = cx + px; pl[1] = cy + py; pl[2] = cz; pl[3] = nx; pl[4] = ny; pl[5] = nz; pl[6] = n; pl[7] = 0.0;
}

// ---------------------------------------------------------------------------------------
// Cloud statistics: per-dimension min / max (frame + error bound) and first/second moments
// about the frame origin (mean, and var(X) for sklearn's tolerance, _kmeans.py:285-293).
// ---------------------------------------------------------------------------------------
__host__ __device__ inline float ord2f(unsigned int u) {
  const unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// out[0..2] = ordered min x,y,z ; out[3..5] = ordered max.  Caller pre-fills min with
// 0xffffffff and max with 0.  One warp per 128-point block of the blocked cloud.
__global__ void __launch_bounds__(kThreads) minmax_kernel(const float* pts, long long n, unsigned int* out) {
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  const int lane = threadIdx.x & 31;
  const long long n_groups = (n + kGroup - 1) / kGroup;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long g = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); g < n_groups; g += stride) {
    const float* blk = pts + g * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup), vz = ldg_stream_f4(blk + 2 * kGroup);
    const float ax[4] = {vx.x, vx.y, vx.z, vx.w}, ay[4] = {vy.x, vy.y, vy.z, vy.w}, az[4] = {vz.x, vz.y, vz.z, vz.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (g * kGroup + lane * 4 + e < n) {
        mn[0] = fminf(mn[0], ax[e]); mx[0] = fmaxf(mx[0], ax[e]);
        mn[1] = fminf(mn[1], ay[e]); mx[1] = fmaxf(mx[1], ay[e]);
        mn[2] = fminf(mn[2], az[e]); mx[2] = fmaxf(mx[2], az[e]);
      }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const unsigned int a = __reduce_min_sync(0xffffffffu, f2ord(mn[d]));
    const unsigned int b = __reduce_max_sync(0xffffffffu, f2ord(mx[d]));
    if (lane == 0) {
      atomicMin(&out[d], a);
      atomicMax(&out[3 + d], b);
    }
  }
}

// partials[b][6] = sum (x-o), sum (x-o)^2 per dim, in FP64, fixed order inside the CTA;
// the last CTA adds the partials in CTA order into out[6].
__global__ void __launch_bounds__(kThreads) moments_kernel(const float* pts, long long n, FrameF f, double* partials,
                                                           unsigned int* ticket, double* out) {
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  double m[6] = {0, 0, 0, 0, 0, 0};
  const int lane = threadIdx.x & 31;
  const long long n_groups = (n + kGroup - 1) / kGroup;
  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long g = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); g < n_groups; g += stride) {
    const float* blk = pts + g * kBlockFloats + lane * 4;
    const float4 vx = ldg_stream_f4(blk), vy = ldg_stream_f4(blk + kGroup), vz = ldg_stream_f4(blk + 2 * kGroup);
    const float ax[4] = {vx.x, vx.y, vx.z, vx.w}, ay[4] = {vy.x, vy.y, vy.z, vy.w}, az[4] = {vz.x, vz.y, vz.z, vz.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (g * kGroup + lane * 4 + e < n) {
        const double X = (double)ax[e] - (double)f.ox, Y = (double)ay[e] - (double)f.oy,
                     Z = (double)az[e] - (double)f.oz;
        m[0] += X; m[1] += Y; m[2] += Z;
        m[3] += X * X; m[4] += Y * Y; m[5] += Z * Z;
      }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = m[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
      partials[(size_t)blockIdx.x * 6 + i] = t;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 6) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += partials[(size_t)b * 6 + threadIdx.x];
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// [n][3] (x,y,z) or SoA x[n] y[n] z[n]  ->  blocked cloud
__global__ void __launch_bounds__(kThreads) to_blocked_kernel(const float* src, long long n, int soa, float* pts) {
  const long long total = n * 3;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
    long long i;
    int c;
    if (soa) {
      c = (int)(t / n);
      i = t - (long long)c * n;
    } else {
      i = t / 3;
      c = (int)(t - i * 3);
    }
    pts[pt_off(i) + c * kGroup] = src[t];  // coalesced reads
  }
}

// out[i] = (x, y, z) of point idx[i]  (X[seeds] of init="random", sklearn/_kmeans.py:1014-1021)
// idx are global indices; a rank owns [rank_offset, rank_offset + n) and writes zeros elsewhere
__global__ void gather_points_kernel(const float* pts, const long long* idx, int m, long long rank_offset,
                                     long long n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long l = idx[i] - rank_offset;
  const bool mine = l >= 0 && l < n;
  const float* q = pts + pt_off(mine ? l : 0);
  out[3 * i + 0] = mine ? q[0] : 0.f;
  out[3 * i + 1] = mine ? q[kGroup] : 0.f;
  out[3 * i + 2] = mine ? q[2 * kGroup] : 0.f;
}

// zero the unused tail of the last block (and nothing else)
__global__ void zero_tail_kernel(float* pts, long long n) {
  const long long cap = (n + kGroup - 1) / kGroup * kGroup;
  for (long long i = n + threadIdx.x; i < cap; i += blockDim.x) {
    float* d = pts + pt_off(i);
    d[0] = 0.f;
    d[kGroup] = 0.f;
    d[2 * kGroup] = 0.f;
  }
}

// blocked cloud -> [n][3] in (x,y,z) or napari (z,y,x) order; coalesced stores
__global__ void __launch_bounds__(kThreads) blocked_to_aos_kernel(const float* pts, long long n, int napari,
                                                                  float* out) {
  const long long total = n * 3;
  for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
    const long long i = t / 3;
    const int c = (int)(t - i * 3);
    const int src = napari ? 2 - c : c;
    out[t] = pts[pt_off(i) + src * kGroup];
  }
}

// same for the points [range[0], range[1]) only (one slab of a streamed unprojection)
__global__ void __launch_bounds__(kThreads) blocked_to_aos_range_kernel(const float* pts, const long long* range,
                                                                        int napari, float* out) {
  const long long t0 = range[0] * 3, t1 = range[1] * 3;
  for (long long t = t0 + (long long)blockIdx.x * kThreads + threadIdx.x; t < t1; t += (long long)gridDim.x * kThreads) {
    const long long i = t / 3;
    const int c = (int)(t - i * 3);
    const int src = napari ? 2 - c : c;
    out[t] = pts[pt_off(i) + src * kGroup];
  }
}

}  // namespace mdkm
