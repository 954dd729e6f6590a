// rcvvote.cu -- CUDA kernels (sm_100a) and the C ABI of librcvvote.so.
//
// Pipeline for one call (all on the caller's stream, no host round trip):
//   K1  k_frame_mask / k_frame_emit / k_points_from_pixels : mask rule + stable stream compaction + depth back-projection
//                                         (AccumulatorSpace.py:603-619, 77-85)      [frames API]
//       k_points_to_voxel               : (N,3) metres -> voxel units               [points API]
//   P   k_prelude   : numpy-pairwise means, recentre, zero boundary, grid side D, tile work list
//                                         (AccumulatorSpace.py:373-401)
//   K2  k_vote      : persistent CTAs, each owning a shared-memory int32 tile of the accumulator;
//                     sphere shells are scattered with shared-memory atomics (raster_core.h) and the
//                     tile's peak is reduced in place, so the volume never touches HBM
//                                         (AccumulatorSpace.py:325-341 + :406)
//   K3  k_argmax_volume : warp-shuffle + grid-level argmax over an HBM-resident int32 volume (:406)
//   F   k_finalize  : un-shift the peak to millimetres (AccumulatorSpace.py:409-415)
//   K4  k_horn      : batched Horn absolute orientation (util/horn.py:75-181)
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/rcvvote.h"
#include "raster_core.h"
#include "runs_core.h"
#include "horn_core.h"

#define RCV_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

using namespace rcv;

#ifndef RCV_RING2_INTERIOR
#define RCV_RING2_INTERIOR 0   // 1: columns well inside a ring skip the ownership test (one more copy of the column body)
#endif
#ifndef RCV_RING2_SYNCWARP
#define RCV_RING2_SYNCWARP 1
#endif
#ifndef RCV_VOTE_THREADS
#define RCV_VOTE_THREADS 512
#endif
constexpr int kVoteThreads = RCV_VOTE_THREADS;
constexpr int kVoteWarps = kVoteThreads / 32;
constexpr int kSmemBytes = 232448;                        // 227 KB: the sm_100 per-CTA maximum
constexpr int kDummyWords = 32 * kVoteWarps;              // one private sink word per lane per warp
constexpr int kTileWords = kSmemBytes / 4 - kDummyWords - 128;   // 128 words left for static shared variables
constexpr int kStVolumeSkipped = 32;

struct ItemMeta {
  long long off;  // first point of the item in the pool
  int n;          // points
  int D, Dp, zb;  // grid side, padded row stride, zero boundary
  int ni, nj;     // tile shape (slices x rows); every tile spans all k
  int status;
  int guard;      // (glo & 0xffff) | ghi << 16, SIGNED 16-bit each.  Gen 2: a tile row holds the cells [-glo, D - 1 + ghi] (+ one spare): guard
                  // cells where the spheres overhang the grid along C, and NEGATIVE values where they stay inside it (cells no sphere
                  // reaches are not stored).  Gen 1: unsigned guards along both in-slice axes of whole-slice tiles, none for row bands.
  int clip;       // gen 2: 1 if run boundaries can leave the tile along C (spheres overhang a grid whose guard band was refused)
  int band;       // 1: row-band tiles (one slice does not fit), 0: whole-slice tiles with guard rows
  int rw0, rwn;   // gen 2: the rows (B = reference x) any sphere of the item can reach: tiles cover [rw0, rw0 + rwn) only
  int sw0, swn;   // gen 2: likewise the slices (A = reference y); the cells (C = z) window is the signed guard pair below
  double mean[3];
  double rmax;
};

struct Unit {
  int item, i0, ni, j0, nj;
};

struct Pool {
  double *X, *Y, *Z, *Rd;  // voxel-unit coordinates (recentred in place by the prelude), radius in voxels
  int* Ri;                 // np.around(radius) -- AccumulatorSpace.py:332
  int* perm;               // per item: its points ordered by (y voxel, R) so that the 32 points of a warp of k_vote draw alike
  int4* rec;               // per point IN VOTE ORDER: the float32 record of the run rasteriser (RunPoint, 2 x int4)
  int* grp;                // per group of 32 points in vote order: the y-slices its spheres reach, lo | hi << 16 (item b starts at off / 32 + b)
  long long cap;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) { unsigned long long o = shfl_xor_u64(v, m); v = o > v ? o : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) v += shfl_xor_u64(v, m);
  return v;
}
__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) { double o = __shfl_xor_sync(0xffffffffu, v, m); v = o < v ? o : v; }
  return v;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) { double o = __shfl_xor_sync(0xffffffffu, v, m); v = o > v ? o : v; }
  return v;
}

// (count, first C-order index) packed so that unsigned max = highest count, then smallest index:
// exactly argwhere(V == V.max())[0] (AccumulatorSpace.py:406).
__device__ __forceinline__ unsigned long long pack_peak(int count, unsigned lin) {
  return ((unsigned long long)(unsigned)count << 32) | (unsigned long long)(0xffffffffu - lin);
}

// ------------------------------------------------------------------------------------------------
// points API front end: xyz (metres) -> voxel units; radius -> voxel units in its own dtype
// (AccumulatorSpace.py:376, :388, :332)
// ------------------------------------------------------------------------------------------------
__global__ void k_items_from_offsets(const long long* __restrict__ offs, int n_items, long long cap, ItemMeta* meta) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_items) return;
  long long o = offs[b] - offs[0], n = offs[b + 1] - offs[b];
  ItemMeta m;
  memset(&m, 0, sizeof(m));
  m.off = o;
  m.n = (int)n;
  if (n < 0 || o + n > cap || n > 0x7fffffffLL) { m.status = RCV_ST_POINT_OVERFLOW; m.n = 0; }
  meta[b] = m;
}

__device__ __forceinline__ void radius_to_voxel_f32(float r, float scale, float unit, double& rd, int& ri) {
  const float rv = __fdiv_rn(__fmul_rn(r, scale), unit);
  rd = (double)rv;
  ri = __float2int_rn(rv);  // half-to-even, like np.around
}
__device__ __forceinline__ void radius_to_voxel_f64(double r, double scale, double unit, double& rd, int& ri) {
  const double rv = __ddiv_rn(__dmul_rn(r, scale), unit);
  rd = rv;
  ri = __double2int_rn(rv);
}

__global__ void k_points_to_voxel(const double* __restrict__ xyz, const void* __restrict__ radii, int radius_dtype,
                                  const long long* __restrict__ offs, int n_items, double acc_unit, double radius_scale,
                                  Pool pool) {
  const long long base = offs[0];
  long long total = offs[n_items] - base;
  if (total > pool.cap) total = pool.cap;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const double* p = xyz + 3 * (base + q);
    pool.X[q] = __ddiv_rn(__dmul_rn(p[0], 1000.0), acc_unit);
    pool.Y[q] = __ddiv_rn(__dmul_rn(p[1], 1000.0), acc_unit);
    pool.Z[q] = __ddiv_rn(__dmul_rn(p[2], 1000.0), acc_unit);
    double rd; int ri;
    if (radius_dtype == RCV_F32) radius_to_voxel_f32(((const float*)radii)[base + q], (float)radius_scale, (float)acc_unit, rd, ri);
    else radius_to_voxel_f64(((const double*)radii)[base + q], radius_scale, acc_unit, rd, ri);
    pool.Rd[q] = rd;
    pool.Ri[q] = ri;
  }
}

// ------------------------------------------------------------------------------------------------
// K1 -- frames API front end: mask rule + back-projection + stable compaction, one CTA per item.
// ------------------------------------------------------------------------------------------------
struct FrameArgs {
  const void* depth; const float* radius; const float* sem; const double* K; const double* max_radii;
  rcv_frame_params fp;
  double acc_unit, radius_scale;
  int n_kpts;
  int vec_ok;  // 128-bit loads allowed: H*W % 8 == 0 and all map pointers 16-byte aligned
};

constexpr int kCompactThreads = 512;
constexpr int kPxPerThread = 8;
constexpr int kPxPerIter = kCompactThreads * kPxPerThread;

__device__ __forceinline__ double load_depth(const void* depth, int dtype, long long idx) {
  if (dtype == RCV_U16) return (double)((const unsigned short*)depth)[idx];
  if (dtype == RCV_F32) return (double)((const float*)depth)[idx];
  return ((const double*)depth)[idx];
}

// Loads 8 consecutive pixels' validity + raw values. Vectorised (128-bit) when the group is whole.
__device__ __forceinline__ unsigned pixel_group_mask(const FrameArgs& a, long long frame_px0, long long item_px0, int px, int npx,
                                                     bool vec_ok, double max_r, double* zraw, float* rad) {
  unsigned mask = 0;
  const int flags = a.fp.mask_flags;
  if (vec_ok && px + kPxPerThread <= npx) {
    float s[kPxPerThread];
    // depth first: every rule needs depth != 0, and most groups of 8 pixels of a masked depth image are empty -- their radius
    // (and seg) values are never requested (12 of the 14 bytes per pixel and keypoint)
    if (a.fp.depth_dtype == RCV_U16) {
      const uint4 d = __ldg(reinterpret_cast<const uint4*>((const unsigned short*)a.depth + frame_px0 + px));
      if ((d.x | d.y | d.z | d.w) == 0u) return 0u;
      const unsigned w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { zraw[2 * q] = (double)(w[q] & 0xffffu); zraw[2 * q + 1] = (double)(w[q] >> 16); }
    } else {
      bool any = false;
#pragma unroll
      for (int q = 0; q < kPxPerThread; ++q) { zraw[q] = load_depth(a.depth, a.fp.depth_dtype, frame_px0 + px + q); any = any || zraw[q] != 0.0; }
      if (!any) return 0u;
    }
    {
      const float4* rp = reinterpret_cast<const float4*>(a.radius + item_px0 + px);
      float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      rad[0] = r0.x; rad[1] = r0.y; rad[2] = r0.z; rad[3] = r0.w; rad[4] = r1.x; rad[5] = r1.y; rad[6] = r1.z; rad[7] = r1.w;
    }
    if (a.sem) {
      const float4* sp = reinterpret_cast<const float4*>(a.sem + item_px0 + px);
      float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
      s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    }
#pragma unroll
    for (int q = 0; q < kPxPerThread; ++q) {
      bool ok = zraw[q] != 0.0;
      if (flags & RCV_MASK_MAX_RADIUS) ok = ok && ((double)rad[q] <= max_r);
      if (flags & RCV_MASK_RADIUS_NONZERO) ok = ok && (rad[q] != 0.f);
      if (flags & RCV_MASK_RADIUS_POSITIVE) ok = ok && (rad[q] > 0.f);
      if (flags & RCV_MASK_SEM_GT) ok = ok && (s[q] > a.fp.sem_threshold);
      if (flags & RCV_MASK_SEM_GE) ok = ok && (s[q] >= a.fp.sem_threshold);
      mask |= (unsigned)ok << q;
    }
  } else {
    for (int q = 0; q < kPxPerThread; ++q) {
      if (px + q >= npx) break;
      zraw[q] = load_depth(a.depth, a.fp.depth_dtype, frame_px0 + px + q);
      rad[q] = a.radius[item_px0 + px + q];
      const float sv = a.sem ? a.sem[item_px0 + px + q] : 0.f;
      bool ok = zraw[q] != 0.0;
      if (flags & RCV_MASK_MAX_RADIUS) ok = ok && ((double)rad[q] <= max_r);
      if (flags & RCV_MASK_RADIUS_NONZERO) ok = ok && (rad[q] != 0.f);
      if (flags & RCV_MASK_RADIUS_POSITIVE) ok = ok && (rad[q] > 0.f);
      if (flags & RCV_MASK_SEM_GT) ok = ok && (sv > a.fp.sem_threshold);
      if (flags & RCV_MASK_SEM_GE) ok = ok && (sv >= a.fp.sem_threshold);
      mask |= (unsigned)ok << q;
    }
  }
  return mask;
}

// ---- K1, streaming form: mask bits -> scan -> pixel indices -> dense conversion ----------------------------------
// The first version (one kernel run twice: count, then re-read + convert inside the divergent compaction loop) read every
// map twice and its write pass ran at 0.95 TB/s.  Here the maps are streamed ONCE (k_frame_mask: 128-bit loads, one survival bit per
// pixel, no block-wide scan), the row-major compaction works on the bit masks only (k_frame_emit: 1/48 of the bytes)
// and the float64 back-projection runs with one thread per surviving pixel, all lanes busy (k_points_from_pixels).
__global__ void __launch_bounds__(kCompactThreads) k_frame_mask(FrameArgs a, int* __restrict__ cnt, unsigned* __restrict__ bits, int words_per_item) {
  const int item = blockIdx.x;
  const int frame = item / a.n_kpts, kp = item % a.n_kpts;
  const int npx = a.fp.height * a.fp.width;
  const long long frame_px0 = (long long)frame * npx, item_px0 = (long long)item * npx;
  const double max_r = a.max_radii ? a.max_radii[(long long)frame * a.fp.max_radii_stride + kp] : 0.0;
  const bool vec_ok = a.vec_ok != 0;
  unsigned* out = bits + (long long)item * words_per_item;
  const int lane = threadIdx.x & 31;
  int n = 0;
  // a warp covers 256 consecutive pixels per step (lane = one group of 8): lanes 4k..4k+3 make up one 32-bit word
  for (int px = threadIdx.x * kPxPerThread; px < words_per_item * 32; px += kPxPerIter) {
    double zraw[kPxPerThread];
    float rad[kPxPerThread];
    unsigned m = 0;
    if (px < npx) m = pixel_group_mask(a, frame_px0, item_px0, px, npx, vec_ok, max_r, zraw, rad);
    n += __popc(m);
    unsigned w = m << ((lane & 3) * 8);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    if ((lane & 3) == 0) out[px >> 5] = w;
  }
  __shared__ int s_n[kCompactThreads / 32];
  n = __reduce_add_sync(0xffffffffu, n);
  if (lane == 0) s_n[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) t += s_n[w];
    cnt[item] = t;
  }
}

// Survivors' pixel indices in row-major order: pix[off .. off + n) for every item (pix aliases the pool's perm array).
__global__ void __launch_bounds__(kCompactThreads) k_frame_emit(const unsigned* __restrict__ bits, int words_per_item, const ItemMeta* __restrict__ meta,
                                                               int* __restrict__ pix, long long item_stride_words) {
  const int item = blockIdx.x;
  const ItemMeta m = meta[item];
  if (m.n == 0) return;   // empty or overflowed item
  const unsigned* in = bits + (long long)item * item_stride_words;
  int* out = pix + m.off;
  __shared__ int s_warp[kCompactThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int w0 = 0; w0 < words_per_item; w0 += kCompactThreads) {
    const int wi = w0 + threadIdx.x;
    unsigned w = wi < words_per_item ? in[wi] : 0u;
    const int c = __popc(w);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int q = 0; q < kCompactThreads / 32; ++q) { const int t = s_warp[q]; if (q < warp) wbase += t; total += t; }
    int o = base + wbase + incl - c;
    while (w) {
      const int b = __ffs((int)w) - 1;
      w &= w - 1u;
      out[o++] = wi * 32 + b;
    }
    base += total;
    __syncthreads();
  }
}

// rgbd_to_point_cloud (AccumulatorSpace.py:77-85), xyz_mm/1000 (:619), *1000/acc_unit (:376) and the radius in voxels
// (:388, :332) for the surviving pixels, one thread per point; reads pix[] (= pool.perm) and writes the pool.
__global__ void __launch_bounds__(256) k_points_from_pixels(FrameArgs a, const ItemMeta* __restrict__ meta, Pool pool) {
  const int item = blockIdx.x;
  const ItemMeta m = meta[item];
  if (m.n == 0) return;
  const int frame = item / a.n_kpts;
  const int W = a.fp.width, npx = a.fp.height * a.fp.width;
  const long long frame_px0 = (long long)frame * npx, item_px0 = (long long)item * npx;
  const double* Kp = a.K + (long long)frame * a.fp.k_stride;
  const double fx = Kp[0], cx = Kp[2], fy = Kp[4], cy = Kp[5];
  for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < m.n; q += gridDim.y * blockDim.x) {
    const long long o = m.off + q;
    const int p = pool.perm[o];
    const int u = p % W, v = p / W;
    const double z = __ddiv_rn(load_depth(a.depth, a.fp.depth_dtype, frame_px0 + p), a.fp.depth_div);
    const double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, cx), z), fx);
    const double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, cy), z), fy);
    pool.X[o] = __ddiv_rn(__dmul_rn(__ddiv_rn(x, a.fp.xyz_div), 1000.0), a.acc_unit);
    pool.Y[o] = __ddiv_rn(__dmul_rn(__ddiv_rn(y, a.fp.xyz_div), 1000.0), a.acc_unit);
    pool.Z[o] = __ddiv_rn(__dmul_rn(__ddiv_rn(z, a.fp.xyz_div), 1000.0), a.acc_unit);
    double rd; int ri;
    radius_to_voxel_f32(a.radius[item_px0 + p], (float)a.radius_scale, (float)a.acc_unit, rd, ri);
    pool.Rd[o] = rd;
    pool.Ri[o] = ri;
  }
}

// Exclusive scan of per-item counts into pool offsets (one block; n_items <= a few 10^4).
__global__ void k_scan_items(const int* __restrict__ cnt, int n_items, long long cap, ItemMeta* meta) {
  __shared__ long long s_part[1024];
  const int per = (n_items + blockDim.x - 1) / blockDim.x;
  const int b0 = threadIdx.x * per, b1 = min(n_items, b0 + per);
  long long s = 0;
  for (int b = b0; b < b1; ++b) s += cnt[b];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) { long long v = s_part[t]; s_part[t] = run; run += v; }
  }
  __syncthreads();
  long long o = s_part[threadIdx.x];
  for (int b = b0; b < b1; ++b) {
    ItemMeta m;
    memset(&m, 0, sizeof(m));
    m.off = o;
    m.n = cnt[b];
    if (o + m.n > cap) { m.status = RCV_ST_POINT_OVERFLOW; m.n = 0; m.off = 0; }
    o += cnt[b];
    meta[b] = m;
  }
}

// ---- scene cloud of a frame: the union over its keypoints of the masked clouds ------------------------------------
// The reference builds the ICP target `xyz_mm_icp` by appending, keypoint after keypoint, the points not seen before
// (an O(N^2) Python loop, AccumulatorSpace.py:620-625, :863-868, :1070-1075).  All of a frame's clouds come from the same
// depth map, so the union is the back-projection of the pixels that survive ANY keypoint's mask rule: one CTA per frame
// ORs the survival bits of the keypoints; compaction and conversion reuse the K1 kernels' scheme.  The points come out in
// row-major pixel order (the reference's order is first appearance by keypoint; ICP does not depend on the order).
__global__ void __launch_bounds__(kCompactThreads) k_scene_mask(FrameArgs a, int* __restrict__ cnt, unsigned* __restrict__ bits, int words_per_item) {
  const int frame = blockIdx.x;
  const int npx = a.fp.height * a.fp.width;
  const long long frame_px0 = (long long)frame * npx;
  const bool vec_ok = a.vec_ok != 0;
  unsigned* out = bits + (long long)frame * words_per_item;
  const int lane = threadIdx.x & 31;
  int n = 0;
  for (int px = threadIdx.x * kPxPerThread; px < words_per_item * 32; px += kPxPerIter) {
    double zraw[kPxPerThread];
    float rad[kPxPerThread];
    unsigned m = 0;
    if (px < npx) {
      for (int kp = 0; kp < a.n_kpts; ++kp) {
        const double max_r = a.max_radii ? a.max_radii[(long long)frame * a.fp.max_radii_stride + kp] : 0.0;
        m |= pixel_group_mask(a, frame_px0, ((long long)frame * a.n_kpts + kp) * npx, px, npx, vec_ok, max_r, zraw, rad);
      }
    }
    n += __popc(m);
    unsigned w = m << ((lane & 3) * 8);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    if ((lane & 3) == 0) out[px >> 5] = w;
  }
  __shared__ int s_n[kCompactThreads / 32];
  n = __reduce_add_sync(0xffffffffu, n);
  if (lane == 0) s_n[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) t += s_n[w];
    cnt[frame] = t;
  }
}

// Exclusive scan of the per-frame counts into offsets[n_frames + 1]; frames that do not fit below `cap` get an empty range
// and status RCV_ST_POINT_OVERFLOW (once one frame overflows all later ones do).
__global__ void k_scene_scan(const int* __restrict__ cnt, int n_frames, long long cap, ItemMeta* meta, long long* __restrict__ offsets,
                             int* __restrict__ status) {
  __shared__ long long s_part[1024];
  __shared__ unsigned long long s_limit;
  const int per = (n_frames + blockDim.x - 1) / blockDim.x;
  const int b0 = threadIdx.x * per, b1 = min(n_frames, b0 + per);
  long long s = 0;
  for (int b = b0; b < b1; ++b) s += cnt[b];
  s_part[threadIdx.x] = s;
  if (threadIdx.x == 0) s_limit = ~0ull;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) { long long v = s_part[t]; s_part[t] = run; run += v; }
    if (run <= cap) s_limit = (unsigned long long)run;   // everything fits: the end offset is the total
  }
  __syncthreads();
  long long o = s_part[threadIdx.x];
  for (int b = b0; b < b1; ++b) {
    if (o + cnt[b] > cap) atomicMin(&s_limit, (unsigned long long)o);
    o += cnt[b];
  }
  __syncthreads();
  const long long limit = (long long)s_limit;
  o = s_part[threadIdx.x];
  for (int b = b0; b < b1; ++b) {
    ItemMeta m;
    memset(&m, 0, sizeof(m));
    const bool fits = o + cnt[b] <= cap;
    m.off = fits ? o : limit;
    m.n = fits ? cnt[b] : 0;
    m.status = fits ? (cnt[b] ? RCV_ST_OK : RCV_ST_EMPTY_MASK) : RCV_ST_POINT_OVERFLOW;
    meta[b] = m;
    offsets[b] = m.off;
    if (status) status[b] = m.status;
    o += cnt[b];
  }
  if (threadIdx.x == 0) offsets[n_frames] = limit;
}

// rgbd_to_point_cloud (AccumulatorSpace.py:77-85) of the union pixels, times `scale` (the YCB evaluator's xyz_icp*1000, :1154)
__global__ void __launch_bounds__(256) k_scene_points(FrameArgs a, const ItemMeta* __restrict__ meta, const int* __restrict__ pix, double scale,
                                                     double* __restrict__ xyz) {
  const int frame = blockIdx.x;
  const ItemMeta m = meta[frame];
  if (m.n == 0) return;
  const int W = a.fp.width, npx = a.fp.height * a.fp.width;
  const long long frame_px0 = (long long)frame * npx;
  const double* Kp = a.K + (long long)frame * a.fp.k_stride;
  const double fx = Kp[0], cx = Kp[2], fy = Kp[4], cy = Kp[5];
  for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < m.n; q += gridDim.y * blockDim.x) {
    const long long o = m.off + q;
    const int p = pix[o];
    const int u = p % W, v = p / W;
    const double z = __ddiv_rn(load_depth(a.depth, a.fp.depth_dtype, frame_px0 + p), a.fp.depth_div);
    const double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, cx), z), fx);
    const double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, cy), z), fy);
    xyz[3 * o + 0] = __dmul_rn(x, scale);
    xyz[3 * o + 1] = __dmul_rn(y, scale);
    xyz[3 * o + 2] = __dmul_rn(z, scale);
  }
}

// single-image compaction for rcv_backproject (AoS float64 output, like the reference's (N,3) array)
__global__ void __launch_bounds__(1024) k_backproject(const double* __restrict__ K, const void* __restrict__ depth, int dtype, int H, int W,
                                                     double* __restrict__ xyz, long long cap, int* __restrict__ n_out) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, npx = H * W;
  const double fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int p0 = 0; p0 < npx; p0 += 1024) {
    const int p = p0 + threadIdx.x;
    double z = 0.0;
    if (p < npx) z = load_depth(depth, dtype, p);
    const bool ok = z != 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int t = s_warp[w]; if (w < warp) wbase += t; total += t; }
    const int base = s_base;
    if (ok) {
      const long long o = base + wbase + __popc(bal & ((1u << lane) - 1));
      if (o < cap) {
        const int u = p % W, v = p / W;
        xyz[3 * o + 0] = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, cx), z), fx);
        xyz[3 * o + 1] = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, cy), z), fy);
        xyz[3 * o + 2] = z;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = s_base;
}

// ------------------------------------------------------------------------------------------------
// P -- Accumulator_3D prelude (AccumulatorSpace.py:373-401), one CTA per item.
// The per-axis means reproduce numpy's pairwise summation (np.mean -> add.reduce, blocks of <=128
// with 8 running partial sums, recursive halving rounded down to a multiple of 8), so the recentred
// coordinates are bit-identical to the reference's.
// ------------------------------------------------------------------------------------------------
struct PwLeaf { int start; short len; short adds; };

__device__ double pw_leaf_sum(const double* __restrict__ a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

// vote order of an item's points: bins of (R, y voxel), R major.  (y major, generation 1's order, leaves 79 % of the run
// kernel's lane-slots live against 84 %: lanes of equal radius keep the same column range in every slice of a chunk.)
#ifdef RCV_SORT_Y_FIRST
#define RCV_SORT_BIN(A, R) ((A) * nRq + (R))
#else
#define RCV_SORT_BIN(A, R) ((R) * nAq + (A))
#endif

struct PreludeArgs {
  Pool pool; ItemMeta* meta; Unit* units; int* counters; int max_units; int max_grid; int policy; int tile_words; int gen; int dp_mod;
  int full_window;   // 1: tiles cover the whole grid (RCV_FULL_WINDOW=1: A/B against the windows)
  PwLeaf* leaves; double* leaf_sums; long long leaf_cap;
};

constexpr int kPreludeThreads = 256;
constexpr int kMaxGuard = 8;     // guard cells per side of a whole-slice tile; beyond that the item uses the clipped passes
constexpr int kSortBins = 8192;   // bins of the (y voxel, R) counting sort that orders an item's points for k_vote

// Row stride of a tile: odd, so that consecutive rows start in different banks; dp_mod >= 0 (experiments) asks for a given
// residue modulo 32.
// gen 2: the 32 lanes of a warp mark ~10 neighbouring rows at nearly the same cell, so the stride decides the bank
// conflicts of the atomics: residues near 0, 16 and +-1..3 modulo 32 fold neighbouring rows onto the same banks (2.0-2.8
// wavefronts per atomic in a replay of the bench geometry), the residues kept here give 1.45-1.6.
__device__ __forceinline__ int row_stride(int need, int dp_mod, int gen) {
  if (dp_mod >= 0) { int dp = need; while ((dp & 31) != (dp_mod & 31)) ++dp; return dp; }
  if (gen < 2) return need | 1;
  const unsigned good = (1u << 5) | (1u << 7) | (1u << 9) | (1u << 11) | (1u << 13) | (1u << 19) | (1u << 21) | (1u << 23) | (1u << 25) | (1u << 27);
  int dp = need;
  while (!((good >> (dp & 31)) & 1u)) ++dp;
  return dp;
}

__global__ void __launch_bounds__(kPreludeThreads) k_prelude(PreludeArgs a) {
  const int item = blockIdx.x;
  ItemMeta m = a.meta[item];
  const int n = m.n;
  __shared__ int s_nleaf;
  __shared__ double s_mean[3];
  __shared__ double s_red[3][kPreludeThreads / 32];
  __shared__ double s_red2[8][kPreludeThreads / 32];
  __shared__ int s_zb, s_ok, s_ubase;
  if (n <= 0) {
    if (threadIdx.x == 0) { if (!(m.status & RCV_ST_POINT_OVERFLOW)) m.status |= RCV_ST_EMPTY_MASK; a.meta[item] = m; }
    return;
  }
  double* X = a.pool.X + m.off; double* Y = a.pool.Y + m.off; double* Z = a.pool.Z + m.off;
  const double* Rd = a.pool.Rd + m.off;
  // disjoint scratch window for this item's leaves: an item with n points has <= n/57 + 1 leaves
  const long long lbase = m.off / 56 + 2LL * item;
  PwLeaf* leaves = a.leaves + lbase;
  double* lsum = a.leaf_sums + 3 * lbase;
  if (threadIdx.x == 0) {
    int st_start[40], st_len[40], st_rc[40], top = 0, nl = 0;
    st_start[0] = 0; st_len[0] = n; st_rc[0] = 0; top = 1;
    while (top) {
      --top;
      const int s = st_start[top], l = st_len[top], rc = st_rc[top];
      if (l <= 128) { PwLeaf lf; lf.start = s; lf.len = (short)l; lf.adds = (short)rc; leaves[nl++] = lf; }
      else {
        int n2 = l / 2; n2 -= n2 % 8;
        st_start[top] = s + n2; st_len[top] = l - n2; st_rc[top] = rc + 1; ++top;  // right child (popped second)
        st_start[top] = s; st_len[top] = n2; st_rc[top] = 0; ++top;                // left child
      }
    }
    s_nleaf = nl;
  }
  __syncthreads();
  const int nleaf = s_nleaf;
  for (int t = threadIdx.x; t < 3 * nleaf; t += blockDim.x) {
    const int l = t / 3, c = t % 3;
    const PwLeaf lf = leaves[l];
    const double* src = (c == 0 ? X : c == 1 ? Y : Z) + lf.start;
    lsum[3 * l + c] = pw_leaf_sum(src, lf.len);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    double st[40]; int top = 0;
    for (int l = 0; l < nleaf; ++l) {
      st[top++] = lsum[3 * l + c];
      for (int q = leaves[l].adds; q > 0; --q) { const double r = st[--top]; const double lft = st[--top]; st[top++] = __dadd_rn(lft, r); }
    }
    s_mean[c] = __ddiv_rn(st[0], (double)n);
  }
  __syncthreads();
  const double mx = s_mean[0], my = s_mean[1], mz = s_mean[2];
  // recentre; global min / max over all three axes; max radius
  double vmin = INFINITY, vmax = -INFINITY, rmax = -INFINITY;
  double elo = INFINITY, ehi = -INFINITY;   // extent of the drawn spheres (coordinate -+ integer radius), for the guard band
  double axl[3] = {INFINITY, INFINITY, INFINITY}, axh[3] = {-INFINITY, -INFINITY, -INFINITY};   // the same per axis (x, y, z): the tile windows
  for (int q = threadIdx.x; q < n; q += blockDim.x) {
    const double x = __dsub_rn(X[q], mx), y = __dsub_rn(Y[q], my), z = __dsub_rn(Z[q], mz);
    X[q] = x; Y[q] = y; Z[q] = z;
    const double lo3 = fmin(x, fmin(y, z)), hi3 = fmax(x, fmax(y, z));
    vmin = fmin(vmin, lo3);
    vmax = fmax(vmax, hi3);
    rmax = fmax(rmax, Rd[q]);
    const int ri = a.pool.Ri[m.off + q];
    if (ri > 0) {
      elo = fmin(elo, lo3 - (double)ri); ehi = fmax(ehi, hi3 + (double)ri);
      axl[0] = fmin(axl[0], x - (double)ri); axh[0] = fmax(axh[0], x + (double)ri);
      axl[1] = fmin(axl[1], y - (double)ri); axh[1] = fmax(axh[1], y + (double)ri);
      axl[2] = fmin(axl[2], z - (double)ri); axh[2] = fmax(axh[2], z + (double)ri);
    }
  }
  for (int c3 = 0; c3 < 3; ++c3) { axl[c3] = warp_min_f64(axl[c3]); axh[c3] = warp_max_f64(axh[c3]); }
  vmin = warp_min_f64(vmin); vmax = warp_max_f64(vmax); rmax = warp_max_f64(rmax);
  elo = warp_min_f64(elo); ehi = warp_max_f64(ehi);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_red[0][warp] = vmin; s_red[1][warp] = vmax; s_red[2][warp] = rmax; s_red2[0][warp] = elo; s_red2[1][warp] = ehi;
    for (int c3 = 0; c3 < 3; ++c3) { s_red2[2 + c3][warp] = axl[c3]; s_red2[5 + c3][warp] = axh[c3]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kPreludeThreads / 32; ++w) {
      vmin = fmin(vmin, s_red[0][w]); vmax = fmax(vmax, s_red[1][w]); rmax = fmax(rmax, s_red[2][w]);
      elo = fmin(elo, s_red2[0][w]); ehi = fmax(ehi, s_red2[1][w]);
      for (int c3 = 0; c3 < 3; ++c3) { axl[c3] = fmin(axl[c3], s_red2[2 + c3][w]); axh[c3] = fmax(axh[c3], s_red2[5 + c3][w]); }
    }
    // zero_boundary = int(xyz_mm_min - radius_max) + 1   (int() truncates toward zero)
    const double zbd = __dsub_rn(vmin, rmax);
    int zb = (int)zbd + 1;  // values beyond int range are rejected below through D
    double pmax = vmax;
    if (zb < 0) pmax = __dsub_rn(vmax, (double)zb);  // max(x - zb) == max(x) - zb: the subtraction is monotone
    const double Dd = (a.policy == RCV_POLICY_YCBGEN) ? trunc(pmax) + 1.0 : trunc(pmax) + trunc(rmax);
    int status = m.status, D = 0, ok = 0;
    if (!(fabs(zbd) < 1e9) || !(fabs(Dd) < 1e9)) status |= RCV_ST_BAD_GRID;
    else {
      D = (int)Dd;
      if (D <= 0) status |= RCV_ST_BAD_GRID;
      else if (D > a.max_grid) status |= RCV_ST_D_EXCEEDS_CAP;
      else ok = 1;
    }
    m.D = D; m.zb = zb; m.rmax = rmax; m.mean[0] = mx; m.mean[1] = my; m.mean[2] = mz; m.status = status;
    int nunits = 0;
    if (ok) {
      // Guard band: how far the candidates of the item's spheres can leave [0, D) along the in-slice axes (they lie within
      // R + 2 of the nearest lattice point, ring_noclip).  The reference sizes the grid so that spheres overhang by at most
      // ~3 voxels; with the guard cells inside the tile no warp of the item needs the clipped passes.
      // nearest lattice point within 0.5 of the (shifted) coordinate, candidates within R + 2 of it
      const double shift = zb < 0 ? (double)zb : 0.0;
      const double need_lo = 2.5 - (elo - shift), need_hi = (ehi - shift) + 2.5 - (double)(D - 1);
      int glo = need_lo > 0.0 ? (need_lo < 1e6 ? (int)ceil(need_lo) : kMaxGuard + 1) : 0;
      int ghi = need_hi > 0.0 ? (need_hi < 1e6 ? (int)ceil(need_hi) : kMaxGuard + 1) : 0;
      if (glo > kMaxGuard || ghi > kMaxGuard) { glo = 0; ghi = 0; }   // YCBGEN-style grids without upper pad: clipped passes
      const bool clipped = (need_lo > 0.0 || need_hi > 0.0) && glo == 0 && ghi == 0;
      // Generation 2: windows.  A sphere reaches the voxels within its integer radius of its centre (+ the margin of the guard
      // rule), so per axis only [min(coord - R) - 2.5, max(coord + R) + 2.5] can hold a mark.  Tiles cover that box only: the rows
      // [rw0, rw0 + rwn) and slices [sw0, sw0 + swn) inside the grid, and along C the cells [-glo, D - 1 + ghi] with SIGNED glo / ghi --
      // guard cells where the spheres overhang the grid, fewer cells than D where they do not (the reference sizes the grid by the
      // widest axis; along the others, typically z for a surface seen by a camera, ~10 % of it is never reached).
      int rw0 = 0, rwn = D, sw0 = 0, swn = D;
      if (a.gen >= 2 && !a.full_window && axl[0] <= axh[0]) {
        auto lo_cell = [&](double v) { const double f = floor(v - shift - 2.5); return f < -1e6 ? -1000000 : (f > 1e6 ? 1000000 : (int)f); };
        auto hi_cell = [&](double v) { const double f = ceil(v - shift + 2.5); return f < -1e6 ? -1000000 : (f > 1e6 ? 1000000 : (int)f); };
        const int x0 = max(lo_cell(axl[0]), 0), x1 = min(hi_cell(axh[0]), D - 1), y0 = max(lo_cell(axl[1]), 0), y1 = min(hi_cell(axh[1]), D - 1);
        if (x0 <= x1) { rw0 = x0; rwn = x1 - x0 + 1; } else { rw0 = 0; rwn = 1; }       // (spheres entirely outside the grid: nothing to draw)
        if (y0 <= y1) { sw0 = y0; swn = y1 - y0 + 1; } else { sw0 = 0; swn = 1; }
        if (!clipped) {
          const int z0 = lo_cell(axl[2]), z1 = hi_cell(axh[2]);
          int g0 = -z0, g1 = z1 - (D - 1);                       // signed guards of the cell window [z0, z1]
          if (g0 <= -D) g0 = -(D - 1);                           // (window entirely above / below the grid: keep one cell)
          if (g1 <= -D) g1 = -(D - 1);
          if (g0 <= kMaxGuard && g1 <= kMaxGuard && D + g0 + g1 >= 1) { glo = g0; ghi = g1; }
        }
      }
      // odd row stride: consecutive rows start in different banks.  The run rasteriser (gen 2) keeps one spare cell above the
      // upper guard: the -1 that closes a run ending at the last cell lands there.
      // gen 2 also keeps TWO planes per slice (run starts, run ends: only `add 1` merges lanes that hit the same address).
      const int spare = a.gen >= 2 ? 1 : 0, planes = a.gen >= 2 ? 2 : 1;
      int Dp = row_stride(D + glo + ghi + spare, a.dp_mod, a.gen);
      // Generation 2 keeps guard CELLS (along z) only: the rows a lane draws are the intersection of its column range with the tile's
      // rows, so spheres that overhang the grid along x simply lose those columns and a slice is the window's rows, not D + guards.
      const int rows_whole = a.gen >= 2 ? rwn : D + glo + ghi;
      long long slice = (long long)planes * rows_whole * Dp;
      // A slice that does not fit is cut into row bands.  Generation 1 does not guard them; generation 2 keeps the guard cells
      // along C (z) -- the rows of a band are restricted by the lanes' column ranges, so with the cells guarded no boundary
      // needs a bounds test and the band runs the same lean column loops as a whole slice.
      if (slice > a.tile_words) {
        if (a.gen < 2) { glo = 0; ghi = 0; Dp = row_stride(D + spare, a.dp_mod, a.gen); slice = (long long)planes * D * Dp; }
      }
      m.band = slice > a.tile_words ? 1 : 0;
      m.Dp = Dp;
      m.guard = (glo & 0xffff) | (ghi << 16);
      m.clip = ((a.gen < 2 && slice > a.tile_words) || clipped) ? 1 : 0;
      m.rw0 = rw0; m.rwn = rwn; m.sw0 = sw0; m.swn = swn;
      if (slice <= a.tile_words) {
        int ni_max = (int)(a.tile_words / slice);
        if (ni_max > 32) ni_max = 32;   // (gen 1: the polar pass keeps one mask bit per slice of a tile)
        if (ni_max > swn) ni_max = swn;
        m.ni = ni_max >= swn ? swn : (a.gen >= 2 ? ni_max : slab_thickness(ni_max));   // gen 1: a multiple of 3 or 4 (its ring passes walk a slab in whole chunks)
        m.nj = a.gen >= 2 ? rwn : D + glo + ghi;
        nunits = (swn + m.ni - 1) / m.ni;
      } else {
        const int nj_max = a.tile_words / (planes * Dp);
        const int nt = (rwn + nj_max - 1) / nj_max;
        m.ni = 1; m.nj = (rwn + nt - 1) / nt;
        nunits = swn * ((rwn + m.nj - 1) / m.nj);
      }
      const int ub = atomicAdd(&a.counters[0], nunits);
      if (ub + nunits > a.max_units) { m.status |= RCV_ST_UNIT_OVERFLOW; ok = 0; atomicSub(&a.counters[0], nunits); }
      s_ubase = ub;
    }
    s_zb = zb; s_ok = ok;
    a.meta[item] = m;
  }
  __syncthreads();
  if (!s_ok) return;
  const int zb = s_zb;
  if (zb < 0) {
    const double zbd = (double)zb;
    for (int q = threadIdx.x; q < n; q += blockDim.x) { X[q] = __dsub_rn(X[q], zbd); Y[q] = __dsub_rn(Y[q], zbd); Z[q] = __dsub_rn(Z[q], zbd); }
  }
  // ---- vote order: counting sort of the item's points by (nearest y voxel, integer radius) ----
  // k_vote gives every lane of a warp one point; lanes that share the slice coordinate and the radius have the same
  // ring class and size in every slice, so no lane idles while another draws.  (Votes are integer adds: order-free.)
  {
    __shared__ int s_hist[kSortBins];
    __shared__ int s_mm[4];
    __shared__ int s_wsum[kPreludeThreads / 32];
    if (threadIdx.x == 0) { s_mm[0] = 0x7fffffff; s_mm[1] = -0x7fffffff; s_mm[2] = 0x7fffffff; s_mm[3] = 0; }
    __syncthreads();
    int amin = 0x7fffffff, amax = -0x7fffffff, rmin = 0x7fffffff, rmx = 0;
    const int* Ri = a.pool.Ri + m.off;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int ya = __double2int_rn(Y[q]), r = max(Ri[q], 0);
      amin = min(amin, ya); amax = max(amax, ya); rmin = min(rmin, r); rmx = max(rmx, r);
    }
    amin = __reduce_min_sync(0xffffffffu, amin); amax = __reduce_max_sync(0xffffffffu, amax);
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmx = __reduce_max_sync(0xffffffffu, rmx);
    if (lane == 0) { atomicMin(&s_mm[0], amin); atomicMax(&s_mm[1], amax); atomicMin(&s_mm[2], rmin); atomicMax(&s_mm[3], rmx); }
    __syncthreads();
    amin = s_mm[0]; amax = s_mm[1]; rmin = s_mm[2]; rmx = s_mm[3];
    const long long nA = (long long)amax - amin + 1;
    int sh = 0, ash = 0;
    while (nA > (kSortBins >> 1) && ((nA >> ash) > (kSortBins >> 1))) ++ash;                       // degenerate extents only
    const int nAq = (int)(((long long)amax - amin) >> ash) + 1;
    while ((long long)nAq * (((rmx - rmin) >> sh) + 1) > kSortBins) ++sh;
    const int nRq = ((rmx - rmin) >> sh) + 1, NB = nAq * nRq;
    for (int b = threadIdx.x; b < NB; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int ya = __double2int_rn(Y[q]), r = max(Ri[q], 0);
      atomicAdd(&s_hist[RCV_SORT_BIN((ya - amin) >> ash, (r - rmin) >> sh)], 1);
    }
    __syncthreads();
    // exclusive scan of the bins: each thread owns a run of consecutive bins
    const int per = (NB + kPreludeThreads - 1) / kPreludeThreads;
    const int b0 = min(NB, (int)threadIdx.x * per), b1 = min(NB, b0 + per);
    int tsum = 0;
    for (int b = b0; b < b1; ++b) tsum += s_hist[b];
    int incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_wsum[w];
    int run = wbase + incl - tsum;
    for (int b = b0; b < b1; ++b) { const int v = s_hist[b]; s_hist[b] = run; run += v; }
    __syncthreads();
    int* perm = a.pool.perm + m.off;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int ya = __double2int_rn(Y[q]), r = max(Ri[q], 0);
      const int pos = atomicAdd(&s_hist[RCV_SORT_BIN((ya - amin) >> ash, (r - rmin) >> sh)], 1);
      perm[pos] = q;
      if (a.gen >= 2) {   // internal axes (A,B,C) = reference (y,x,z)
        RunPoint rp;
        run_point_setup(rp, Y[q], X[q], Z[q], Ri[q]);
        // a point that draws nothing (R <= 0) still leaves its four cancelling marks at its own cell: it must be one the rows store
        if (rp.R <= 0) { const int g0 = (short)(a.meta[item].guard & 0xffff); rp.ipc = g0 < 0 ? -g0 : 0; rp.fc = 0.f; }
        int4* rec = a.pool.rec + 2 * (m.off + pos);
        rec[0] = make_int4(rp.ipa, rp.ipb, rp.ipc, rp.R);
        rec[1] = make_int4(__float_as_int(rp.fa), __float_as_int(rp.fb), __float_as_int(rp.fc), __float_as_int(rp.W));
      }
    }
  }
  // gen 2: slice range reached by each group of 32 points in vote order (the vote kernel builds a tile's work list from these)
  if (a.gen >= 2) {
    __syncthreads();
    int* grp = a.pool.grp + (m.off >> 5) + item;
    const int4* rec = a.pool.rec + 2 * m.off;
    for (int g = warp; g * 32 < n; g += kPreludeThreads / 32) {
      const int q = g * 32 + lane;
      int lo = 0x7fff, hi = -0x8000;
      if (q < n) {
        const int4 r0 = rec[2 * q];
        if (r0.w > 0) { lo = max(r0.x - r0.w - 1, -0x8000); hi = min(r0.x + r0.w + 1, 0x7fff); }
      }
      lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
      if (lane == 0) grp[g] = (lo & 0xffff) | (hi << 16);
    }
  }
  // tile work list
  m = a.meta[item];
  const int glo = (short)(m.guard & 0xffff);
  const bool whole = !m.band;       // whole-slice tiles; otherwise row bands over the row window
  const int tj = whole ? 1 : (m.rwn + m.nj - 1) / m.nj, ti = (m.swn + m.ni - 1) / m.ni;
  for (int t = threadIdx.x; t < ti * tj; t += blockDim.x) {
    Unit u;
    u.item = item;
    u.i0 = m.sw0 + (t / tj) * m.ni; u.ni = min(m.ni, m.sw0 + m.swn - u.i0);
    if (whole) { u.j0 = m.nj > m.D ? -glo : m.rw0; u.nj = m.nj; }   // gen 1: guard rows inside the tile; gen 2: the row window
    else { u.j0 = m.rw0 + (t % tj) * m.nj; u.nj = min(m.nj, m.rw0 + m.rwn - u.j0); }
    a.units[s_ubase + t] = u;
  }
}

// ------------------------------------------------------------------------------------------------
// K2 -- the vote kernel.  Persistent: one CTA per SM pulls (item, tile) units from a global queue.
// ------------------------------------------------------------------------------------------------
struct VoteArgs {
  Pool pool; const ItemMeta* meta; const Unit* units; int* counters;
  unsigned long long* best; unsigned long long* votes;
  int32_t* volume; long long volume_cap;
};

// Shared-memory atomics by 32-bit shared-window address (offsets are bytes): the fast path adds 1 to
// the voxel or, for a lane with no vote, to the lane's private sink word, so the instruction is never
// predicated or branched around (a predicated ATOMS compiles to a divergent branch, 3.6x slower).
__device__ __forceinline__ void smem_inc(unsigned addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
struct SmemEmit {   // offsets are bytes relative to the tile; `sink` is the lane's private word after the tile
  unsigned base; int sink;
  __device__ __forceinline__ void operator()(int off) const { smem_inc(base + (unsigned)off); }
};
// Internal axes.  The rasteriser walks "slices" along its first axis A; the kernel feeds it (A,B,C) = (y,x,z)
// of the reference so that the 32 points of a warp -- consecutive surviving pixels of an image row -- share
// their slice geometry (same y up to a voxel).  exact_hit() gets the coordinates back in reference order
// because the float64 sum dx^2 + dy^2 + dz^2 is order-sensitive.
// The exact float64 path lives behind a call so the compiler cannot hoist any of its arithmetic into the
// hot loop (it did: 40% of the first version's instructions were speculated DADD/DMUL/DSQRT).
__device__ __noinline__ bool exact_hit_call(double px, double py, double pz, int R, int i, int j, int k) {
  return exact_hit(px, py, pz, R, i, j, k);
}
struct SlowExactCall {   // arguments are internal (A,B,C) indices
  double pa, pb, pc; int R;
  __device__ __forceinline__ bool operator()(int ia, int ib, int ic) const { return exact_hit_call(pb, pa, pc, R, ib, ia, ic); }
};
__device__ __noinline__ void ring_slow_call(double px, double py, double pz, int R, int ipy, int ipz, float hw_m, float fv, float thr,
                                            int vrel0, int vn, int ucoord, int pass, int sv, int i, int ub, int arc, float q, float fl, int vt,
                                            unsigned base) {
  PointCtx c;
  c.px = px; c.py = py; c.pz = pz; c.R = R; c.ipy = ipy; c.ipz = ipz; c.hw_m = hw_m;
  LaneTask L;
  L.fv = fv; L.thr = thr; L.vrel0 = vrel0; L.vn = vn; L.ucoord = ucoord; L.pass = pass != 0; L.sv = sv;
  SlowExactCall slow{px, py, pz, R};
  SmemEmit es{base, 0};
  ring_slow(c, L, i, ub, arc, q, fl, vt, slow, es);
}
struct SlowArcCall {
  unsigned base;
  __device__ __forceinline__ void operator()(const PointCtx& c, const LaneTask& L, int i, int ub, int arc, float q, float fl, int vt) const {
    ring_slow_call(c.px, c.py, c.pz, c.R, c.ipy, c.ipz, c.hw_m, L.fv, L.thr, L.vrel0, L.vn, L.ucoord, L.pass ? 1 : 0, L.sv, i, ub, arc, q, fl, vt, base);
  }
};

__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(0xffffffffu, v); }

// One pass over a chunk of NC consecutive slices for a warp whose 32 lanes each own one point: every lane walks the
// columns (rows) u = -H..H of ITS thin rings (one candidate per arc); the lane task of column u is shared by the
// chunk's slices.  a4[s] is NaN for a slice that is not a thin ring of this lane (NaN never votes, never asks for
// the exact path).  H is the maximum over the warp.  CLIP = some lane's candidates may leave the tile.
template <bool CLIP, int NC>
__device__ __forceinline__ void ring_pass(const PointCtx& c, const Tile& t, const float (&a4)[NC], int H, int i0c, int sbase0, int slice_bytes,
                                          const SmemEmit& emit_c, const SlowArcCall& slowarc_c) {
  SmemEmit emit = emit_c;
  SlowArcCall slowarc = slowarc_c;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {   // one copy of the code for both passes (instruction cache)
    float uf = (float)(-H);
#pragma unroll 1
    for (int u = -H; u <= H; ++u, uf += 1.0f) {
      LaneTask L;
      lane_setup_pu(c, t, pass != 0, u, uf, 4, L);
      ThinOut o[NC];
#pragma unroll
      for (int sidx = 0; sidx < NC; ++sidx) thin_fast<CLIP>(c, a4[sidx], L, sbase0 + sidx * slice_bytes, emit.sink, emit, o[sidx]);
      bool any = false;
#pragma unroll
      for (int sidx = 0; sidx < NC; ++sidx) any |= o[sidx].t0 | o[sidx].t1;
      if (__builtin_expect(any, 0)) {
#pragma unroll
        for (int sidx = 0; sidx < NC; ++sidx)
          if (o[sidx].t0 || o[sidx].t1) thin_slow(c, L, i0c + sidx, o[sidx], slowarc);
      }
    }
  }
}

// ---- fast ring pass (raster_core.h "ring2"): one PTX block per column of a chunk ---------------------------------
// The two arcs of a slice run through the same float sequence with mirrored constants, so they are issued as packed
// fp32 pairs (sm_100 FADD2 / FFMA2: one issue slot for two IEEE-rounded results -- measured in tools/ubench_f32x2.cu:
// the packed op takes the FMA pipe for two cycles but one issue slot, which is what this issue-bound loop is short of).
// Pair operands (lo = top arc, hi = bottom arc):
//   %1 (du, du)     %2 (mu0, mu1)   %3 (hW, hW)   %4 (cp, cm)   %5 (-fv, +fv)   a2[s] = (a_s, a_s)
// Scalars: %0 flags (in/out) | %6 hw_m %7 -hw_p %8 sink %9 sv %10 thr %11 -sv | a2[NC], K0[NC], K1[NC], bit.
// g = a - du^2 is ONE packed fma (ptxas would contract a packed multiply + subtract anyway, .rn or not; the host mirror
// ring2_fast() uses the fused form too).  Per slice: packed fma + packed subtract + MUFU.SQRT + 5 packed ops (+3 for the second candidates); per candidate: one
// multiply-add for the address, two or three FSETP, one select, one shared-memory atomic, one predicate op.
#define R2_DECL \
  "{\n\t.reg .pred pv, pl, po, pa;\n\t.reg .f32 zs, q0, q1, d0, d1, aq, ad, glo, ghi;\n\t.reg .b32 b0, b1, adr, adu;\n\t" \
  ".reg .b64 g2, zs2, hwg2, t2, tm2, fl2, d2, q2, m1, nd2, z2;\n\tsetp.ne.u32 pa, %8, %8;\n\tmov.b64 m1, 0xBF800000BF800000;\n\t" \
  "mov.b64 z2, 0;\n\tsub.rn.f32x2 nd2, z2, %1;\n\t"
#define R2_SLICE_HEAD(A) \
  "fma.rn.f32x2 g2, nd2, %1, " A ";\n\tmov.b64 {glo, ghi}, g2;\n\tsqrt.approx.ftz.f32 zs, glo;\n\tmov.b64 zs2, {zs, zs};\n\tsub.rn.f32x2 hwg2, %3, g2;\n\t" \
  "add.rn.f32x2 t2, zs2, %4;\n\tadd.rn.f32x2 tm2, t2, %2;\n\tsub.rn.f32x2 fl2, tm2, %2;\n\tadd.rn.f32x2 d2, fl2, %5;\n\t" \
  "fma.rn.f32x2 q2, d2, d2, hwg2;\n\tmov.b64 {q0, q1}, q2;\n\tmov.b64 {d0, d1}, d2;\n\tmov.b64 {b0, b1}, tm2;\n\t"
// candidates one voxel inwards of both arcs: (integer - 1) -+ fraction, one rounding like every coordinate difference
#define R2_SLICE_HEAD2 \
  "add.rn.f32x2 fl2, fl2, m1;\n\tadd.rn.f32x2 d2, fl2, %5;\n\tfma.rn.f32x2 q2, d2, d2, hwg2;\n\tmov.b64 {q0, q1}, q2;\n\tmov.b64 {d0, d1}, d2;\n\t"
#define R2_VOTE_INT(Q) "abs.f32 aq, " Q ";\n\tsetp.lt.f32 pv, aq, %6;\n\tsetp.gt.f32 pl, " Q ", %7;\n\t"
#define R2_VOTE_OWN(Q, D) \
  "abs.f32 aq, " Q ";\n\tabs.f32 ad, " D ";\n\tsetp.gt.f32 po, ad, %10;\n\tsetp.lt.and.f32 pv, aq, %6, po;\n\tsetp.gt.and.f32 pl, " Q ", %7, po;\n\t"
#define R2_EMIT_TAIL "selp.u32 adr, adu, %8, pv;\n\tred.shared.add.u32 [adr], 1;\n\txor.pred pl, pl, pv;\n\tor.pred pa, pa, pl;\n\t"
#define R2_EMIT(B, SV, K) "mad.lo.u32 adu, " B ", " SV ", " K ";\n\t" R2_EMIT_TAIL
#define R2_EMIT_INWARD(ADU, STEP) "add.u32 " ADU ", " ADU ", " STEP ";\n\tmov.b32 adu, " ADU ";\n\t" R2_EMIT_TAIL
// VOTE is a macro name taking (Q, D): R2_V_INT or R2_V_OWN
#define R2_V_INT(Q, D) R2_VOTE_INT(Q)
#define R2_V_OWN(Q, D) R2_VOTE_OWN(Q, D)
#define R2_SLICE1(A, K0, K1, VOTE) R2_SLICE_HEAD(A) VOTE("q0", "d0") R2_EMIT("b0", "%9", K0) VOTE("q1", "d1") R2_EMIT("b1", "%11", K1)
// two candidates per arc: the unselected addresses of the first candidates are kept (au0 / au1) and stepped inwards
#define R2_SLICE2(A, K0, K1, VOTE) \
  R2_SLICE_HEAD(A) VOTE("q0", "d0") "mad.lo.u32 au0, b0, %9, " K0 ";\n\tmov.b32 adu, au0;\n\t" R2_EMIT_TAIL \
  VOTE("q1", "d1") "mad.lo.u32 au1, b1, %11, " K1 ";\n\tmov.b32 adu, au1;\n\t" R2_EMIT_TAIL \
  R2_SLICE_HEAD2 VOTE("q0", "d0") R2_EMIT_INWARD("au0", "%11") VOTE("q1", "d1") R2_EMIT_INWARD("au1", "%9")
#define R2_TAIL(BIT) "@pa or.b32 %0, %0, " BIT ";\n\t}"
#define R2_DECL2 ".reg .b32 au0, au1;\n\t"
#define R2_BODY1(SL, VOTE) R2_DECL R2_DECL2 SL("%12", "%13", "%14", VOTE) R2_TAIL("%15")
#define R2_BODY2(SL, VOTE) R2_DECL R2_DECL2 SL("%12", "%14", "%16", VOTE) SL("%13", "%15", "%17", VOTE) R2_TAIL("%18")
#define R2_BODY3(SL, VOTE) R2_DECL R2_DECL2 SL("%12", "%15", "%18", VOTE) SL("%13", "%16", "%19", VOTE) SL("%14", "%17", "%20", VOTE) R2_TAIL("%21")
#define R2_BODY4(SL, VOTE) \
  R2_DECL R2_DECL2 SL("%12", "%16", "%20", VOTE) SL("%13", "%17", "%21", VOTE) SL("%14", "%18", "%22", VOTE) SL("%15", "%19", "%23", VOTE) R2_TAIL("%24")
#define R2_COMMON_IN "l"(duf2), "l"(mu2), "l"(hW2), "l"(cpm2), "l"(sfv2), "f"(hw_m), "f"(nhw_p), "r"(sink), "r"(sv), "f"(thr), "r"(nsv)
#define R2_IN1 R2_COMMON_IN, "l"(a2[0]), "r"(K0[0]), "r"(K1[0]), "r"(bit)
#define R2_IN2 R2_COMMON_IN, "l"(a2[0]), "l"(a2[1]), "r"(K0[0]), "r"(K0[1]), "r"(K1[0]), "r"(K1[1]), "r"(bit)
#define R2_IN3 R2_COMMON_IN, "l"(a2[0]), "l"(a2[1]), "l"(a2[2]), "r"(K0[0]), "r"(K0[1]), "r"(K0[2]), "r"(K1[0]), "r"(K1[1]), "r"(K1[2]), "r"(bit)
#define R2_IN4 \
  R2_COMMON_IN, "l"(a2[0]), "l"(a2[1]), "l"(a2[2]), "l"(a2[3]), "r"(K0[0]), "r"(K0[1]), "r"(K0[2]), "r"(K0[3]), "r"(K1[0]), "r"(K1[1]), "r"(K1[2]), \
      "r"(K1[3]), "r"(bit)

typedef unsigned long long f32x2_t;   // two floats in an aligned register pair: lo = top arc, hi = bottom arc
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float lo2(f32x2_t a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); (void)hi; return lo; }

// Adds `bit` to `flags` if a candidate of the column needs the exact path.
template <bool OWN, int NC, int M>
__device__ __forceinline__ void ring2_asm(unsigned& flags, unsigned bit, f32x2_t duf2, f32x2_t mu2, f32x2_t hW2, f32x2_t cpm2, f32x2_t sfv2, float hw_m,
                                          float nhw_p, unsigned sink, unsigned sv, unsigned nsv, float thr, const f32x2_t (&a2)[NC],
                                          const unsigned (&K0)[NC], const unsigned (&K1)[NC]) {
  if constexpr (NC == 1) {
    if constexpr (M == 1) {
      if constexpr (OWN) asm volatile(R2_BODY1(R2_SLICE1, R2_V_OWN) : "+r"(flags) : R2_IN1 : "memory");
      else asm volatile(R2_BODY1(R2_SLICE1, R2_V_INT) : "+r"(flags) : R2_IN1 : "memory");
    } else {
      if constexpr (OWN) asm volatile(R2_BODY1(R2_SLICE2, R2_V_OWN) : "+r"(flags) : R2_IN1 : "memory");
      else asm volatile(R2_BODY1(R2_SLICE2, R2_V_INT) : "+r"(flags) : R2_IN1 : "memory");
    }
  } else if constexpr (NC == 2) {
    if constexpr (M == 1) {
      if constexpr (OWN) asm volatile(R2_BODY2(R2_SLICE1, R2_V_OWN) : "+r"(flags) : R2_IN2 : "memory");
      else asm volatile(R2_BODY2(R2_SLICE1, R2_V_INT) : "+r"(flags) : R2_IN2 : "memory");
    } else {
      if constexpr (OWN) asm volatile(R2_BODY2(R2_SLICE2, R2_V_OWN) : "+r"(flags) : R2_IN2 : "memory");
      else asm volatile(R2_BODY2(R2_SLICE2, R2_V_INT) : "+r"(flags) : R2_IN2 : "memory");
    }
  } else if constexpr (NC == 3) {
    if constexpr (M == 1) {
      if constexpr (OWN) asm volatile(R2_BODY3(R2_SLICE1, R2_V_OWN) : "+r"(flags) : R2_IN3 : "memory");
      else asm volatile(R2_BODY3(R2_SLICE1, R2_V_INT) : "+r"(flags) : R2_IN3 : "memory");
    } else {
      if constexpr (OWN) asm volatile(R2_BODY3(R2_SLICE2, R2_V_OWN) : "+r"(flags) : R2_IN3 : "memory");
      else asm volatile(R2_BODY3(R2_SLICE2, R2_V_INT) : "+r"(flags) : R2_IN3 : "memory");
    }
  } else {
    if constexpr (M == 1) {
      if constexpr (OWN) asm volatile(R2_BODY4(R2_SLICE1, R2_V_OWN) : "+r"(flags) : R2_IN4 : "memory");
      else asm volatile(R2_BODY4(R2_SLICE1, R2_V_INT) : "+r"(flags) : R2_IN4 : "memory");
    } else {
      if constexpr (OWN) asm volatile(R2_BODY4(R2_SLICE2, R2_V_OWN) : "+r"(flags) : R2_IN4 : "memory");
      else asm volatile(R2_BODY4(R2_SLICE2, R2_V_INT) : "+r"(flags) : R2_IN4 : "memory");
    }
  }
}

// The rare branch of the fast ring pass (one lane, out of line): everything it needs arrives in scalars so that the hot
// loop keeps nothing alive for it and nothing is re-derived from the point.
__device__ __noinline__ void ring2_slow_call(double pa, double pb, double pc, int R, int ipl, int ipc, int u, int i0c, int pass_nc, float hW,
                                             float hw_m, float hw_p, float duf, float cp, float cm, float fv, float mu0, float mu1, unsigned sv,
                                             float a0, float a1, float a2, float a3, unsigned k00, unsigned k01, unsigned k02, unsigned k03,
                                             unsigned k10, unsigned k11, unsigned k12, unsigned k13) {
  const float a[4] = {a0, a1, a2, a3};
  const unsigned K0[4] = {k00, k01, k02, k03}, K1[4] = {k10, k11, k12, k13};
  SlowExactCall slow{pa, pb, pc, R};
  SmemEmit es{0u, 0};   // addresses are absolute
  ring2_slow_lane((pass_nc & 1) != 0, (pass_nc >> 1) & 7, pass_nc >> 4, ipl, ipc, u, i0c, hW, hw_m, hw_p, duf, cp, cm, fv, mu0, mu1, sv, a, K0, K1, slow, es);
}

// Fast ring pass of one chunk for a warp whose candidates cannot leave the tile.  H = half-width (warp maximum),
// Hin = interior half-width (warp minimum, -1 = none).
template <bool PASS, int NC, int M>
__device__ __forceinline__ void ring_pass2(const PointCtx& c, const Tile& t, const float (&a4)[NC], int H, int Hin, float dbm, int i0c, int sbase0,
                                           int slice_bytes, unsigned base, unsigned sink_abs) {
  const float nhw_p = -c.hw_p;
  const int ioff = Hin >= 0 ? Hin : 0x40000000;
  const unsigned ispan = Hin >= 0 ? 2u * (unsigned)Hin : 0u;
  // PASS = false: Z-pass (lane axis B, candidates along C; the column is carried by the magic constants, which step by
  // Dp), true: Y-pass (lane axis C, candidates along B; the column is carried by K0/K1, which step by 4 bytes).
  const float fu = PASS ? c.fz : c.fy, fv = PASS ? c.fy : c.fz;
  const float cp = f_add(fv, dbm), cm = f_sub(dbm, fv);
  const float mstep = (float)t.Dp;
  const f32x2_t mstep2 = pack2(mstep, -mstep), fu2 = pack2(fu, fu), one2 = pack2(1.0f, 1.0f);
  const f32x2_t hW2 = pack2(c.hW, c.hW), cpm2 = pack2(cp, cm), sfv2 = pack2(-fv, fv);
  unsigned K0[NC], K1[NC], sv = 0;
  f32x2_t a2[NC];
#pragma unroll
  for (int sidx = 0; sidx < NC; ++sidx) a2[sidx] = pack2(a4[sidx], a4[sidx]);
#pragma unroll
  for (int sidx = 0; sidx < NC; ++sidx) ring2_consts(c, t, PASS, -H, base + (unsigned)(sbase0 + sidx * slice_bytes), 4u, K0[sidx], K1[sidx], sv);
  const unsigned nsv = 0u - sv;
  float mu0, mu1;
  ring2_magic(PASS, -H, t.Dp, mu0, mu1);
  f32x2_t mu2 = pack2(mu0, mu1);
  f32x2_t uf2 = pack2((float)(-H), (float)(-H));
  // Columns are walked in blocks of 32; a column with an undecided candidate only sets its bit, and the exact path runs
  // after the block for all flagged (lane, column) pairs at once: the hot loop has one branch (its back edge) and the
  // lanes of a warp that need the exact path in the same block take it together.
#pragma unroll 1
  for (int ub0 = -H; ub0 <= H; ub0 += 32) {
    const int ue = min(ub0 + 31, H);
    unsigned flags = 0u, bit = 1u;
#pragma unroll 1
    for (int u = ub0; u <= ue; ++u, bit <<= 1) {
      const f32x2_t duf2 = sub2(uf2, fu2);           // (u - fu) in both halves; uf2 holds exact integers
      if (RCV_RING2_INTERIOR && (unsigned)(u + ioff) <= ispan)
        ring2_asm<false, NC, M>(flags, bit, duf2, mu2, hW2, cpm2, sfv2, c.hw_m, nhw_p, sink_abs, sv, nsv, 0.f, a2, K0, K1);
      else
        ring2_asm<true, NC, M>(flags, bit, duf2, mu2, hW2, cpm2, sfv2, c.hw_m, nhw_p, sink_abs, sv, nsv, ring2_thr(PASS, lo2(duf2)), a2, K0, K1);
      if (PASS) {
#pragma unroll
        for (int sidx = 0; sidx < NC; ++sidx) { K0[sidx] += 4u; K1[sidx] += 4u; }
      } else {
        mu2 = add2(mu2, mstep2);                     // (mu0 + Dp, mu1 - Dp): exact integers
      }
      uf2 = add2(uf2, one2);
    }
    if (__builtin_expect(flags != 0u, 0)) {
      do {
        const int j = __ffs((int)flags) - 1;
        flags &= flags - 1u;
        const int u = ub0 + j;
        const unsigned kback = PASS ? 4u * (unsigned)(ue + 1 - u) : 0u;   // K0/K1 now stand at column ue + 1
        float m0, m1;
        ring2_magic(PASS, u, t.Dp, m0, m1);
        ring2_slow_call(c.px, c.py, c.pz, c.R, PASS ? c.ipz : c.ipy, PASS ? c.ipy : c.ipz, u, i0c, (PASS ? 1 : 0) | (NC << 1) | (M << 4), c.hW,
                        c.hw_m, c.hw_p, f_sub((float)u, fu), cp, cm, fv, m0, m1, sv, a4[0], a4[NC > 1 ? 1 : 0], a4[NC > 2 ? 2 : 0], a4[NC > 3 ? 3 : 0],
                        K0[0] - kback, K0[NC > 1 ? 1 : 0] - kback, K0[NC > 2 ? 2 : 0] - kback, K0[NC > 3 ? 3 : 0] - kback, K1[0] - kback,
                        K1[NC > 1 ? 1 : 0] - kback, K1[NC > 2 ? 2 : 0] - kback, K1[NC > 3 ? 3 : 0] - kback);
      } while (flags);
    }
    __syncwarp();   // lanes that took the exact path rejoin here (measured: they otherwise walk the rest of the pass alone)
  }
}

__device__ __noinline__ void polar_slow_call(double pa, double pb, double pc, int R, float hw_m, int ti0, float q0, float q1, int vt0, int vt1,
                                             int t01, int jb, int kc, int cell, unsigned mplus, unsigned mminus, int sstride, unsigned base) {
  PolarOut o;
  o.q0 = q0; o.q1 = q1; o.vt0 = vt0; o.vt1 = vt1; o.t0 = (t01 & 1) != 0; o.t1 = (t01 & 2) != 0;
  SlowExactCall slow{pa, pb, pc, R};
  SmemEmit es{base, 0};
  polar_slow(hw_m, ti0, o, jb, kc, cell, mplus, mminus, sstride, slow, es);
}

// Polar pass of a warp: rows ub = -Hp..Hp around each lane's own point; in every row each lane walks the two column
// segments [-co,-ci] and [max(ci,1),co] of ITS OWN annulus (polar_row_range); the trip count is the warp maximum.
template <int SIDES>
__device__ __forceinline__ void polar_segment(const PointCtx& c, const Tile& t, float cpx, float cmx, int uc0, int ucl, int T, float db2, int jb,
                                              bool row_ok, int rowoff, unsigned mplus, unsigned mminus, int slice_bytes, const SmemEmit& emit) {
  float ucf = (float)uc0;
  int kc = c.ipz + uc0;
  const int kcl = c.ipz + ucl;
#pragma unroll 2
  for (int tt = 0; tt < T; ++tt, ucf += 1.0f, ++kc) {
    const float dc = f_sub(ucf, c.fz);
    const float s2 = f_fma(dc, dc, db2);
    const bool ok = row_ok & (kc <= kcl) & ((unsigned)kc < (unsigned)t.D);
    const int cell = rowoff + kc * 4;
    PolarOut o;
    polar_fast<SIDES>(c, t, cpx, cmx, s2, cell, ok, mplus, mminus, slice_bytes, emit.sink, emit, o);
    if (__builtin_expect(o.t0 || o.t1, 0))
      polar_slow_call(c.px, c.py, c.pz, c.R, c.hw_m, t.i0, o.q0, o.q1, o.vt0, o.vt1, (o.t0 ? 1 : 0) | (o.t1 ? 2 : 0), jb, kc, cell, mplus, mminus,
                      slice_bytes, emit.base);
  }
}

template <int SIDES>
__device__ __forceinline__ void polar_pass(const PointCtx& c, const Tile& t, int Hp, float s_lo, float s_hi, unsigned mplus, unsigned mminus,
                                           int slice_bytes, const SmemEmit& emit_c) {
  SmemEmit emit = emit_c;
  const bool lane_on = (mplus | mminus) != 0u;
  const float cpx = f_add(c.fx, c.dbias_m05), cmx = f_sub(c.dbias_m05, c.fx);
  float ubf = (float)(-Hp);
#pragma unroll 1
  for (int ub = -Hp; ub <= Hp; ++ub, ubf += 1.0f) {
    const float db = f_sub(ubf, c.fy);
    const float db2 = f_mul(db, db);
    int ci, co;
    polar_row_range(s_lo, s_hi, c.eps, db2, lane_on, ci, co);
    const int c1 = ci > 1 ? ci : 1;
    const int T0 = warp_max_i32(co >= 0 ? co - ci + 1 : 0), T1 = warp_max_i32(co >= 0 ? co - c1 + 1 : 0);
    if (T0 <= 0) continue;
    const int jb = c.ipy + ub;
    const bool row_ok = (co >= 0) & ((unsigned)(jb - t.j0) < (unsigned)t.nj);
    const int rowoff = (jb - t.j0) * t.Dp * 4;
    polar_segment<SIDES>(c, t, cpx, cmx, co >= 0 ? -co : 0, co >= 0 ? -ci : -1, T0, db2, jb, row_ok, rowoff, mplus, mminus, slice_bytes, emit);
    if (T1 > 0)
      polar_segment<SIDES>(c, t, cpx, cmx, co >= 0 ? c1 : 0, co >= 0 ? co : -1, T1, db2, jb, row_ok, rowoff, mplus, mminus, slice_bytes, emit);
  }
}

// ---- fast polar pass (raster_core.h "polar2"): one PTX block per PAIR of adjacent cells of a row, one side of the pole --
// Operands: %0 flags (in/out) | pairs: %1 (uc, uc+1) %2 (fz, fz) %3 (R^2 - dB^2) x2 %4 (hW, hW) %5 (cx, cx) %6 (sfx, sfx)
// %7 (nmidw, nmidw) | %8 nhalfw %9 nhi %10 hw_m %11 -hw_p %12 sink %13 smul %14 K (cell uc) %15 bit.
#define P2_CELL(R, Q, F, B, OFS, ACC) \
  "abs.f32 ar, " R ";\n\tsetp.le.f32 pw, ar, %8;\n\tabs.f32 aq, " Q ";\n\tsetp.lt.and.f32 ps, aq, %10, pw;\n\tsetp.le.and.f32 pv, " F ", %9, ps;\n\t" \
  "setp.gt.and.f32 pl, " Q ", %11, pw;\n\t" ACC "mad.lo.u32 adu, " B ", %13, %14;\n\t" OFS "selp.u32 adr, adu, %12, pv;\n\tred.shared.add.u32 [adr], 1;\n\t"
#define P2_BODY \
  "{\n\t.reg .pred pw, ps, pv, pl, pa;\n\t.reg .f32 glo, ghi, z0, z1, q0, q1, f0, f1, r0, r1, aq, ar;\n\t.reg .b32 b0, b1, adu, adr;\n\t" \
  ".reg .b64 dc2, nd2, g2, zs2, hwg2, t2, tm2, fl2, d2, q2, rr2, mg, z2;\n\t" \
  "mov.b64 mg, 0x4B4000004B400000;\n\tmov.b64 z2, 0;\n\tsub.rn.f32x2 dc2, %1, %2;\n\tsub.rn.f32x2 nd2, z2, dc2;\n\tfma.rn.f32x2 g2, nd2, dc2, %3;\n\t" \
  "mov.b64 {glo, ghi}, g2;\n\tsqrt.approx.ftz.f32 z0, glo;\n\tsqrt.approx.ftz.f32 z1, ghi;\n\tmov.b64 zs2, {z0, z1};\n\tsub.rn.f32x2 hwg2, %4, g2;\n\t" \
  "add.rn.f32x2 t2, zs2, %5;\n\tadd.rn.f32x2 tm2, t2, mg;\n\tsub.rn.f32x2 fl2, tm2, mg;\n\tadd.rn.f32x2 d2, fl2, %6;\n\tfma.rn.f32x2 q2, d2, d2, hwg2;\n\t" \
  "sub.rn.f32x2 rr2, fl2, %7;\n\tmov.b64 {q0, q1}, q2;\n\tmov.b64 {f0, f1}, fl2;\n\tmov.b64 {r0, r1}, rr2;\n\tmov.b64 {b0, b1}, tm2;\n\t" \
  P2_CELL("r0", "q0", "f0", "b0", "", "xor.pred pa, pl, ps;\n\t") \
  P2_CELL("r1", "q1", "f1", "b1", "add.u32 adu, adu, 4;\n\t", "xor.pred pl, pl, ps;\n\tor.pred pa, pa, pl;\n\t") "@pa or.b32 %0, %0, %15;\n\t}"

// The rare branch of the fast polar pass: exact decisions for the two cells of a flagged pair (one lane, out of line).
__device__ __noinline__ void polar2_slow_call(double pa, double pb, double pc, int R, int ipx, int jb, int kc0, int nlo, int nhi, int sgn, float fz,
                                              float hW, float hw_m, float hw_p, float cx, float sfx, float ucf0, float r2m, unsigned smul, unsigned K) {
  PointCtx c;
  c.px = pa; c.py = pb; c.pz = pc; c.R = R; c.ipx = ipx; c.fz = fz; c.hW = hW; c.hw_m = hw_m; c.hw_p = hw_p;
  Polar2Side S;
  S.cx = cx; S.sfx = sfx; S.nlo = nlo; S.nhi_i = nhi; S.sgn = sgn; S.nhi = (float)nhi;
  S.nmidw = f_mul((float)(nlo + nhi + 1), 0.5f);
  S.nhalfw = f_mul((float)(nhi + 1 - nlo), 0.5f);
  SlowExactCall slow{pa, pb, pc, R};
  SmemEmit es{0u, 0};   // addresses are absolute
  polar2_slow_cell(c, S, ucf0, r2m, jb, kc0, smul, K, slow, es);
  polar2_slow_cell(c, S, f_add(ucf0, 1.0f), r2m, jb, kc0 + 1, smul, K + 4u, slow, es);
}

// Fast polar pass of a warp that cannot leave the tile, both sides of the pole through ONE copy of the loops.
// Rows ub = -Hp..Hp around each lane's own point; in a row every lane walks the two column segments of ITS annulus,
// [..,-ci] and [max(ci,1),..], in pairs of cells; the trip count is the warp maximum and surplus cells lie OUTWARDS of the
// annulus, where the candidate's slice is never one of the lane's polar slices (so no per-cell validity test is needed).
__device__ __forceinline__ void polar_pass2(const PointCtx& c, const Tile& t, int Hp, float s_lo, float s_hi, bool lane_on, const Polar2Side& Sp,
                                            const Polar2Side& Sm, bool anyp, bool anym, int slice_bytes, unsigned base, unsigned sink_abs) {
  const f32x2_t hW2 = pack2(c.hW, c.hW), fz2 = pack2(c.fz, c.fz), two2 = pack2(2.0f, 2.0f);
  const float nhw_p = -c.hw_p;
  const unsigned MB = (unsigned)RCV_MAGIC_BITS, vrel0 = (unsigned)(c.ipx - t.i0);
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    if (side == 0 ? !anyp : !anym) continue;
    const Polar2Side& S = side == 0 ? Sp : Sm;
    const float cx = side == 0 ? Sp.cx : Sm.cx, sfx = side == 0 ? Sp.sfx : Sm.sfx, nmidw = side == 0 ? Sp.nmidw : Sm.nmidw;
    const float nhalfw = side == 0 ? Sp.nhalfw : Sm.nhalfw, nhi = side == 0 ? Sp.nhi : Sm.nhi;
    const int nlo_i = side == 0 ? Sp.nlo : Sm.nlo, nhi_i = side == 0 ? Sp.nhi_i : Sm.nhi_i;
    const f32x2_t cx2 = pack2(cx, cx), sfx2 = pack2(sfx, sfx), nmidw2 = pack2(nmidw, nmidw);
    const unsigned smul = side == 0 ? (unsigned)slice_bytes : 0u - (unsigned)slice_bytes;
    const unsigned Kside = base + (side == 0 ? vrel0 - MB : vrel0 + MB) * (unsigned)slice_bytes;
    (void)S;
    float ubf = (float)(-Hp);
#pragma unroll 1
    for (int ub = -Hp; ub <= Hp; ++ub, ubf += 1.0f) {
      const float db = f_sub(ubf, c.fy);
      const float db2 = f_mul(db, db);
      int ci, co;
      polar_row_range(s_lo, s_hi, c.eps, db2, lane_on, ci, co);
      const int c1 = ci > 1 ? ci : 1;
      const int T0 = warp_max_i32(co >= 0 ? co - ci + 1 : 0), T1 = warp_max_i32(co >= 0 ? co - c1 + 1 : 0);
      if (T0 <= 0) continue;
      const float r2m = f_sub(c.R2, db2);
      const f32x2_t r2m2 = pack2(r2m, r2m);
      const int jb = c.ipy + ub;
      const unsigned rowK = Kside + (unsigned)((jb - t.j0) * t.Dp) * 4u;
#pragma unroll 1
      for (int seg = 0; seg < 2; ++seg) {
        const int np = ((seg ? T1 : T0) + 1) >> 1;
        if (np <= 0) continue;
        const int start = seg == 0 ? (co >= 0 ? -ci : -1) - 2 * np + 1 : (co >= 0 ? c1 : 1);
        unsigned K = rowK + (unsigned)(c.ipz + start) * 4u;
        f32x2_t uc2 = pack2((float)start, (float)(start + 1));
        unsigned flags = 0u, bit = 1u;
#pragma unroll 1
        for (int p = 0; p < np; ++p, bit <<= 1, K += 8u) {
          asm volatile(P2_BODY : "+r"(flags) : "l"(uc2), "l"(fz2), "l"(r2m2), "l"(hW2), "l"(cx2), "l"(sfx2), "l"(nmidw2), "f"(nhalfw), "f"(nhi),
                       "f"(c.hw_m), "f"(nhw_p), "r"(sink_abs), "r"(smul), "r"(K), "r"(bit) : "memory");
          uc2 = add2(uc2, two2);
        }
        if (__builtin_expect(flags != 0u, 0)) {
          do {
            const int j = __ffs((int)flags) - 1;
            flags &= flags - 1u;
            const int uc = start + 2 * j;
            polar2_slow_call(c.px, c.py, c.pz, c.R, c.ipx, jb, c.ipz + uc, nlo_i, nhi_i, side == 0 ? 1 : -1, c.fz, c.hW, c.hw_m, c.hw_p, cx, sfx,
                             (float)uc, r2m, smul, K - 8u * (unsigned)(np - j));
          } while (flags);
        }
        __syncwarp();
      }
    }
  }
}

// Ring work of one warp for ONE chunk of NC slices of the slab (chunks tile the slab from its first slice): thin rings
// by the two ring passes, spheres too small for the polar pass by a dense scan.
template <int NC>
__device__ __forceinline__ void ring_chunk_work(const PointCtx& c, const Tile& t, int ia, int ib, int i0c, int slice_bytes, bool noclip,
                                                const SlowExactCall& slow_c, const SmemEmit& emit, const SlowArcCall& slowarc) {
  SlowExactCall slow = slow_c;
  const bool polar_lane = c.R >= RCV_POLAR_MIN_R;
  const int maxcode = noclip ? RCV_RING2_MAX_CODE : 1;
  float a4[NC];
  int abits = 0, mc = 1;
#pragma unroll
  for (int sidx = 0; sidx < NC; ++sidx) {
    const int i = i0c + sidx;
    float ar = 0.f; int code = 0;
    if (i >= ia && i <= ib) slice_setup(c, i, ar, code);
    const bool thin = code >= 1 && code <= maxcode;      // rings the ring passes draw; the rest belongs to the polar pass
    a4[sidx] = thin ? ar : __int_as_float(0x7fc00000);
    if (thin) { abits = max(abits, __float_as_int(ar)); mc = max(mc, code); }   // thin => a > 36 > 0: bit order = value order
    const int dl = polar_lane ? 0 : -code;
    const int dmax = warp_max_i32(dl);
    if (dmax > 0) {   // spheres too small for the polar pass (R < RCV_POLAR_MIN_R): bounding-box scan of the slice
      const int sbase = (i - t.i0) * slice_bytes;
#pragma unroll 1
      for (int rr = -dmax; rr <= dmax; ++rr)
#pragma unroll 1
        for (int kk = -dmax; kk <= dmax; ++kk) {
          const bool ok = dl > 0 && rr >= -dl && rr <= dl && kk >= -dl && kk <= dl;
          dense_cell(c, ar, t, i, sbase, 4, emit.sink, rr, kk, ok, emit, slow);
        }
    }
  }
  const int amax_bits = warp_max_i32(abits);
  if (amax_bits > 0) {
    const int H = ring_half_width(__int_as_float(amax_bits));
    const int sbase0 = (i0c - t.i0) * slice_bytes;
    if (noclip) {
      // interior half-width: warp minimum over the lanes that draw (a lane without thin rings does not constrain it)
      float amin = 3.0e38f;
#pragma unroll
      for (int sidx = 0; sidx < NC; ++sidx) amin = fminf(amin, a4[sidx]);   // fminf ignores the NaN of a non-thin slice
      const int hin_l = amin < 3.0e38f ? ring2_interior(f_sub(amin, c.W)) : 0x7fffffff;
      const int Hin = __reduce_min_sync(0xffffffffu, hin_l);
      const float dbm = amin < 3.0e38f ? ring2_dbias_m05(c, amin) : c.dbias_m05;   // floor bias of this lane for this chunk
      const unsigned sink_abs = emit.base + (unsigned)emit.sink;
      if (warp_max_i32(mc) == 1) {   // one candidate per arc
        ring_pass2<false, NC, 1>(c, t, a4, H, Hin, dbm, i0c, sbase0, slice_bytes, emit.base, sink_abs);
        ring_pass2<true, NC, 1>(c, t, a4, H, Hin, dbm, i0c, sbase0, slice_bytes, emit.base, sink_abs);
      } else {                       // some ring of the chunk crosses a column in two voxels: two candidates per arc
        ring_pass2<false, NC, 2>(c, t, a4, H, Hin, dbm, i0c, sbase0, slice_bytes, emit.base, sink_abs);
        ring_pass2<true, NC, 2>(c, t, a4, H, Hin, dbm, i0c, sbase0, slice_bytes, emit.base, sink_abs);
      }
    } else {
      ring_pass<true, NC>(c, t, a4, H, i0c, sbase0, slice_bytes, emit, slowarc);
    }
  }
}

// The lane's non-thin slices in the slab: masks (bit v = slice t.i0 + v, by side of the pole) and the annulus
// s_lo < dB^2 + dC^2 < s_hi that holds their rings.
__device__ __forceinline__ void polar_collect(const PointCtx& c, const Tile& t, int ia, int ib, int maxcode, unsigned& mplus, unsigned& mminus,
                                              float& s_lo, float& s_hi) {
  mplus = 0u; mminus = 0u; s_hi = 0.f; s_lo = 3.0e38f;
  if (c.R < RCV_POLAR_MIN_R) return;
#pragma unroll 1
  for (int i = ia; i <= ib; ++i) {
    float ar; int code;
    slice_setup(c, i, ar, code);
    if (code > maxcode || code < 0) {
      if (i > c.ipx) mplus |= 1u << (i - t.i0); else mminus |= 1u << (i - t.i0);
      s_hi = fmaxf(s_hi, ar);
      s_lo = fminf(s_lo, f_sub(ar, c.W));
    }
  }
}

__global__ void __launch_bounds__(kVoteThreads, 1) k_vote(VoteArgs a) {
  extern __shared__ __align__(16) int smem[];
  int* tile = smem;
  __shared__ int s_unit, s_next;
  __shared__ unsigned long long s_key[kVoteWarps], s_sum[kVoteWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
  asm volatile("mov.u32 %0, %0;" : "+r"(tile_s));   // keep the shared-window base in a register (ptxas otherwise rematerialises it per use)
  const int n_units = a.counters[0];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { s_unit = atomicAdd(&a.counters[1], 1); s_next = 0; }
    __syncthreads();
    const int ui = s_unit;
    if (ui >= n_units) break;
    const Unit u = a.units[ui];
    const ItemMeta& m = a.meta[u.item];
    const int D = m.D, Dp = m.Dp, n = m.n;
    const long long off = m.off;
    const int glo = (short)(m.guard & 0xffff), ghi = m.guard >> 16;     // signed: negative where the spheres stay inside the grid along C
    const Tile t{u.i0, u.ni, u.j0, u.nj, D, Dp, glo, ghi};
    // voxel (i, j, k) of the tile sits at word ((i - i0) * nj + (j - j0)) * Dp + k of the pointer shifted by the low guard
    const SmemEmit emit{tile_s + 4u * (unsigned)glo, 4 * (kTileWords + warp * 32 + lane - glo)};
    const SlowArcCall slowarc{tile_s + 4u * (unsigned)glo};
    const int words = u.ni * u.nj * Dp;
    {
      int4* t4 = reinterpret_cast<int4*>(tile);
      const int n4 = (words + 3) >> 2;
      for (int w = threadIdx.x; w < n4; w += kVoteThreads) t4[w] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    // ---- scatter: a work item is (phase, group of 32 consecutive points of the vote order), one point per lane.
    // Phase 0 draws the polar caps of the group's spheres inside the slab, phase 1 + c the thin rings of chunk c.
    const int slice_bytes = u.nj * Dp * 4;
    const int NC = ring_chunk(u.ni);
    const int ngroups = (n + 31) >> 5, nwork = ngroups * (1 + (u.ni + NC - 1) / NC);
    for (;;) {
      int w = 0;
      if (lane == 0) w = atomicAdd(&s_next, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= nwork) break;
      const int phase = w / ngroups, cur = (w - phase * ngroups) << 5;
      double pa = 0.0, pb = 0.0, pc = 0.0; int R = 0;
      if (cur + lane < n) {
        const long long q = off + a.pool.perm[off + cur + lane];
        pb = a.pool.X[q]; pa = a.pool.Y[q]; pc = a.pool.Z[q]; R = a.pool.Ri[q];
      }
      PointCtx c;
      point_setup(c, pa, pb, pc, R);
      int ia, ib;
      slice_range(c, t, ia, ib);
      if (phase == 0) {
        unsigned mplus, mminus;
        float s_lo, s_hi;
        const bool noclip = __all_sync(0xffffffffu, ring_noclip(c, t));   // same partition as the ring phases of this group
        polar_collect(c, t, ia, ib, noclip ? RCV_RING2_MAX_CODE : 1, mplus, mminus, s_lo, s_hi);
        const int Hp = warp_max_i32((mplus | mminus) ? polar_half_width(s_hi, c.eps) : -1);
        if (Hp >= 0) {
          const bool anyp = __any_sync(0xffffffffu, mplus != 0u), anym = __any_sync(0xffffffffu, mminus != 0u);
          Polar2Side Sp, Sm;
          const bool cont = polar2_side(c, t, mplus, true, Sp) & polar2_side(c, t, mminus, false, Sm);   // consecutive slices on each side
          if (noclip && __all_sync(0xffffffffu, cont))
            polar_pass2(c, t, Hp, s_lo, s_hi, (mplus | mminus) != 0u, Sp, Sm, anyp, anym, slice_bytes, emit.base, emit.base + (unsigned)emit.sink);
          else
            polar_pass<3>(c, t, Hp, s_lo, s_hi, mplus, mminus, slice_bytes, emit);   // clipped warps: the general polar pass
        }
      } else {
        const int i0c = t.i0 + (phase - 1) * NC;
        const bool here = ia <= ib && ib >= i0c && ia < i0c + NC;
        if (!__any_sync(0xffffffffu, here)) continue;
        const bool noclip = __all_sync(0xffffffffu, ring_noclip(c, t));
        SlowExactCall slow{pa, pb, pc, R};
        if (NC == 3) ring_chunk_work<3>(c, t, ia, ib, i0c, slice_bytes, noclip, slow, emit, slowarc);
        else if (NC == 4) ring_chunk_work<4>(c, t, ia, ib, i0c, slice_bytes, noclip, slow, emit, slowarc);
        else if (NC == 2) ring_chunk_work<2>(c, t, ia, ib, i0c, slice_bytes, noclip, slow, emit, slowarc);   // thin slabs of large grids
        else ring_chunk_work<1>(c, t, ia, ib, i0c, slice_bytes, noclip, slow, emit, slowarc);
      }
    }
    __syncthreads();
    // ---- peak of the tile (K3 fused) + vote tally + optional volume dump ----
    // tile row r = (A - i0) * nj + (B - j0) holds reference voxels (i, j, k) = (B, A, k)
    unsigned long long key = 0, sum = 0;
    const int jr0 = u.j0 < 0 ? 0 : u.j0, njr = min(u.j0 + u.nj, D) - jr0;   // the real rows of a slice (guard rows are not read back)
    const int rows = u.ni * njr;
    for (int r = warp; r < rows; r += kVoteWarps) {
      const int sa = r / njr, gb = jr0 + r - sa * njr, ga = u.i0 + sa;
      const unsigned lin0 = ((unsigned)gb * (unsigned)D + (unsigned)ga) * (unsigned)D;
      const int* row = tile + glo + (sa * u.nj + (gb - u.j0)) * Dp;
      for (int k = lane; k < D; k += 32) {
        const int v = row[k];
        const unsigned long long kk = pack_peak(v, lin0 + k);
        key = kk > key ? kk : key;
        sum += (unsigned)v;
        if (a.volume && (long long)D * D * D <= a.volume_cap) a.volume[(long long)lin0 + k] = v;
      }
    }
    key = warp_max_u64(key); sum = warp_sum_u64(sum);
    if (lane == 0) { s_key[warp] = key; s_sum[warp] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kVoteWarps; ++w) { key = s_key[w] > key ? s_key[w] : key; sum += s_sum[w]; }
      atomicMax(&a.best[u.item], key);
      if (sum) atomicAdd(&a.votes[u.item], sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2, second generation -- the run-length vote kernel (runs_core.h).
// Same persistent structure (one CTA per SM, (item, tile) units from a global queue, the tile never leaves shared
// memory), but the tile is a DIFFERENCE ARRAY along C (reference z): a lane (= one point) walks the columns
// u = -H..H of its sphere and records, per column and slice, the two runs of shell voxels as marks at their ends (a
// plane of run starts, a plane of run ends) -- four atomics per (column, slice) whatever the run lengths; when every
// point has been drawn, one prefix sum per row of (starts - ends) gives the counts, fused with the peak search and the
// vote tally.
// ------------------------------------------------------------------------------------------------
#ifndef RCV_RUNS_THREADS
#define RCV_RUNS_THREADS 640
#endif
#ifndef RCV_RUNS_SMALL_FRAC
#define RCV_RUNS_SMALL_FRAC 0.36f   // a work item is "short" if its chord is below 0.6 of the item's widest
#endif
#ifndef RCV_RUNS_CTAS
#define RCV_RUNS_CTAS 1            // resident CTAs per SM (each gets 1/RCV_RUNS_CTAS of the shared memory)
#endif
constexpr int kRunsThreads = RCV_RUNS_THREADS;
constexpr int kRunsWarps = kRunsThreads / 32;
constexpr int kRunsWorkList = 1024;
constexpr int kRunsSmemBytes = RCV_RUNS_CTAS == 1 ? kSmemBytes : (228 * 1024 / RCV_RUNS_CTAS - 1024) & ~15;   // 228 KB per SM, 1 KB reserved per CTA
#ifndef RCV_RUNS_QUEUE
#define RCV_RUNS_QUEUE 40             // deferred exact decisions a warp can hold (8 bytes each); drained 32 at a time
#endif
constexpr int kRunsTileWords = kRunsSmemBytes / 4 - 192 - kRunsWorkList / 2 - 2 * RCV_RUNS_QUEUE * (RCV_RUNS_THREADS / 32);    // static shared variables: 192 words + the per-warp queues of deferred exact decisions
// The two planes of a tile sit at a fixed distance (half the tile memory), so the address of an end mark is the start
// plane's address plus an immediate.
constexpr int kRunsPlaneWords = (kRunsTileWords / 2) & ~3;
constexpr unsigned kRunsPlaneBytes = (unsigned)kRunsPlaneWords * 4u;
__device__ __forceinline__ void smem_inc_e(unsigned addr) { asm volatile("red.shared.add.u32 [%0+%1], 1;" ::"r"(addr), "n"(kRunsPlaneBytes) : "memory"); }

// Only `add 1` (ATOMS.POPC.INC) merges the lanes of a warp that hit the same address; every other shared-memory atomic
// (add of a register or of -1, inc, dec) replays once per duplicate lane (tools/ubench_atoms2.cu,
// profiles/r02c_ubench_atoms2.jsonl: 282 G warp-instr/s against 9 G with 32 equal addresses).  Adjacent pixels share
// most of their run boundaries, so the difference array is kept as two planes of plain counters -- run starts and run
// ends -- and every mark is an `add 1`.

// The rare path of one (lane, column): exact decisions for the voxels next to flagged run boundaries, for every slice
// of the chunk.  Everything is re-derived from the point record (the same arithmetic as the fast path); the float64
// coordinates are fetched here.  (A,B,C) = reference (y,x,z): exact_hit() wants the reference order.
struct RunSlowCtx {
  const double *X, *Y, *Z;   // pool (item base already added)
  const int* perm;           // vote order -> pool index (item base already added)
};
template <bool CLIP>
__device__ __noinline__ void runs_slow_column(RunPoint c, int pidx, RunSlowCtx sc, int u, int uc, int i0c, int nsl, int Dp, unsigned K0, unsigned slice_bytes,
                                              int clo, int chi) {
  RunLane L;
  run_lane_setup(c, L);
  RunCol C;
  run_col_setup(c, u, uc, Dp, C);
  const unsigned base = (unsigned)RCV_MAGIC_BITS + (unsigned)(uc * Dp);
  const int q = sc.perm[pidx];
  const double px = sc.X[q], py = sc.Y[q], pz = sc.Z[q];
  const int row = c.ipb + u;
  for (int s = 0; s < nsl; ++s) {
    const int i = i0c + s;
    const f2 aa = run_slice_consts(c, L, i, true);
    const unsigned K = K0 + (unsigned)s * slice_bytes;
    auto exact = [&](int m) { return exact_hit_call(px, py, pz, c.R, row, i, c.ipc + m); };
    auto fix = [&](unsigned nbits, int delta) {   // voxel m joins (delta = +1) or leaves (-1) the set: a run [m, m+1) added to / taken from the planes
      int b0 = (int)nbits, b1 = (int)nbits + 1;
      if (CLIP) { b0 = min(max(b0, (int)base + clo), (int)base + chi); b1 = min(max(b1, (int)base + clo), (int)base + chi); }
      smem_inc((unsigned)b0 * 4u + K + (delta > 0 ? 0u : kRunsPlaneBytes));
      smem_inc((unsigned)b1 * 4u + K + (delta > 0 ? kRunsPlaneBytes : 0u));
    };
    run_slow_slice(L, C, aa, base, exact, fix);
  }
}

// ---- deferred exact decisions -----------------------------------------------------------------------------------
// A flagged (point, column) is not decided where it is found -- one lane would walk the float64 path while 31 wait --
// but queued per warp; the queue is drained 32 items at a time (every lane takes one) and at the end of the tile.  An item
// carries what the rare path needs to re-derive the column: the point's index in vote order, the column, the chunk.
constexpr int kRunsQueue = RCV_RUNS_QUEUE;
struct RunTile {               // per-tile constants of the rare path
  Tile t;
  unsigned tile_s, slice_bytes;
  const int4* rec;             // point records of the item (item base already added)
  RunSlowCtx sc;
  bool clip;
};
__device__ __forceinline__ unsigned run_K0(const RunTile& rt, const RunPoint& c, int i0c) {
  return rt.tile_s + 4u * (unsigned)(rt.t.glo + ((i0c - rt.t.i0) * rt.t.nj + (c.ipb - rt.t.j0)) * rt.t.Dp + c.ipc - RCV_MAGIC_BITS);
}
__device__ __forceinline__ RunPoint run_load_point(const int4* rec, int pidx) {
  const int4 r0 = __ldg(rec + 2 * pidx), r1 = __ldg(rec + 2 * pidx + 1);
  RunPoint c;
  c.ipa = r0.x; c.ipb = r0.y; c.ipc = r0.z; c.R = r0.w;
  c.fa = __int_as_float(r1.x); c.fb = __int_as_float(r1.y); c.fc = __int_as_float(r1.z); c.W = __int_as_float(r1.w);
  return c;
}
__device__ __noinline__ void runs_slow_item(const RunTile& rt, unsigned long long item) {
  const int pidx = (int)(unsigned)(item & 0xffffffffu);
  const unsigned hi = (unsigned)(item >> 32);
  const int u = (int)(hi & 0xfffu) - 2048, nsl = (int)((hi >> 12) & 0xfu), i0c = (int)(hi >> 16);
  const RunPoint c = run_load_point(rt.rec, pidx);
  const int clo = -rt.t.glo - c.ipc, chi = rt.t.D + rt.t.ghi - c.ipc;
  if (rt.clip) runs_slow_column<true>(c, pidx, rt.sc, u, u, i0c, nsl, rt.t.Dp, run_K0(rt, c, i0c), rt.slice_bytes, clo, chi);
  else runs_slow_column<false>(c, pidx, rt.sc, u, u, i0c, nsl, rt.t.Dp, run_K0(rt, c, i0c), rt.slice_bytes, clo, chi);
}
// Drains the warp's queue: whole batches of 32 items, or everything when `all`.  Called by all 32 lanes.
__device__ __forceinline__ void runs_drain(const RunTile& rt, unsigned long long* q, int* qn, bool all) {
  __syncwarp();
  int n = min(*qn, kRunsQueue);
  __syncwarp();                       // every lane has read the count before lane 0 may write it back
  const int lane = threadIdx.x & 31;
  while (n >= 32 || (all && n > 0)) {
    const int take = min(n, 32);
    if (lane < take) runs_slow_item(rt, q[n - take + lane]);
    n -= take;
    __syncwarp();
  }
  if (lane == 0) *qn = n;
  __syncwarp();
}
// Pushes the flagged columns of a block of 32 (bit j of `flags` = column u0 + j) of this lane's point.
__device__ __forceinline__ void runs_push(const RunTile& rt, unsigned long long* q, int* qn, unsigned flags, int pidx, int u0, int i0c, int nsl) {
  while (flags) {
    const int j = __ffs((int)flags) - 1;
    flags &= flags - 1u;
    const unsigned long long item = (unsigned long long)(unsigned)pidx |
                                    ((unsigned long long)((unsigned)(u0 + j + 2048) | ((unsigned)nsl << 12) | ((unsigned)i0c << 16)) << 32);
    const int slot = atomicAdd(qn, 1);
    if (slot < kRunsQueue) q[slot] = item;
    else runs_slow_item(rt, item);          // queue full (a pathological block): decide at once
  }
}

// One column of the chunk for this lane: the four marks per slice; returns 1 if a boundary of the column is flagged.
template <int NC, bool CLIP>
__device__ __forceinline__ unsigned runs_one_column(const RunLane& L, const RunCol& C, const f2 (&aa)[NC], const unsigned (&K)[NC], int lo, int hi) {
  unsigned amb = 0u;
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    RunOut o;
    run_slice(L, C, aa[s], o);
    int b1 = (int)o.b1, b2 = (int)o.b2, b3 = (int)o.b3, b4 = (int)o.b4;
    if (CLIP) { b1 = min(max(b1, lo), hi); b2 = min(max(b2, lo), hi); b3 = min(max(b3, lo), hi); b4 = min(max(b4, lo), hi); }
    smem_inc((unsigned)b1 * 4u + K[s]);
    smem_inc_e((unsigned)b2 * 4u + K[s]);
    smem_inc((unsigned)b3 * 4u + K[s]);
    smem_inc_e((unsigned)b4 * 4u + K[s]);
    run_flag_acc(amb, o);
  }
  return amb;
}

// A stretch [ua, ub] of columns of one chunk of NC slices, 32 lanes = 32 points.  PARK: some lane's own column range
// [ulo, uhi] does not cover the stretch; beyond its range a lane parks on its edge row and the column is empty.
// Without PARK (every lane draws every column of the stretch) columns go two at a time when NC <= 2: twice the
// independent dependency chains per warp and half the loop overhead.
//   K0 = shared-window address constant of the chunk's first slice for this lane (run_K0): the cell of boundary bits b
//        (= MAGIC_BITS + uc * Dp + lattice offset) is  b * 4 + K0 + s * slice_bytes.
template <int NC, bool CLIP, bool PARK>
__device__ __forceinline__ void runs_columns(const RunPoint& c, const RunLane& L, const RunTile& rt, const f2 (&aa)[NC], const unsigned (&K)[NC],
                                             int ua, int ub, int ulo, int uhi, int pidx, int i0c, int nsl,
                                             unsigned long long* q, int* qn) {
  const Tile& t = rt.t;
  const int clo = -t.glo - c.ipc, chi = t.D + t.ghi - c.ipc;   // clip bounds of the boundary bits relative to `base` (CLIP)
  const float Dpf = (float)t.Dp;
#ifdef RCV_RUNS_NO_PAIRS
  constexpr bool PAIRS = false;
#else
  constexpr bool PAIRS = !PARK && !CLIP && NC <= 2;
#endif
#pragma unroll 1
  for (int u0 = ua; u0 <= ub; u0 += 32) {
    unsigned flags = 0;
    const int ue = min(u0 + 31, ub);
    float uf = (float)u0;
    unsigned bit = 1u;
    int u = u0;
    if (PAIRS) {
#pragma unroll 1
      for (; u < ue; u += 2, uf += 2.0f, bit <<= 2) {
        RunCol C0, C1;
        const float du0 = f_sub(uf, c.fb), uf1 = f_add(uf, 1.0f), du1 = f_sub(uf1, c.fb);
        C0.mu = f2_dup(f_fma(uf, Dpf, RCV_MAGIC)); C0.du = f2_dup(du0); C0.ndu = f2_dup(-du0);
        C1.mu = f2_dup(f_fma(uf1, Dpf, RCV_MAGIC)); C1.du = f2_dup(du1); C1.ndu = f2_dup(-du1);
        const unsigned a0 = runs_one_column<NC, false>(L, C0, aa, K, 0, 0);
        const unsigned a1 = runs_one_column<NC, false>(L, C1, aa, K, 0, 0);
        if (a0) flags |= bit;
        if (a1) flags |= bit << 1;
      }
    }
#pragma unroll 1
    for (; u <= ue; ++u, uf += 1.0f, bit <<= 1) {
      float du = f_sub(uf, c.fb);
      RunCol C;
      int uc = u;
      if (PARK) {
        uc = min(max(u, ulo), uhi);                              // beyond its own range the lane parks on its edge row
        if (uc != u) du = 1.0e18f;                               // ... and the parked column is empty (g'' = -huge)
        C.mu = f2_dup(f_fma((float)uc, Dpf, RCV_MAGIC));          // exact: integers below 2^24
      } else {
        C.mu = f2_dup(f_fma(uf, Dpf, RCV_MAGIC));
      }
      C.du = f2_dup(du); C.ndu = f2_dup(-du);
      const int base = RCV_MAGIC_BITS + uc * t.Dp;
      if (runs_one_column<NC, CLIP>(L, C, aa, K, base + clo, base + chi)) flags |= bit;
    }
    if (__any_sync(0xffffffffu, flags != 0u)) {
      runs_push(rt, q, qn, flags, pidx, u0, i0c, nsl);            // (a flagged column is never a parked one)
      runs_drain(rt, q, qn, false);
    }
  }
}

// One chunk of NC slices for a warp whose 32 lanes each own one point.
template <int NC, bool CLIP>
__device__ __forceinline__ void runs_chunk(const RunPoint& c, const RunLane& L, const RunTile& rt, int i0c, int nsl, int pidx, unsigned long long* q, int* qn) {
  const Tile& t = rt.t;
  f2 aa[NC];
  float amax = -1.f;
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    aa[s] = run_slice_consts(c, L, i0c + s, s < nsl);
    amax = fmaxf(amax, f2_lo(aa[s]));
  }
  const int Hl = run_half_width(amax);
  // the lane's own column range [ulo, uhi], inside the tile's rows (a no-op for whole-slice tiles, whose guard rows hold every
  // sphere; the restriction that makes a row band a band); empty: ulo > uhi
  int ulo = max(-Hl, t.j0 - c.ipb), uhi = min(Hl, t.j0 + t.nj - 1 - c.ipb);
  if (Hl < 0) { ulo = 1; uhi = 0; }
  const bool some = ulo <= uhi;
  int wlo = __reduce_min_sync(0xffffffffu, some ? ulo : 0x7fffffff), whi = __reduce_max_sync(0xffffffffu, some ? uhi : -0x7fffffff);
  if (wlo > whi) return;
  // the stretch every lane covers with its own range: there the column body needs no parking logic
  int mlo = __reduce_max_sync(0xffffffffu, some ? ulo : 0x7fffffff), mhi = __reduce_min_sync(0xffffffffu, some ? uhi : -0x7fffffff);
  mlo = max(mlo, wlo); mhi = min(mhi, whi);
  // a lane without columns parks on a row that is certainly inside the tile
  const int upark = min(max(0, t.j0 - c.ipb), t.j0 + t.nj - 1 - c.ipb);
  if (!some) { ulo = upark; uhi = upark; }
  // starts go to plane 0, ends to plane 1; a slice of the chunk beyond the tile (its columns are all empty) re-uses the
  // last real slice's cells, where its four marks cancel
  const unsigned K0 = run_K0(rt, c, i0c);
  unsigned K[NC];
#pragma unroll
  for (int s = 0; s < NC; ++s) K[s] = K0 + (unsigned)min(s, nsl - 1) * rt.slice_bytes;
  if (CLIP || mlo > mhi) {
    runs_columns<NC, CLIP, true>(c, L, rt, aa, K, wlo, whi, ulo, uhi, pidx, i0c, nsl, q, qn);
  } else {
    if (wlo < mlo) runs_columns<NC, CLIP, true>(c, L, rt, aa, K, wlo, mlo - 1, ulo, uhi, pidx, i0c, nsl, q, qn);
    runs_columns<NC, CLIP, false>(c, L, rt, aa, K, mlo, mhi, ulo, uhi, pidx, i0c, nsl, q, qn);
    if (mhi < whi) runs_columns<NC, CLIP, true>(c, L, rt, aa, K, mhi + 1, whi, ulo, uhi, pidx, i0c, nsl, q, qn);
  }
}

__global__ void __launch_bounds__(kRunsThreads, RCV_RUNS_CTAS) k_vote_runs(VoteArgs a) {
  extern __shared__ __align__(16) int smem[];
  int* tile = smem;
  __shared__ int s_unit, s_next;
  __shared__ unsigned long long s_key[kRunsWarps], s_sum[kRunsWarps];
  __shared__ unsigned long long s_q[kRunsWarps][kRunsQueue];
  __shared__ int s_qn[kRunsWarps];
  __shared__ unsigned short s_work[kRunsWorkList];   // the tile's work items (chunk * ngroups + group) whose spheres reach the chunk
  __shared__ int s_nwork, s_nsmall;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
  asm volatile("mov.u32 %0, %0;" : "+r"(tile_s));
  const int n_units = a.counters[0];
  if (lane == 0) s_qn[warp] = 0;
  for (;;) {
    __syncthreads();                     // the previous tile is finished
    if (threadIdx.x == 0) { s_unit = atomicAdd(&a.counters[1], 1); s_next = 0; s_nwork = 0; s_nsmall = 0; }
    __syncthreads();
    const int ui = s_unit;
    if (ui >= n_units) break;
    const Unit u = a.units[ui];
    const ItemMeta& m = a.meta[u.item];
    const int D = m.D, Dp = m.Dp, n = m.n;
    const long long off = m.off;
    const int glo = (short)(m.guard & 0xffff), ghi = m.guard >> 16;     // signed: negative where the spheres stay inside the grid along C
    const Tile t{u.i0, u.ni, u.j0, u.nj, D, Dp, glo, ghi};
    // whole-slice tiles whose guard band holds every sphere of the item need no clipping (the prelude sized the guards);
    // row bands and unguarded grids (guard = 0 although spheres overhang) take the clipped variant
    const bool clip = m.clip != 0;
    const int plane_words = (u.ni * u.nj * Dp + 3) & ~3;   // plane 0: run starts at word 0, plane 1: run ends at word kRunsPlaneWords
    {
      int4* t4 = reinterpret_cast<int4*>(tile);
      int4* e4 = reinterpret_cast<int4*>(tile + kRunsPlaneWords);
      const int n4 = plane_words >> 2;
      for (int w = threadIdx.x; w < n4; w += kRunsThreads) { t4[w] = make_int4(0, 0, 0, 0); e4[w] = make_int4(0, 0, 0, 0); }
    }
    // ---- scatter: a work item is (chunk of NC slices, group of 32 consecutive points of the vote order) ----
    const unsigned slice_bytes = (unsigned)(u.nj * Dp * 4);
#ifdef RCV_RUNS_MAX_NC
    const int NC = min(ring_chunk(u.ni), RCV_RUNS_MAX_NC);
#else
    const int NC = ring_chunk(u.ni);
#endif
    const int ngroups = (n + 31) >> 5, nchunks = (u.ni + NC - 1) / NC, nall = ngroups * nchunks;
    // Work list: only the (chunk, group) pairs whose spheres reach the chunk's slices (group summaries from the prelude), so
    // that no warp fetches the records of a group just to find that it has nothing to draw.  An entry is chunk << 11 | group.
    // Long items (chunks near the groups' centre slices: many columns) fill the list from the front, short ones (near the
    // poles of the spheres along A) from the back, and the warps pull front first: the items still running when the first
    // warp reaches the end-of-tile barrier are short ones.  Too many pairs for the list (huge items): every pair is a work
    // item and the reach test is made on the records.
    const bool listed = ngroups <= 2048 && nchunks <= 32;
    if (listed) {
      const int* grp = a.pool.grp + (off >> 5) + u.item;
      const float hm = 2.0f * (float)m.rmax + 2.0f, small_below = RCV_RUNS_SMALL_FRAC * hm * hm;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int i0c = u.i0 + ch * NC, i1c = min(i0c + NC, u.i0 + u.ni) - 1;
        for (int g = threadIdx.x; g < ngroups; g += kRunsThreads) {
          const int r = __ldg(grp + g);
          const int lo = (short)(r & 0xffff), hi = r >> 16;
          if (hi >= i0c && lo <= i1c) {
            // (twice the) chord of the group's largest sphere at the chunk's middle slice, squared
            const int e = i0c + i1c - (hi + lo), c4 = (hi - lo) * (hi - lo) - e * e;
            const bool small = (float)c4 < small_below;
            const int slot = atomicAdd(small ? &s_nsmall : &s_nwork, 1);
            if (slot < kRunsWorkList) s_work[small ? kRunsWorkList - 1 - slot : slot] = (unsigned short)(ch << 11 | g);
          }
        }
      }
    }
    __syncthreads();
    const int nbig = s_nwork;
    const bool use_list = listed && nbig + s_nsmall <= kRunsWorkList;
    // (Measured and dropped, r02q/r02r: splitting the tile's last work items into column halves, fetching the next tile's
    // unit index during the scatter, and requesting the next item's records one item ahead -- each 0.5-4% slower.)
    const int nwork = use_list ? nbig + s_nsmall : nall;
    const RunSlowCtx sc{a.pool.X + off, a.pool.Y + off, a.pool.Z + off, a.pool.perm + off};
    const RunTile rt{t, tile_s, slice_bytes, a.pool.rec + 2 * off, sc, clip};
    unsigned long long* q = s_q[warp];
    int* qn = &s_qn[warp];
    auto pull = [&](int& ch, int& g, int4& r0, int4& r1) {
      int k = 0;
      if (lane == 0) k = atomicAdd(&s_next, 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      ch = -1; g = 0;
      r0 = make_int4(0, 0, glo < 0 ? -glo : 0, 0); r1 = make_int4(0, 0, 0, 0);     // padding lanes: R = 0, at a cell the rows store
      if (k >= nwork) return;
      if (use_list) { const int e = (int)s_work[k < nbig ? k : kRunsWorkList - 1 - (k - nbig)]; ch = e >> 11; g = e & 2047; }
      else { ch = k / ngroups; g = k - ch * ngroups; }
      const int pi = (g << 5) + lane;
      if (pi < n) { r0 = __ldg(rt.rec + 2 * pi); r1 = __ldg(rt.rec + 2 * pi + 1); }
    };
    for (;;) {
      int ch, g; int4 n0, n1;
      pull(ch, g, n0, n1);
      if (ch < 0) break;
      RunPoint c;
      c.ipa = n0.x; c.ipb = n0.y; c.ipc = n0.z; c.R = n0.w;
      c.fa = __int_as_float(n1.x); c.fb = __int_as_float(n1.y); c.fc = __int_as_float(n1.z); c.W = __int_as_float(n1.w);
      const int cur = g << 5;
      const int i0c = u.i0 + ch * NC, nsl = min(NC, u.i0 + u.ni - i0c);
      // does any sphere of the group reach the chunk's slices?
      const bool here = c.R > 0 && (c.ipa + c.R + 1 >= i0c) && (c.ipa - c.R - 1 <= i0c + nsl - 1);
      if (!__any_sync(0xffffffffu, here)) continue;
      if (!here) c.R = 0;     // (a lane that draws nothing leaves cancelling marks at its own cell: inside the window for a real point,
                              // and the prelude / the padding below give the others a stored cell)
      RunLane L;
      run_lane_setup(c, L);
      const int pidx = cur + lane;
      if (!clip) {
#if !defined(RCV_RUNS_MAX_NC) || RCV_RUNS_MAX_NC >= 3
        if (NC == 4) runs_chunk<4, false>(c, L, rt, i0c, nsl, pidx, q, qn);
        else if (NC == 3) runs_chunk<3, false>(c, L, rt, i0c, nsl, pidx, q, qn);
        else
#endif
        if (NC == 2) runs_chunk<2, false>(c, L, rt, i0c, nsl, pidx, q, qn);
        else runs_chunk<1, false>(c, L, rt, i0c, nsl, pidx, q, qn);
      } else {
        if (NC >= 3) { runs_chunk<2, true>(c, L, rt, i0c, min(nsl, 2), pidx, q, qn); if (nsl > 2) runs_chunk<2, true>(c, L, rt, i0c + 2, nsl - 2, pidx, q, qn); }
        else if (NC == 2) runs_chunk<2, true>(c, L, rt, i0c, nsl, pidx, q, qn);
        else runs_chunk<1, true>(c, L, rt, i0c, nsl, pidx, q, qn);
      }
    }
    runs_drain(rt, q, qn, true);     // the exact decisions still queued belong to this tile
    __syncthreads();
    // ---- prefix sum per row (differences -> counts) fused with the peak of the tile (K3) and the vote tally ----
    // tile row r = (A - i0) * nj + (B - j0) holds reference voxels (i, j, k) = (B, A, k); cell k sits at word glo + k
    unsigned long long key = 0, sum = 0;
    const int jr0 = u.j0 < 0 ? 0 : u.j0, njr = min(u.j0 + u.nj, D) - jr0;   // the real rows of a slice (guard rows are not read back)
    const int rows = u.ni * njr;
    const bool dump = a.volume && (long long)D * D * D <= a.volume_cap;
    const int klo = glo < 0 ? -glo : 0, khi = ghi < 0 ? D - 1 + ghi : D - 1;     // the grid's cells that the tile stores
    for (int r = threadIdx.x; r < rows; r += kRunsThreads) {
      const int sa = r / njr, gb = jr0 + r - sa * njr, ga = u.i0 + sa;
      const unsigned lin0 = ((unsigned)gb * (unsigned)D + (unsigned)ga) * (unsigned)D;
      int* row = tile + (sa * u.nj + (gb - u.j0)) * Dp;
      const int* rowe = row + kRunsPlaneWords;
      int run = 0;
      for (int k = 0; k < glo; ++k) run += row[k] - rowe[k];     // guard cells below the grid
      int best = 0, bestk = 0;                                   // (cells outside the window hold no vote: count 0 at k = 0 is the row's floor)
      row += glo; rowe += glo;                                   // cell k of the row is row[k], for k in [klo, khi]
      for (int k = klo; k <= khi; ++k) {
        run += row[k] - rowe[k];
        sum += (unsigned)run;
        if (run > best) { best = run; bestk = k; }
        if (dump) row[k] = run;
      }
      const unsigned long long kk = pack_peak(best, lin0 + (unsigned)bestk);
      key = kk > key ? kk : key;
    }
    if (dump) {   // parity / debug: the counts go to HBM, coalesced (the volume was zeroed by the host: cells outside the window stay 0)
      __syncthreads();
      for (int r = warp; r < rows; r += kRunsWarps) {
        const int sa = r / njr, gb = jr0 + r - sa * njr, ga = u.i0 + sa;
        const unsigned lin0 = ((unsigned)gb * (unsigned)D + (unsigned)ga) * (unsigned)D;
        const int* row = tile + glo + (sa * u.nj + (gb - u.j0)) * Dp;
        for (int k = klo + lane; k <= khi; k += 32) a.volume[(long long)lin0 + k] = row[k];
      }
    }
    key = warp_max_u64(key); sum = warp_sum_u64(sum);
    if (lane == 0) { s_key[warp] = key; s_sum[warp] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kRunsWarps; ++w) { key = s_key[w] > key ? s_key[w] : key; sum += s_sum[w]; }
      atomicMax(&a.best[u.item], key);
      if (sum) atomicAdd(&a.votes[u.item], sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// F -- peak -> millimetres (AccumulatorSpace.py:409-415); scatter of per-item outputs
// ------------------------------------------------------------------------------------------------
struct FinalArgs {
  const ItemMeta* meta; const unsigned long long* best; const unsigned long long* votes; int n_items;
  double acc_unit; int policy; int want_volume; long long volume_cap;
  double* centre_mm; int* peak; long long* votes_out; int* n_points; int* grid; int* zb; int* status;
};

__global__ void k_finalize(FinalArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.n_items) return;
  const ItemMeta m = a.meta[b];
  int st = m.status;
  double c[3] = {0.0, 0.0, 0.0};
  int pk = 0;
  long long nv = 0;
  if (st == RCV_ST_OK) {
    const unsigned long long key = a.best[b];
    pk = (int)(key >> 32);
    nv = (long long)a.votes[b];
    // no vote at all: argwhere(V == 0)[0] is voxel (0, 0, 0), which a tile need not cover (tiles span the box the spheres can reach)
    const unsigned lin = nv == 0 ? 0u : 0xffffffffu - (unsigned)(key & 0xffffffffu);
    if (nv == 0) pk = 0;
    const unsigned D = (unsigned)m.D;
    const int idx[3] = {(int)(lin / (D * D)), (int)((lin / D) % D), (int)(lin % D)};
    for (int q = 0; q < 3; ++q) {
      double v = (double)idx[q];
      if (m.zb < 0) v = __dadd_rn(v, (double)m.zb);
      if (a.policy == RCV_POLICY_YCBGEN) c[q] = __dadd_rn(__dmul_rn(__dadd_rn(v, m.mean[q]), a.acc_unit), 0.5);
      else c[q] = __dmul_rn(__dadd_rn(__dadd_rn(v, m.mean[q]), 0.5), a.acc_unit);
    }
    nv = (long long)a.votes[b];
    if (a.want_volume && (long long)m.D * m.D * m.D > a.volume_cap) st |= kStVolumeSkipped;
  }
  a.centre_mm[3 * b] = c[0]; a.centre_mm[3 * b + 1] = c[1]; a.centre_mm[3 * b + 2] = c[2];
  if (a.peak) a.peak[b] = pk;
  if (a.votes_out) a.votes_out[b] = nv;
  if (a.n_points) a.n_points[b] = m.n;
  if (a.grid) a.grid[b] = m.D;
  if (a.zb) a.zb[b] = m.zb;
  if (a.status) a.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// K3 -- standalone peak search over an HBM volume (AccumulatorSpace.py:406)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_argmax_volume(const int32_t* __restrict__ vol, long long total, unsigned long long* best) {
  unsigned long long key = 0;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; q < total; q += stride) {
    if (q + 3 < total && ((reinterpret_cast<uintptr_t>(vol + q) & 15) == 0)) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(vol + q));
      unsigned long long k0 = pack_peak(v.x, (unsigned)q), k1 = pack_peak(v.y, (unsigned)q + 1), k2 = pack_peak(v.z, (unsigned)q + 2),
                         k3 = pack_peak(v.w, (unsigned)q + 3);
      k0 = k1 > k0 ? k1 : k0; k2 = k3 > k2 ? k3 : k2; k0 = k2 > k0 ? k2 : k0;
      key = k0 > key ? k0 : key;
    } else {
      for (int e = 0; e < 4 && q + e < total; ++e) { const unsigned long long kk = pack_peak(vol[q + e], (unsigned)(q + e)); key = kk > key ? kk : key; }
    }
  }
  key = warp_max_u64(key);
  __shared__ unsigned long long s_k[8];
  if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) key = s_k[w] > key ? s_k[w] : key;
    atomicMax(best, key);
  }
}

__global__ void k_argmax_unpack(const unsigned long long* best, int D, int* idx_out, int* max_out) {
  const unsigned long long key = *best;
  const unsigned lin = 0xffffffffu - (unsigned)(key & 0xffffffffu);
  idx_out[0] = (int)(lin / ((unsigned)D * D)); idx_out[1] = (int)((lin / D) % D); idx_out[2] = (int)(lin % D);
  *max_out = (int)(key >> 32);
}

// ------------------------------------------------------------------------------------------------
// K4 -- batched Horn absolute orientation (util/horn.py:75-181), one thread per frame.
// Quaternion method: eigenvector of the largest eigenvalue of the symmetric 4x4 N built from the
// cross-covariance sums; cyclic Jacobi with the reference's sweep order, thresholds and 50-sweep cap.
// ------------------------------------------------------------------------------------------------
__global__ void k_horn(const double* __restrict__ model, long long model_stride, const double* __restrict__ est, int n, int n_frames,
                       double* __restrict__ RT) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const double* P1 = model + (long long)f * model_stride;
  const double* P2 = est + (long long)f * 3 * n;
  double C1[3] = {0, 0, 0}, C2[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) { C1[j] += P1[3 * i + j]; C2[j] += P2[3 * i + j]; }
  for (int j = 0; j < 3; ++j) { C1[j] /= n; C2[j] /= n; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; ++i) {
    double a[3], b[3];
    for (int j = 0; j < 3; ++j) { a[j] = P1[3 * i + j] - C1[j]; b[j] = P2[3 * i + j] - C2[j]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] += a[r] * b[c];
  }
  double R[3][3];
  horn_rotation_from_S(S, R);
  double* o = RT + (long long)f * 16;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) o[4 * r + c] = R[r][c];
    o[4 * r + 3] = C2[r] - (R[r][0] * C1[0] + R[r][1] * C1[1] + R[r][2] * C1[2]);
    o[12 + r] = 0.0;
  }
  o[15] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// ADD(-S) distance before ICP (AccumulatorSpace.py:664-702): for every ground-truth-transformed CAD point the distance
// to the nearest estimate-transformed CAD point, brute force in float64.  One CTA per (frame, tile of 256 ground-truth
// points): a thread owns one ground-truth point, the estimated cloud streams through shared memory in tiles.
// ------------------------------------------------------------------------------------------------
constexpr int kAddThreads = 256;
__device__ __forceinline__ void rt_apply(const double* __restrict__ RT, double x, double y, double z, double& ox, double& oy, double& oz) {
  // np.dot(xyz, R.T) + t (project(), :71): row . point, accumulated left to right
  ox = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, RT[0]), __dmul_rn(y, RT[1])), __dmul_rn(z, RT[2])), RT[3]);
  oy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, RT[4]), __dmul_rn(y, RT[5])), __dmul_rn(z, RT[6])), RT[7]);
  oz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, RT[8]), __dmul_rn(y, RT[9])), __dmul_rn(z, RT[10])), RT[11]);
}
__global__ void __launch_bounds__(kAddThreads) k_add_nn(const double* __restrict__ model, int n_model, const double* __restrict__ RT_est,
                                                       const double* __restrict__ RT_gt, double* __restrict__ part_sum,
                                                       double* __restrict__ part_min, int tiles) {
  const int frame = blockIdx.y, tile = blockIdx.x;
  __shared__ double s_e[3][kAddThreads];
  __shared__ double s_rt[2][12];
  __shared__ double s_red[2][kAddThreads / 32];
  if (threadIdx.x < 12) { s_rt[0][threadIdx.x] = RT_est[16LL * frame + threadIdx.x]; s_rt[1][threadIdx.x] = RT_gt[16LL * frame + threadIdx.x]; }
  __syncthreads();
  const int g = tile * kAddThreads + threadIdx.x;
  double gx = 0, gy = 0, gz = 0;
  if (g < n_model) rt_apply(s_rt[1], model[3 * g], model[3 * g + 1], model[3 * g + 2], gx, gy, gz);
  double best = INFINITY;
  for (int e0 = 0; e0 < n_model; e0 += kAddThreads) {
    const int e = e0 + threadIdx.x;
    double ex = INFINITY, ey = INFINITY, ez = INFINITY;   // padding never wins the minimum
    if (e < n_model) rt_apply(s_rt[0], model[3 * e], model[3 * e + 1], model[3 * e + 2], ex, ey, ez);
    __syncthreads();
    s_e[0][threadIdx.x] = ex; s_e[1][threadIdx.x] = ey; s_e[2][threadIdx.x] = ez;
    __syncthreads();
    const int cnt = min(kAddThreads, n_model - e0);
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) {
      const double dx = gx - s_e[0][q], dy = gy - s_e[1][q], dz = gz - s_e[2][q];
      const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
      best = fmin(best, d2);
    }
  }
  double dist = g < n_model ? sqrt(best) : 0.0, dmin = g < n_model ? dist : INFINITY;
  // deterministic tile reduction: warp shuffles in a fixed pattern, then warp partials in order
#pragma unroll
  for (int m = 16; m; m >>= 1) { dist += __shfl_xor_sync(0xffffffffu, dist, m); dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, m)); }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = dist; s_red[1][threadIdx.x >> 5] = dmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, mn = INFINITY;
    for (int w = 0; w < kAddThreads / 32; ++w) { s += s_red[0][w]; mn = fmin(mn, s_red[1][w]); }
    part_sum[(long long)frame * tiles + tile] = s;
    part_min[(long long)frame * tiles + tile] = mn;
  }
}
__global__ void k_add_finish(const double* __restrict__ part_sum, const double* __restrict__ part_min, int tiles, int n_model, int n_frames,
                             double* __restrict__ mean_out, double* __restrict__ min_out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  double s = 0, mn = INFINITY;
  for (int t = 0; t < tiles; ++t) { s += part_sum[(long long)f * tiles + t]; mn = fmin(mn, part_min[(long long)f * tiles + t]); }
  mean_out[f] = s / (double)n_model;
  min_out[f] = mn;
}

// ------------------------------------------------------------------------------------------------
// Roofline denominator: conflict-free shared-memory atomic rate of this GPU (one ATOMS per warp
// instruction, 32 distinct banks), measured the same way as tools/ubench_atoms.cu.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) k_ubench_atoms(int iters, unsigned* out) {
  extern __shared__ __align__(16) int smem[];
  unsigned* s = reinterpret_cast<unsigned*>(smem);
  constexpr unsigned kWords = 32768;
  for (unsigned i = threadIdx.x; i < kWords; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = warp * 997u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (unsigned u = 0; u < 8; ++u) {
      atomicAdd(&s[(base + ((lane + u) & 31) + 32u * u) & (kWords - 1)], 1u);
      base += 1031u;
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (unsigned i = threadIdx.x; i < kWords; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeefu) out[blockIdx.x] = acc;
}

}  // namespace

// ================================================================================================
// Context and C ABI
// ================================================================================================
struct rcv_ctx {
  int device, sms;
  int dp_mod;   // experiments: residue of the tile row stride modulo 32 (RCV_DP_MOD), -1 = any odd stride
  int gen;   // vote kernel generation: 2 = run-length difference arrays (k_vote_runs), 1 = per-candidate rasteriser (k_vote, RCV_VOTE_GEN=1)
  rcv_config cfg;
  Pool pool;
  ItemMeta* meta;
  Unit* units;
  int* counters;  // [0] units queued, [1] queue cursor
  int* cnt;
  unsigned* mask_bits; long long mask_words;
  int last_mask_frames, last_mask_kpts, last_mask_words_per_item;   // survival bits left by the most recent frames call (rcv_scene_clouds_last)
  float* head_radius; long long head_radius_cap;   // fused head: radius planes of the items of a call
  double* icp_scratch; long long icp_cap;         // ICP scratch: per-frame state + per-(frame, tile) partial sums
  int full_window;                                // RCV_FULL_WINDOW=1: gen-2 tiles cover the whole grid instead of the reachable box
  void* icp_grid; long long icp_grid_cap;         // ICP uniform grids (bytes): sized by the scene points of a call
  void* add_grid; long long add_grid_cap;         // ADD(-S): grid over the CAD model + the estimated points of a call's frames (bytes)
  double* add_part; long long add_part_cap;     // ADD metric scratch: per-(frame, tile) partial sums and minima   // K1 scratch: one survival bit per pixel of every item of a call (grown on demand)
  unsigned long long *best, *votes;
  PwLeaf* leaves; double* leaf_sums; long long leaf_cap;
  cudaEvent_t evr[64][2]; long long ev_count;   // ring of (start, stop) events around the vote kernel
  cudaEvent_t ev_ubench[2];                     // rcv_ubench_smem_atomics
  long long launches;
  // staging for the _host entry points
  cudaStream_t s_in, s_in2, s_run; cudaEvent_t ev_in[2], ev_in2[2], ev_done[2], ev_user;
  long long h2d_bytes;   // bytes the last rcv_vote_frames_host call copied to the device (after row-range cropping)
  void* st_depth[2]; float* st_radius[2]; float* st_sem[2]; double* st_K; double* st_maxr;
  double* st_centre; int* st_peak; long long* st_votes; int* st_np; int* st_grid; int* st_status;
  long long st_frames, st_kpts, st_px, st_depth_bytes; int st_has_sem; long long st_total_items;
  double* st_horn_in; long long st_horn_cap;
  char err[512];
};

static char g_create_err[512] = "";
extern "C" long long rcv_icp_scratch_doubles(int n_frames, int n_model);   // refine.cu

#define CK(ctx, call)                                                                                         \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return RCV_E_CUDA;                                                                                      \
    }                                                                                                         \
  } while (0)

#define FAIL(ctx, code, ...)                                  \
  do {                                                        \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
    return (code);                                            \
  } while (0)

RCV_EXPORT int rcv_abi_version(void) { return RCV_ABI_VERSION; }

RCV_EXPORT const char* rcv_last_error(const rcv_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

RCV_EXPORT long long rcv_launch_count(const rcv_ctx* ctx) { return ctx ? ctx->launches : 0; }

RCV_EXPORT long long rcv_last_h2d_bytes(const rcv_ctx* ctx) { return ctx ? ctx->h2d_bytes : 0; }

RCV_EXPORT void rcv_destroy(rcv_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->pool.X); cudaFree(c->pool.Y); cudaFree(c->pool.Z); cudaFree(c->pool.Rd); cudaFree(c->pool.Ri); cudaFree(c->pool.perm); cudaFree(c->pool.rec); cudaFree(c->pool.grp);
  cudaFree(c->meta); cudaFree(c->units); cudaFree(c->counters); cudaFree(c->cnt); cudaFree(c->best); cudaFree(c->votes);
  cudaFree(c->leaves); cudaFree(c->leaf_sums); cudaFree(c->mask_bits); cudaFree(c->add_part); cudaFree(c->icp_scratch); cudaFree(c->icp_grid); cudaFree(c->add_grid); cudaFree(c->head_radius);
  for (int s = 0; s < 2; ++s) { cudaFree(c->st_depth[s]); cudaFree(c->st_radius[s]); cudaFree(c->st_sem[s]); }
  cudaFree(c->st_K); cudaFree(c->st_maxr); cudaFree(c->st_centre); cudaFree(c->st_peak); cudaFree(c->st_votes);
  cudaFree(c->st_np); cudaFree(c->st_grid); cudaFree(c->st_status); cudaFree(c->st_horn_in);
  for (int e = 0; e < 64; ++e) { if (c->evr[e][0]) cudaEventDestroy(c->evr[e][0]); if (c->evr[e][1]) cudaEventDestroy(c->evr[e][1]); }
  if (c->ev_user) cudaEventDestroy(c->ev_user);
  for (int q = 0; q < 2; ++q) if (c->ev_ubench[q]) cudaEventDestroy(c->ev_ubench[q]);
  for (int s = 0; s < 2; ++s) { if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]); if (c->ev_done[s]) cudaEventDestroy(c->ev_done[s]); }
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_in2) cudaStreamDestroy(c->s_in2);
  for (int q = 0; q < 2; ++q) if (c->ev_in2[q]) cudaEventDestroy(c->ev_in2[q]);
  if (c->s_run) cudaStreamDestroy(c->s_run);
  free(c);
}

RCV_EXPORT int rcv_create(int device, const rcv_config* cfg, rcv_ctx** out) {
  if (!cfg || !out || cfg->abi_version != RCV_ABI_VERSION || cfg->max_items <= 0 || cfg->max_points_total <= 0 || cfg->max_grid <= 0 ||
      cfg->max_grid > 1024) {
    snprintf(g_create_err, sizeof(g_create_err), "rcv_create: invalid configuration");
    return RCV_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_create_err, sizeof(g_create_err), "rcv_create: no CUDA device %d (found %d); there is no CPU fallback", device, ndev);
    return RCV_E_NOGPU;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    snprintf(g_create_err, sizeof(g_create_err), "rcv_create: device %d is sm_%d%d; librcvvote is built for sm_100a only", device, prop.major,
             prop.minor);
    return RCV_E_NOGPU;
  }
  rcv_ctx* c = (rcv_ctx*)calloc(1, sizeof(rcv_ctx));
  if (!c) return RCV_E_INVALID;
  c->device = device; c->cfg = *cfg; c->sms = prop.multiProcessorCount;
  { const char* g = getenv("RCV_VOTE_GEN"); c->gen = (g && atoi(g) == 1) ? 1 : 2; }
  { const char* g = getenv("RCV_DP_MOD"); c->dp_mod = g ? atoi(g) : -1; }
  { const char* g = getenv("RCV_FULL_WINDOW"); c->full_window = (g && atoi(g)) ? 1 : 0; }
  if (c->cfg.max_units <= 0) {
    // worst case per item: D slices x ceil(D / rows-per-tile) tiles
    const long long per_item = (long long)cfg->max_grid * ((cfg->max_grid * (long long)(cfg->max_grid | 1)) / kTileWords + 1);
    long long mu = per_item * cfg->max_items;
    if (mu > (1LL << 26)) mu = 1LL << 26;
    c->cfg.max_units = (int)mu;
  }
  *out = c;
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "%s failed: %s", #call, cudaGetErrorString(e_)); rcv_destroy(c); *out = NULL; return RCV_E_CUDA; } } while (0)
  CKC(cudaSetDevice(device));
  const long long cap = cfg->max_points_total;
  c->pool.cap = cap;
  CKC(cudaMalloc(&c->pool.X, cap * 8)); CKC(cudaMalloc(&c->pool.Y, cap * 8)); CKC(cudaMalloc(&c->pool.Z, cap * 8));
  CKC(cudaMalloc(&c->pool.Rd, cap * 8)); CKC(cudaMalloc(&c->pool.Ri, cap * 4)); CKC(cudaMalloc(&c->pool.perm, cap * 4));
  CKC(cudaMalloc(&c->pool.rec, cap * 32));
  CKC(cudaMalloc(&c->pool.grp, (size_t)(cap / 32 + cfg->max_items + 8) * 4));
  CKC(cudaMalloc(&c->meta, sizeof(ItemMeta) * (size_t)cfg->max_items));
  CKC(cudaMalloc(&c->units, sizeof(Unit) * (size_t)c->cfg.max_units));
  CKC(cudaMalloc(&c->counters, 64 + 4096));
  CKC(cudaMalloc(&c->cnt, 4 * (size_t)cfg->max_items));
  CKC(cudaMalloc(&c->best, 8 * (size_t)cfg->max_items)); CKC(cudaMalloc(&c->votes, 8 * (size_t)cfg->max_items));
  c->leaf_cap = cap / 56 + 2LL * cfg->max_items + 8;
  CKC(cudaMalloc(&c->leaves, sizeof(PwLeaf) * (size_t)c->leaf_cap)); CKC(cudaMalloc(&c->leaf_sums, 24 * (size_t)c->leaf_cap));
  for (int e = 0; e < 64; ++e) { CKC(cudaEventCreate(&c->evr[e][0])); CKC(cudaEventCreate(&c->evr[e][1])); }
  CKC(cudaEventCreate(&c->ev_ubench[0])); CKC(cudaEventCreate(&c->ev_ubench[1]));
  // scratch whose size depends on the image / model: allocated here when the configuration names the sizes (no allocation on a hot call)
  if (cfg->image_pixels > 0) {
    c->mask_words = (long long)cfg->max_items * ((cfg->image_pixels + 255) / 256) * 8;
    CKC(cudaMalloc(&c->mask_bits, (size_t)c->mask_words * 4));
    if (cfg->head_items > 0) {
      c->head_radius_cap = (long long)cfg->head_items * cfg->image_pixels;
      CKC(cudaMalloc(&c->head_radius, (size_t)c->head_radius_cap * 4));
    }
  }
  if (cfg->max_model_points > 0) {
    c->add_part_cap = 2LL * cfg->max_items * ((cfg->max_model_points + kAddThreads - 1) / kAddThreads);
    CKC(cudaMalloc(&c->add_part, (size_t)c->add_part_cap * 8));
    c->icp_cap = rcv_icp_scratch_doubles(cfg->max_items, cfg->max_model_points);
    CKC(cudaMalloc(&c->icp_scratch, (size_t)c->icp_cap * 8));
  }
  CKC(cudaFuncSetAttribute(k_ubench_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  CKC(cudaFuncSetAttribute(k_vote, cudaFuncAttributeMaxDynamicSharedMemorySize, (kTileWords + kDummyWords) * 4));
  CKC(cudaFuncSetAttribute(k_vote_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, kRunsTileWords * 4));
#undef CKC
  return RCV_OK;
}

// prelude -> vote -> finalize over items already in the pool (meta[].off/n/status set)
static int run_items(rcv_ctx* c, int n_items, const rcv_vote_params* vp, double* centre_mm, int* peak, long long* votes, int* n_points,
                     int* grid, int* zb, int* status, int32_t* volume, long long volume_cap, cudaStream_t st) {
  CK(c, cudaMemsetAsync(c->counters, 0, 64, st));
  CK(c, cudaMemsetAsync(c->best, 0, 8 * (size_t)n_items, st));
  CK(c, cudaMemsetAsync(c->votes, 0, 8 * (size_t)n_items, st));
  PreludeArgs pa{c->pool, c->meta, c->units, c->counters, c->cfg.max_units, c->cfg.max_grid, vp->grid_policy,
                 c->gen >= 2 ? 2 * kRunsPlaneWords : kTileWords, c->gen, c->dp_mod, c->full_window, c->leaves, c->leaf_sums, c->leaf_cap};
  if (volume) CK(c, cudaMemsetAsync(volume, 0, (size_t)volume_cap * 4, st));   // tiles only cover the box the spheres can reach: the rest of a dumped volume is zero
  k_prelude<<<n_items, kPreludeThreads, 0, st>>>(pa);
  VoteArgs va{c->pool, c->meta, c->units, c->counters, c->best, c->votes, volume, volume_cap};
  const int slot = (int)(c->ev_count % 64);
  CK(c, cudaEventRecord(c->evr[slot][0], st));
  if (c->gen >= 2) k_vote_runs<<<c->sms * RCV_RUNS_CTAS, kRunsThreads, kRunsTileWords * 4, st>>>(va);
  else k_vote<<<c->sms, kVoteThreads, (kTileWords + kDummyWords) * 4, st>>>(va);
  CK(c, cudaEventRecord(c->evr[slot][1], st));
  c->ev_count += 1;
  FinalArgs fa{c->meta, c->best, c->votes, n_items, vp->acc_unit, vp->grid_policy, volume != nullptr, volume_cap,
               centre_mm, peak, votes, n_points, grid, zb, status};
  k_finalize<<<(n_items + 127) / 128, 128, 0, st>>>(fa);
  c->launches += 3;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

static int check_vote_params(rcv_ctx* c, const rcv_vote_params* vp) {
  if (!vp || !(vp->acc_unit > 0) || !(vp->radius_scale > 0)) FAIL(c, RCV_E_INVALID, "vote params: acc_unit and radius_scale must be positive");
  if (vp->grid_policy != RCV_POLICY_LM && vp->grid_policy != RCV_POLICY_YCBGEN) FAIL(c, RCV_E_INVALID, "vote params: unknown grid policy %d", vp->grid_policy);
  if (vp->radius_dtype != RCV_F32 && vp->radius_dtype != RCV_F64) FAIL(c, RCV_E_INVALID, "vote params: radius dtype must be RCV_F32 or RCV_F64");
  return RCV_OK;
}

RCV_EXPORT int rcv_vote_points(rcv_ctx* c, const double* xyz, const void* radii, const long long* item_offsets, int n_items,
                               const rcv_vote_params* vp, double* centre_mm, int* peak, long long* votes, int* grid, int* zero_boundary,
                               int* status, int32_t* volume_out, long long volume_capacity, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!xyz || !radii || !item_offsets || !centre_mm || !status) FAIL(c, RCV_E_INVALID, "rcv_vote_points: null pointer");
  if (n_items <= 0 || n_items > c->cfg.max_items) FAIL(c, RCV_E_CAPACITY, "rcv_vote_points: n_items %d outside (0, %d]", n_items, c->cfg.max_items);
  if (volume_out && n_items != 1) FAIL(c, RCV_E_INVALID, "rcv_vote_points: volume_out needs n_items == 1");
  int rc = check_vote_params(c, vp);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  k_items_from_offsets<<<(n_items + 127) / 128, 128, 0, st>>>(item_offsets, n_items, c->pool.cap, c->meta);
  k_points_to_voxel<<<c->sms * 4, 256, 0, st>>>(xyz, radii, vp->radius_dtype, item_offsets, n_items, vp->acc_unit, vp->radius_scale, c->pool);
  c->launches += 2;
  return run_items(c, n_items, vp, centre_mm, peak, votes, nullptr, grid, zero_boundary, status, volume_out, volume_capacity, st);
}

static int check_frame_params(rcv_ctx* c, int n_frames, int n_kpts, const rcv_frame_params* fp, bool has_sem, const double* max_radii) {
  if (!fp || fp->height <= 0 || fp->width <= 0) FAIL(c, RCV_E_INVALID, "frame params: bad image size");
  if (fp->depth_dtype != RCV_U16 && fp->depth_dtype != RCV_F32 && fp->depth_dtype != RCV_F64) FAIL(c, RCV_E_INVALID, "frame params: bad depth dtype");
  if (!(fp->depth_div > 0) || !(fp->xyz_div > 0)) FAIL(c, RCV_E_INVALID, "frame params: depth_div and xyz_div must be positive");
  if ((fp->mask_flags & (RCV_MASK_SEM_GT | RCV_MASK_SEM_GE)) && !has_sem) FAIL(c, RCV_E_INVALID, "frame params: sem rule without a sem map");
  if ((fp->mask_flags & RCV_MASK_MAX_RADIUS) && !max_radii) FAIL(c, RCV_E_INVALID, "frame params: max-radius rule without max_radii");
  if (n_frames <= 0 || n_kpts <= 0 || (long long)n_frames * n_kpts > c->cfg.max_items)
    FAIL(c, RCV_E_CAPACITY, "n_frames*n_kpts = %lld exceeds max_items %d", (long long)n_frames * n_kpts, c->cfg.max_items);
  return RCV_OK;
}

RCV_EXPORT int rcv_vote_frames(rcv_ctx* c, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem, const double* K,
                               const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp, double* centre_mm, int* peak,
                               long long* votes, int* n_points, int* grid, int* status, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!depth || !radius || !K || !centre_mm || !status) FAIL(c, RCV_E_INVALID, "rcv_vote_frames: null pointer");
  int rc = check_vote_params(c, vp);
  if (rc) return rc;
  rc = check_frame_params(c, n_frames, n_kpts, fp, sem != nullptr, max_radii);
  if (rc) return rc;
  if (vp->radius_dtype != RCV_F32) FAIL(c, RCV_E_INVALID, "rcv_vote_frames: radius maps are float32");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const int n_items = n_frames * n_kpts;
  const long long npx = (long long)fp->height * fp->width;
  const int vec_ok = (npx % 8 == 0) && (((uintptr_t)depth | (uintptr_t)radius | (uintptr_t)sem) % 16 == 0);
  FrameArgs fa{depth, radius, sem, K, max_radii, *fp, vp->acc_unit, vp->radius_scale, n_kpts, vec_ok};
  const int words = (int)((npx + 255) / 256) * 8;   // whole warp-steps of 256 pixels
  const long long need = (long long)n_items * words;
  if (need > c->mask_words) {   // only when rcv_config.image_pixels was left 0 or is exceeded: first call of this size
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->mask_bits); c->mask_bits = nullptr; c->mask_words = 0;
    CK(c, cudaMalloc(&c->mask_bits, (size_t)need * 4));
    c->mask_words = need;
  }
  k_frame_mask<<<n_items, kCompactThreads, 0, st>>>(fa, c->cnt, c->mask_bits, words);
  c->last_mask_frames = n_frames; c->last_mask_kpts = n_kpts; c->last_mask_words_per_item = words;
  k_scan_items<<<1, 1024, 0, st>>>(c->cnt, n_items, c->pool.cap, c->meta);
  k_frame_emit<<<n_items, kCompactThreads, 0, st>>>(c->mask_bits, words, c->meta, c->pool.perm, words);
  k_points_from_pixels<<<dim3(n_items, 4), 256, 0, st>>>(fa, c->meta, c->pool);
  c->launches += 4;
  return run_items(c, n_items, vp, centre_mm, peak, votes, n_points, grid, nullptr, status, nullptr, 0, st);
}

// true iff the `n` bytes at p are all zero (a depth row: u16 / f32 / f64 zeros are all-zero bytes; -0.0 counts as non-zero,
// which only costs the copy of that row)
static bool row_is_zero(const char* p, long long n) {
  long long i = 0;
  for (; i < n && ((uintptr_t)(p + i) & 7); ++i) if (p[i]) return false;
  const unsigned long long* w = (const unsigned long long*)(p + i);
  const long long nw = (n - i) / 8;
  unsigned long long acc = 0;
  for (long long q = 0; q < nw; q += 16) {            // a cache line or two at a time, early exit
    const long long e = q + 16 < nw ? q + 16 : nw;
    for (long long j = q; j < e; ++j) acc |= w[j];
    if (acc) return false;
  }
  for (i += nw * 8; i < n; ++i) if (p[i]) return false;
  return true;
}

// helper threads of the host entry points (row-range scan); RCV_HOST_THREADS overrides, 0 = scan on the calling thread
static int host_helper_threads() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RCV_HOST_THREADS");
    v = e ? atoi(e) : 3;
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && v > hw - 1) v = hw - 1;
    if (v < 0) v = 0;
    if (v > 16) v = 16;
  }
  return v;
}

static int ensure_staging(rcv_ctx* c, int frames, int kpts, long long px, long long depth_bytes, int has_sem, long long total_items) {
  if (c->st_frames >= frames && c->st_kpts == kpts && c->st_px == px && c->st_depth_bytes == depth_bytes && c->st_has_sem >= has_sem &&
      c->st_total_items >= total_items)
    return RCV_OK;
  for (int s = 0; s < 2; ++s) { cudaFree(c->st_depth[s]); cudaFree(c->st_radius[s]); cudaFree(c->st_sem[s]); c->st_depth[s] = nullptr; c->st_radius[s] = nullptr; c->st_sem[s] = nullptr; }
  cudaFree(c->st_K); cudaFree(c->st_maxr); cudaFree(c->st_centre); cudaFree(c->st_peak); cudaFree(c->st_votes); cudaFree(c->st_np);
  cudaFree(c->st_grid); cudaFree(c->st_status);
  c->st_K = c->st_maxr = c->st_centre = nullptr; c->st_peak = c->st_np = c->st_grid = c->st_status = nullptr; c->st_votes = nullptr;
  c->st_frames = 0;
  for (int s = 0; s < 2; ++s) {
    CK(c, cudaMalloc(&c->st_depth[s], (size_t)(frames * px * depth_bytes)));
    CK(c, cudaMalloc(&c->st_radius[s], (size_t)(frames * kpts * px * 4)));
    if (has_sem) CK(c, cudaMalloc(&c->st_sem[s], (size_t)(frames * kpts * px * 4)));
  }
  const long long nf = total_items / kpts + 1;
  CK(c, cudaMalloc(&c->st_K, (size_t)(nf * 9 * 8))); CK(c, cudaMalloc(&c->st_maxr, (size_t)(total_items + kpts) * 8));
  CK(c, cudaMalloc(&c->st_centre, (size_t)total_items * 24)); CK(c, cudaMalloc(&c->st_peak, (size_t)total_items * 4));
  CK(c, cudaMalloc(&c->st_votes, (size_t)total_items * 8)); CK(c, cudaMalloc(&c->st_np, (size_t)total_items * 4));
  CK(c, cudaMalloc(&c->st_grid, (size_t)total_items * 4)); CK(c, cudaMalloc(&c->st_status, (size_t)total_items * 4));
  if (!c->s_in) {
    CK(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)); CK(c, cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
    CK(c, cudaStreamCreateWithFlags(&c->s_in2, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) CK(c, cudaEventCreateWithFlags(&c->ev_in2[q], cudaEventDisableTiming));
    for (int s = 0; s < 2; ++s) { CK(c, cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming)); CK(c, cudaEventCreateWithFlags(&c->ev_done[s], cudaEventDisableTiming)); }
    CK(c, cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming));
  }
  c->st_frames = frames; c->st_kpts = kpts; c->st_px = px; c->st_depth_bytes = depth_bytes; c->st_has_sem = has_sem; c->st_total_items = total_items;
  return RCV_OK;
}

RCV_EXPORT int rcv_vote_frames_host(rcv_ctx* c, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem,
                                    const double* K, const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp,
                                    double* centre_mm, int* peak, long long* votes, int* n_points, int* grid, int* status,
                                    int frames_per_chunk, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!depth || !radius || !K || !centre_mm || !status) FAIL(c, RCV_E_INVALID, "rcv_vote_frames_host: null pointer");
  int rc = check_vote_params(c, vp);
  if (rc) return rc;
  if (frames_per_chunk <= 0) frames_per_chunk = 256;
  if ((long long)frames_per_chunk * n_kpts > c->cfg.max_items) frames_per_chunk = c->cfg.max_items / (n_kpts > 0 ? n_kpts : 1);
  if (frames_per_chunk > n_frames) frames_per_chunk = n_frames;
  rc = check_frame_params(c, frames_per_chunk, n_kpts, fp, sem != nullptr, max_radii);
  if (rc) return rc;
  CK(c, cudaSetDevice(c->device));
  const long long px = (long long)fp->height * fp->width;
  const long long dbytes = fp->depth_dtype == RCV_U16 ? 2 : fp->depth_dtype == RCV_F32 ? 4 : 8;
  const long long total_items = (long long)n_frames * n_kpts;
  rc = ensure_staging(c, frames_per_chunk, n_kpts, px, dbytes, sem != nullptr, total_items);
  if (rc) return rc;
  cudaStream_t user = (cudaStream_t)stream;
  CK(c, cudaEventRecord(c->ev_user, user));
  CK(c, cudaStreamWaitEvent(c->s_in, c->ev_user, 0));
  CK(c, cudaStreamWaitEvent(c->s_in2, c->ev_user, 0));
  CK(c, cudaStreamWaitEvent(c->s_run, c->ev_user, 0));
  const long long nK = fp->k_stride ? n_frames : 1, nM = fp->max_radii_stride ? n_frames : 1;
  c->h2d_bytes = (nK * 9 + (max_radii ? nM * n_kpts : 0)) * 8;
  CK(c, cudaMemcpyAsync(c->st_K, K, (size_t)(nK * 9 * 8), cudaMemcpyHostToDevice, c->s_in));
  if (max_radii) CK(c, cudaMemcpyAsync(c->st_maxr, max_radii, (size_t)(nM * n_kpts * 8), cudaMemcpyHostToDevice, c->s_in));
  // Row-range cropping.  Every mask rule keeps a pixel only where depth != 0, so the rows above the first and
  // below the last non-zero depth row of a frame cannot contribute: only rows [r0, r1) of the depth image and of each radius /
  // seg plane cross the bus; the depth slot is cleared first, and whatever the other planes still hold outside the range sits
  // under zero depth.  Results are bit-identical to copying whole images; a dense depth image costs one row of scanning.
  // The scan reads most of every depth image from host DRAM, so helper threads run ahead of the thread that issues the copies
  // (helpers claim frames in order from a shared counter; RCV_HOST_THREADS, default 3; small calls scan inline).
  const long long row_b = (long long)fp->width * dbytes;
  const int height = fp->height;
  auto scan_frame = [depth, px, dbytes, row_b, height](int f, int& r0, int& r1) {
    const char* dsrc = (const char*)depth + (size_t)f * (size_t)(px * dbytes);
    r0 = 0; r1 = height;
    while (r0 < r1 && row_is_zero(dsrc + (size_t)r0 * row_b, row_b)) ++r0;
    while (r1 > r0 && row_is_zero(dsrc + (size_t)(r1 - 1) * row_b, row_b)) --r1;
  };
  struct RowRange { int r0, r1; std::atomic<int> ready; };
  int n_helpers = host_helper_threads();
  if (n_frames < 64) n_helpers = 0;
  std::vector<RowRange> ranges;
  std::vector<std::thread> helpers;
  std::atomic<int> next_frame{0};
  struct Joiner { std::vector<std::thread>& t; ~Joiner() { for (auto& h : t) if (h.joinable()) h.join(); } } joiner{helpers};   // also on the error returns below
  if (n_helpers) {
    try {                                               // no exception may cross the C ABI: without helpers the caller scans
      ranges = std::vector<RowRange>((size_t)n_frames);
      for (auto& r : ranges) r.ready.store(0, std::memory_order_relaxed);
      RowRange* rr = ranges.data();
      std::atomic<int>* nf = &next_frame;
      for (int t = 0; t < n_helpers; ++t)
        helpers.emplace_back([=]() {                    // frames are claimed in order, so any number of helpers (>= 1) finishes them all
          for (int f = nf->fetch_add(1); f < n_frames; f = nf->fetch_add(1)) { scan_frame(f, rr[f].r0, rr[f].r1); rr[f].ready.store(1, std::memory_order_release); }
        });
    } catch (...) {
      if (helpers.empty()) n_helpers = 0;
    }
  }
  // The first copy is not hidden behind any voting: the first two chunks are a quarter and a half of the regular size.
  int chunk = 0;
  for (int f0 = 0, nf = 0; f0 < n_frames; f0 += nf, ++chunk) {
    int want = frames_per_chunk;
    if (n_frames >= 4 * frames_per_chunk && frames_per_chunk >= 128) want = chunk == 0 ? frames_per_chunk / 4 : chunk == 1 ? frames_per_chunk / 2 : frames_per_chunk;
    nf = n_frames - f0 < want ? n_frames - f0 : want;
    const int s = chunk & 1;
    if (chunk >= 2) { CK(c, cudaStreamWaitEvent(c->s_in, c->ev_done[s], 0)); CK(c, cudaStreamWaitEvent(c->s_in2, c->ev_done[s], 0)); }  // slot reuse: previous occupant finished voting
    CK(c, cudaMemsetAsync(c->st_depth[s], 0, (size_t)(nf * px * dbytes), c->s_in));
    for (int f = 0; f < nf; ++f) {
      const char* dsrc = (const char*)depth + (size_t)((f0 + f) * px * dbytes);
      int r0, r1;
      if (n_helpers) {
        RowRange& r = ranges[(size_t)(f0 + f)];
        while (!r.ready.load(std::memory_order_acquire)) std::this_thread::yield();
        r0 = r.r0; r1 = r.r1;
      } else {
        scan_frame(f0 + f, r0, r1);
      }
      if (r1 <= r0) continue;
      const long long p0 = (long long)r0 * fp->width, np = (long long)(r1 - r0) * fp->width;
      CK(c, cudaMemcpyAsync((char*)c->st_depth[s] + (size_t)((f * px + p0) * dbytes), dsrc + (size_t)(p0 * dbytes), (size_t)(np * dbytes), cudaMemcpyHostToDevice, c->s_in));
      c->h2d_bytes += np * dbytes;
      // the same rows of the frame's n_kpts radius (and seg) planes: one strided copy (n_kpts "rows" of np floats, a plane apart)
      const long long plane = (long long)f * n_kpts * px + p0, src = (long long)(f0 + f) * n_kpts * px + p0;
      CK(c, cudaMemcpy2DAsync(c->st_radius[s] + plane, (size_t)px * 4, radius + src, (size_t)px * 4, (size_t)np * 4, (size_t)n_kpts, cudaMemcpyHostToDevice, c->s_in2));
      if (sem) CK(c, cudaMemcpy2DAsync(c->st_sem[s] + plane, (size_t)px * 4, sem + src, (size_t)px * 4, (size_t)np * 4, (size_t)n_kpts, cudaMemcpyHostToDevice, c->s_in2));
      c->h2d_bytes += np * 4 * n_kpts * (sem ? 2 : 1);
    }
    CK(c, cudaEventRecord(c->ev_in[s], c->s_in));
    CK(c, cudaEventRecord(c->ev_in2[s], c->s_in2));
    CK(c, cudaStreamWaitEvent(c->s_run, c->ev_in[s], 0));
    CK(c, cudaStreamWaitEvent(c->s_run, c->ev_in2[s], 0));
    const long long i0 = (long long)f0 * n_kpts;
    rc = rcv_vote_frames(c, nf, n_kpts, c->st_depth[s], c->st_radius[s], sem ? c->st_sem[s] : nullptr,
                         c->st_K + (fp->k_stride ? (long long)f0 * fp->k_stride : 0),
                         max_radii ? c->st_maxr + (fp->max_radii_stride ? (long long)f0 * fp->max_radii_stride : 0) : nullptr, fp, vp,
                         c->st_centre + 3 * i0, c->st_peak + i0, c->st_votes + i0, c->st_np + i0, c->st_grid + i0, c->st_status + i0, c->s_run);
    if (rc) return rc;
    CK(c, cudaEventRecord(c->ev_done[s], c->s_run));
  }
  CK(c, cudaMemcpyAsync(centre_mm, c->st_centre, (size_t)total_items * 24, cudaMemcpyDeviceToHost, c->s_run));
  CK(c, cudaMemcpyAsync(status, c->st_status, (size_t)total_items * 4, cudaMemcpyDeviceToHost, c->s_run));
  if (peak) CK(c, cudaMemcpyAsync(peak, c->st_peak, (size_t)total_items * 4, cudaMemcpyDeviceToHost, c->s_run));
  if (votes) CK(c, cudaMemcpyAsync(votes, c->st_votes, (size_t)total_items * 8, cudaMemcpyDeviceToHost, c->s_run));
  if (n_points) CK(c, cudaMemcpyAsync(n_points, c->st_np, (size_t)total_items * 4, cudaMemcpyDeviceToHost, c->s_run));
  if (grid) CK(c, cudaMemcpyAsync(grid, c->st_grid, (size_t)total_items * 4, cudaMemcpyDeviceToHost, c->s_run));
  CK(c, cudaStreamSynchronize(c->s_in));
  CK(c, cudaStreamSynchronize(c->s_in2));
  CK(c, cudaStreamSynchronize(c->s_run));
  return RCV_OK;
}

RCV_EXPORT int rcv_backproject(rcv_ctx* c, const double* K, const void* depth, int depth_dtype, int height, int width, double* xyz_out,
                               long long xyz_capacity, int* n_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!K || !depth || !xyz_out || !n_out || height <= 0 || width <= 0) FAIL(c, RCV_E_INVALID, "rcv_backproject: bad argument");
  if (depth_dtype != RCV_U16 && depth_dtype != RCV_F32 && depth_dtype != RCV_F64) FAIL(c, RCV_E_INVALID, "rcv_backproject: bad depth dtype");
  CK(c, cudaSetDevice(c->device));
  k_backproject<<<1, 1024, 0, (cudaStream_t)stream>>>(K, depth, depth_dtype, height, width, xyz_out, xyz_capacity, n_out);
  c->launches += 1;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

RCV_EXPORT int rcv_argmax_volume(rcv_ctx* c, const int32_t* volume, int grid, int* idx_out, int* max_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!volume || !idx_out || !max_out || grid <= 0 || grid > 1024) FAIL(c, RCV_E_INVALID, "rcv_argmax_volume: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaMemsetAsync(c->best, 0, 8, st));
  const long long total = (long long)grid * grid * grid;
  long long blocks = (total / 4 + 255) / 256;
  if (blocks > c->sms * 8) blocks = c->sms * 8;
  if (blocks < 1) blocks = 1;
  k_argmax_volume<<<(int)blocks, 256, 0, st>>>(volume, total, c->best);
  k_argmax_unpack<<<1, 1, 0, st>>>(c->best, grid, idx_out, max_out);
  c->launches += 2;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

RCV_EXPORT int rcv_horn_batch(rcv_ctx* c, const double* model, long long model_stride, const double* est, int n, int n_frames, double* RT,
                              void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!model || !est || !RT || n <= 0 || n_frames <= 0) FAIL(c, RCV_E_INVALID, "rcv_horn_batch: bad argument");
  if (model_stride != 0 && model_stride != 3LL * n) FAIL(c, RCV_E_INVALID, "rcv_horn_batch: model_stride must be 0 or 3*n");
  CK(c, cudaSetDevice(c->device));
  k_horn<<<(n_frames + 63) / 64, 64, 0, (cudaStream_t)stream>>>(model, model_stride, est, n, n_frames, RT);
  c->launches += 1;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

extern "C" long long rcv_add_grid_bytes(int n_frames, int n_model);
extern "C" int rcv_add_grid_launch(const double* model, int n_model, const double* RT_est, const double* RT_gt, int n_frames, double* part_sum,
                                   double* part_min, int tiles, int threads_per_tile, void* scratch, void* stream, long long* launches);

RCV_EXPORT int rcv_add_metric_batch(rcv_ctx* c, const double* model_mm, int n_model, const double* RT_est, const double* RT_gt, int n_frames,
                                    double* mean_out, double* min_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!model_mm || !RT_est || !RT_gt || !mean_out || !min_out || n_model <= 0 || n_frames <= 0)
    FAIL(c, RCV_E_INVALID, "rcv_add_metric_batch: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const int tiles = (n_model + kAddThreads - 1) / kAddThreads;
  const long long need = 2LL * n_frames * tiles;
  if (need > c->add_part_cap) {   // first call of this size only
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->add_part); c->add_part = nullptr; c->add_part_cap = 0;
    CK(c, cudaMalloc(&c->add_part, (size_t)need * 8));
    c->add_part_cap = need;
  }
  double* ps = c->add_part;
  double* pm = c->add_part + (long long)n_frames * tiles;
  // Nearest neighbours through a grid over the CAD model (refine.cu: same distances, same reduction, bit-identical) unless the
  // model is small or RCV_ADD_BRUTE=1 asks for the all-pairs kernel.
  const char* brute_env = getenv("RCV_ADD_BRUTE");
  if (n_model >= 1024 && !(brute_env && atoi(brute_env))) {
    const long long gb = rcv_add_grid_bytes(n_frames, n_model);
    if (gb > c->add_grid_cap) {   // first call of this size only
      CK(c, cudaStreamSynchronize(st));
      cudaFree(c->add_grid); c->add_grid = nullptr; c->add_grid_cap = 0;
      CK(c, cudaMalloc(&c->add_grid, (size_t)gb));
      c->add_grid_cap = gb;
    }
    CK(c, (cudaError_t)rcv_add_grid_launch(model_mm, n_model, RT_est, RT_gt, n_frames, ps, pm, tiles, kAddThreads, c->add_grid, stream, &c->launches));
  } else {
    k_add_nn<<<dim3(tiles, n_frames), kAddThreads, 0, st>>>(model_mm, n_model, RT_est, RT_gt, ps, pm, tiles);
    c->launches += 1;
  }
  k_add_finish<<<(n_frames + 127) / 128, 128, 0, st>>>(ps, pm, tiles, n_model, n_frames, mean_out, min_out);
  c->launches += 1;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

RCV_EXPORT int rcv_scene_clouds(rcv_ctx* c, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem,
                                const double* K, const double* max_radii, const rcv_frame_params* fp, double scale, double* xyz_out,
                                long long xyz_capacity, long long* offsets_out, int* status_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!depth || !radius || !K || !xyz_out || !offsets_out || xyz_capacity <= 0) FAIL(c, RCV_E_INVALID, "rcv_scene_clouds: bad argument");
  int rc = check_frame_params(c, n_frames, n_kpts, fp, sem != nullptr, max_radii);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const long long npx = (long long)fp->height * fp->width;
  const int vec_ok = (npx % 8 == 0) && (((uintptr_t)depth | (uintptr_t)radius | (uintptr_t)sem) % 16 == 0);
  FrameArgs fa{depth, radius, sem, K, max_radii, *fp, 1.0, 1.0, n_kpts, vec_ok};
  const int words = (int)((npx + 255) / 256) * 8;
  const long long need = (long long)n_frames * words;
  if (need > c->mask_words) {
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->mask_bits); c->mask_bits = nullptr; c->mask_words = 0;
    CK(c, cudaMalloc(&c->mask_bits, (size_t)need * 4));
    c->mask_words = need;
  }
  const long long cap = xyz_capacity < c->pool.cap ? xyz_capacity : c->pool.cap;   // the pixel list lives in the pool's index array
  c->last_mask_frames = 0;   // the item bits of an earlier frames call are overwritten
  k_scene_mask<<<n_frames, kCompactThreads, 0, st>>>(fa, c->cnt, c->mask_bits, words);
  k_scene_scan<<<1, 1024, 0, st>>>(c->cnt, n_frames, cap, c->meta, offsets_out, status_out);
  k_frame_emit<<<n_frames, kCompactThreads, 0, st>>>(c->mask_bits, words, c->meta, c->pool.perm, words);
  k_scene_points<<<dim3(n_frames, 4), 256, 0, st>>>(fa, c->meta, c->pool.perm, scale, xyz_out);
  c->launches += 4;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

// OR of the keypoints' survival bits of each frame, left in the frame's first plane; count of survivors per frame
__global__ void __launch_bounds__(256) k_scene_or_bits(unsigned* __restrict__ bits, int words_per_item, int n_kpts, int* __restrict__ cnt) {
  const int frame = blockIdx.x;
  unsigned* base = bits + (long long)frame * n_kpts * words_per_item;
  int n = 0;
  for (int w = threadIdx.x; w < words_per_item; w += blockDim.x) {
    unsigned m = base[w];
    for (int k = 1; k < n_kpts; ++k) m |= base[(long long)k * words_per_item + w];
    base[w] = m;
    n += __popc(m);
  }
  __shared__ int s_n[8];
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_n[w]; cnt[frame] = t; }
}

// rcv_scene_clouds for the frames of the MOST RECENT rcv_vote_frames / rcv_head_vote_frames call on this context: the union of
// the keypoints' masks is the OR of the survival bits that call left behind, so no map is read again (the fused head never
// wrote the seg plane to memory).  Consumes those bits.
RCV_EXPORT int rcv_scene_clouds_last(rcv_ctx* c, int n_frames, int n_kpts, const void* depth, const double* K, const rcv_frame_params* fp,
                                     double scale, double* xyz_out, long long xyz_capacity, long long* offsets_out, int* status_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!depth || !K || !fp || !xyz_out || !offsets_out || xyz_capacity <= 0) FAIL(c, RCV_E_INVALID, "rcv_scene_clouds_last: bad argument");
  const long long npx = (long long)fp->height * fp->width;
  const int words = (int)((npx + 255) / 256) * 8;
  if (n_frames != c->last_mask_frames || n_kpts != c->last_mask_kpts || words != c->last_mask_words_per_item)
    FAIL(c, RCV_E_INVALID, "rcv_scene_clouds_last: no survival bits of %d frames x %d keypoints of %d x %d pixels on this context", n_frames, n_kpts,
         fp->height, fp->width);
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  FrameArgs fa{depth, nullptr, nullptr, K, nullptr, *fp, 1.0, 1.0, n_kpts, 0};
  const long long cap = xyz_capacity < c->pool.cap ? xyz_capacity : c->pool.cap;
  k_scene_or_bits<<<n_frames, 256, 0, st>>>(c->mask_bits, words, n_kpts, c->cnt);
  k_scene_scan<<<1, 1024, 0, st>>>(c->cnt, n_frames, cap, c->meta, offsets_out, status_out);
  k_frame_emit<<<n_frames, kCompactThreads, 0, st>>>(c->mask_bits, words, c->meta, c->pool.perm, (long long)n_kpts * words);
  k_scene_points<<<dim3(n_frames, 4), 256, 0, st>>>(fa, c->meta, c->pool.perm, scale, xyz_out);
  c->last_mask_frames = 0;
  c->launches += 4;
  CK(c, cudaGetLastError());
  return RCV_OK;
}

extern "C" long long rcv_icp_scratch_doubles(int n_frames, int n_model);
extern "C" int rcv_icp_launch(const double* model, int n_model, const double* scene, const long long* scene_off, const double* RT_init,
                              const double* max_dist, int n_frames, int max_iter, double rel_fitness, double rel_rmse, double* scratch,
                              double* RT_out, double* fitness_out, double* rmse_out, int* iters_out, void* stream, long long* launches,
                              void* grid_scratch, long long n_scene);
extern "C" long long rcv_icp_grid_bytes(int n_frames, long long n_scene, int n_model);

RCV_EXPORT int rcv_icp_batch(rcv_ctx* c, const double* model, int n_model, const double* scene, const long long* scene_offsets,
                             const double* RT_init, const double* max_dist, int n_frames, int max_iter, double rel_fitness, double rel_rmse,
                             double* RT_out, double* fitness_out, double* rmse_out, int* iters_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!model || !scene || !scene_offsets || !RT_init || !max_dist || !RT_out || !fitness_out || !rmse_out || !iters_out || n_model <= 0 ||
      n_frames <= 0 || n_frames > 65535 || max_iter < 0)
    FAIL(c, RCV_E_INVALID, "rcv_icp_batch: bad argument (n_frames in 1..65535, max_iter >= 0)");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const long long need = rcv_icp_scratch_doubles(n_frames, n_model);
  if (need > c->icp_cap) {   // first call of this size only
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->icp_scratch); c->icp_scratch = nullptr; c->icp_cap = 0;
    CK(c, cudaMalloc(&c->icp_scratch, (size_t)need * 8));
    c->icp_cap = need;
  }
  // Uniform grid over each frame's scene (refine.cu): its scratch depends on the number of scene points, which only the device
  // knows (scene_offsets): one 8-byte read per call.  RCV_ICP_BRUTE=1, or more frames than the grid tables are worth
  // (2 x 128 KB each), keeps the all-pairs search.
  const char* brute_env = getenv("RCV_ICP_BRUTE");          // read per call: the parity test switches it
  const int brute = (brute_env && atoi(brute_env)) ? 1 : 0;
  void* grid = nullptr;
  long long n_scene = 0;
  if (!brute && n_frames <= 2048) {
    CK(c, cudaMemcpyAsync(&n_scene, scene_offsets + n_frames, 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (n_scene < 0) FAIL(c, RCV_E_INVALID, "rcv_icp_batch: negative scene offset");
    const long long gb = rcv_icp_grid_bytes(n_frames, n_scene, n_model);
    if (gb > c->icp_grid_cap) {
      cudaFree(c->icp_grid); c->icp_grid = nullptr; c->icp_grid_cap = 0;
      CK(c, cudaMalloc(&c->icp_grid, (size_t)gb));
      c->icp_grid_cap = gb;
    }
    grid = c->icp_grid;
  }
  CK(c, (cudaError_t)rcv_icp_launch(model, n_model, scene, scene_offsets, RT_init, max_dist, n_frames, max_iter, rel_fitness, rel_rmse,
                                    c->icp_scratch, RT_out, fitness_out, rmse_out, iters_out, stream, &c->launches, grid, n_scene));
  return RCV_OK;
}

RCV_EXPORT int rcv_horn_batch_host(rcv_ctx* c, const double* model, long long model_stride, const double* est, int n, int n_frames,
                                   double* RT, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!model || !est || !RT || n <= 0 || n_frames <= 0) FAIL(c, RCV_E_INVALID, "rcv_horn_batch_host: bad argument");
  CK(c, cudaSetDevice(c->device));
  const long long nm = model_stride ? (long long)n_frames * 3 * n : 3LL * n, ne = (long long)n_frames * 3 * n, nr = (long long)n_frames * 16;
  if (c->st_horn_cap < nm + ne + nr) {
    cudaFree(c->st_horn_in); c->st_horn_in = nullptr; c->st_horn_cap = 0;
    CK(c, cudaMalloc(&c->st_horn_in, (size_t)(nm + ne + nr) * 8));
    c->st_horn_cap = nm + ne + nr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  double *dm = c->st_horn_in, *de = dm + nm, *dr = de + ne;
  CK(c, cudaMemcpyAsync(dm, model, (size_t)nm * 8, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(de, est, (size_t)ne * 8, cudaMemcpyHostToDevice, st));
  int rc = rcv_horn_batch(c, dm, model_stride, de, n, n_frames, dr, stream);
  if (rc) return rc;
  CK(c, cudaMemcpyAsync(RT, dr, (size_t)nr * 8, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  return RCV_OK;
}

extern "C" int rcv_head1x1_launch(const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw, int sms,
                                  void* stream);

RCV_EXPORT int rcv_head_1x1(rcv_ctx* c, const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw,
                            void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!up_bf16 || !weight || !bias || !out || n_images <= 0 || hw <= 0) FAIL(c, RCV_E_INVALID, "rcv_head_1x1: bad argument");
  if (hw % 8 != 0 || ((uintptr_t)up_bf16 & 15) != 0) FAIL(c, RCV_E_INVALID, "rcv_head_1x1: hw must be a multiple of 8 and `up` 16-byte aligned");
  CK(c, cudaSetDevice(c->device));
  const int rc = rcv_head1x1_launch(up_bf16, weight, bias, out, n_images, hw, c->sms, stream);
  if (rc != 0) FAIL(c, RCV_E_CUDA, "rcv_head_1x1: launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  c->launches += 1;
  return RCV_OK;
}

extern "C" int rcv_head1x1_fused_launch(const void* up_bf16, const float* weight, const float* bias, float* radius_out, int n_items, long long hw,
                                        int sms, const void* depth, int depth_dtype, int n_kpts, const double* max_radii, int max_radii_stride,
                                        int flags, float sem_threshold, unsigned* bits, int words_per_item, int* cnt, void* stream);

RCV_EXPORT int rcv_head_vote_frames(rcv_ctx* c, int n_frames, int n_kpts, const void* up_bf16, const float* weight, const float* bias,
                                    const void* depth, const double* K, const double* max_radii, const rcv_frame_params* fp,
                                    const rcv_vote_params* vp, double* centre_mm, int* peak, long long* votes, int* n_points, int* grid, int* status,
                                    float* radius_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!up_bf16 || !weight || !bias || !depth || !K || !centre_mm || !status) FAIL(c, RCV_E_INVALID, "rcv_head_vote_frames: null pointer");
  int rc = check_vote_params(c, vp);
  if (rc) return rc;
  rc = check_frame_params(c, n_frames, n_kpts, fp, true /* the seg values come from the head */, max_radii);
  if (rc) return rc;
  const long long npx = (long long)fp->height * fp->width;
  if (n_kpts > 8 || npx % 8 != 0 || ((uintptr_t)up_bf16 & 15)) FAIL(c, RCV_E_INVALID, "rcv_head_vote_frames: n_kpts <= 8, H*W %% 8 == 0, up 16-byte aligned");
  if (vp->radius_dtype != RCV_F32) FAIL(c, RCV_E_INVALID, "rcv_head_vote_frames: the head's radius maps are float32");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const int n_items = n_frames * n_kpts;
  const int words = (int)((npx + 255) / 256) * 8;
  const long long need = (long long)n_items * words;
  if (need > c->mask_words) {
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->mask_bits); c->mask_bits = nullptr; c->mask_words = 0;
    CK(c, cudaMalloc(&c->mask_bits, (size_t)need * 4));
    c->mask_words = need;
  }
  float* rad = radius_out;
  if (!rad) {   // context-owned radius planes (the vote only gathers them at the surviving pixels)
    if ((long long)n_items * npx > c->head_radius_cap) {
      CK(c, cudaStreamSynchronize(st));
      cudaFree(c->head_radius); c->head_radius = nullptr; c->head_radius_cap = 0;
      CK(c, cudaMalloc(&c->head_radius, (size_t)n_items * npx * 4));
      c->head_radius_cap = (long long)n_items * npx;
    }
    rad = c->head_radius;
  }
  CK(c, cudaMemsetAsync(c->cnt, 0, 4 * (size_t)n_items, st));
  CK(c, (cudaError_t)rcv_head1x1_fused_launch(up_bf16, weight, bias, rad, n_items, npx, c->sms, depth, fp->depth_dtype, n_kpts, max_radii,
                                              fp->max_radii_stride, fp->mask_flags, fp->sem_threshold, c->mask_bits, words, c->cnt, stream));
  c->last_mask_frames = n_frames; c->last_mask_kpts = n_kpts; c->last_mask_words_per_item = words;
  FrameArgs fa{depth, rad, nullptr, K, max_radii, *fp, vp->acc_unit, vp->radius_scale, n_kpts, 0};
  k_scan_items<<<1, 1024, 0, st>>>(c->cnt, n_items, c->pool.cap, c->meta);
  k_frame_emit<<<n_items, kCompactThreads, 0, st>>>(c->mask_bits, words, c->meta, c->pool.perm, words);
  k_points_from_pixels<<<dim3(n_items, 4), 256, 0, st>>>(fa, c->meta, c->pool);
  c->launches += 4;
  return run_items(c, n_items, vp, centre_mm, peak, votes, n_points, grid, nullptr, status, nullptr, 0, st);
}

// ---- K6: conv7 + BN + ReLU + conv8 (+ mask rule) in one kernel (conv7head.cu) ----
extern "C" int rcv_conv7_head_launch(const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                                     const float* b8, float* out, int n_images, int H, int W, int sms, void* stream);
extern "C" int rcv_conv7_head_fused_launch(const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                                           const float* b8, float* radius_out, int n_frames, int H, int W, int sms, const void* depth,
                                           int depth_dtype, int n_kpts, int kp, const double* max_radii, int max_radii_stride, int flags,
                                           float sem_threshold, unsigned* bits, int words_per_item, int* cnt, void* stream);

RCV_EXPORT int rcv_conv7_head(rcv_ctx* c, const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                              const float* b8, float* out, int n_images, int height, int width, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!x_nhwc_bf16 || !w7 || !bn_scale || !bn_shift || !w8 || !b8 || !out || n_images <= 0 || height <= 0 || width <= 0)
    FAIL(c, RCV_E_INVALID, "rcv_conv7_head: bad argument");
  if (width % 128 != 0 || ((uintptr_t)x_nhwc_bf16 & 15) != 0) FAIL(c, RCV_E_INVALID, "rcv_conv7_head: width must be a multiple of 128 and x 16-byte aligned");
  CK(c, cudaSetDevice(c->device));
  const int rc = rcv_conv7_head_launch(x_nhwc_bf16, w7, bn_scale, bn_shift, w8, b8, out, n_images, height, width, c->sms, stream);
  if (rc != 0) FAIL(c, RCV_E_CUDA, "rcv_conv7_head: launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  c->launches += 1;
  return RCV_OK;
}

RCV_EXPORT int rcv_conv7_head_vote_frames(rcv_ctx* c, int n_frames, int n_kpts, const void* const* x_nhwc_bf16, const float* w7, const float* bn_scale,
                                          const float* bn_shift, const float* w8, const float* b8, const void* depth, const double* K,
                                          const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp, double* centre_mm,
                                          int* peak, long long* votes, int* n_points, int* grid, int* status, float* radius_out, void* stream) {
  if (!c) return RCV_E_INVALID;
  if (!x_nhwc_bf16 || !w7 || !bn_scale || !bn_shift || !w8 || !b8 || !depth || !K || !centre_mm || !status)
    FAIL(c, RCV_E_INVALID, "rcv_conv7_head_vote_frames: null pointer");
  int rc = check_vote_params(c, vp);
  if (rc) return rc;
  rc = check_frame_params(c, n_frames, n_kpts, fp, true /* the seg values come from the head */, max_radii);
  if (rc) return rc;
  const long long npx = (long long)fp->height * fp->width;
  if (n_kpts > 8 || fp->width % 128 != 0) FAIL(c, RCV_E_INVALID, "rcv_conv7_head_vote_frames: n_kpts <= 8, width %% 128 == 0");
  for (int k = 0; k < n_kpts; ++k)
    if (!x_nhwc_bf16[k] || ((uintptr_t)x_nhwc_bf16[k] & 15)) FAIL(c, RCV_E_INVALID, "rcv_conv7_head_vote_frames: x[%d] null or not 16-byte aligned", k);
  if (vp->radius_dtype != RCV_F32) FAIL(c, RCV_E_INVALID, "rcv_conv7_head_vote_frames: the head's radius maps are float32");
  cudaStream_t st = (cudaStream_t)stream;
  CK(c, cudaSetDevice(c->device));
  const int n_items = n_frames * n_kpts;
  const int words = (int)((npx + 255) / 256) * 8;
  const long long need = (long long)n_items * words;
  if (need > c->mask_words) {
    CK(c, cudaStreamSynchronize(st));
    cudaFree(c->mask_bits); c->mask_bits = nullptr; c->mask_words = 0;
    CK(c, cudaMalloc(&c->mask_bits, (size_t)need * 4));
    c->mask_words = need;
  }
  float* rad = radius_out;
  if (!rad) {   // context-owned radius planes (the vote only gathers them at the surviving pixels)
    if ((long long)n_items * npx > c->head_radius_cap) {
      CK(c, cudaStreamSynchronize(st));
      cudaFree(c->head_radius); c->head_radius = nullptr; c->head_radius_cap = 0;
      CK(c, cudaMalloc(&c->head_radius, (size_t)n_items * npx * 4));
      c->head_radius_cap = (long long)n_items * npx;
    }
    rad = c->head_radius;
  }
  CK(c, cudaMemsetAsync(c->cnt, 0, 4 * (size_t)n_items, st));
  for (int k = 0; k < n_kpts; ++k)
    CK(c, (cudaError_t)rcv_conv7_head_fused_launch(x_nhwc_bf16[k], w7 + (size_t)k * 32 * 64 * 9, bn_scale + k * 32, bn_shift + k * 32, w8 + k * 64, b8 + k * 2,
                                                   rad, n_frames, fp->height, fp->width, c->sms, depth, fp->depth_dtype, n_kpts, k, max_radii,
                                                   fp->max_radii_stride, fp->mask_flags, fp->sem_threshold, c->mask_bits, words, c->cnt, stream));
  c->last_mask_frames = n_frames; c->last_mask_kpts = n_kpts; c->last_mask_words_per_item = words;
  FrameArgs fa{depth, rad, nullptr, K, max_radii, *fp, vp->acc_unit, vp->radius_scale, n_kpts, 0};
  k_scan_items<<<1, 1024, 0, st>>>(c->cnt, n_items, c->pool.cap, c->meta);
  k_frame_emit<<<n_items, kCompactThreads, 0, st>>>(c->mask_bits, words, c->meta, c->pool.perm, words);
  k_points_from_pixels<<<dim3(n_items, 4), 256, 0, st>>>(fa, c->meta, c->pool);
  c->launches += 3 + n_kpts;
  return run_items(c, n_items, vp, centre_mm, peak, votes, n_points, grid, nullptr, status, nullptr, 0, st);
}

RCV_EXPORT int rcv_vote_kernel_times(rcv_ctx* c, float* ms_out, int n) {
  if (!c || !ms_out || n <= 0) return RCV_E_INVALID;
  if (n > 64) n = 64;
  if ((long long)n > c->ev_count) n = (int)c->ev_count;
  for (int q = 0; q < n; ++q) {
    const int slot = (int)((c->ev_count - n + q) % 64);
    CK(c, cudaEventSynchronize(c->evr[slot][1]));
    CK(c, cudaEventElapsedTime(&ms_out[q], c->evr[slot][0], c->evr[slot][1]));
  }
  return n;
}

RCV_EXPORT float rcv_last_vote_kernel_ms(rcv_ctx* c) {
  float ms = -1.f;
  if (!c || rcv_vote_kernel_times(c, &ms, 1) != 1) return -1.f;
  return ms;
}

RCV_EXPORT int rcv_ubench_smem_atomics(rcv_ctx* c, double* atomics_per_second) {
  if (!c || !atomics_per_second) return RCV_E_INVALID;
  CK(c, cudaSetDevice(c->device));
  unsigned* out = reinterpret_cast<unsigned*>(c->counters + 8);
  const int iters = 4096;
  cudaEvent_t a = c->ev_ubench[0], b = c->ev_ubench[1];   // its own pair: the ring evr[] belongs to the vote kernel's timings
  float best = 1e30f;
  for (int r = 0; r < 7; ++r) {
    CK(c, cudaEventRecord(a, 0));
    k_ubench_atoms<<<c->sms, 1024, 32768 * 4, 0>>>(iters, out);
    CK(c, cudaEventRecord(b, 0));
    CK(c, cudaEventSynchronize(b));
    float ms;
    CK(c, cudaEventElapsedTime(&ms, a, b));
    if (r >= 2 && ms < best) best = ms;
  }
  c->launches += 7;
  *atomics_per_second = (double)c->sms * 1024.0 * iters * 8.0 / (best * 1e-3);
  return RCV_OK;
}
