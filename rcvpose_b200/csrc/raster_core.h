// raster_core.h -- per-lane arithmetic of the sphere-shell vote rasteriser (kernel K2).
//
// What it renders.  The reference's vote loop (AccumulatorSpace.py:325-341) increments voxel
// (i,j,k) for point p with integer radius R iff
//        0 < R - sqrt((i-px)^2 + (j-py)^2 + (k-pz)^2) < sqrt(3)/4        (float64, strict)
// evaluated for EVERY voxel of the D^3 cube.  This file emits exactly that voxel set by scattering:
// for the x-slice i of a sphere the set is a ring in the (j,k) plane with outer radius^2
// a = R^2 - dx^2 and inner radius^2 b = a - W, W = R^2 - (R - sqrt3/4)^2.  The ring is split into
// four arcs by the dominant axis of (dy,dz):
//     Z-pass: one lane per column j, candidates k = topmost voxel under the outer circle and the
//             m-1 below it (mirror image for the bottom arc); owns voxels with |dz| >= |dy|;
//     Y-pass: one lane per row k, candidates along j; owns |dy| > |dz|.
// In its own pass an arc crosses each column in fewer than m voxels (m = 1 for most slices), so a
// lane tests m candidates per arc and there is no data-dependent loop.  Rings too close to the pole
// of the sphere (b < 4 or a <= 36) are rendered densely over their small bounding box instead.
//
// Exactness.  Candidates are classified in float32 with a proven error bound eps; a candidate whose
// float32 residual lies within eps of either shell boundary is re-decided by exact_hit(), which
// repeats the reference's float64 operation sequence (no FMA).  The voxel set is therefore the
// reference's, bit for bit; tests/test_raster_hostsim.py fuzzes this file (compiled for the host)
// against the brute-force oracle, and the -m gpu tests check the CUDA build the same way.
//
// The header is shared by the CUDA kernel (rcvvote.cu) and the host-side test harness
// (tests/hostsim.cpp): all float arithmetic goes through the RCV_F* wrappers, which are the
// never-contracted intrinsics on the device and plain IEEE operations (-ffp-contract=off) on the host.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RCV_HD __host__ __device__ __forceinline__
#else
#define RCV_HD inline
#endif

namespace rcv {

// 3**0.5/4 as float64 (AccumulatorSpace.py:328): 0x1.bb67ae8584caap-2
#define RCV_SHELL 0.4330127018922193

#if defined(__CUDA_ARCH__)
RCV_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
RCV_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
RCV_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
RCV_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RCV_HD float f_sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RCV_HD int f_bits(float x) { return __float_as_int(x); }
RCV_HD float f_from_bits(int x) { return __int_as_float(x); }
RCV_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
RCV_HD double d_sub(double a, double b) { return __dsub_rn(a, b); }
RCV_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
RCV_HD double d_sqrt(double a) { return __dsqrt_rn(a); }
RCV_HD int d_rint(double a) { return __double2int_rn(a); }
#else
// Host build (tests only).  g_sqrt_perturb lets the fuzz tests emulate the <=2-ulp error of the
// device's MUFU.SQRT so the exactness argument is exercised, not just IEEE sqrtf.
static int g_sqrt_perturb = 0;
static uint32_t g_sqrt_rng = 12345u;
RCV_HD float f_add(float a, float b) { return a + b; }
RCV_HD float f_sub(float a, float b) { return a - b; }
RCV_HD float f_mul(float a, float b) { return a * b; }
RCV_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
RCV_HD int f_bits(float x) { int i; memcpy(&i, &x, 4); return i; }
RCV_HD float f_from_bits(int x) { float f; memcpy(&f, &x, 4); return f; }
RCV_HD float f_sqrt_fast(float x) {
  float r = sqrtf(x);
  if (g_sqrt_perturb && r > 0.f) {
    g_sqrt_rng = g_sqrt_rng * 1664525u + 1013904223u;
    r = f_from_bits(f_bits(r) + (int)((g_sqrt_rng >> 16) % 5u) - 2);
  }
  return r;
}
RCV_HD double d_add(double a, double b) { return a + b; }
RCV_HD double d_sub(double a, double b) { return a - b; }
RCV_HD double d_mul(double a, double b) { return a * b; }
RCV_HD double d_sqrt(double a) { return sqrt(a); }
RCV_HD int d_rint(double a) { return (int)nearbyint(a); }
#endif

#define RCV_MAGIC 12582912.0f        // 1.5 * 2^23: adding it rounds a float to the nearest integer
#define RCV_MAGIC_BITS 0x4B400000

// The reference predicate, operation for operation (AccumulatorSpace.py:337-338).
RCV_HD bool exact_hit(double px, double py, double pz, int R, int i, int j, int k) {
  const double dx = d_sub((double)i, px), dy = d_sub((double)j, py), dz = d_sub((double)k, pz);
  const double s = d_add(d_add(d_mul(dx, dx), d_mul(dy, dy)), d_mul(dz, dz));
  const double t = d_sub((double)R, d_sqrt(s));
  return (t < RCV_SHELL) && (t > 0.0);
}

// A tile of the accumulator owned by one CTA: slices [i0,i0+ni) x rows [j0,j0+nj) x all k in [0,D).
// Word offset of voxel (i,j,k) = ((i-i0)*nj + (j-j0))*Dp + k, Dp >= D odd (bank spreading).
struct Tile {
  int i0, ni, j0, nj, D, Dp;
};

// Per-point constants (warp-uniform).
struct PointCtx {
  double px, py, pz;  // exact shifted voxel-unit coordinates (what the reference passes to fast_for)
  int R;
  int ipx, ipy, ipz;  // nearest lattice point
  float fx, fy, fz;   // p - ip, |f| <= 0.5
  float R2, W, eps, dbias_m05;
  float hW, hw_m, hw_p;  // W/2, W/2 - eps, W/2 + eps
};

RCV_HD void point_setup(PointCtx& c, double px, double py, double pz, int R) {
  c.px = px; c.py = py; c.pz = pz; c.R = R;
  c.ipx = d_rint(px); c.ipy = d_rint(py); c.ipz = d_rint(pz);
  c.fx = (float)d_sub(px, (double)c.ipx);
  c.fy = (float)d_sub(py, (double)c.ipy);
  c.fz = (float)d_sub(pz, (double)c.ipz);
  c.R2 = (float)(R * R);  // exact for R < 4096
  const double rin = (double)R - RCV_SHELL;
  c.W = (float)((double)R * (double)R - rin * rin);
  // |e_float32 - e_exact| <= 2^-21 (R+2)^2 ("error bound"); eps carries a 2x margin.
  const float rp2 = (float)(R + 2);
  c.eps = f_mul(f_mul(rp2, rp2), 9.5367431640625e-07f);        // 2^-20
  const float dbias = f_add(f_mul(c.eps, 0.125f), f_mul(rp2, 9.5367431640625e-07f));
  c.dbias_m05 = f_sub(dbias, 0.5f);
  c.hW = f_mul(c.W, 0.5f);
  c.hw_m = f_sub(c.hW, c.eps);
  c.hw_p = f_add(c.hW, c.eps);
}

// Per-(point, slice) classification (warp-uniform; the kernel evaluates 32 slices at once, one per
// lane, and broadcasts (a, code) with shuffles).  code > 0: ring slice, code = candidates per arc;
// code < 0: dense slice, -code = half-width of its bounding box; code == 0: nothing to draw.
RCV_HD void slice_setup(const PointCtx& c, int i, float& a, int& code) {
  const float dxf = f_sub((float)(i - c.ipx), c.fx);
  a = f_sub(c.R2, f_mul(dxf, dxf));
  const float b = f_sub(a, c.W);
  if (!(a > -c.eps)) { code = 0; return; }
  if (b < 4.0f || a <= 36.0f) {
    code = -(int)f_add(f_sqrt_fast(fmaxf(a, 0.f)), 1.5f);  // floor(ro + 1.5) >= ro + 0.5, with slack
    return;
  }
  // Longest run of an arc inside its own pass: sqrt(a - b/2) - sqrt(b/2)  (column |du| = sqrt(b/2)).
  const float hb = f_mul(b, 0.5f);
  const float lmax = f_sub(f_sqrt_fast(f_sub(a, hb)), f_sqrt_fast(hb));
  code = (int)f_add(lmax, 0.02f) + 1;
}

// Per-lane task: one column (Z-pass) or row (Y-pass) of the ring, both arcs.
struct LaneTask {
  float du2, thr, fv;
  int ubase;   // offset of (lane coordinate, candidate coordinate 0) inside a slice of the tile, in `unit`s
  int sv;      // stride of the candidate axis (1 word for Z-pass, Dp words for Y-pass), in `unit`s
  int vrel0;   // lattice base of the candidate axis relative to the tile origin
  int vn;      // extent of the candidate axis in the tile
  int ucoord;  // global index along the lane axis
  bool pass;   // false: Z-pass (lane axis j, candidates along k); true: Y-pass
  bool active;
};

// Half-width (in lanes) needed so that every owned voxel of a ring with outer radius^2 <= amax
// lies in a lane: owned => du^2 <= dv^2 and du^2 + dv^2 < a  =>  |du| < sqrt(a/2); |u| <= |du| + 0.5.
RCV_HD int ring_half_width(float amax) { return (int)f_add(f_sqrt_fast(f_mul(amax, 0.5f)), 0.5f) + 1; }

RCV_HD void lane_setup(const PointCtx& c, const Tile& t, int H, int tau, int unit, LaneTask& L) {
  const int Wc = 2 * H + 1;
  L.pass = tau >= Wc;
  const int u = tau - (L.pass ? Wc : 0) - H;
  const float fu = L.pass ? c.fz : c.fy;
  L.fv = L.pass ? c.fy : c.fz;
  const float duf = f_sub((float)u, fu);
  L.du2 = f_mul(duf, duf);
  const float ad = fabsf(duf);
  // Z-pass owns |dv| >= |du|  <=>  |dv| > pred(|du|); Y-pass owns |dv| > |du|: complementary.
  L.thr = L.pass ? ad : (ad > 0.f ? f_from_bits(f_bits(ad) - 1) : -1.0f);
  bool ok;
  if (!L.pass) {
    L.ucoord = c.ipy + u;
    ok = (unsigned)(L.ucoord - t.j0) < (unsigned)t.nj;
    L.ubase = (L.ucoord - t.j0) * t.Dp * unit;
    L.sv = unit; L.vrel0 = c.ipz; L.vn = t.D;
  } else {
    L.ucoord = c.ipz + u;
    ok = (unsigned)L.ucoord < (unsigned)t.D;
    L.ubase = L.ucoord * unit;
    L.sv = t.Dp * unit; L.vrel0 = c.ipy - t.j0; L.vn = t.nj;
  }
  L.active = ok && tau < 2 * Wc;
}

// One lane, one ring slice with outer radius^2 `a` and `m` candidates per arc (THIN: m == 1).
//   emit(offset, vote)     -- called exactly 2*m times (vote may be false): the unconditional
//                             shared-memory atomic of the fast path;
//   slow(i, j, k) -> bool  -- the exact float64 predicate, called only for candidates whose float32
//                             residual is within eps of a shell boundary (rare);
//   emit_slow(offset)      -- a vote decided on the slow path.
// Offsets are in the units lane_setup() was given (words on the host, bytes on the device);
// `slice_base` = (i - i0) * nj * Dp in the same units.
template <bool THIN, class Emit, class Slow, class EmitSlow>
RCV_HD void ring_lane(const PointCtx& c, float a, int m, const LaneTask& L, int i, int slice_base, Emit& emit, Slow& slow,
                      EmitSlow& emit_slow) {
  const float g = f_sub(a, L.du2);
  const bool lane_ok = L.active && (g > 0.f);
  const float zs = f_sqrt_fast(fmaxf(g, 0.f));
  const float hWg = f_sub(c.hW, g);
  const int ub = L.ubase + slice_base;
  if (THIN) m = 1;
#pragma unroll
  for (int arc = 0; arc < 2; ++arc) {
    const float fvs = arc ? -L.fv : L.fv;
    const float tm = f_add(f_add(f_add(zs, fvs), c.dbias_m05), RCV_MAGIC);
    const int kq = f_bits(tm) - RCV_MAGIC_BITS;  // topmost candidate (>= true topmost voxel inside the outer circle)
    const float flr = f_sub(tm, RCV_MAGIC);
    const int vtop = arc ? (L.vrel0 - kq) : (L.vrel0 + kq);
#pragma unroll 1
    for (int cc = 0; cc < m; ++cc) {
      const float dvfs = f_sub(THIN ? flr : f_sub(flr, (float)cc), fvs);
      const float q = f_fma(dvfs, dvfs, hWg);              // e + W/2, e = |v-p|^2 - R^2 in float32
      const bool own = fabsf(dvfs) > L.thr;
      const int vrel = arc ? (vtop + cc) : (vtop - cc);
      const bool inb = (unsigned)vrel < (unsigned)L.vn;
      const bool sure = fabsf(q) < c.hw_m;                 // -W + eps < e < -eps
      emit(ub + vrel * L.sv, sure && own && inb && lane_ok);
      if (!sure && (q > -c.hw_p) && lane_ok) {
        // within eps of the inner boundary, or not surely inside the outer one: decide exactly
        const int dv = arc ? (cc - kq) : (kq - cc);
        if (own && inb && slow(i, L.pass ? (c.ipy + dv) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv))) emit_slow(ub + vrel * L.sv);
        if (cc == 0 && q >= c.hw_m) {
          // The top candidate may lie outside the outer sphere; the run can then reach one voxel lower.
          const int dv2 = arc ? (m - kq) : (kq - m);
          const float dvfs2 = f_sub(f_sub(flr, (float)m), fvs);
          const int vrel2 = L.vrel0 + dv2;
          if ((fabsf(dvfs2) > L.thr) && ((unsigned)vrel2 < (unsigned)L.vn) &&
              slow(i, L.pass ? (c.ipy + dv2) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv2)))
            emit_slow(ub + vrel2 * L.sv);
        }
      }
      if (THIN) break;
    }
  }
}

// One lane, one cell (dj, dk) of a dense slice's bounding box.
template <class Emit, class Slow>
RCV_HD void dense_cell(const PointCtx& c, float a, const Tile& t, int i, int slice_base, int unit, int dj, int dk, bool cell_ok,
                       Emit& emit, Slow& slow) {
  const float dyf = f_sub((float)dj, c.fy), dzf = f_sub((float)dk, c.fz);
  const float q = f_add(f_fma(dzf, dzf, f_fma(dyf, dyf, -a)), c.hW);
  const int vj = c.ipy + dj, vk = c.ipz + dk;
  const bool inb = cell_ok && ((unsigned)(vj - t.j0) < (unsigned)t.nj) && ((unsigned)vk < (unsigned)t.D);
  const bool sure = fabsf(q) < c.hw_m;
  const bool amb = !sure && (fabsf(q) <= c.hw_p);
  bool vote = sure && inb;
  if (amb && inb) vote = slow(i, vj, vk);
  emit(slice_base + ((vj - t.j0) * t.Dp + vk) * unit, vote);
}

// Slice range of point c inside tile t (inclusive); empty if ia > ib.
RCV_HD void slice_range(const PointCtx& c, const Tile& t, int& ia, int& ib) {
  ia = c.ipx - c.R - 1; if (ia < t.i0) ia = t.i0;
  ib = c.ipx + c.R + 1; if (ib > t.i0 + t.ni - 1) ib = t.i0 + t.ni - 1;
  if (c.R <= 0) { ia = 1; ib = 0; }
}

}  // namespace rcv
