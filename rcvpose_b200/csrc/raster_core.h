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
  if (!L.active) L.thr = 3.0e38f;   // an inactive lane owns nothing
}

// Fast-path vote decision of one candidate: sure = residual strictly inside the shell by more than eps,
// own = the candidate belongs to this pass, inb = inside the tile.  Returns the offset to increment:
// the voxel's or, when there is no vote, the lane's private sink.
RCV_HD int pick_offset(float q, float adv, int vrel, int off, int sink, float hw_m, float thr, int vn) {
#if defined(__CUDA_ARCH__)
  int r;
  asm("{\n\t.reg .pred p;\n\t.reg .f32 aq;\n\t"
      "abs.f32 aq, %1;\n\t"
      "setp.gt.f32 p, %2, %3;\n\t"
      "setp.lt.and.f32 p, aq, %4, p;\n\t"
      "setp.lt.and.u32 p, %5, %6, p;\n\t"
      "selp.b32 %0, %7, %8, p;\n\t}"
      : "=r"(r) : "f"(q), "f"(adv), "f"(thr), "f"(hw_m), "r"(vrel), "r"(vn), "r"(off), "r"(sink));
  return r;
#else
  return ((adv > thr) && (fabsf(q) < hw_m) && ((unsigned)vrel < (unsigned)vn)) ? off : sink;
#endif
}

// Slow path of ring_lane for one arc's candidate `cc` whose float32 residual q is not decisive.
// (fl, vt) describe the arc: fl = float offset of the top candidate from the lattice base (mirrored for
// the bottom arc), vt = its index along the candidate axis relative to the tile.
template <class Slow, class EmitSlow>
RCV_HD void ring_slow(const PointCtx& c, const LaneTask& L, int i, int ub, int m, int arc, int cc, float q, float fl, int vt, Slow& slow,
                      EmitSlow& emit_slow) {
  const int v = arc ? (vt + cc) : (vt - cc), dv = v - L.vrel0;
  if (((unsigned)v < (unsigned)L.vn) && slow(i, L.pass ? (c.ipy + dv) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv))) emit_slow(ub + v * L.sv);
  if (cc == 0 && q >= c.hw_m) {
    // The top candidate may lie outside the outer sphere; the run can then reach one voxel lower.
    const int v2 = arc ? (vt + m) : (vt - m), dv2 = v2 - L.vrel0;
    const float d2 = arc ? f_add(f_sub(fl, (float)m), L.fv) : f_sub(f_sub(fl, (float)m), L.fv);
    if ((fabsf(d2) > L.thr) && ((unsigned)v2 < (unsigned)L.vn) &&
        slow(i, L.pass ? (c.ipy + dv2) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv2)))
      emit_slow(ub + v2 * L.sv);
  }
}

// ---- thin ring slices (one candidate per arc): fast part and rare part split so that the kernel can
// interleave the fast parts of two slices (two independent dependency chains each) before a single
// rarely-taken branch.
struct ThinOut {
  float q0, q1, fl0, fl1;
  int vt0, vt1, ub;
  bool t0, t1;   // candidate of the top / bottom arc needs the exact path
};

//   emit(offset) -- the unconditional shared-memory atomic of the fast path, called exactly twice; the
//                   offset is the voxel's, or `sink` when there is no vote.
// Offsets are in the units lane_setup() was given (words on the host, bytes on the device);
// `slice_base` = (i - i0) * nj * Dp in the same units.
template <class Emit>
RCV_HD void thin_fast(const PointCtx& c, float a, const LaneTask& L, int slice_base, int sink, Emit& emit, ThinOut& o) {
  const float g = f_sub(a, L.du2);
  const float zs = f_sqrt_fast(fmaxf(g, 0.f));
  const float hWg = f_sub(c.hW, g);        // g <= 0  =>  q >= W/2: never a fast-path vote
  o.ub = L.ubase + slice_base;
  const float tm0 = f_add(f_add(f_add(zs, L.fv), c.dbias_m05), RCV_MAGIC);   // top arc
  const float tm1 = f_add(f_add(f_sub(zs, L.fv), c.dbias_m05), RCV_MAGIC);   // bottom arc, mirrored
  o.fl0 = f_sub(tm0, RCV_MAGIC); o.fl1 = f_sub(tm1, RCV_MAGIC);
  o.vt0 = L.vrel0 + (f_bits(tm0) - RCV_MAGIC_BITS);   // topmost candidate (>= true topmost voxel under the outer circle)
  o.vt1 = L.vrel0 - (f_bits(tm1) - RCV_MAGIC_BITS);
  const float d0 = f_sub(o.fl0, L.fv), d1 = f_add(o.fl1, L.fv);              // dv of the two candidates
  o.q0 = f_fma(d0, d0, hWg); o.q1 = f_fma(d1, d1, hWg);                      // e + W/2, e = |v-p|^2 - R^2 in float32
  emit(pick_offset(o.q0, fabsf(d0), o.vt0, o.ub + o.vt0 * L.sv, sink, c.hw_m, L.thr, L.vn));
  emit(pick_offset(o.q1, fabsf(d1), o.vt1, o.ub + o.vt1 * L.sv, sink, c.hw_m, L.thr, L.vn));
  // not surely inside the shell, but not surely beyond the inner boundary either: decide exactly
  o.t0 = (fabsf(d0) > L.thr) && !(fabsf(o.q0) < c.hw_m) && (o.q0 > -c.hw_p);
  o.t1 = (fabsf(d1) > L.thr) && !(fabsf(o.q1) < c.hw_m) && (o.q1 > -c.hw_p);
}

template <class SlowArc>
RCV_HD void thin_slow(const PointCtx& c, const LaneTask& L, int i, const ThinOut& o, SlowArc& slowarc) {
  if (o.t0) slowarc(c, L, i, o.ub, 1, 0, 0, o.q0, o.fl0, o.vt0);
  if (o.t1) slowarc(c, L, i, o.ub, 1, 1, 0, o.q1, o.fl1, o.vt1);
}

// One lane, one ring slice with outer radius^2 `a` and `m` >= 1 candidates per arc (general form).
//   slowarc(...) -- ring_slow() behind a call: decides candidates whose float32 residual is within eps
//                   of a shell boundary with the exact float64 predicate (rare).
template <class Emit, class SlowArc>
RCV_HD void ring_lane(const PointCtx& c, float a, int m, const LaneTask& L, int i, int slice_base, int sink, Emit& emit, SlowArc& slowarc) {
  const float g = f_sub(a, L.du2);
  const float zs = f_sqrt_fast(fmaxf(g, 0.f));
  const float hWg = f_sub(c.hW, g);
  const int ub = L.ubase + slice_base;
  const float tm0 = f_add(f_add(f_add(zs, L.fv), c.dbias_m05), RCV_MAGIC);
  const float tm1 = f_add(f_add(f_sub(zs, L.fv), c.dbias_m05), RCV_MAGIC);
  const float fl0 = f_sub(tm0, RCV_MAGIC), fl1 = f_sub(tm1, RCV_MAGIC);
  const int vt0 = L.vrel0 + (f_bits(tm0) - RCV_MAGIC_BITS);
  const int vt1 = L.vrel0 - (f_bits(tm1) - RCV_MAGIC_BITS);
#pragma unroll 1
  for (int cc = 0; cc < m; ++cc) {
    const float fc = (float)cc;
    const float d0 = f_sub(f_sub(fl0, fc), L.fv), d1 = f_add(f_sub(fl1, fc), L.fv);
    const float q0 = f_fma(d0, d0, hWg), q1 = f_fma(d1, d1, hWg);
    const int v0 = vt0 - cc, v1 = vt1 + cc;
    emit(pick_offset(q0, fabsf(d0), v0, ub + v0 * L.sv, sink, c.hw_m, L.thr, L.vn));
    emit(pick_offset(q1, fabsf(d1), v1, ub + v1 * L.sv, sink, c.hw_m, L.thr, L.vn));
    const bool t0 = (fabsf(d0) > L.thr) && !(fabsf(q0) < c.hw_m) && (q0 > -c.hw_p);
    const bool t1 = (fabsf(d1) > L.thr) && !(fabsf(q1) < c.hw_m) && (q1 > -c.hw_p);
    if (t0 || t1) {
      if (t0) slowarc(c, L, i, ub, m, 0, cc, q0, fl0, vt0);
      if (t1) slowarc(c, L, i, ub, m, 1, cc, q1, fl1, vt1);
    }
  }
}

// One lane, one cell (dj, dk) of a dense slice's bounding box.
template <class Emit, class Slow>
RCV_HD void dense_cell(const PointCtx& c, float a, const Tile& t, int i, int slice_base, int unit, int sink, int dj, int dk, bool cell_ok,
                       Emit& emit, Slow& slow) {
  const float dyf = f_sub((float)dj, c.fy), dzf = f_sub((float)dk, c.fz);
  const float q = f_add(f_fma(dzf, dzf, f_fma(dyf, dyf, -a)), c.hW);
  const int vj = c.ipy + dj, vk = c.ipz + dk;
  const bool inb = cell_ok && ((unsigned)(vj - t.j0) < (unsigned)t.nj) && ((unsigned)vk < (unsigned)t.D);
  const bool sure = fabsf(q) < c.hw_m;
  const bool amb = !sure && (fabsf(q) <= c.hw_p);
  bool vote = sure && inb;
  if (amb && inb) vote = slow(i, vj, vk);
  emit(vote ? slice_base + ((vj - t.j0) * t.Dp + vk) * unit : sink);
}

// Slice range of point c inside tile t (inclusive); empty if ia > ib.
RCV_HD void slice_range(const PointCtx& c, const Tile& t, int& ia, int& ib) {
  ia = c.ipx - c.R - 1; if (ia < t.i0) ia = t.i0;
  ib = c.ipx + c.R + 1; if (ib > t.i0 + t.ni - 1) ib = t.i0 + t.ni - 1;
  if (c.R <= 0) { ia = 1; ib = 0; }
}

}  // namespace rcv
