// raster_core.h -- per-lane arithmetic of the sphere-shell vote rasteriser (kernel K2).
//
// What it renders.  The reference's vote loop (AccumulatorSpace.py:325-341) increments voxel
// (i,j,k) for point p with integer radius R iff
//        0 < R - sqrt((i-px)^2 + (j-py)^2 + (k-pz)^2) < sqrt(3)/4        (float64, strict)
// evaluated for EVERY voxel of the D^3 cube.  This file emits exactly that voxel set by scattering.
// One lane owns one point.  Axes here are the rasteriser's own (A,B,C) = (x,y,z) of this file; the
// kernel feeds it a permutation of the reference's axes (see rcvvote.cu).  For the A-slice i of a
// sphere the set is a ring in the (B,C) plane with outer radius^2 a = R^2 - dx^2 and inner radius^2
// b = a - W, W = R^2 - (R - sqrt3/4)^2.
//
//   thin rings (slice_setup code 1: no arc crosses a column of its own pass in more than one voxel,
//   about 80 % of a sphere's slices) are drawn by two ring passes.  The ring is split into four arcs
//   by the dominant axis of (dy,dz):
//     Z-pass: one task per column j, candidate k = topmost voxel under the outer circle (mirror image
//             for the bottom arc); owns voxels with |dz| >= |dy|;
//     Y-pass: one task per row k, candidates along j; owns |dy| > |dz|.
//   all other slices (the two polar caps of the slice axis) are drawn by the polar pass: for
//   R >= RCV_POLAR_MIN_R a column ALONG the slice axis crosses the shell of such a slice in less than
//   one voxel, so the thin-ring arithmetic applies with the axes rotated; a candidate votes iff its
//   slice is one of the lane's non-thin slices (bit mask), so the two kinds of pass partition the
//   shell exactly.  Spheres with R < RCV_POLAR_MIN_R are scanned densely over their bounding box.
//
// Exactness.  Candidates are classified in float32 with a proven error bound eps; a candidate whose
// float32 residual lies within eps of either shell boundary is re-decided by exact_hit(), which
// repeats the reference's float64 operation sequence (no FMA).  The voxel set is therefore the
// reference's, bit for bit; tests/test_raster_hostsim.py fuzzes this file (compiled for the host)
// against the brute-force oracle, and the -m gpu tests check the CUDA build the same way.
//
// The header is shared by the CUDA kernel (rcvvote.cu) and the host-side test harness
// (tests/hostsim.cpp): all float arithmetic goes through the f_* / d_* wrappers, which are the
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
RCV_HD float f_sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // NaN for x < 0
RCV_HD int f_bits(float x) { return __float_as_int(x); }
RCV_HD float f_from_bits(int x) { return __int_as_float(x); }
RCV_HD unsigned u_shr(unsigned m, int v) { unsigned r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(v)); return r; }   // 0 for v outside [0,32)
RCV_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
RCV_HD double d_sub(double a, double b) { return __dsub_rn(a, b); }
RCV_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
RCV_HD double d_sqrt(double a) { return __dsqrt_rn(a); }
RCV_HD int d_rint(double a) { return __double2int_rn(a); }
#else
// Host build (tests only).  g_sqrt_perturb lets the fuzz tests emulate the <=2-ulp error of the
// device's MUFU.SQRT so the exactness argument is exercised, not just IEEE sqrtf.
static int g_sqrt_perturb = 0;
RCV_HD float f_add(float a, float b) { return a + b; }
RCV_HD float f_sub(float a, float b) { return a - b; }
RCV_HD float f_mul(float a, float b) { return a * b; }
RCV_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
RCV_HD int f_bits(float x) { int i; memcpy(&i, &x, 4); return i; }
RCV_HD float f_from_bits(int x) { float f; memcpy(&f, &x, 4); return f; }
RCV_HD float f_sqrt_fast(float x) {
  float r = sqrtf(x);   // NaN for x < 0, like the device instruction
  if (g_sqrt_perturb && r > 0.f) {
    // deterministic in the argument (a fast path and the rare path that re-derives it must see the same value, as on the device)
    uint32_t h = (uint32_t)f_bits(x) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    r = f_from_bits(f_bits(r) + (int)(h % 5u) - 2);
  }
  return r;
}
RCV_HD unsigned u_shr(unsigned m, int v) { return ((unsigned)v < 32u) ? (m >> v) : 0u; }
RCV_HD double d_add(double a, double b) { return a + b; }
RCV_HD double d_sub(double a, double b) { return a - b; }
RCV_HD double d_mul(double a, double b) { return a * b; }
RCV_HD double d_sqrt(double a) { return sqrt(a); }
RCV_HD int d_rint(double a) { return (int)nearbyint(a); }
#endif

#define RCV_MAGIC 12582912.0f        // 1.5 * 2^23: adding it rounds a float to the nearest integer
#define RCV_MAGIC_BITS 0x4B400000
#define RCV_POLAR_MIN_R 7            // smallest R drawn by ring + polar passes (see polar_fast)
#ifndef RCV_RING2_MAX_CODE
#define RCV_RING2_MAX_CODE 2         // the fast ring pass draws rings of slice_setup code 1..2 (up to 2 candidates per arc)
#endif

// The reference predicate, operation for operation (AccumulatorSpace.py:337-338).
RCV_HD bool exact_hit(double px, double py, double pz, int R, int i, int j, int k) {
  const double dx = d_sub((double)i, px), dy = d_sub((double)j, py), dz = d_sub((double)k, pz);
  const double s = d_add(d_add(d_mul(dx, dx), d_mul(dy, dy)), d_mul(dz, dz));
  const double t = d_sub((double)R, d_sqrt(s));
  return (t < RCV_SHELL) && (t > 0.0);
}

// A tile of the accumulator owned by one CTA: slices [i0,i0+ni) x rows [j0,j0+nj) x all k in [0,D).
// Word offset of voxel (i,j,k) = ((i-i0)*nj + (j-j0))*Dp + k, Dp >= D odd (bank spreading).
// A guarded tile also holds glo / ghi guard cells below 0 / above D - 1 along C (the tile pointer is shifted by glo cells,
// Dp >= D + glo + ghi) and along B (j0 = -glo, nj = D + glo + ghi): votes of a sphere that overhangs the grid land in guard
// cells, which are never read back, so no candidate of such a tile needs a bounds test (ring_noclip).
struct Tile {
  int i0, ni, j0, nj, D, Dp;
  int glo, ghi;
};

// Per-point constants (one lane).
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
  // bias of the floor that picks the top candidate: covers the rounding of sqrt and of the additions before the floor
  const float dbias = f_add(f_mul(c.eps, 0.125f), f_mul(rp2, 9.5367431640625e-07f));
  c.dbias_m05 = f_sub(dbias, 0.5f);
  c.hW = f_mul(c.W, 0.5f);
  c.hw_m = f_sub(c.hW, c.eps);
  c.hw_p = f_add(c.hW, c.eps);
}

// Per-(point, slice) classification.  code == 1: thin ring; code > 1: thick ring (code = longest arc run);
// code < 0: small ring, -code = half-width of its bounding box; code == 0: nothing to draw.
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

// One column (Z-pass) or row (Y-pass) of a lane's rings, both arcs.
struct LaneTask {
  float du2, thr, fv;
  float cp, cm;  // fv + dbias - 0.5 and -fv + dbias - 0.5: what is added to the half-chord before the floor
  int ubase;     // offset of (lane coordinate, candidate coordinate 0) inside a slice of the tile, in `unit`s
  int sv;        // stride of the candidate axis (1 word for Z-pass, Dp words for Y-pass), in `unit`s
  int vrel0;     // lattice base of the candidate axis relative to the tile origin
  int vn;        // extent of the candidate axis in the tile
  int ucoord;    // global index along the lane axis
  bool pass;     // false: Z-pass (lane axis j, candidates along k); true: Y-pass
};

// Half-width (in columns) needed so that every owned voxel of a ring with outer radius^2 <= amax
// lies in a task: owned => du^2 <= dv^2 and du^2 + dv^2 < a  =>  |du| < sqrt(a/2); |u| <= |du| + 0.5.
RCV_HD int ring_half_width(float amax) { return (int)f_add(f_sqrt_fast(f_mul(amax, 0.5f)), 0.5f) + 1; }

// Task of column/row `u` (lattice offset from the point's nearest lattice point) in pass `pass`; uf == (float)u.
RCV_HD void lane_setup_pu(const PointCtx& c, const Tile& t, bool pass, int u, float uf, int unit, LaneTask& L) {
  L.pass = pass;
  const float fu = pass ? c.fz : c.fy;
  L.fv = pass ? c.fy : c.fz;
  L.cp = f_add(L.fv, c.dbias_m05);
  L.cm = f_sub(c.dbias_m05, L.fv);
  const float duf = f_sub(uf, fu);
  L.du2 = f_mul(duf, duf);
  const float ad = fabsf(duf);
  // Z-pass owns |dv| >= |du|  <=>  |dv| > pred(|du|); Y-pass owns |dv| > |du|: complementary.
  L.thr = pass ? ad : (ad > 0.f ? f_from_bits(f_bits(ad) - 1) : -1.0f);
  bool ok;
  if (!pass) {
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
  if (!ok) L.thr = 3.0e38f;   // a task outside the tile owns nothing
}

// Fast-path vote decision of one candidate: sure = residual strictly inside the shell by more than eps,
// own = the candidate belongs to this pass, inb = inside the tile (tested only if CLIP).  Returns the offset to
// increment: the voxel's or, when there is no vote, the lane's private sink.
template <bool CLIP>
RCV_HD int pick_offset(float q, float adv, int vrel, int off, int sink, float hw_m, float thr, int vn) {
#if defined(__CUDA_ARCH__)
  int r;
  if (CLIP)
    asm("{\n\t.reg .pred p;\n\t.reg .f32 aq;\n\t"
        "abs.f32 aq, %1;\n\t"
        "setp.gt.f32 p, %2, %3;\n\t"
        "setp.lt.and.f32 p, aq, %4, p;\n\t"
        "setp.lt.and.u32 p, %5, %6, p;\n\t"
        "selp.b32 %0, %7, %8, p;\n\t}"
        : "=r"(r) : "f"(q), "f"(adv), "f"(thr), "f"(hw_m), "r"(vrel), "r"(vn), "r"(off), "r"(sink));
  else
    asm("{\n\t.reg .pred p;\n\t.reg .f32 aq;\n\t"
        "abs.f32 aq, %1;\n\t"
        "setp.gt.f32 p, %2, %3;\n\t"
        "setp.lt.and.f32 p, aq, %4, p;\n\t"
        "selp.b32 %0, %5, %6, p;\n\t}"
        : "=r"(r) : "f"(q), "f"(adv), "f"(thr), "f"(hw_m), "r"(off), "r"(sink));
  return r;
#else
  return ((adv > thr) && (fabsf(q) < hw_m) && (!CLIP || ((unsigned)vrel < (unsigned)vn))) ? off : sink;
#endif
}

// Slow path for one arc's top candidate whose float32 residual q is not decisive.
// (fl, vt) describe the arc: fl = float offset of the top candidate from the lattice base (mirrored for
// the bottom arc), vt = its index along the candidate axis relative to the tile.
template <class Slow, class EmitSlow>
RCV_HD void ring_slow(const PointCtx& c, const LaneTask& L, int i, int ub, int arc, float q, float fl, int vt, Slow& slow, EmitSlow& emit_slow) {
  const int dv = vt - L.vrel0;
  if (((unsigned)vt < (unsigned)L.vn) && slow(i, L.pass ? (c.ipy + dv) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv))) emit_slow(ub + vt * L.sv);
  if (q >= c.hw_m) {
    // The top candidate may lie outside the outer sphere; the run can then reach one voxel lower.
    const int v2 = arc ? (vt + 1) : (vt - 1), dv2 = v2 - L.vrel0;
    const float d2 = arc ? f_add(f_sub(fl, 1.0f), L.fv) : f_sub(f_sub(fl, 1.0f), L.fv);
    if ((fabsf(d2) > L.thr) && ((unsigned)v2 < (unsigned)L.vn) &&
        slow(i, L.pass ? (c.ipy + dv2) : L.ucoord, L.pass ? L.ucoord : (c.ipz + dv2)))
      emit_slow(ub + v2 * L.sv);
  }
}

// ---- thin rings: fast part and rare part split so that the kernel can issue the fast parts of several
// slices (independent dependency chains) before a single rarely-taken branch.
struct ThinOut {
  float q0, q1, fl0, fl1;
  int vt0, vt1, ub;
  bool t0, t1;   // candidate of the top / bottom arc needs the exact path
};

//   emit(offset) -- the unconditional shared-memory atomic of the fast path, called exactly twice; the
//                   offset is the voxel's, or `sink` when there is no vote.
// Offsets are in the units lane_setup_pu() was given (words on the host, bytes on the device);
// `slice_base` = (i - i0) * nj * Dp in the same units.  `a` = outer radius^2 of the lane's ring in this slice,
// or NaN if the slice is not a thin ring of this lane: NaN never votes and never asks for the exact path
// (every comparison with it is false), and neither does a column outside the outer circle (sqrt of a negative).
template <bool CLIP, class Emit>
RCV_HD void thin_fast(const PointCtx& c, float a, const LaneTask& L, int slice_base, int sink, Emit& emit, ThinOut& o) {
  const float g = f_sub(a, L.du2);
  const float zs = f_sqrt_fast(g);
  const float hWg = f_sub(c.hW, g);
  o.ub = L.ubase + slice_base;
  const float tm0 = f_add(f_add(zs, L.cp), RCV_MAGIC);   // top arc
  const float tm1 = f_add(f_add(zs, L.cm), RCV_MAGIC);   // bottom arc, mirrored
  o.fl0 = f_sub(tm0, RCV_MAGIC); o.fl1 = f_sub(tm1, RCV_MAGIC);
  o.vt0 = L.vrel0 + (f_bits(tm0) - RCV_MAGIC_BITS);   // topmost candidate (>= true topmost voxel under the outer circle)
  o.vt1 = L.vrel0 - (f_bits(tm1) - RCV_MAGIC_BITS);
  const float d0 = f_sub(o.fl0, L.fv), d1 = f_add(o.fl1, L.fv);              // dv of the two candidates
  o.q0 = f_fma(d0, d0, hWg); o.q1 = f_fma(d1, d1, hWg);                      // e + W/2, e = |v-p|^2 - R^2 in float32
  emit(pick_offset<CLIP>(o.q0, fabsf(d0), o.vt0, o.ub + o.vt0 * L.sv, sink, c.hw_m, L.thr, L.vn));
  emit(pick_offset<CLIP>(o.q1, fabsf(d1), o.vt1, o.ub + o.vt1 * L.sv, sink, c.hw_m, L.thr, L.vn));
  // not surely inside the shell, but not surely beyond the inner boundary either: decide exactly
  o.t0 = (fabsf(d0) > L.thr) && !(fabsf(o.q0) < c.hw_m) && (o.q0 > -c.hw_p);
  o.t1 = (fabsf(d1) > L.thr) && !(fabsf(o.q1) < c.hw_m) && (o.q1 > -c.hw_p);
}

template <class SlowArc>
RCV_HD void thin_slow(const PointCtx& c, const LaneTask& L, int i, const ThinOut& o, SlowArc& slowarc) {
  if (o.t0) slowarc(c, L, i, o.ub, 0, o.q0, o.fl0, o.vt0);
  if (o.t1) slowarc(c, L, i, o.ub, 1, o.q1, o.fl1, o.vt1);
}

// ---- fast ring pass ("ring2"): thin rings of a warp whose candidates cannot leave the tile (ring_noclip) ----------
// Same candidates and the same float32 arithmetic as thin_fast(), reorganised so that one candidate costs about a
// dozen instructions on the device (rcvvote.cu issues it as one PTX block per column):
//   * g = a - du^2 is formed with one rounding (fused); ptxas contracts the packed multiply and subtract into an
//     FFMA2 even when both carry .rn, so the fused form is the one both sides of the parity use;
//   * the address is  bits(tm) * sv + K : the float -> int conversion of the magic-number floor, the lattice base and
//     the tile offset are folded into one multiply-add.  In the Z-pass the column is folded in as well by adding
//     MAGIC + u*Dp instead of MAGIC (both are integers below 2^24, so the rounding to an integer is unchanged except
//     for the parity of exact ties, which the bias of the floor makes irrelevant: a tie candidate is never a sure vote);
//   * columns with du^2 < b/2 - 1 (b = smallest inner radius^2 among the chunk's rings) are "interior": a
//     candidate that is surely inside the shell has dv^2 > b - du^2 > du^2, so the ownership test |dv| > |du| is
//     implied and only evaluated in the few boundary columns (OWN);
//   * a candidate that needs the exact path is flagged as  lo != sure  (sure implies lo); the rare branch re-derives
//     the column with ring2_slow_column(), which applies the ownership test in every case.
struct Ring2Cand {
  float q, fl, d;
  unsigned bits;
};
RCV_HD void ring2_magic(bool pass, int u, int Dp, float& m0, float& m1) {
  const float s = pass ? 0.f : (float)(u * Dp);   // exact: |u * Dp| < 2^22
  m0 = f_add(RCV_MAGIC, s);
  m1 = f_sub(RCV_MAGIC, s);
}
// cpm = +-fv + dbias - 0.5, mu = magic constant of the arc, sfv = -fv (top arc) or +fv (bottom arc, mirrored)
RCV_HD void ring2_cand(float zs, float cpm, float mu, float sfv, float hWg, Ring2Cand& o) {
  const float tm = f_add(f_add(zs, cpm), mu);
  o.bits = (unsigned)f_bits(tm);
  o.fl = f_sub(tm, mu);
  o.d = f_add(o.fl, sfv);
  o.q = f_fma(o.d, o.d, hWg);
}
// Address constants of slice `slice_base` (units as in lane_setup_pu) for column u0: top arc address = bits*sv + K0,
// bottom arc address = K1 - bits*sv.  In the Y-pass K0/K1 advance by `unit` per column, in the Z-pass they are constant.
RCV_HD void ring2_consts(const PointCtx& c, const Tile& t, bool pass, int u0, unsigned slice_base, unsigned unit, unsigned& K0, unsigned& K1,
                         unsigned& sv) {
  const unsigned MB = (unsigned)RCV_MAGIC_BITS;
  if (!pass) {
    sv = unit;
    const unsigned row = (unsigned)((c.ipy - t.j0) * t.Dp + c.ipz);
    K0 = slice_base + (row - MB) * unit;
    K1 = slice_base + (row + MB) * unit;
  } else {
    sv = (unsigned)t.Dp * unit;
    const unsigned col = slice_base + (unsigned)(c.ipz + u0) * unit;
    K0 = col + ((unsigned)(c.ipy - t.j0) - MB) * sv;
    K1 = col + ((unsigned)(c.ipy - t.j0) + MB) * sv;
  }
}
// Ownership threshold of a column (as in lane_setup_pu, tile bounds aside): Z-pass owns |dv| >= |du|, Y-pass |dv| > |du|.
RCV_HD float ring2_thr(bool pass, float duf) {
  // Y-pass: |du|.  Z-pass: the float just below |du| (so that > means >=), and -1 for du == 0; branch-free: the bit
  // pattern of |du| minus one wraps to 0xffffffff for du == 0 and the unsigned minimum maps that to -1.0f.
  const unsigned b = (unsigned)f_bits(fabsf(duf)) - (pass ? 0u : 1u);
  return f_from_bits((int)(b < 0xBF800000u ? b : 0xBF800000u));
}
// Bias of the magic-number floor for one lane and one chunk, minus 0.5 (what point_setup() stores in dbias_m05 for the
// general case).  The floor must not miss the topmost voxel under the outer circle, so the bias has to cover the error of
// zs = sqrt(g): 2 ulp of the approximate square root and the two additions before the floor ((R + 2) 2^-20, as in
// point_setup) plus |g_float - g_exact| / (2 zs) with |g_float - g_exact| <= eps / 2.  point_setup() assumes zs >= 2; an
// OWNED candidate of a ring with outer radius^2 a has dv^2 >= du^2, hence g = a - du^2 >= a/2 - 1 and zs >= sqrt(amin/2 - 1)
// for the smallest ring of the chunk (amin > 36).  A candidate with g < amin/2 - 1 has du^2 - dv^2 > 2: it and the voxel
// above it belong to the other pass, so a miss there is harmless.  The smaller bias makes "candidate above the outer
// circle" -- 90 % of the trips to the exact path -- about three times rarer.  (1.25 = margin on the eps/2 bound.)
RCV_HD float ring2_dbias_m05(const PointCtx& c, float amin) {
  const float zmin = f_sub(f_sqrt_fast(f_sub(f_mul(amin, 0.5f), 1.0f)), 0.1f);   // amin > 36: zmin > 4
  const float rp2 = (float)(c.R + 2);
  const float dbias = f_add(f_mul(c.eps, 0.3125f) / zmin, f_mul(rp2, 9.5367431640625e-07f));
  return f_sub(dbias, 0.5f);
}
// Interior half-width of a lane: columns |u| <= Hin have du^2 < bmin/2 - 1 (|du| <= |u| + 0.5); -1 if there is none.
RCV_HD int ring2_interior(float bmin) {
  const float h = f_sub(f_mul(bmin, 0.5f), 1.0f);
  if (!(h > 0.3f)) return -1;
  const float v = f_sub(f_sqrt_fast(h), 0.51f);
  return v >= 0.f ? (int)v : -1;
}
// One slice of one column, both arcs, M candidates per arc (the topmost voxel under the outer circle and the M - 1
// voxels inwards of it: a ring with slice_setup code <= M crosses a column of its own pass in at most M voxels):
// 2 M emits (vote address or sink); returns true if a candidate needs the exact path.
template <bool OWN, int M, class Emit>
RCV_HD bool ring2_fast(const PointCtx& c, float a, float duf, float thr, float cp, float cm, float fv, float mu0, float mu1, unsigned K0,
                       unsigned K1, unsigned sv, unsigned sink, Emit& emit) {
  const float g = f_fma(-duf, duf, a);   // a - du^2 with ONE rounding (the device issues it as a packed FFMA2)
  const float zs = f_sqrt_fast(g);
  const float hWg = f_sub(c.hW, g);
  bool any = false;
  for (int arc = 0; arc < 2; ++arc) {
    Ring2Cand cd;
    const float sfv = arc ? fv : -fv;
    ring2_cand(zs, arc ? cm : cp, arc ? mu1 : mu0, sfv, hWg, cd);
    unsigned addr = arc ? K1 - cd.bits * sv : cd.bits * sv + K0;
    float fl = cd.fl, d = cd.d, q = cd.q;
    for (int m = 0; m < M; ++m) {
      if (m) {   // one voxel inwards: (integer - 1) -+ fraction, one rounding like every other coordinate difference
        fl = f_sub(fl, 1.0f);
        d = f_add(fl, sfv);
        q = f_fma(d, d, hWg);
        addr = arc ? addr + sv : addr - sv;
      }
      const bool o = !OWN || (fabsf(d) > thr);
      const bool sure = o && (fabsf(q) < c.hw_m);
      const bool lo = o && (q > -c.hw_p);
      emit((int)(sure ? addr : sink));
      any |= (lo != sure);
    }
  }
  return any;
}

// The rare branch of the fast ring pass, for one lane: exact decisions for the candidates of column u that the fast
// path could not decide.  Re-derives the candidates of the chunk's slices with the same arithmetic (fast and slow path
// must agree on which voxel is the candidate), applies the ownership test in every case, and decides like
// ring_slow(): the undecided candidates exactly, and -- if the topmost one may lie outside the outer sphere -- the
// voxel one step inwards of the last candidate.
//   ipl / ipc = nearest lattice coordinate of the point along the lane / candidate axis, a[s] = outer radius^2 of
//   slice i0c + s (NaN: not a thin ring of this lane), K0/K1/sv as in ring2_consts.  No bounds tests: the pass is
//   only used when no candidate can leave the tile (ring_noclip).  slow(iA, iB, iC) is the exact predicate.
template <class Slow, class EmitSlow>
RCV_HD void ring2_slow_lane(bool pass, int nc, int M, int ipl, int ipc, int u, int i0c, float hW, float hw_m, float hw_p, float duf, float cp, float cm,
                            float fv, float mu0, float mu1, unsigned sv, const float* a, const unsigned* K0, const unsigned* K1, Slow& slow,
                            EmitSlow& emit_slow) {
  const float thr = ring2_thr(pass, duf);
  const int lc = ipl + u;
  for (int sidx = 0; sidx < nc; ++sidx) {
    const float as = a[sidx];
    if (!(as == as)) continue;
    const int i = i0c + sidx;
    const float g = f_fma(-duf, duf, as);   // like ring2_fast
    const float zs = f_sqrt_fast(g);
    const float hWg = f_sub(hW, g);
    for (int arc = 0; arc < 2; ++arc) {
      Ring2Cand cd;
      const float sfv = arc ? fv : -fv;
      ring2_cand(zs, arc ? cm : cp, arc ? mu1 : mu0, sfv, hWg, cd);
      const int n = (int)cd.fl, dir = arc ? 1 : -1;
      const unsigned addr0 = arc ? K1[sidx] - cd.bits * sv : cd.bits * sv + K0[sidx];
      // candidates m = 0 .. M-1 were seen by the fast path: decide those it could not.  If the topmost one may lie
      // outside the outer sphere the run can reach one voxel further inwards (m = M), which the fast path never saw.
      const int mend = (cd.q >= hw_m) ? M : M - 1;
      for (int m = 0; m <= mend; ++m) {
        const float fl = f_sub(cd.fl, (float)m);                 // exact: small integers
        const float d = m ? f_add(fl, sfv) : cd.d;
        if (!(fabsf(d) > thr)) continue;
        const float q = m ? f_fma(d, d, hWg) : cd.q;
        const bool sure = fabsf(q) < hw_m;
        if (!(q > -hw_p) || q >= hw_p) continue;                 // surely below the inner or above the outer sphere: no vote
        const unsigned addr = addr0 + (unsigned)(m * dir) * sv;
        if (sure) {                                              // surely inside the shell
          if (m == M) emit_slow((int)addr);                      // (m < M: the fast path has cast this vote already)
          continue;
        }
        const int cc = arc ? ipc - n + m : ipc + n - m;          // within eps of a shell boundary: the reference's float64 sequence
        if (slow(i, pass ? cc : lc, pass ? lc : cc)) emit_slow((int)addr);
      }
    }
  }
}

// One lane, one cell (dj, dk) of a small sphere's slice bounding box (R < RCV_POLAR_MIN_R only).
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

// ---- polar pass --------------------------------------------------------------------------------------
// Slices whose ring is not thin (slice_setup code != 1) lie in the polar caps of the slice axis.  For
// R >= RCV_POLAR_MIN_R every such slice has |dx| >= (W+1)/2, so a column ALONG the slice axis through the
// in-plane cell (j,k) crosses the shell in less than one voxel: the only candidate of the + side is the topmost
// voxel under the outer sphere (mirror image for the - side).  A candidate votes iff its slice is one of the
// lane's non-thin slices in this tile (bit masks mplus / mminus, bit v = slice t.i0 + v; a tile has <= 32 slices).
struct PolarOut {
  float q0, q1;
  int vt0, vt1;
  bool t0, t1;
};

RCV_HD bool mask_bit(unsigned m, int v) { return (u_shr(m, v) & 1u) != 0u; }

// s = dB^2 + dC^2 of the cell, cell = its offset inside a slice, ok = the cell is inside the tile and the lane
// takes part; sstride = slice stride (same units as cell).  SIDES: bit 0 = + side, bit 1 = - side.
// cpx / cmx = fx + dbias - 0.5 and -fx + dbias - 0.5.
// vote = |q| < hw_m  &&  ok  &&  bit v of mask  (one predicate chain, one select)
RCV_HD int polar_pick(float q, float hw_m, bool ok, unsigned mask, int v, int off, int sink) {
#if defined(__CUDA_ARCH__)
  int r;
  asm("{\n\t.reg .pred p;\n\t.reg .f32 aq;\n\t.reg .b32 sh;\n\t"
      "abs.f32 aq, %1;\n\t"
      "shr.u32 sh, %4, %5;\n\t"
      "and.b32 sh, sh, %3;\n\t"
      "setp.ne.b32 p, sh, 0;\n\t"
      "setp.lt.and.f32 p, aq, %2, p;\n\t"
      "selp.b32 %0, %6, %7, p;\n\t}"
      : "=r"(r) : "f"(q), "f"(hw_m), "r"(ok ? 1 : 0), "r"(mask), "r"(v), "r"(off), "r"(sink));
  return r;
#else
  return ((fabsf(q) < hw_m) && ok && mask_bit(mask, v)) ? off : sink;
#endif
}

template <int SIDES, class Emit>
RCV_HD void polar_fast(const PointCtx& c, const Tile& t, float cpx, float cmx, float s, int cell, bool ok, unsigned mplus, unsigned mminus,
                       int sstride, int sink, Emit& emit, PolarOut& o) {
  const float g = f_sub(c.R2, s);
  const float zs = f_sqrt_fast(g);         // NaN outside the sphere's shadow: no vote, no exact path
  const float hWg = f_sub(c.hW, g);
  const int vrel0 = c.ipx - t.i0;
  o.t0 = false; o.t1 = false; o.q0 = 0.f; o.q1 = 0.f; o.vt0 = 0; o.vt1 = 0;
  if (SIDES & 1) {
    const float tm0 = f_add(f_add(zs, cpx), RCV_MAGIC);
    const float fl0 = f_sub(tm0, RCV_MAGIC);
    o.vt0 = vrel0 + (f_bits(tm0) - RCV_MAGIC_BITS);
    const float d0 = f_sub(fl0, c.fx);
    o.q0 = f_fma(d0, d0, hWg);
    emit(polar_pick(o.q0, c.hw_m, ok, mplus, o.vt0, cell + o.vt0 * sstride, sink));
    o.t0 = ok & !(fabsf(o.q0) < c.hw_m) & (o.q0 > -c.hw_p);
  }
  if (SIDES & 2) {
    const float tm1 = f_add(f_add(zs, cmx), RCV_MAGIC);
    const float fl1 = f_sub(tm1, RCV_MAGIC);
    o.vt1 = vrel0 - (f_bits(tm1) - RCV_MAGIC_BITS);
    const float d1 = f_add(fl1, c.fx);
    o.q1 = f_fma(d1, d1, hWg);
    emit(polar_pick(o.q1, c.hw_m, ok, mminus, o.vt1, cell + o.vt1 * sstride, sink));
    o.t1 = ok & !(fabsf(o.q1) < c.hw_m) & (o.q1 > -c.hw_p);
  }
}

// Column range of row `db2` (= dB^2) of the polar pass: the lane's non-thin rings of this tile lie in the annulus
// s_lo < dB^2 + dC^2 < s_hi (s_hi = largest outer radius^2, s_lo = smallest inner radius^2), so only cells with
// ci <= |uc| <= co can hold a candidate (|uc| differs from |dC| by at most 0.5; co < 0: the row is empty).
RCV_HD void polar_row_range(float s_lo, float s_hi, float eps, float db2, bool lane_on, int& ci, int& co) {
  // the annulus is widened by 2 eps: a cell whose float32 s sits exactly on a ring boundary can still hold a vote
  // (it is then decided by the exact path)
  const float m = f_add(eps, eps);
  const float rem_hi = f_sub(f_add(s_hi, m), db2), rem_lo = f_sub(f_sub(s_lo, m), db2);
  co = -1; ci = 0x7fffffff;
  if (lane_on && rem_hi > 0.f) {
    co = (int)f_add(f_sqrt_fast(rem_hi), 0.6f);
    ci = rem_lo > 1.0f ? (int)f_sub(f_sqrt_fast(rem_lo), 0.6f) : 0;
    if (ci < 0) ci = 0;
  }
}

// Exact decision of the candidates polar_fast() could not decide; (jb, kc) = in-plane lattice coordinates of the cell.
template <class Slow, class EmitSlow>
RCV_HD void polar_slow(float hw_m, int ti0, const PolarOut& o, int jb, int kc, int cell, unsigned mplus, unsigned mminus, int sstride, Slow& slow,
                       EmitSlow& emit_slow) {
  if (o.t0) {
    if (mask_bit(mplus, o.vt0) && slow(ti0 + o.vt0, jb, kc)) emit_slow(cell + o.vt0 * sstride);
    if (o.q0 >= hw_m) {   // the top candidate may lie outside the outer sphere; the voxel below it can then be inside
      const int v2 = o.vt0 - 1;
      if (mask_bit(mplus, v2) && slow(ti0 + v2, jb, kc)) emit_slow(cell + v2 * sstride);
    }
  }
  if (o.t1) {
    if (mask_bit(mminus, o.vt1) && slow(ti0 + o.vt1, jb, kc)) emit_slow(cell + o.vt1 * sstride);
    if (o.q1 >= hw_m) {
      const int v2 = o.vt1 + 1;
      if (mask_bit(mminus, v2) && slow(ti0 + v2, jb, kc)) emit_slow(cell + v2 * sstride);
    }
  }
}

// ---- fast polar pass ("polar2"): the polar caps of a warp whose candidates cannot leave the tile ------------------
// Same candidates as polar_fast() for ONE side of the pole at a time, reorganised like ring2: the address is
// bits(tm) * smul + K, and "the candidate's slice is one of the lane's polar slices of this tile" is a range test on the
// integer-valued float fl = n instead of a bit-mask test (the polar slices of one side are consecutive; the kernel
// checks that and otherwise falls back to polar_fast()).  n counts slices away from the point's own slice: slice =
// ipx + sgn * n.  Votes need n in [nlo, nhi]; the exact path is asked for when n is in [nlo, nhi + 1] (the candidate may
// lie outside the outer sphere, then the voxel one step inwards, n - 1, can be a vote: polar_slow()) and the float
// residual is not decisive.  g = R^2 - dB^2 - dC^2 is formed as fma(-dC, dC, R^2 - dB^2).
struct Polar2Side {
  float cx, sfx;          // +-fx + dbias - 0.5 and -+fx (d = n + sfx)
  float nmidw, nhalfw;    // centre and half-width of the wide range [nlo, nhi + 1]; NaN centre = the lane has no slice here
  float nhi;              // (float)nhi
  int nlo, nhi_i, sgn;
};
// mask = polar slices of the lane on this side (bit v = slice t.i0 + v).  Returns false if the set bits are not consecutive.
RCV_HD bool polar2_side(const PointCtx& c, const Tile& t, unsigned mask, bool plus, Polar2Side& s) {
  s.sgn = plus ? 1 : -1;
  s.cx = plus ? f_add(c.fx, c.dbias_m05) : f_sub(c.dbias_m05, c.fx);
  s.sfx = plus ? -c.fx : c.fx;
  if (mask == 0u) {
    s.nlo = 1; s.nhi_i = 0; s.nhi = 0.f; s.nmidw = f_from_bits(0x7fc00000); s.nhalfw = 0.f;
    return true;
  }
#if defined(__CUDA_ARCH__)
  const int vlo = __ffs((int)mask) - 1, vhi = 31 - __clz((int)mask);
#else
  int vlo = 0, vhi = 31;
  while (!((mask >> vlo) & 1u)) ++vlo;
  while (!((mask >> vhi) & 1u)) --vhi;
#endif
  const unsigned full = (vhi - vlo == 31) ? 0xffffffffu : (((1u << (vhi - vlo + 1)) - 1u) << vlo);
  const int vrel0 = c.ipx - t.i0;
  s.nlo = plus ? vlo - vrel0 : vrel0 - vhi;
  s.nhi_i = plus ? vhi - vrel0 : vrel0 - vlo;
  s.nhi = (float)s.nhi_i;
  s.nmidw = f_mul((float)(s.nlo + s.nhi_i + 1), 0.5f);
  s.nhalfw = f_mul((float)(s.nhi_i + 1 - s.nlo), 0.5f);
  return mask == full;
}
struct Polar2Cell {
  float q, fl, hWg;
  unsigned bits;
  bool wide, sure, vote, amb;
};
// One cell (lattice column offset ucf from the point's nearest lattice point along C), r2mdb2 = R^2 - dB^2 of its row.
RCV_HD void polar2_cell(const PointCtx& c, const Polar2Side& s, float ucf, float r2mdb2, Polar2Cell& o) {
  const float dc = f_sub(ucf, c.fz);
  const float g = f_fma(-dc, dc, r2mdb2);
  const float zs = f_sqrt_fast(g);          // NaN outside the sphere's shadow: no vote, no exact path
  const float hWg = f_sub(c.hW, g);
  o.hWg = hWg;
  const float tm = f_add(f_add(zs, s.cx), RCV_MAGIC);
  o.bits = (unsigned)f_bits(tm);
  o.fl = f_sub(tm, RCV_MAGIC);
  const float d = f_add(o.fl, s.sfx);
  o.q = f_fma(d, d, hWg);
  const float r = f_sub(o.fl, s.nmidw);
  o.wide = fabsf(r) <= s.nhalfw;
  o.sure = o.wide && (fabsf(o.q) < c.hw_m);
  o.vote = o.sure && (o.fl <= s.nhi);
  const bool lo = o.wide && (o.q > -c.hw_p);
  o.amb = lo != o.sure;
}
// Exact decisions for one flagged cell: (jb, kc) = its lattice coordinates, K = its address constant (address of
// slice n = bits * smul + K with smul = sgn * slice stride).  slow(iA, iB, iC) is the exact predicate.
template <class Slow, class EmitSlow>
RCV_HD void polar2_slow_cell(const PointCtx& c, const Polar2Side& s, float ucf, float r2mdb2, int jb, int kc, unsigned smul, unsigned K, Slow& slow,
                             EmitSlow& emit_slow) {
  Polar2Cell o;
  polar2_cell(c, s, ucf, r2mdb2, o);
  if (!o.amb) return;
  const int n = (int)o.fl;
  const unsigned addr = o.bits * smul + K;
  // the candidate itself: the float64 sequence only if its residual is within eps of a shell boundary (beyond hw_p it is
  // surely outside the outer sphere)
  if (o.q < c.hw_p && n >= s.nlo && n <= s.nhi_i && slow(c.ipx + s.sgn * n, jb, kc)) emit_slow((int)addr);
  if (o.q >= c.hw_m) {   // the candidate may lie outside the outer sphere: the voxel one step inwards can then be inside
    const int n2 = n - 1;
    if (n2 >= s.nlo && n2 <= s.nhi_i) {
      const float d2 = f_add(f_sub(o.fl, 1.0f), s.sfx);          // (integer - 1) -+ fraction, one rounding
      const float q2 = f_fma(d2, d2, o.hWg);
      if (fabsf(q2) < c.hw_m) emit_slow((int)(addr - smul));     // surely inside the shell
      else if ((q2 > -c.hw_p) && (q2 < c.hw_p) && slow(c.ipx + s.sgn * n2, jb, kc)) emit_slow((int)(addr - smul));
    }
  }
}

// Half-width (in rows) of the polar pass for a lane whose largest non-thin ring has outer radius^2 amax.
RCV_HD int polar_half_width(float amax, float eps) { return (int)f_add(f_sqrt_fast(fmaxf(f_add(amax, f_add(eps, eps)), 0.f)), 0.6f); }

// Slice range of point c inside tile t (inclusive); empty if ia > ib.
RCV_HD void slice_range(const PointCtx& c, const Tile& t, int& ia, int& ib) {
  ia = c.ipx - c.R - 1; if (ia < t.i0) ia = t.i0;
  ib = c.ipx + c.R + 1; if (ib > t.i0 + t.ni - 1) ib = t.i0 + t.ni - 1;
  if (c.R <= 0) { ia = 1; ib = 0; }
}

// Slices per chunk of the ring passes for a slab of ni slices: 3 or 4, whichever pads the slab less (thin slabs of large
// grids: 1 or 2, the slab itself).
RCV_HD int ring_chunk(int ni) { return ni <= 2 ? (ni < 1 ? 1 : ni) : ((((ni + 2) / 3) * 3 < ((ni + 3) / 4) * 4) ? 3 : 4); }
// Largest slab thickness <= ni_max that is a multiple of 3 or 4 (no padding in the ring passes).
RCV_HD int slab_thickness(int ni_max) {
  if (ni_max < 3) return ni_max;
  const int a3 = (ni_max / 3) * 3, a4 = (ni_max / 4) * 4;
  return a3 > a4 ? a3 : a4;
}

// True if no candidate of the lane's ring passes can fall outside the tile along a candidate axis
// (candidates lie within R + 1 of the nearest lattice point), so the per-candidate bounds test can be skipped.
RCV_HD bool ring_noclip(const PointCtx& c, const Tile& t) {
  if (c.R <= 0) return true;   // draws nothing (padding lanes of a warp's last group)
  return (c.ipz - c.R - 2 >= -t.glo) && (c.ipz + c.R + 2 < t.D + t.ghi) && (c.ipy - c.R - 2 >= t.j0) && (c.ipy + c.R + 2 < t.j0 + t.nj);
}

}  // namespace rcv
