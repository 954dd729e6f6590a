// runs_core.h -- per-lane arithmetic of the run-length ("difference array") vote rasteriser (kernel K2, second generation).
//
// What it renders.  The reference's vote loop (AccumulatorSpace.py:325-341) increments voxel (i,j,k) for a point p with
// integer radius R iff  0 < R - sqrt((i-px)^2 + (j-py)^2 + (k-pz)^2) < sqrt(3)/4  (float64, strict), for EVERY voxel of
// the D^3 cube.  Seen from one column of the lattice (two coordinates fixed, the third -- C, the fastest axis of the
// tile -- running) that set is
//        { n :  gi < (n - fc)^2 < g },     g = R^2 - dA^2 - dB^2,   gi = g - W,   W = R^2 - (R - sqrt3/4)^2
// i.e. at most two runs of consecutive voxels, mirror images of each other around the point:
//        ( fc - sqrt(g), fc - sqrt(gi) )   and   ( fc + sqrt(gi), fc + sqrt(g) )
// (one run when gi <= 0, nothing when g <= 0).  Instead of one shared-memory atomic per voxel of a run -- the first
// generation tested one candidate voxel per column and arc, ~11 instructions per candidate at a 36 % hit rate -- a run
// [b, e) is recorded as +1 at b and -1 at e in a DIFFERENCE ARRAY along C; when all points of a tile have been drawn,
// one prefix sum per row turns the differences into the vote counts (rcvvote.cu, tile epilogue).  A column therefore
// costs four atomics whatever the length of its runs, there is no per-voxel work, no ownership test between passes, no
// polar / dense / thick-ring special case (an annulus is an annulus), and integer adds commute, so the counts are the
// reference's, bit for bit.
//
// Exactness.  The four run boundaries are computed in float32: b = first integer at or above the transition
// x = fc -+ sqrt(.), as round-to-nearest of x + 0.5 through the magic-number trick.  |g_float32 - g_exact| <= E (bound
// below), so the transition is known to within E / sqrt(g) (+ rounding); if an integer lies that close to it the
// boundary is "flagged" and the one voxel in question is re-decided with exact_hit(), the reference's float64
// operation sequence, and the difference array corrected (+-1 at the voxel, -+1 behind it).  Everything else is decided
// by float32 alone.  tests/test_runs_hostsim.py fuzzes this file (compiled for the host, square roots perturbed by
// +-2 ulp) against the brute-force oracle; the -m gpu tests check the CUDA build the same way.
#pragma once
#include "raster_core.h"   // f_* / d_* wrappers (never contracted), exact_hit, RCV_MAGIC

namespace rcv {

// ---- packed pairs of float32 ------------------------------------------------------------------------------------
// On the device a pair lives in one 64-bit register and add / sub / fma are single FADD2 / FFMA2 issues (sm_100);
// they round each half exactly like the scalar .rn instructions (tools/check_f32x2.cu).  On the host: two floats.
#if defined(__CUDA_ARCH__)
typedef unsigned long long f2;
RCV_HD f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
RCV_HD float f2_lo(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); (void)hi; return lo; }
RCV_HD float f2_hi(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); (void)lo; return hi; }
RCV_HD f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
RCV_HD f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
RCV_HD f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
RCV_HD float f_max0(float x) { return fmaxf(x, 0.f); }   // NaN -> 0 (FMNMX returns the non-NaN operand)
#else
struct f2 { float lo, hi; };
RCV_HD f2 f2_make(float lo, float hi) { f2 r; r.lo = lo; r.hi = hi; return r; }
RCV_HD float f2_lo(f2 a) { return a.lo; }
RCV_HD float f2_hi(f2 a) { return a.hi; }
RCV_HD f2 f2_add(f2 a, f2 b) { return f2_make(a.lo + b.lo, a.hi + b.hi); }
RCV_HD f2 f2_sub(f2 a, f2 b) { return f2_make(a.lo - b.lo, a.hi - b.hi); }
RCV_HD f2 f2_fma(f2 a, f2 b, f2 c) { return f2_make(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
RCV_HD float f_max0(float x) { return (x > 0.f) ? x : 0.f; }   // NaN -> 0
#endif
RCV_HD f2 f2_dup(float x) { return f2_make(x, x); }

// square root for the run boundaries: NaN for a negative argument, and the NaN is POSITIVE (sign bit clear) on both the
// device (MUFU returns the canonical 0x7fffffff) and the host, so that "flagged" (a sign bit, see run_slice) is never
// raised by a column that lies outside the sphere
RCV_HD float f_sqrt_run(float x) {
#if defined(__CUDA_ARCH__)
  return f_sqrt_fast(x);
#else
  return (x >= 0.f) ? f_sqrt_fast(x) : f_from_bits(0x7fffffff);
#endif
}

// reciprocal square root for the run boundaries (MUFU.RSQ, relative error below 2^-22): NaN for a negative argument,
// +inf for zero; s' = g * rsqrt(g)
RCV_HD float f_rsqrt_run(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  if (!(x >= 0.f)) return f_from_bits(0x7fffffff);
  const float s = f_sqrt_fast(x);            // (perturbed by +-2 ulp under g_sqrt_perturb, like the device instruction's error)
  return s > 0.f ? 1.0f / s : f_from_bits(0x7f800000);
#endif
}

// ---- per point ----------------------------------------------------------------------------------------------------
// Internal axes (A,B,C): A = slice axis of the tile (slabs of A-slices), B = row axis, C = the run axis (fastest).
struct RunPoint {
  int ipa, ipb, ipc;   // nearest lattice point
  float fa, fb, fc;    // p - ip, |f| <= 0.5
  int R;               // <= 0: draws nothing
  float W;             // R^2 - (R - sqrt3/4)^2, rounded to float32
};
// (what the prelude stores per point, 32 bytes, in vote order)
RCV_HD void run_point_setup(RunPoint& c, double pa, double pb, double pc, int R) {
  c.ipa = d_rint(pa); c.ipb = d_rint(pb); c.ipc = d_rint(pc);
  c.fa = (float)d_sub(pa, (double)c.ipa);
  c.fb = (float)d_sub(pb, (double)c.ipb);
  c.fc = (float)d_sub(pc, (double)c.ipc);
  c.R = R;
  const double rin = (double)R - RCV_SHELL;
  c.W = (float)((double)R * (double)R - rin * rin);
}

// Error budget (all in voxel^2).  With dA' = fl((i - ipa) - fa), dB' likewise, a' = fl(R^2 - fl(dA'^2)),
// g' = fma(-dB', dB', a'):   |g' - g_exact| <= 2^-21 (R + 2)^2 =: E0   (coordinates rounded to float32: 2^-23 (dA^2 + dB^2);
// three roundings at magnitude <= (R+1)^2: 3 * 2^-24 (R+1)^2).  The slice constant carries a bias of +1.5 E0, so that the
// biased g'' >= g_exact and gi'' >= gi_exact always (a NEGATIVE g'' / gi'' then proves the disc empty) and
// |g'' - g_exact|, |gi'' - gi_exact| <= 2.75 E0.  A boundary computed as x' = fl(c -+ s'), s' = sqrt.approx(g''), is within
// 2.75 E0 / s' + (R + 2) 2^-21 of the exact transition; multiplied by s' <= R + 2:  dist(x', Z) * s' <= 3.75 E0 must be
// flagged (3.875 E0 with the rsqrt-and-multiply form of s', relative error 1.25 * 2^-22).  epsh = 4.5 E0.
struct RunLane {
  f2 cc;      // (c, c), c = fc + 0.5: the +0.5 turns round-to-nearest into "first integer at or above"
  f2 nepsh;   // (-epsh, -epsh)
  float E0;
};
RCV_HD float run_E0(int R) { const float rp2 = (float)(R + 2); return f_mul(f_mul(rp2, rp2), 4.76837158203125e-07f); }   // 2^-21 (R+2)^2
RCV_HD void run_lane_setup(const RunPoint& c, RunLane& L) {
  L.E0 = run_E0(c.R);
  L.cc = f2_dup(f_add(c.fc, 0.5f));
  L.nepsh = f2_dup(-f_mul(L.E0, 4.5f));
}
// Slice constants of A-slice i: (a'', aW'') = biased outer radius^2 of the ring in that slice and the same minus W.
// A slice the sphere does not reach (or a lane that draws nothing) gets a negative pair: every column is then empty.
RCV_HD f2 run_slice_consts(const RunPoint& c, const RunLane& L, int i, bool live) {
  const float dA = f_sub((float)(i - c.ipa), c.fa);
  const float a = f_sub((float)(c.R * c.R), f_mul(dA, dA));   // R^2 exact for R < 4096
  const float a2 = f_add(a, f_mul(L.E0, 1.5f));
  if (!(live && c.R > 0)) return f2_dup(-1.0f);
  return f2_make(a2, f_sub(a2, c.W));
}
// Half-width in columns: every column with g'' >= 0 has |dB| <= sqrt(a''), |u| <= |dB| + 0.5.  -1: no column at all.
RCV_HD int run_half_width(float a2max) { return a2max > 0.f ? (int)f_add(f_mul(f_sqrt_fast(a2max), 1.000001f), 0.5f) + 1 : -1; }

// ---- one column of one slice -------------------------------------------------------------------------------------------
struct RunCol {
  f2 du, ndu;   // (dB, dB), (-dB, -dB) of this column
  f2 mu;        // (mu, mu), mu = MAGIC + (column offset inside the tile row block) * Dp: folds the row into the address
};
RCV_HD void run_col_setup(const RunPoint& c, int u, int uc, int Dp, RunCol& C) {
  const float du = f_sub((float)u, c.fb);
  C.du = f2_dup(du); C.ndu = f2_dup(-du);
  C.mu = f2_dup(f_add(RCV_MAGIC, (float)(uc * Dp)));   // exact: |uc * Dp| < 2^22
}
// Boundaries as raw float bits of x + 0.5 + mu (an integer-valued float in [2^23, 2^24): bits = MAGIC_BITS + uc*Dp + b):
//   lower run [b1, b2), upper run [b3, b4); starts b1 and b3, ends b2 and b4.
// dL = (d1, d2), dU = (d4, d3): d = b - (x + 0.5) in [-0.5, 0.5], so the transition lies 0.5 - |d| from the nearest integer.
// thr = (thr_outer, thr_inner): a boundary is FLAGGED -- an integer may lie between its float32 transition and the exact
// one -- iff |d| >= thr, thr = 0.5 - epsh / s'  (dist * s' <= epsh, error budget above).  NaN thr (empty disc): never.
struct RunOut {
  unsigned b1, b2, b3, b4;
  f2 thr, dU, dL;
};
RCV_HD void run_slice(const RunLane& L, const RunCol& C, f2 aa, RunOut& o) {
  const f2 gg = f2_fma(C.ndu, C.du, aa);                               // (g'', gi''), one rounding each
  const f2 rr = f2_make(f_rsqrt_run(f2_lo(gg)), f_rsqrt_run(f2_hi(gg)));   // 1 / sqrt: NaN for a negative argument, +inf for 0
  const f2 sn = f2_fma(gg, rr, f2_dup(0.f));                           // s' = g'' / sqrt(g''): NaN for g'' <= 0
  const f2 sc = f2_make(f_max0(f2_lo(sn)), f_max0(f2_hi(sn)));         // NaN -> 0: the run collapses
  const f2 xU = f2_add(L.cc, sc), xL = f2_sub(L.cc, sc);               // transitions + 0.5: (x4, x3), (x1, x2)
  const f2 tU = f2_add(xU, C.mu), tL = f2_add(xL, C.mu);               // rounded to integers by the magic number
  o.b4 = (unsigned)f_bits(f2_lo(tU)); o.b3 = (unsigned)f_bits(f2_hi(tU));
  o.b1 = (unsigned)f_bits(f2_lo(tL)); o.b2 = (unsigned)f_bits(f2_hi(tL));
  o.dU = f2_sub(f2_sub(tU, C.mu), xU); o.dL = f2_sub(f2_sub(tL, C.mu), xL);
  o.thr = f2_fma(L.nepsh, rr, f2_dup(0.5f));                           // 0.5 - epsh / s'
}
RCV_HD bool run_amb(float d, float thr) { return fabsf(d) >= thr; }     // false for a NaN threshold
RCV_HD bool run_flagged(const RunOut& o) {
  // (bitwise |: no short-circuit branches in the hot loop)
  return run_amb(f2_lo(o.dL), f2_lo(o.thr)) | run_amb(f2_hi(o.dL), f2_hi(o.thr)) | run_amb(f2_hi(o.dU), f2_hi(o.thr)) |
         run_amb(f2_lo(o.dU), f2_lo(o.thr));
}

// acc |= flagged(o), as one predicate chain on the device (four FSETP with the |d| modifier, no materialised booleans)
RCV_HD void run_flag_acc(unsigned& acc, const RunOut& o) {
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .pred p;\n\t.reg .f32 a;\n\t"
      "setp.ne.u32 p, %0, 0;\n\t"
      "abs.f32 a, %1;\n\tsetp.ge.or.f32 p, a, %5, p;\n\t"
      "abs.f32 a, %2;\n\tsetp.ge.or.f32 p, a, %6, p;\n\t"
      "abs.f32 a, %3;\n\tsetp.ge.or.f32 p, a, %6, p;\n\t"
      "abs.f32 a, %4;\n\tsetp.ge.or.f32 p, a, %5, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "+r"(acc) : "f"(f2_lo(o.dL)), "f"(f2_hi(o.dL)), "f"(f2_hi(o.dU)), "f"(f2_lo(o.dU)), "f"(f2_lo(o.thr)), "f"(f2_hi(o.thr)));
#else
  if (run_flagged(o)) acc = 1u;
#endif
}

// ---- the rare path: exact decision of the voxels next to flagged boundaries ----------------------------------------------
// Re-derives the column with the same arithmetic (fast and slow path must agree on b1..b4), and for each flagged boundary
// takes the integer m nearest to the float32 transition: the only voxel whose side of that boundary can be in doubt.
// Its count in the difference array (from b1..b4) is compared with exact(m) and the array corrected.
//   base = MAGIC_BITS + uc * Dp (what the boundaries' bits are relative to),  exact(n) = reference predicate of the voxel
//   n lattice steps from the point's nearest lattice point along C,  fix(nbits, delta): voxel `nbits` joins (delta = +1)
//   or leaves (-1) the set.
template <class Exact, class Fix>
RCV_HD int run_slow_slice(const RunLane& L, const RunCol& C, f2 aa, unsigned base, Exact& exact, Fix& fix) {
  RunOut o;
  run_slice(L, C, aa, o);
  const float d[4] = {f2_lo(o.dL), f2_hi(o.dL), f2_hi(o.dU), f2_lo(o.dU)};       // boundaries 1..4: outer, inner, inner, outer
  const float thr[4] = {f2_lo(o.thr), f2_hi(o.thr), f2_hi(o.thr), f2_lo(o.thr)};
  const int b[4] = {(int)(o.b1 - base), (int)(o.b2 - base), (int)(o.b3 - base), (int)(o.b4 - base)};
  int done[4], nd = 0, nfix = 0;
  for (int e = 0; e < 4; ++e) {
    if (!run_amb(d[e], thr[e])) continue;
    const int m = d[e] > 0.f ? b[e] - 1 : b[e];     // transition x = b - 0.5 - d: nearest integer is b - 1 for d > 0, else b
    bool seen = false;
    for (int q = 0; q < nd; ++q) seen = seen || (done[q] == m);
    if (seen) continue;
    done[nd++] = m;
    const int fast = (m >= b[0]) - (m >= b[1]) + (m >= b[2]) - (m >= b[3]);
    const int want = exact(m) ? 1 : 0;
    if (want != fast) { fix(base + (unsigned)m, want - fast); ++nfix; }
  }
  return nfix;
}

}  // namespace rcv
