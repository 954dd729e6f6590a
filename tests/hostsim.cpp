// Host-side harness for rcvpose_b200/csrc/raster_core.h (TEST INFRASTRUCTURE ONLY).
// Executes the rasteriser's per-lane code on the CPU with 32 simulated lanes in lockstep, in the same loop
// structure as the CUDA kernel (k_vote in rcvvote.cu: lane = point, slice chunks, ring passes, polar pass), with
// the trip counts taken as warp maxima exactly as the kernel's warp reductions do, so the exactness of the voxel
// set can be fuzzed against the oracle in the CPU-only container.  It is never loaded by the product.
#include <cstdint>
#include <cstdio>
#include "../rcvpose_b200/csrc/raster_core.h"

using namespace rcv;

static long long* g_cat = nullptr;
extern "C" __attribute__((visibility("default"))) void hostsim_set_census(long long* p) { g_cat = p; }

struct HostEmit {
  int32_t* tile; long words; long long votes = 0, calls = 0, oob = 0, slow_calls = 0;
  void operator()(int off) {
    ++calls;
    if (off == -1) return;                      // the sink
    if (off < 0 || off >= words) { ++oob; return; }
    tile[off] += 1; ++votes;
  }
};
struct HostEmitSlow {
  HostEmit* e;
  void operator()(int off) { (*e)(off); --e->calls; }
};
// Internal axes are (A,B,C) = (y,x,z) of the reference: the tile is a slab of A-slices, buffer layout
// [A-i0][B-j0][C].  exact_hit() needs the reference's operand order (dx^2 + dy^2 + dz^2 is order-sensitive).
struct HostSlowPerm {
  const PointCtx* c; HostEmit* e;
  bool operator()(int i, int j, int k) { ++e->slow_calls; return exact_hit(c->py, c->px, c->pz, c->R, j, i, k); }
};
struct HostSlowArcPerm {
  HostSlowPerm* slow; HostEmitSlow* es;
  void operator()(const PointCtx& c, const LaneTask& L, int i, int ub, int arc, float q, float fl, int vt) {
    ring_slow(c, L, i, ub, arc, q, fl, vt, *slow, *es);
  }
};

template <int NC>
static void ring_chunks(PointCtx* c, const int* ia, const int* ib, const Tile& t, int wa, int wb, unsigned* mplus, unsigned* mminus,
                        float* smax, float* slo, HostEmit& emit, HostEmitSlow& emit_slow, long long* counters) {
  const int slice_words = t.nj * t.Dp;
  const float kNaN = __builtin_nanf("");
  bool noclip = true;
  for (int l = 0; l < 32; ++l) noclip = noclip && ring_noclip(c[l], t);
  const int maxcode = noclip ? RCV_RING2_MAX_CODE : 1;   // rings the ring passes draw; the rest goes to the polar pass
  const int first = t.i0 + ((wa - t.i0) / NC) * NC;          // chunks tile the slab from its first slice
  for (int i0c = first; i0c <= wb; i0c += NC) {
    float a4[32][NC]; float amax = 0.f; bool any_thin = false; int mcand = 1;
    for (int sidx = 0; sidx < NC; ++sidx) {
      const int i = i0c + sidx;
      int dmax = 0; float ad[32]; int code[32];
      for (int l = 0; l < 32; ++l) {
        ad[l] = 0.f; code[l] = 0;
        if (i >= ia[l] && i <= ib[l]) slice_setup(c[l], i, ad[l], code[l]);
        const bool thin = code[l] >= 1 && code[l] <= maxcode;
        a4[l][sidx] = thin ? ad[l] : kNaN;
        if (thin) { any_thin = true; if (ad[l] > amax) amax = ad[l]; if (code[l] > mcand) mcand = code[l]; }
        if (c[l].R >= RCV_POLAR_MIN_R) {
          if (code[l] > maxcode || code[l] < 0) {
            if (i > c[l].ipx) mplus[l] |= 1u << (i - t.i0); else mminus[l] |= 1u << (i - t.i0);
            if (ad[l] > smax[l]) smax[l] = ad[l];
            const float bl = f_sub(ad[l], c[l].W);
            if (bl < slo[l]) slo[l] = bl;
          }
        } else if (-code[l] > dmax) dmax = -code[l];
      }
      if (dmax > 0) {
        ++counters[1];
        const int sbase = (i - t.i0) * slice_words;
        for (int rr = -dmax; rr <= dmax; ++rr)
          for (int kk = -dmax; kk <= dmax; ++kk)
            for (int l = 0; l < 32; ++l) {
              HostSlowPerm slow{&c[l], &emit};
              const int hb = c[l].R < RCV_POLAR_MIN_R ? -code[l] : 0;
              const bool ok = hb > 0 && rr >= -hb && rr <= hb && kk >= -hb && kk <= hb;
              dense_cell(c[l], ad[l], t, i, sbase, 1, -1, rr, kk, ok, emit, slow);
            }
      }
    }
    if (any_thin) {
      ++counters[0];
      const int H = ring_half_width(amax);
      if (noclip) {
        // fast ring pass (ring2): interior columns skip the ownership test; Hin is the warp minimum
        int Hin = 1 << 30;
        float dbm[32];
        for (int l = 0; l < 32; ++l) {
          float amin = 3.0e38f; bool any_l = false;
          for (int sidx = 0; sidx < NC; ++sidx) if (a4[l][sidx] == a4[l][sidx]) { any_l = true; if (a4[l][sidx] < amin) amin = a4[l][sidx]; }
          if (any_l) { const int h = ring2_interior(f_sub(amin, c[l].W)); if (h < Hin) Hin = h; }
          dbm[l] = any_l ? ring2_dbias_m05(c[l], amin) : c[l].dbias_m05;
        }
        for (int pass = 0; pass < 2; ++pass)
          for (int u = -H; u <= H; ++u) {
            const bool interior = Hin >= 0 && u >= -Hin && u <= Hin;
            for (int l = 0; l < 32; ++l) {
              HostSlowPerm slow{&c[l], &emit};
              HostSlowArcPerm slowarc{&slow, &emit_slow};
              const float fu = pass ? c[l].fz : c[l].fy, fv = pass ? c[l].fy : c[l].fz;
              const float cp = f_add(fv, dbm[l]), cm = f_sub(dbm[l], fv);
              const float duf = f_sub((float)u, fu);
              const float thr = ring2_thr(pass != 0, duf);
              float mu0, mu1;
              ring2_magic(pass != 0, u, t.Dp, mu0, mu1);
              ++counters[2]; ++counters[3];
              bool any = false;
              unsigned K0[NC], K1[NC], sv = 0;
              float arow[NC];
              for (int sidx = 0; sidx < NC; ++sidx) {
                arow[sidx] = a4[l][sidx];
                if (g_cat) {   // candidate census (analysis only)
                  const float gg = f_fma(-duf, duf, arow[sidx]);
                  const int ncand = 2 * mcand;
                  if (!(arow[sidx] == arow[sidx])) g_cat[c[l].R > 0 ? 0 : 4] += ncand;      // slice not drawn by this lane (or padding lane)
                  else if (!(gg >= 0.f)) g_cat[1] += ncand;                                    // column outside the ring
                  else g_cat[2] += ncand;                                                      // live candidates
                }
                ring2_consts(c[l], t, pass != 0, u, (unsigned)((i0c + sidx - t.i0) * slice_words), 1u, K0[sidx], K1[sidx], sv);
                if (mcand == 1) {
                  if (interior) any |= ring2_fast<false, 1>(c[l], arow[sidx], duf, thr, cp, cm, fv, mu0, mu1, K0[sidx], K1[sidx], sv, 0xffffffffu, emit);
                  else any |= ring2_fast<true, 1>(c[l], arow[sidx], duf, thr, cp, cm, fv, mu0, mu1, K0[sidx], K1[sidx], sv, 0xffffffffu, emit);
                } else {
                  if (interior) any |= ring2_fast<false, 2>(c[l], arow[sidx], duf, thr, cp, cm, fv, mu0, mu1, K0[sidx], K1[sidx], sv, 0xffffffffu, emit);
                  else any |= ring2_fast<true, 2>(c[l], arow[sidx], duf, thr, cp, cm, fv, mu0, mu1, K0[sidx], K1[sidx], sv, 0xffffffffu, emit);
                }
              }
              if (any)
                ring2_slow_lane(pass != 0, NC, mcand, pass ? c[l].ipz : c[l].ipy, pass ? c[l].ipy : c[l].ipz, u, i0c, c[l].hW, c[l].hw_m, c[l].hw_p, duf,
                                cp, cm, fv, mu0, mu1, sv, arow, K0, K1, slow, emit_slow);
            }
          }
      } else {
        for (int pass = 0; pass < 2; ++pass)
          for (int u = -H; u <= H; ++u)
            for (int l = 0; l < 32; ++l) {
              HostSlowPerm slow{&c[l], &emit};
              HostSlowArcPerm slowarc{&slow, &emit_slow};
              LaneTask L;
              lane_setup_pu(c[l], t, pass != 0, u, (float)u, 1, L);
              ++counters[2];
              for (int sidx = 0; sidx < NC; ++sidx) {
                const int i = i0c + sidx;
                ThinOut o;
                thin_fast<true>(c[l], a4[l][sidx], L, (i - t.i0) * slice_words, -1, emit, o);
                if (o.t0 || o.t1) thin_slow(c[l], L, i, o, slowarc);
              }
            }
      }
    }
  }
}

extern "C" __attribute__((visibility("default")))
int hostsim_render_v6(const double* p, const int* R, long n, int D, int Dp, int i0, int ni, int j0, int nj,
                      int32_t* tile, int sqrt_perturb, long long* stats) {
  g_sqrt_perturb = sqrt_perturb;
  if (ni > 32) return 2;                       // the polar pass keeps one bit per slice of the tile
  Tile t{i0, ni, j0, nj, D, Dp};
  HostEmit emit{tile, (long)ni * nj * Dp};
  HostEmitSlow emit_slow{&emit};
  long long counters[4] = {0, 0, 0, 0}, polar_cells = 0;
  const int slice_words = nj * Dp;
  for (long g0 = 0; g0 < n; g0 += 32) {
    PointCtx c[32]; int ia[32], ib[32];
    unsigned mplus[32], mminus[32]; float smax[32], slo[32];
    int wa = 1 << 30, wb = -(1 << 30);
    for (int l = 0; l < 32; ++l) {
      const long q = g0 + l;
      if (q < n) point_setup(c[l], p[3 * q + 1], p[3 * q], p[3 * q + 2], R[q]);   // (A,B,C) = (y,x,z)
      else point_setup(c[l], 0.0, 0.0, 0.0, 0);
      slice_range(c[l], t, ia[l], ib[l]);
      if (ia[l] <= ib[l]) { if (ia[l] < wa) wa = ia[l]; if (ib[l] > wb) wb = ib[l]; }
      mplus[l] = mminus[l] = 0u; smax[l] = 0.f; slo[l] = 3.0e38f;
    }
    if (wa > wb) continue;
    if (ring_chunk(ni) == 1) ring_chunks<1>(c, ia, ib, t, wa, wb, mplus, mminus, smax, slo, emit, emit_slow, counters);
    else if (ring_chunk(ni) == 2) ring_chunks<2>(c, ia, ib, t, wa, wb, mplus, mminus, smax, slo, emit, emit_slow, counters);
    else if (ring_chunk(ni) == 3) ring_chunks<3>(c, ia, ib, t, wa, wb, mplus, mminus, smax, slo, emit, emit_slow, counters);
    else ring_chunks<4>(c, ia, ib, t, wa, wb, mplus, mminus, smax, slo, emit, emit_slow, counters);
    // polar pass over the tile's polar slices
    int Hp = -1; bool anyp = false, anym = false;
    for (int l = 0; l < 32; ++l) {
      if (mplus[l] | mminus[l]) { const int h = polar_half_width(smax[l], c[l].eps); if (h > Hp) Hp = h; }
      anyp |= mplus[l] != 0u; anym |= mminus[l] != 0u;
    }
    bool p2 = true;   // fast polar pass: nobody can leave the tile and everybody's polar slices are consecutive
    Polar2Side sp[32], sm[32];
    for (int l = 0; l < 32; ++l) {
      p2 = p2 && ring_noclip(c[l], t);
      p2 = polar2_side(c[l], t, mplus[l], true, sp[l]) && p2;
      p2 = polar2_side(c[l], t, mminus[l], false, sm[l]) && p2;
    }
    if (Hp >= 0 && p2) {
      const unsigned MB = (unsigned)RCV_MAGIC_BITS;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !anyp : !anym) continue;
        for (int ub = -Hp; ub <= Hp; ++ub) {
          float r2m[32]; int ci_l[32], co_l[32], T[2] = {0, 0};
          for (int l = 0; l < 32; ++l) {
            const float db = f_sub((float)ub, c[l].fy);
            const float db2 = f_mul(db, db);
            r2m[l] = f_sub(c[l].R2, db2);
            polar_row_range(slo[l], smax[l], c[l].eps, db2, (mplus[l] | mminus[l]) != 0u, ci_l[l], co_l[l]);
            if (co_l[l] >= 0) {
              const int c1 = ci_l[l] > 1 ? ci_l[l] : 1;
              if (co_l[l] - ci_l[l] + 1 > T[0]) T[0] = co_l[l] - ci_l[l] + 1;
              if (co_l[l] - c1 + 1 > T[1]) T[1] = co_l[l] - c1 + 1;
            }
          }
          if (T[0] <= 0) continue;
          for (int seg = 0; seg < 2; ++seg) {
            const int ncell = 2 * ((T[seg] + 1) / 2);      // the device walks pairs of cells
            if (ncell <= 0) continue;
            for (int tt = 0; tt < ncell; ++tt)
              for (int l = 0; l < 32; ++l) {
                HostSlowPerm slow{&c[l], &emit};
                const Polar2Side& S = side == 0 ? sp[l] : sm[l];
                // segments end (seg 0) or start (seg 1) at the lane's own inner bound; surplus cells lie outwards
                const int start = seg == 0 ? (co_l[l] >= 0 ? -ci_l[l] : -1) - ncell + 1 : (co_l[l] >= 0 ? (ci_l[l] > 1 ? ci_l[l] : 1) : 1);
                const int uc = start + tt;
                const int jb = c[l].ipy + ub, kc = c[l].ipz + uc;
                const unsigned cellbase = (unsigned)((jb - j0) * Dp + kc);
                const unsigned vrel0 = (unsigned)(c[l].ipx - t.i0);
                const unsigned smul = side == 0 ? (unsigned)slice_words : 0u - (unsigned)slice_words;
                const unsigned K = cellbase + (side == 0 ? (vrel0 - MB) : (vrel0 + MB)) * (unsigned)slice_words;
                Polar2Cell o;
                ++polar_cells;
                polar2_cell(c[l], S, (float)uc, r2m[l], o);
                emit(o.vote ? (int)(o.bits * smul + K) : -1);
                if (o.amb) polar2_slow_cell(c[l], S, (float)uc, r2m[l], jb, kc, smul, K, slow, emit_slow);
              }
          }
        }
      }
    } else if (Hp >= 0) {
      for (int ub = -Hp; ub <= Hp; ++ub) {
        float db2[32]; int CI = 0x7fffffff, CO = -1, ci_l[32], co_l[32];
        for (int l = 0; l < 32; ++l) {
          const float db = f_sub((float)ub, c[l].fy);
          db2[l] = f_mul(db, db);
          polar_row_range(slo[l], smax[l], c[l].eps, db2[l], (mplus[l] | mminus[l]) != 0u, ci_l[l], co_l[l]);
          if (ci_l[l] < CI) CI = ci_l[l];
          if (co_l[l] > CO) CO = co_l[l];
        }
        if (CO < 0) continue;
        // every lane walks ITS OWN annulus segments [-co,-ci] and [max(ci,1),co]; trip counts are warp maxima
        for (int seg = 0; seg < 2; ++seg) {
          int T = 0, uc0[32], ucl[32];
          for (int l = 0; l < 32; ++l) {
            uc0[l] = 0; ucl[l] = -1;
            if (co_l[l] >= 0) {
              if (seg == 0) { uc0[l] = -co_l[l]; ucl[l] = -ci_l[l]; }
              else { uc0[l] = ci_l[l] > 1 ? ci_l[l] : 1; ucl[l] = co_l[l]; }
            }
            if (ucl[l] - uc0[l] + 1 > T) T = ucl[l] - uc0[l] + 1;
          }
          for (int tt = 0; tt < T; ++tt)
            for (int l = 0; l < 32; ++l) {
              HostSlowPerm slow{&c[l], &emit};
              const float cpx = f_add(c[l].fx, c[l].dbias_m05), cmx = f_sub(c[l].dbias_m05, c[l].fx);
              const int uc = uc0[l] + tt;
              const float dc = f_sub((float)uc, c[l].fz);
              const float s2 = f_fma(dc, dc, db2[l]);
              const int jb = c[l].ipy + ub, kc = c[l].ipz + uc;
              const bool ok = uc <= ucl[l] && ((unsigned)(jb - j0) < (unsigned)nj) && ((unsigned)kc < (unsigned)D);
              const int cell = (jb - j0) * Dp + kc;
              PolarOut o;
              ++polar_cells;
              if (anyp && anym) polar_fast<3>(c[l], t, cpx, cmx, s2, cell, ok, mplus[l], mminus[l], slice_words, -1, emit, o);
              else if (anyp) polar_fast<1>(c[l], t, cpx, cmx, s2, cell, ok, mplus[l], mminus[l], slice_words, -1, emit, o);
              else polar_fast<2>(c[l], t, cpx, cmx, s2, cell, ok, mplus[l], mminus[l], slice_words, -1, emit, o);
              if (o.t0 || o.t1) polar_slow(c[l].hw_m, t.i0, o, jb, kc, cell, mplus[l], mminus[l], slice_words, slow, emit_slow);
            }
        }
      }
    }
  }
  if (stats) { stats[0] = emit.votes; stats[1] = emit.calls; stats[2] = emit.oob; stats[3] = counters[0]; stats[4] = counters[1]; stats[5] = counters[2]; stats[6] = emit.slow_calls; stats[7] = polar_cells; stats[8] = counters[3]; }
  return emit.oob ? 1 : 0;
}
