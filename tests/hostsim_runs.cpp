// Host-side harness for rcvpose_b200/csrc/runs_core.h (TEST INFRASTRUCTURE ONLY).
// Executes the run-length rasteriser's per-lane code on the CPU with 32 simulated lanes in lockstep, in the loop
// structure of the CUDA kernel (k_vote in rcvvote.cu: lane = point, chunks of NC slices, columns u = -H..H with H the
// warp maximum, flagged columns re-decided after each block of 32 columns, prefix sum per row at the end), so that the
// exactness of the voxel set can be fuzzed against the oracle in the CPU-only container.  Never loaded by the product.
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../rcvpose_b200/csrc/runs_core.h"

using namespace rcv;

namespace {
struct Stats { long long cols = 0, live_cols = 0, flagged_cols = 0, exact_calls = 0, fixes = 0, oob = 0, atomics = 0; };

// Tile of the difference array: slices [i0, i0+ni) of A, rows [j0, j0+nj) of B, cells [-glo, D+ghi] of C
// (one spare cell above the upper guard receives the -1 of a run that ends at the clip bound).
struct HTile {
  int i0, ni, j0, nj, D, Dp, glo, ghi;
  int32_t* w; long words;
  long cell(int i, int j, int k) const { return ((long)(i - i0) * nj + (j - j0)) * Dp + (k + glo); }
};

template <int NC, bool CLIP>
void render_group(const RunPoint* c_in, const double (*pd)[3], const HTile& t, Stats& st) {
  RunLane L[32];
  RunPoint cc[32];
  for (int l = 0; l < 32; ++l) {
    cc[l] = c_in[l];
    if (cc[l].R <= 0) { cc[l].ipc = t.glo < 0 ? -t.glo : 0; cc[l].fc = 0.f; }     // k_vote_runs: a lane that draws nothing marks a cell the row stores
    run_lane_setup(cc[l], L[l]);
  }
  const RunPoint* c = cc;
  for (int i0c = t.i0; i0c < t.i0 + t.ni; i0c += NC) {
    f2 aa[32][NC];
    // runs_chunk (rcvvote.cu): every lane's own column range [ulo, uhi] restricted to the tile's rows (a no-op for guarded
    // whole-slice tiles, the band restriction otherwise); a lane without columns parks on a row inside the tile
    int ulo[32], uhi[32], upark[32]; bool some[32];
    int wlo = 0x7fffffff, whi = -0x7fffffff;
    for (int l = 0; l < 32; ++l) {
      float amax = -1.f;
      for (int s = 0; s < NC; ++s) {
        const int i = i0c + s;
        aa[l][s] = run_slice_consts(c[l], L[l], i, i < t.i0 + t.ni);
        if (f2_lo(aa[l][s]) > amax) amax = f2_lo(aa[l][s]);
      }
      const int Hl = run_half_width(amax);
      ulo[l] = -Hl > t.j0 - c[l].ipb ? -Hl : t.j0 - c[l].ipb;
      uhi[l] = Hl < t.j0 + t.nj - 1 - c[l].ipb ? Hl : t.j0 + t.nj - 1 - c[l].ipb;
      if (Hl < 0) { ulo[l] = 1; uhi[l] = 0; }
      some[l] = ulo[l] <= uhi[l];
      if (some[l]) { if (ulo[l] < wlo) wlo = ulo[l]; if (uhi[l] > whi) whi = uhi[l]; }
      int up = 0 > t.j0 - c[l].ipb ? 0 : t.j0 - c[l].ipb;
      if (up > t.j0 + t.nj - 1 - c[l].ipb) up = t.j0 + t.nj - 1 - c[l].ipb;
      upark[l] = up;
      if (!some[l]) { ulo[l] = up; uhi[l] = up; }
    }
    if (wlo > whi) continue;
    for (int u0 = wlo; u0 <= whi; u0 += 32) {
      unsigned flags[32] = {0};
      for (int jj = 0; jj < 32 && u0 + jj <= whi; ++jj) {
        const int u = u0 + jj;
        for (int l = 0; l < 32; ++l) {
          // beyond its own range the lane parks on its edge row and the column is empty by construction
          const int uc = u < ulo[l] ? ulo[l] : (u > uhi[l] ? uhi[l] : u);
          const bool skip = uc != u || !some[l];
          const int row = c[l].ipb + uc;
          RunCol C;
          run_col_setup(c[l], u, uc, t.Dp, C);
          if (skip) { C.du = f2_dup(1.0e18f); C.ndu = f2_dup(-1.0e18f); }     // g'' = -huge: the column is empty
          const unsigned base = (unsigned)RCV_MAGIC_BITS + (unsigned)(uc * t.Dp);
          ++st.cols;
          for (int s = 0; s < NC; ++s) {
            const int i = i0c + s;
            if (i >= t.i0 + t.ni) continue;
            RunOut o;
            run_slice(L[l], C, aa[l][s], o);
            int b[4] = {(int)(o.b1 - base), (int)(o.b2 - base), (int)(o.b3 - base), (int)(o.b4 - base)};
            if (b[0] != b[3]) ++st.live_cols;
            const long rowbase = t.cell(i, c[l].ipb, 0) + c[l].ipc;            // cell of lattice offset 0 in row ipb
            for (int e = 0; e < 4; ++e) {
              int n = b[e];
              if (CLIP) { const int lo = -t.glo - c[l].ipc, hi = t.D + t.ghi - c[l].ipc; n = n < lo ? lo : (n > hi ? hi : n); }
              const long off = rowbase + (long)uc * t.Dp + n;
              ++st.atomics;
              if (off < 0 || off >= t.words || (c[l].ipc + n) < -t.glo || (c[l].ipc + n) > t.D + t.ghi || row < t.j0 || row >= t.j0 + t.nj) { ++st.oob; continue; }
              t.w[off] += (e & 1) ? -1 : 1;
            }
            if (run_flagged(o)) flags[l] |= 1u << jj;
          }
        }
      }
      // flagged columns: exact decisions
      for (int l = 0; l < 32; ++l) {
        for (int jj = 0; jj < 32; ++jj) {
          if (!((flags[l] >> jj) & 1u)) continue;
          ++st.flagged_cols;
          const int u = u0 + jj;
          const int uc = u < ulo[l] ? ulo[l] : (u > uhi[l] ? uhi[l] : u);
          if (uc != u || !some[l]) { ++st.oob; continue; }   // (a parked column is empty and must never be flagged)
          const int row = c[l].ipb + u;
          RunCol C;
          run_col_setup(c[l], u, uc, t.Dp, C);
          const unsigned base = (unsigned)RCV_MAGIC_BITS + (unsigned)(uc * t.Dp);
          for (int s = 0; s < NC; ++s) {
            const int i = i0c + s;
            if (i >= t.i0 + t.ni) continue;
            const long rowbase = t.cell(i, c[l].ipb, 0) + c[l].ipc + (long)uc * t.Dp;
            auto exact = [&](int m) {
              ++st.exact_calls;
              // internal (A,B,C) = reference (y,x,z); exact_hit takes the reference order
              return exact_hit(pd[l][0], pd[l][1], pd[l][2], c[l].R, row, i, c[l].ipc + m);
            };
            auto fix = [&](unsigned nbits, int delta) {
              const int m = (int)(nbits - base);
              for (int q = 0; q < 2; ++q) {
                int n = m + q;
                if (CLIP) { const int lo = -t.glo - c[l].ipc, hi = t.D + t.ghi - c[l].ipc; n = n < lo ? lo : (n > hi ? hi : n); }
                const long off = rowbase + n;
                if (off < 0 || off >= t.words || (c[l].ipc + n) < -t.glo || (c[l].ipc + n) > t.D + t.ghi) { ++st.oob; continue; }
                t.w[off] += q ? -delta : delta;
              }
            };
            st.fixes += run_slow_slice(L[l], C, aa[l][s], base, exact, fix);
          }
        }
      }
    }
  }
}
}  // namespace

// p: n x 3 float64 in REFERENCE order (x,y,z), R: n int32.  Renders the tile (internal axes: A = y slices [a0, a0+na),
// B = x rows [j0, j0+nj), C = z) and returns the COUNTS (prefix-summed) in out[na][nj][D] (reference voxel (x, y, z) =
// (j, a, k)).  clip != 0 exercises the clipped variant (required when guards are too small or rows are a band).
extern "C" __attribute__((visibility("default")))
int hostsim_runs_render(const double* p, const int* R, long n, int D, int glo, int ghi, int a0, int na, int j0, int nj, int NC, int clip,
                        int32_t* out, int sqrt_perturb, long long* stats) {
  g_sqrt_perturb = sqrt_perturb;
  HTile t;
  t.i0 = a0; t.ni = na; t.j0 = j0; t.nj = nj; t.D = D; t.glo = glo; t.ghi = ghi;
  t.Dp = (D + glo + ghi + 1) | 1;
  t.words = (long)na * nj * t.Dp;
  std::vector<int32_t> buf(t.words, 0);
  t.w = buf.data();
  Stats st;
  for (long g0 = 0; g0 < n; g0 += 32) {
    RunPoint c[32]; double pd[32][3];
    for (int l = 0; l < 32; ++l) {
      const long q = g0 + l;
      if (q < n) { pd[l][0] = p[3 * q]; pd[l][1] = p[3 * q + 1]; pd[l][2] = p[3 * q + 2]; run_point_setup(c[l], p[3 * q + 1], p[3 * q], p[3 * q + 2], R[q]); }
      else { pd[l][0] = pd[l][1] = pd[l][2] = 0.0; run_point_setup(c[l], 0.0, 0.0, 0.0, 0); }
    }
#define RG(N) do { if (clip) render_group<N, true>(c, pd, t, st); else render_group<N, false>(c, pd, t, st); } while (0)
    if (NC == 1) RG(1); else if (NC == 2) RG(2); else if (NC == 3) RG(3); else RG(4);
#undef RG
  }
  // epilogue: prefix sum along C per row, from the first guard cell
  for (int a = 0; a < na; ++a)
    for (int j = 0; j < nj; ++j) {
      const int32_t* row = t.w + ((long)a * nj + j) * t.Dp;
      int32_t run = 0;
      for (int k = -glo; k < D + ghi; ++k) {
        run += row[k + glo];
        if (k >= 0 && k < D) out[((long)a * nj + j) * D + k] = run;
      }
    }
  if (stats) { stats[0] = st.cols; stats[1] = st.live_cols; stats[2] = st.flagged_cols; stats[3] = st.exact_calls; stats[4] = st.fixes; stats[5] = st.oob; stats[6] = st.atomics; }
  return st.oob ? 1 : 0;
}
