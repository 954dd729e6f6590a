// Host-side harness for rcvpose_b200/csrc/raster_core.h (TEST INFRASTRUCTURE ONLY).
// Executes the rasteriser's per-lane code on the CPU, one simulated lane at a time, in the same
// loop structure as the CUDA kernel (k_vote in rcvvote.cu: 32-slice groups, lane chunks, ring slices
// then dense slices), so the exactness of the voxel set can be fuzzed against the oracle in the
// CPU-only container.  It is never loaded by the product.
#include <cstdint>
#include <cstdio>
#include "../rcvpose_b200/csrc/raster_core.h"

using namespace rcv;

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
struct HostSlow {
  const PointCtx* c; HostEmit* e;
  bool operator()(int i, int j, int k) { ++e->slow_calls; return exact_hit(c->px, c->py, c->pz, c->R, i, j, k); }
};

struct HostSlowArc {
  HostSlow* slow; HostEmitSlow* es;
  void operator()(const PointCtx& c, const LaneTask& L, int i, int ub, int m, int arc, int cc, float q, float fl, int vt) {
    ring_slow(c, L, i, ub, m, arc, cc, q, fl, vt, *slow, *es);
  }
};

extern "C" __attribute__((visibility("default")))
int hostsim_render(const double* p, const int* R, long n, int D, int Dp, int i0, int ni, int j0, int nj,
                   int32_t* tile, int sqrt_perturb, long long* stats) {
  g_sqrt_perturb = sqrt_perturb;
  Tile t{i0, ni, j0, nj, D, Dp};
  HostEmit emit{tile, (long)ni * nj * Dp};
  HostEmitSlow emit_slow{&emit};
  long long ring_slices = 0, dense_slices = 0, lane_tasks = 0;
  for (long q = 0; q < n; ++q) {
    PointCtx c;
    point_setup(c, p[3 * q], p[3 * q + 1], p[3 * q + 2], R[q]);
    HostSlow slow{&c, &emit};
    HostSlowArc slowarc{&slow, &emit_slow};
    int ia, ib;
    slice_range(c, t, ia, ib);
    for (int sb = ia; sb <= ib; sb += 32) {
      float a_l[32]; int code_l[32];
      float amax = 0.f; bool any_ring = false;
      for (int lane = 0; lane < 32; ++lane) {
        a_l[lane] = 0.f; code_l[lane] = 0;
        if (sb + lane <= ib) slice_setup(c, sb + lane, a_l[lane], code_l[lane]);
        if (code_l[lane] > 0) { any_ring = true; if (a_l[lane] > amax) amax = a_l[lane]; }
      }
      if (any_ring) {
        const int H = ring_half_width(amax);
        const int ntask = 2 * (2 * H + 1);
        for (int base = 0; base < ntask; base += 32)
          for (int lane = 0; lane < 32; ++lane) {
            LaneTask L;
            lane_setup(c, t, H, base + lane, 1, L);
            // thin slices first, in pairs (as the kernel does), then the general ones
            int thin[32], nthin = 0, thick[32], nthick = 0;
            for (int sl = 0; sl < 32; ++sl) { if (code_l[sl] == 1) thin[nthin++] = sl; else if (code_l[sl] > 1) thick[nthick++] = sl; }
            for (int s = 0; s < nthin; s += 2) {
              ThinOut oa, ob;
              const int ia_ = sb + thin[s];
              thin_fast(c, a_l[thin[s]], L, (ia_ - i0) * nj * Dp, -1, emit, oa);
              const bool two = s + 1 < nthin;
              const int ib_ = two ? sb + thin[s + 1] : 0;
              if (two) thin_fast(c, a_l[thin[s + 1]], L, (ib_ - i0) * nj * Dp, -1, emit, ob);
              if (oa.t0 || oa.t1) thin_slow(c, L, ia_, oa, slowarc);
              if (two && (ob.t0 || ob.t1)) thin_slow(c, L, ib_, ob, slowarc);
              lane_tasks += two ? 2 : 1;
            }
            for (int s = 0; s < nthick; ++s) {
              const int i = sb + thick[s];
              ++lane_tasks;
              ring_lane(c, a_l[thick[s]], code_l[thick[s]], L, i, (i - i0) * nj * Dp, -1, emit, slowarc);
            }
          }
      }
      for (int sl = 0; sl < 32; ++sl) {
        if (code_l[sl] >= 0) continue;
        const int i = sb + sl, hb = -code_l[sl], side = 2 * hb + 1;
        ++dense_slices;
        const int lpr = side <= 16 ? 16 : 32, rpi = 32 / lpr;   // lanes per row, rows per warp iteration
        for (int r0 = 0; r0 < side; r0 += rpi)
          for (int k0 = 0; k0 < side; k0 += lpr)
            for (int lane = 0; lane < 32; ++lane) {
              const int rr = r0 + lane / lpr, kk = k0 + lane % lpr;
              dense_cell(c, a_l[sl], t, i, (i - i0) * nj * Dp, 1, -1, rr - hb, kk - hb, rr < side && kk < side, emit, slow);
            }
      }
    }
  }
  if (stats) { stats[0] = emit.votes; stats[1] = emit.calls; stats[2] = emit.oob; stats[3] = ring_slices; stats[4] = dense_slices; stats[5] = lane_tasks; stats[6] = emit.slow_calls; }
  return emit.oob ? 1 : 0;
}
