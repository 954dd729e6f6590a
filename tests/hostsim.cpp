// Host-side harness for rcvpose_b200/csrc/raster_core.h (TEST INFRASTRUCTURE ONLY).
// Executes the rasteriser's per-lane code on the CPU, one simulated lane at a time, in the same
// loop structure as the CUDA kernel (vote_kernel in rcvvote.cu), so the exactness of the voxel set
// can be fuzzed against the oracle in the CPU-only container.  It is never loaded by the product.
#include <cstdint>
#include <cstdio>
#include "../rcvpose_b200/csrc/raster_core.h"

using namespace rcv;

struct HostEmit {
  int32_t* tile; long words; long long votes = 0, calls = 0, oob = 0;
  void operator()(int off, bool vote) {
    ++calls;
    if (!vote) return;
    if (off < 0 || off >= words) { ++oob; return; }
    tile[off] += 1; ++votes;
  }
};

extern "C" __attribute__((visibility("default")))
int hostsim_render(const double* p, const int* R, long n, int D, int Dp, int i0, int ni, int j0, int nj,
                   int32_t* tile, int sqrt_perturb, long long* stats) {
  g_sqrt_perturb = sqrt_perturb;
  Tile t{i0, ni, j0, nj, D, Dp};
  HostEmit emit{tile, (long)ni * nj * Dp};
  long long ring_slices = 0, dense_slices = 0, lane_tasks = 0;
  for (long q = 0; q < n; ++q) {
    PointCtx c;
    point_setup(c, p[3 * q], p[3 * q + 1], p[3 * q + 2], R[q]);
    int ia, ib;
    slice_range(c, t, ia, ib);
    if (ia > ib) continue;
    int istar = c.ipx < ia ? ia : (c.ipx > ib ? ib : c.ipx);
    SliceCtx s0;
    slice_setup(c, istar, s0);
    if (s0.a > 36.0f) {
      const int H = ring_half_width(s0.a);
      const int ntask = 2 * (2 * H + 1);
      for (int base = 0; base < ntask; base += 32)
        for (int lane = 0; lane < 32; ++lane) {
          LaneTask L;
          lane_setup(c, t, H, base + lane, L);
          for (int i = ia; i <= ib; ++i) {
            SliceCtx s;
            slice_setup(c, i, s);
            if (s.kind != SLICE_RING) continue;
            if (lane == 0 && base == 0) ++ring_slices;
            ++lane_tasks;
            ring_lane(c, s, L, i, (i - i0) * nj * Dp, emit);
          }
        }
    }
    for (int i = ia; i <= ib; ++i) {
      SliceCtx s;
      slice_setup(c, i, s);
      if (s.kind != SLICE_DENSE) continue;
      ++dense_slices;
      const int side = 2 * s.m + 1;
      for (int cell = 0; cell < ((side * side + 31) / 32) * 32; ++cell) dense_cell(c, s, t, i, (i - i0) * nj * Dp, cell, emit);
    }
  }
  if (stats) { stats[0] = emit.votes; stats[1] = emit.calls; stats[2] = emit.oob; stats[3] = ring_slices; stats[4] = dense_slices; stats[5] = lane_tasks; }
  return emit.oob ? 1 : 0;
}
