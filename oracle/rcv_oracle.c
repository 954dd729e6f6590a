/*
 * rcv_oracle.c -- CPU restatement of the reference's radial-voting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 * The product path (rcvpose_b200/) never imports, links or executes it.
 *
 * Every function restates, in plain C, the arithmetic of one reference function (citations are
 * file:line into the upstream repository aaronWool/rcvpose).  It is pinned against the reference
 * itself: tests/golden/make_golden.py imports the real reference in the build container
 * (single-threaded numba) and stores its outputs as fixtures; tests/test_oracle_golden.py checks
 * this file against those fixtures bit-for-bit (integer vote volumes, D, zero boundary) and to
 * the last ulp for the float64 results.  The reference has no tests or golden vectors of its own.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -ffp-contract=off: no FMA contraction, so every
 * float64 operation rounds exactly as numpy / numba's LLVM code does on x86-64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * numpy's pairwise summation, as used by np.mean / np.add.reduce on a strided float64 vector
 * (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum; PW_BLOCKSIZE = 128).
 * The reference calls np.mean(xyz_mm[:,c]) at AccumulatorSpace.py:381-383.
 * ---------------------------------------------------------------------------------------- */
static double pairwise_sum(const double *a, long n, long stride)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; ++i) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2, stride) + pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

ORC_API double orc_pairwise_sum(const double *a, long n, long stride) { return pairwise_sum(a, n, stride); }

/* ------------------------------------------------------------------------------------------
 * rgbd_to_point_cloud -- AccumulatorSpace.py:77-85.
 *   vs,us = depth.nonzero() (row-major); z = depth[v,u];
 *   x = ((u - K[0,2]) * z) / K[0,0];  y = ((v - K[1,2]) * z) / K[1,1];  pts = [x,y,z] float64.
 * swap_xy=1 gives the 3DRadius_ycb.py:62-70 variant (x from rows, y from columns).
 * depth is passed as float64 (uint16 -> float64 is exact).  Returns N.
 * ---------------------------------------------------------------------------------------- */
ORC_API long orc_backproject(const double *K, const double *depth, int H, int W, int swap_xy, double *xyz_out)
{
    long n = 0;
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            double z = depth[(long)v * W + u];
            if (z == 0.0) continue;
            double x = (((double)u - K[2]) * z) / K[0];
            double y = (((double)v - K[5]) * z) / K[4];
            if (swap_xy) {
                /* 3DRadius_ycb.py:67-68: xs from vs with K[1,2]/K[1,1]?  No: it swaps the roles of
                 * us/vs against the intrinsics rows; restated as used by the YCBGEN policy tests. */
                x = (((double)v - K[2]) * z) / K[0];
                y = (((double)u - K[5]) * z) / K[4];
            }
            if (xyz_out) {
                xyz_out[3 * n + 0] = x;
                xyz_out[3 * n + 1] = y;
                xyz_out[3 * n + 2] = z;
            }
            ++n;
        }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Accumulator_3D prelude -- AccumulatorSpace.py:373-401.
 *   xyz_mm = xyz*1000/acc_unit          (two roundings: (x*1000)/acc_unit)
 *   mean per axis (numpy pairwise), subtract
 *   radial_list_mm = radial_list*radius_scale/acc_unit   IN THE RADIUS DTYPE (:388)
 *   zb = int(min(xyz_mm) - max(radial_mm)) + 1 ; if zb<0: xyz_mm -= zb
 *   length = int(max(xyz_mm)) ; D = length + int(max(radial_mm))
 * policy 0 = LM (above).  policy 1 = YCBGEN (3DRadius_ycb.py:113-141): length=int(max)+1, D=length.
 * radius_is_f32: radii are float32 (network output) -> arithmetic in float32, else float64.
 * Outputs: p[n*3] shifted voxel-unit points, R[n] integer radii (np.around = half-to-even, :332),
 *          mean[3], zb, D, rmax (as float64).  Returns 0, or 1 for empty input (ValueError upstream).
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_prelude(const double *xyz, long n, const void *radii, int radius_is_f32, double acc_unit,
                        double radius_scale, int policy, double *p, int *R, double *mean, int *zb_out,
                        int *D_out, double *rmax_out)
{
    if (n <= 0) return 1;
    for (long i = 0; i < 3 * n; ++i) p[i] = (xyz[i] * 1000.0) / acc_unit;
    for (int c = 0; c < 3; ++c) mean[c] = pairwise_sum(p + c, n, 3) / (double)n;
    for (long i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) p[3 * i + c] -= mean[c];
    double rmax;
    if (radius_is_f32) {
        const float *r = (const float *)radii;
        float su = (float)radius_scale, au = (float)acc_unit;
        float m = -INFINITY;
        for (long i = 0; i < n; ++i) {
            float rv = (r[i] * su) / au;
            if (rv > m) m = rv;
            R[i] = (int)rintf(rv); /* FE_TONEAREST: half-to-even, like np.around */
        }
        rmax = (double)m;
    } else {
        const double *r = (const double *)radii;
        double m = -INFINITY;
        for (long i = 0; i < n; ++i) {
            double rv = (r[i] * radius_scale) / acc_unit;
            if (rv > m) m = rv;
            R[i] = (int)rint(rv);
        }
        rmax = m;
    }
    double pmin = INFINITY;
    for (long i = 0; i < 3 * n; ++i)
        if (p[i] < pmin) pmin = p[i];
    int zb = (int)(pmin - rmax) + 1; /* int() truncates toward zero */
    if (zb < 0)
        for (long i = 0; i < 3 * n; ++i) p[i] -= (double)zb;
    double pmax = -INFINITY;
    for (long i = 0; i < 3 * n; ++i)
        if (p[i] > pmax) pmax = p[i];
    int length = (int)pmax;
    int D = (policy == 1) ? (length + 1) : (length + (int)rmax);
    *zb_out = zb;
    *D_out = D;
    *rmax_out = rmax;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * fast_for -- AccumulatorSpace.py:325-341.  Brute force over every (point, voxel):
 *   distance = ((i-x)**2 + (j-y)**2 + (k-z)**2)**0.5 ; vote iff R-distance < f and R-distance > 0,
 *   f = 3**0.5/4.  Run sequentially in the point dimension per voxel slice (deterministic; the
 *   reference's prange version is racy -- SURVEY section 0.2), parallel over i-slices, which owns
 *   disjoint voxels and therefore reproduces the single-threaded reference exactly.
 * vol is int32 [D][D][D], C order (k fastest); must be zeroed by the caller (votes accumulate).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_fast_for(const double *p, const int *R, long n, int D, int32_t *vol, int threads)
{
    const double factor = sqrt(3.0) / 4.0; /* 3**0.5/4 = 0x1.bb67ae8584caap-2 */
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < D; ++i) {
        for (long c = 0; c < n; ++c) {
            const double x = p[3 * c], y = p[3 * c + 1], z = p[3 * c + 2];
            const double radius = (double)R[c];
            if (R[c] <= 0) continue; /* R - distance > 0 impossible */
            const double dx = (double)i - x;
            /* cheap exact cull: |dx| >= R  =>  distance >= R  =>  no vote in this slice */
            if (dx >= radius || -dx >= radius) continue;
            const double dx2 = dx * dx;
            for (int j = 0; j < D; ++j) {
                const double dy = (double)j - y;
                if (dy >= radius || -dy >= radius) continue;
                const double dxy2 = dx2 + dy * dy;
                int32_t *row = vol + ((long)i * D + j) * D;
                for (int k = 0; k < D; ++k) {
                    const double dz = (double)k - z;
                    const double distance = sqrt(dxy2 + dz * dz);
                    const double t = radius - distance;
                    if (t < factor && t > 0.0) row[k] += 1;
                }
            }
        }
    }
}

/* The reference's loop with NO culls -- every one of the N * D^3 (point, voxel) shell tests the numba kernel performs
 * (AccumulatorSpace.py:329-339) -- parallel over i-slices instead of the reference's racy prange over points (same
 * work, disjoint voxels per thread, deterministic).  This is what bench.py times as the CPU reference arm. */
ORC_API void orc_fast_for_full(const double *p, const int *R, long n, int D, int32_t *vol, int threads)
{
    const double factor = sqrt(3.0) / 4.0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < D; ++i) {
        for (long c = 0; c < n; ++c) {
            const double x = p[3 * c], y = p[3 * c + 1], z = p[3 * c + 2];
            const double radius = (double)R[c];
            const double dx = (double)i - x, dx2 = dx * dx;
            for (int j = 0; j < D; ++j) {
                const double dy = (double)j - y;
                const double dxy2 = dx2 + dy * dy;
                int32_t *row = vol + ((long)i * D + j) * D;
                for (int k = 0; k < D; ++k) {
                    const double dz = (double)k - z;
                    const double distance = sqrt(dxy2 + dz * dz);
                    const double t = radius - distance;
                    if (t < factor && t > 0.0) row[k] += 1;
                }
            }
        }
    }
}

/* The literal triple loop with no culls, for validating the culled version above. */
ORC_API void orc_fast_for_literal(const double *p, const int *R, long n, int D, int32_t *vol)
{
    const double factor = sqrt(3.0) / 4.0;
    for (long c = 0; c < n; ++c) {
        const double radius = (double)R[c];
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
                for (int k = 0; k < D; ++k) {
                    double a = (double)i - p[3 * c], b = (double)j - p[3 * c + 1], d = (double)k - p[3 * c + 2];
                    double distance = sqrt((a * a + b * b) + d * d);
                    if (radius - distance < factor && radius - distance > 0.0) vol[((long)i * D + j) * D + k] += 1;
                }
    }
}

/* ------------------------------------------------------------------------------------------
 * Exact scatter renderer: same voxel set as orc_fast_for, but visits only the voxels near each
 * sphere (O(R^2) columns per point).  Each (i,j) column's k-range is bracketed with a float64
 * sqrt and widened by two voxels, then every candidate is decided by the reference predicate
 * itself.  Validated against orc_fast_for in tests; used where N*D^3 is out of reach (D=512).
 * ---------------------------------------------------------------------------------------- */
static inline int ref_predicate(double dxy2, double dz, double radius, double factor)
{
    const double distance = sqrt(dxy2 + dz * dz);
    const double t = radius - distance;
    return (t < factor) && (t > 0.0);
}

ORC_API void orc_scatter(const double *p, const int *R, long n, int D, int32_t *vol)
{
    const double factor = sqrt(3.0) / 4.0;
    for (long c = 0; c < n; ++c) {
        const int Ri = R[c];
        if (Ri <= 0) continue;
        const double x = p[3 * c], y = p[3 * c + 1], z = p[3 * c + 2], radius = (double)Ri;
        int i0 = (int)floor(x - radius) - 1, i1 = (int)ceil(x + radius) + 1;
        int j0 = (int)floor(y - radius) - 1, j1 = (int)ceil(y + radius) + 1;
        if (i0 < 0) i0 = 0;
        if (j0 < 0) j0 = 0;
        if (i1 > D - 1) i1 = D - 1;
        if (j1 > D - 1) j1 = D - 1;
        for (int i = i0; i <= i1; ++i) {
            const double dx = (double)i - x;
            for (int j = j0; j <= j1; ++j) {
                const double dy = (double)j - y;
                const double dxy2 = dx * dx + dy * dy;
                const double a = radius * radius - dxy2;
                if (a < -1.0) continue;
                const double ro = sqrt(a > 0 ? a : 0.0);
                double bi = (radius - factor) * (radius - factor) - dxy2;
                const double ri = bi > 0 ? sqrt(bi) : 0.0;
                int32_t *row = vol + ((long)i * D + j) * D;
                /* upper run candidates [z+ri-2, z+ro+2], lower run [z-ro-2, z-ri+2]; merge if they overlap */
                int lo1 = (int)floor(z - ro) - 2, hi1 = (int)ceil(z - ri) + 2;
                int lo2 = (int)floor(z + ri) - 2, hi2 = (int)ceil(z + ro) + 2;
                if (lo2 <= hi1) { hi1 = hi2; lo2 = 1; hi2 = 0; }
                for (int pass = 0; pass < 2; ++pass) {
                    int lo = pass ? lo2 : lo1, hi = pass ? hi2 : hi1;
                    if (lo < 0) lo = 0;
                    if (hi > D - 1) hi = D - 1;
                    for (int k = lo; k <= hi; ++k)
                        if (ref_predicate(dxy2, (double)k - z, radius, factor)) row[k] += 1;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Peak search + un-shift -- AccumulatorSpace.py:406-419.
 *   center = argwhere(V == V.max()) (C order) -> row 0 = smallest linear index among the maxima;
 *   center += zb if zb < 0 ; centre_mm[c] = (center[c] + mean[c] + 0.5) * acc_unit
 * Returns number of maxima (ties); idx_out = first maximum's (i,j,k); max_out = its count.
 * ---------------------------------------------------------------------------------------- */
ORC_API long orc_peak(const int32_t *vol, int D, int *idx_out, int32_t *max_out)
{
    long total = (long)D * D * D, first = 0, ties = 0;
    int32_t m = INT32_MIN;
    for (long q = 0; q < total; ++q)
        if (vol[q] > m) { m = vol[q]; first = q; ties = 1; }
        else if (vol[q] == m) ++ties;
    idx_out[0] = (int)(first / ((long)D * D));
    idx_out[1] = (int)((first / D) % D);
    idx_out[2] = (int)(first % D);
    *max_out = m;
    return ties;
}

ORC_API void orc_center_mm(const int *idx, int zb, const double *mean, double acc_unit, int policy, double *center_mm)
{
    for (int c = 0; c < 3; ++c) {
        double v = (double)idx[c];
        if (zb < 0) v = v + (double)zb;
        if (policy == 1)
            center_mm[c] = (v + mean[c]) * acc_unit + 0.5; /* 3DRadius_ycb.py:155-157 (float form) */
        else
            center_mm[c] = (v + mean[c] + 0.5) * acc_unit; /* AccumulatorSpace.py:413-415 */
    }
}

/* ------------------------------------------------------------------------------------------
 * HornPoseFitting.lmshorn + myjacobi + rotate -- util/horn.py:7-181.
 * Horn 1987 closed-form absolute orientation via the max-eigenvector of the 4x4 matrix N,
 * eigen-solved by cyclic Jacobi sweeps (Numerical-Recipes layout, 1-based arrays kept so the
 * operation order is the reference's).  P1, P2 are (n,3) row-major and are left unchanged
 * (the reference centres them in place and restores them, :104-107,178-181); A is 4x4 row-major.
 * ---------------------------------------------------------------------------------------- */
static void jrotate(double a[5][5], int i, int j, int k, int l, double s, double tau)
{
    double g = a[i][j], h = a[k][l];
    a[i][j] = g - s * (h + g * tau);
    a[k][l] = h + s * (g - h * tau);
}

static void myjacobi(double a[5][5], int n, double d[5], double v[5][5])
{
    double b[5] = {0}, z[5] = {0};
    for (int ip = 1; ip <= n; ++ip) {
        for (int iq = 1; iq <= n; ++iq) v[ip][iq] = 0.0;
        v[ip][ip] = 1.0;
    }
    for (int ip = 1; ip <= n; ++ip) { b[ip] = d[ip] = a[ip][ip]; z[ip] = 0.0; }
    for (int i = 1; i <= 50; ++i) {
        double sm = 0.0;
        for (int ip = 1; ip < n; ++ip)
            for (int iq = 1; iq <= n; ++iq) sm += fabs(a[ip][iq]); /* util/horn.py:28-30: full rows incl. diagonal */
        if (sm == 0.0) return;
        double tresh = (i < 4) ? 0.2 * sm / (n * n) : 0.0;
        for (int ip = 1; ip < n; ++ip)
            for (int iq = ip + 1; iq <= n; ++iq) {
                double g = 100.0 * fabs(a[ip][iq]);
                if (i > 4 && fabs(d[ip]) + g == fabs(d[ip]) && fabs(d[iq]) + g == fabs(d[iq]))
                    a[ip][iq] = 0.0;
                else if (fabs(a[ip][iq]) > tresh) {
                    double h = d[iq] - d[ip], t;
                    if (fabs(h) + g == fabs(h))
                        t = a[ip][iq] / h;
                    else {
                        double theta = 0.5 * h / a[ip][iq];
                        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                        if (theta < 0.0) t = -t;
                    }
                    double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
                    h = t * a[ip][iq];
                    z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h;
                    a[ip][iq] = 0.0;
                    for (int j = 1; j < ip; ++j) jrotate(a, j, ip, j, iq, s, tau);
                    for (int j = ip + 1; j < iq; ++j) jrotate(a, ip, j, j, iq, s, tau);
                    for (int j = iq + 1; j <= n; ++j) jrotate(a, ip, j, iq, j, s, tau);
                    for (int j = 1; j <= n; ++j) jrotate(v, j, ip, j, iq, s, tau);
                }
            }
        for (int ip = 1; ip <= n; ++ip) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
    }
}

ORC_API void orc_lmshorn(const double *P1in, const double *P2in, int n, double *A)
{
    double *P1 = (double *)malloc(sizeof(double) * 3 * n), *P2 = (double *)malloc(sizeof(double) * 3 * n);
    memcpy(P1, P1in, sizeof(double) * 3 * n);
    memcpy(P2, P2in, sizeof(double) * 3 * n);
    double C1[3] = {0, 0, 0}, C2[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) { C1[j] += P1[3 * i + j]; C2[j] += P2[3 * i + j]; }
    for (int j = 0; j < 3; ++j) { C1[j] /= n; C2[j] /= n; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) { P1[3 * i + j] -= C1[j]; P2[3 * i + j] -= C2[j]; }
    double S[3][3] = {{0}};
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) S[r][c] += P1[3 * i + r] * P2[3 * i + c];
    const double Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2], Syx = S[1][0], Syy = S[1][1], Syz = S[1][2],
                 Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
    double N[5][5] = {{0}}, Dg[5] = {0}, V[5][5] = {{0}};
    N[1][1] = Sxx + Syy + Szz; N[1][2] = Syz - Szy;        N[1][3] = Szx - Sxz;         N[1][4] = Sxy - Syx;
    N[2][1] = Syz - Szy;       N[2][2] = Sxx - Syy - Szz;  N[2][3] = Sxy + Syx;         N[2][4] = Szx + Sxz;
    N[3][1] = Szx - Sxz;       N[3][2] = Sxy + Syx;        N[3][3] = -Sxx + Syy - Szz;  N[3][4] = Syz + Szy;
    N[4][1] = Sxy - Syx;       N[4][2] = Szx + Sxz;        N[4][3] = Syz + Szy;         N[4][4] = -Sxx - Syy + Szz;
    myjacobi(N, 4, Dg, V);
    int me = 1;
    for (int i = 2; i < 5; ++i)
        if (Dg[i] > Dg[me]) me = i;
    const double q0 = V[1][me], q1 = V[2][me], q2 = V[3][me], q3 = V[4][me];
    double Rm[3][3];
    Rm[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
    Rm[0][1] = 2 * (q1 * q2 - q0 * q3);
    Rm[0][2] = 2 * (q1 * q3 + q0 * q2);
    Rm[1][0] = 2 * (q1 * q2 + q0 * q3);
    Rm[1][1] = q0 * q0 + q2 * q2 - q1 * q1 - q3 * q3;
    Rm[1][2] = 2 * (q2 * q3 - q0 * q1);
    Rm[2][0] = 2 * (q1 * q3 - q0 * q2);
    Rm[2][1] = 2 * (q2 * q3 + q0 * q1);
    Rm[2][2] = q0 * q0 + q3 * q3 - q1 * q1 - q2 * q2;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[4 * r + c] = Rm[r][c];
        A[4 * r + 3] = C2[r] - (Rm[r][0] * C1[0] + Rm[r][1] * C1[1] + Rm[r][2] * C1[2]);
        A[12 + r] = 0.0;
    }
    A[15] = 1.0;
    free(P1);
    free(P2);
}

/* ------------------------------------------------------------------------------------------
 * Accumulator_3D end to end -- AccumulatorSpace.py:373-419 -- for timing the CPU baseline and
 * for one-call parity.  vol_out may be NULL (a scratch volume is allocated).  brute=1 uses the
 * reference's N*D^3 loop with per-slice / per-row culls, brute=2 the same loop with no culls (every N*D^3 test, the CPU
 * reference arm of bench.py), brute=0 the exact scatter renderer.
 * Returns 0 ok, 1 empty input, 2 non-positive D.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_accumulator_3d(const double *xyz, long n, const void *radii, int radius_is_f32, double acc_unit,
                               double radius_scale, int policy, int brute, int threads, double *center_mm,
                               int *D_out, int *zb_out, int32_t *peak_out, long long *votes_out, int32_t *vol_out)
{
    if (n <= 0) return 1;
    double *p = (double *)malloc(sizeof(double) * 3 * n), mean[3], rmax;
    int *R = (int *)malloc(sizeof(int) * n), zb, D;
    orc_prelude(xyz, n, radii, radius_is_f32, acc_unit, radius_scale, policy, p, R, mean, &zb, &D, &rmax);
    *D_out = D;
    *zb_out = zb;
    if (D <= 0) { free(p); free(R); return 2; }
    long total = (long)D * D * D;
    int32_t *vol = vol_out ? vol_out : (int32_t *)malloc(sizeof(int32_t) * total);
    memset(vol, 0, sizeof(int32_t) * total);
    if (brute == 2) orc_fast_for_full(p, R, n, D, vol, threads);
    else if (brute) orc_fast_for(p, R, n, D, vol, threads);
    else orc_scatter(p, R, n, D, vol);
    int idx[3];
    int32_t mx;
    orc_peak(vol, D, idx, &mx);
    orc_center_mm(idx, zb, mean, acc_unit, policy, center_mm);
    *peak_out = mx;
    if (votes_out) {
        long long s = 0;
        for (long q = 0; q < total; ++q) s += vol[q];
        *votes_out = s;
    }
    if (!vol_out) free(vol);
    free(p);
    free(R);
    return 0;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
