// Checks that the packed fp32 instructions of sm_100 (add/sub/mul/fma .rn.f32x2) round exactly like the scalar ones.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(unsigned long long p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ float rnd(unsigned& s, int mode) {
  const unsigned r = rng(s);
  if (mode == 0) return __uint_as_float(r);                                  // any bit pattern
  if (mode == 1) return (float)(int)(r >> 8) * 5.9604645e-08f * 64.f - 32.f;   // [-32, 32)
  if (mode == 2) return 12582912.0f + (float)((int)(r >> 12) - 500000);       // near the magic constant
  return (float)((int)(r % 4001u) - 2000) * 0.25f;                           // quarter integers
}
__global__ void k(unsigned long long* bad, int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned long long nb[4] = {0, 0, 0, 0};
  for (int i = 0; i < iters; ++i) {
    const int m = i & 3;
    const float a0 = rnd(s, m), a1 = rnd(s, m), b0 = rnd(s, (m + 1) & 3), b1 = rnd(s, (m + 1) & 3), c0 = rnd(s, m), c1 = rnd(s, m);
    unsigned long long A = pk(a0, a1), B = pk(b0, b1), C = pk(c0, c1), R;
    float r0, r1;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B)); up(R, r0, r1);
    if (__float_as_uint(r0) != __float_as_uint(__fadd_rn(a0, b0)) || __float_as_uint(r1) != __float_as_uint(__fadd_rn(a1, b1))) ++nb[0];
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B)); up(R, r0, r1);
    if (__float_as_uint(r0) != __float_as_uint(__fsub_rn(a0, b0)) || __float_as_uint(r1) != __float_as_uint(__fsub_rn(a1, b1))) ++nb[1];
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B)); up(R, r0, r1);
    if (__float_as_uint(r0) != __float_as_uint(__fmul_rn(a0, b0)) || __float_as_uint(r1) != __float_as_uint(__fmul_rn(a1, b1))) ++nb[2];
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(A), "l"(B), "l"(C)); up(R, r0, r1);
    if (__float_as_uint(r0) != __float_as_uint(__fmaf_rn(a0, b0, c0)) || __float_as_uint(r1) != __float_as_uint(__fmaf_rn(a1, b1, c1))) ++nb[3];
  }
  for (int q = 0; q < 4; ++q) if (nb[q]) atomicAdd(&bad[q], nb[q]);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  k<<<148, 256>>>(d, 20000);
  unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("{\"check\": \"f32x2_vs_scalar\", \"trials_per_op\": %lld, \"mismatch_add\": %llu, \"mismatch_sub\": %llu, \"mismatch_mul\": %llu, \"mismatch_fma\": %llu}\n",
         148LL * 256 * 20000, h[0], h[1], h[2], h[3]);
  return 0;
}
