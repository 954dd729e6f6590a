// Micro-benchmark: issue rate of the packed fp32 instructions of sm_100 (FADD2 / FFMA2) against scalar FADD / FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_f32x2 tools/ubench_f32x2.cu ; prints JSON lines.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, float* out, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float c = seed * 1.0001f, d = seed * 0.5f;
  unsigned long long p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), pc = pk(c, c), pd = pk(d, d);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {        // 8 scalar FFMA
      a0 = fmaf(a0, c, d); a1 = fmaf(a1, c, d); a2 = fmaf(a2, c, d); a3 = fmaf(a3, c, d);
      a4 = fmaf(a4, c, d); a5 = fmaf(a5, c, d); a6 = fmaf(a6, c, d); a7 = fmaf(a7, c, d);
    } else if (MODE == 1) { // 4 FFMA2 (same flops)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pc), "l"(pd));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pc), "l"(pd));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pc), "l"(pd));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pc), "l"(pd));
    } else if (MODE == 2) { // 8 scalar FADD
      a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c);
      a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c);
    } else if (MODE == 3) { // 4 FADD2
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pc));
    } else {                // 4 FADD2 + 4 FSETP-class ALU ops (does the packed op leave issue slots for the ALU pipe?)
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pc));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pc));
      a0 = fmaxf(a0, c); a1 = fminf(a1, d); a2 = fmaxf(a2, c); a3 = fminf(a3, d);
    }
  }
  float lo, hi, s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0 ^ p1 ^ p2 ^ p3));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + lo + hi;
}
template <int MODE> void run(const char* name, int ops_per_iter, float* out) {
  const int iters = 20000, blocks = 148 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 512>>>(100, out, 1.0f);
  cudaEventRecord(e0); k<MODE><<<blocks, 512>>>(iters, out, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winst = (double)blocks * 16 * iters * ops_per_iter;
  printf("{\"bench\": \"%s\", \"ms\": %.3f, \"Gwarp_instr_per_s\": %.1f, \"warp_instr_per_clk_per_smsp\": %.3f}\n", name, ms, winst / ms / 1e6,
         winst / (ms * 1e-3) / (148.0 * 4 * 1.965e9));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  run<0>("ffma_scalar_x8", 8, out); run<1>("ffma2_x4", 4, out); run<2>("fadd_scalar_x8", 8, out); run<3>("fadd2_x4", 4, out);
  run<4>("fadd2_x4_plus_fmnmx_x4", 8, out);
  return 0;
}
