// Micro-benchmark 2: which shared-memory atomic forms merge lanes that hit the SAME address?
// (The run-length vote kernel issues +1 and -1 at run ends; adjacent points often share an end.)
// ops: 0 red.add imm 1 (ATOMS.POPC.INC), 1 red.add reg -1 (ATOMS.ADD), 2 red.add reg +1, 3 red.inc 0xffffffff, 4 red.dec 0xffffffff,
//      5 red.add imm -1
// patterns: 0 conflict-free distinct, 1 groups of 2 lanes share an address (16 distinct, distinct banks), 2 groups of 4, 3 all same,
//           4 groups of 3 consecutive lanes share + rows of odd stride (vote-like), 5 random
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int SMEM_WORDS = 32768;

template <int OP>
__device__ __forceinline__ void op(unsigned addr, unsigned m1, unsigned p1) {
  if (OP == 0) asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
  if (OP == 1) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(m1) : "memory");
  if (OP == 2) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(p1) : "memory");
  if (OP == 3) asm volatile("red.shared.inc.u32 [%0], %1;" :: "r"(addr), "r"(m1) : "memory");
  if (OP == 4) asm volatile("red.shared.dec.u32 [%0], %1;" :: "r"(addr), "r"(m1) : "memory");
  if (OP == 5) asm volatile("red.shared.add.u32 [%0], -1;" :: "r"(addr) : "memory");
}

template <int OP, int PAT>
__global__ void __launch_bounds__(1024, 1) k(int iters, unsigned* out, unsigned m1, unsigned p1) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(s);
  unsigned base = warp * 997u;
  unsigned rnd = threadIdx.x * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a;
      if (PAT == 0) a = base + ((lane + u) & 31) + 32u * u;
      if (PAT == 1) a = base + (lane >> 1) + 32u * u;
      if (PAT == 2) a = base + (lane >> 2) + 32u * u;
      if (PAT == 3) a = base + u;
      if (PAT == 4) a = base + (lane / 3) * 121u + ((lane / 3) * (lane / 3) >> 3) + u;
      if (PAT == 5) { rnd = rnd * 1664525u + 1013904223u; a = rnd >> 12; }
      a &= (SMEM_WORDS - 1);
      op<OP>(sb + 4u * a, m1, p1);
      base += 1031u;
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeef) out[blockIdx.x] = acc;
}

template <int OP, int PAT>
void run(const char* on, const char* pn, int sms, unsigned* out) {
  CK(cudaFuncSetAttribute(k<OP, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WORDS * 4));
  const int iters = 2000;
  k<OP, PAT><<<sms, 1024, SMEM_WORDS * 4>>>(10, out, 0xffffffffu, 1u);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP, PAT><<<sms, 1024, SMEM_WORDS * 4>>>(iters, out, 0xffffffffu, 1u);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)sms * 32 * iters * 8;
  printf("{\"op\": \"%s\", \"pattern\": \"%s\", \"ms\": %.4f, \"Gwarp_instr_per_s\": %.2f, \"Glanes_per_s\": %.1f}\n", on, pn, ms, instr / ms / 1e6, instr * 32 / ms / 1e6);
}

#define ROW(OP, ON) run<OP, 0>(ON, "distinct", sms, out); run<OP, 1>(ON, "pairs_same", sms, out); run<OP, 2>(ON, "quads_same", sms, out); \
  run<OP, 3>(ON, "all_same", sms, out); run<OP, 4>(ON, "vote_like_triples", sms, out); run<OP, 5>(ON, "random", sms, out);
int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  unsigned* out; CK(cudaMalloc(&out, 4096));
  ROW(0, "add_imm_1(POPC.INC)") ROW(1, "add_reg_-1") ROW(2, "add_reg_+1") ROW(3, "inc_wrap") ROW(4, "dec_wrap") ROW(5, "add_imm_-1")
  return 0;
}
