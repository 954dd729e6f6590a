// Micro-benchmark: shared-memory atomic (ATOMS/RED.shared) throughput on sm_100a, plus the
// MUFU.SQRT / F2I / packed f32x2 rates the vote rasteriser's instruction budget depends on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_atoms tools/ubench_atoms.cu
// The conflict-free full-warp figure is the denominator of roofline.frac for the vote kernel.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int SMEM_WORDS = 32768;   // 128 KB tile

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_atoms(int iters, unsigned* out) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = warp * 997u;
  unsigned rnd = threadIdx.x * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a;
      bool act = true;
      if (MODE == 0) a = base + ((lane + u) & 31) + 32u * u;                  // conflict-free, 32 lanes
      if (MODE == 1) { a = base + ((lane + u) & 31) + 32u * u; act = lane < 19; }   // 19 lanes active
      if (MODE == 2) { a = base + ((lane + u) & 31) + 32u * u; act = lane < 8; }    // 8 lanes active
      if (MODE == 3) { rnd = rnd * 1664525u + 1013904223u; a = rnd >> 12; }    // random word (bank conflicts)
      if (MODE == 4) a = base + (lane & 15) * 2 + 64u * u;                     // 2 lanes per address (same-address pairs)
      if (MODE == 5) a = base + (lane >> 1) * 32 + (lane & 1) + 7u * u;        // 16-way bank conflict, distinct addresses
      if (MODE == 6) a = base + lane * 87u + ((lane * lane) >> 5) + u;         // z-pass like: odd row stride + smooth k
      if (MODE == 7) a = base + u;                                            // all lanes same address
      if (MODE == 8) { a = base + ((lane + u) & 31) + 32u * u; }               // packed u16: add 1<<16 or 1
      a &= (SMEM_WORDS - 1);
      unsigned v = (MODE == 8) ? (1u << ((lane & 1) * 16)) : 1u;
      if (act) atomicAdd(&s[a], v);
      base += 1031u;
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeef) out[blockIdx.x] = acc;
}


// Predication variants: how should a rasteriser issue an atomic for only the lanes that hit?
//  0: C++ `if (hit) atomicAdd` (compiler emits BSSY/BRA/BSYNC around each ATOMS)
//  1: inline-PTX predicated `@p red.shared.add.u32`
//  2: unconditional atomicAdd(addr, hit ? 1 : 0)   (ATOMS.ADD of a register value)
//  3: unconditional atomicAdd(addr, v) with v==1 loaded from a register (ATOMS.ADD, all lanes)
// hit pattern: pseudo-random ~59% of lanes, different every instruction.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_pred(int iters, unsigned* out, unsigned one) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = warp * 997u;
  unsigned rnd = threadIdx.x * 2654435761u + 12345u;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a = (base + ((lane + u) & 31) + 32u * u) & (SMEM_WORDS - 1);
      rnd = rnd * 1664525u + 1013904223u;
      bool hit = (rnd >> 8) < (unsigned)(0.59 * 16777216.0);
      if (MODE == 0) { if (hit) atomicAdd(&s[a], 1u); }
      if (MODE == 1) {
        unsigned addr = sbase + a * 4u;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p red.shared.add.u32 [%0], %2; }" :: "r"(addr), "r"((unsigned)hit), "r"(one) : "memory");
      }
      if (MODE == 2) atomicAdd(&s[a], hit ? 1u : 0u);
      if (MODE == 3) atomicAdd(&s[a], one);
      base += 1031u;
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeef) out[blockIdx.x] = acc;
}

// plain LDS+IADD+STS (non-atomic) for comparison
__global__ void __launch_bounds__(1024, 1) k_ldssts(int iters, unsigned* out) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = warp * 1024u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned a = (base + ((lane + u) & 31) + 32u * u) & (SMEM_WORDS - 1);
      s[a] = s[a] + 1;
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int i = threadIdx.x; i < SMEM_WORDS; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeef) out[blockIdx.x] = acc;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_alu(int iters, float* out) {
  float x0 = threadIdx.x * 1.0001f + 1.0f, x1 = x0 + 0.5f, x2 = x0 + 0.25f, x3 = x0 + 0.125f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {   // sqrt.approx
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x2));
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x3));
        x0 += 3.0f; x1 += 3.0f; x2 += 3.0f; x3 += 3.0f;
      }
      if (MODE == 1) {   // F2I floor + I2F
        int i0, i1, i2, i3;
        asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(i0) : "f"(x0));
        asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(i1) : "f"(x1));
        asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(i2) : "f"(x2));
        asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(i3) : "f"(x3));
        x0 += __int_as_float(i0 & 0x3fffffff) * 1e-30f; x1 += __int_as_float(i1 & 0x3fffffff) * 1e-30f;
        x2 += __int_as_float(i2 & 0x3fffffff) * 1e-30f; x3 += __int_as_float(i3 & 0x3fffffff) * 1e-30f;
      }
      if (MODE == 2) {   // scalar FFMA chain x4
        x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.5f);
        x2 = fmaf(x2, 1.0001f, 0.5f); x3 = fmaf(x3, 1.0001f, 0.5f);
      }
      if (MODE == 3) {   // packed fma.rn.f32x2 x2 (4 flops-lanes per 2 instr)
        unsigned long long p0, p1, c, d;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(1.0001f));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(0.5f));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(c), "l"(d));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(c), "l"(d));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
      }
      if (MODE == 4) {   // fp64 DFMA chain x4
        double d0 = x0, d1 = x1; 
        d0 = fma(d0, 1.0001, 0.5); d1 = fma(d1, 1.0001, 0.5);
        d0 = fma(d0, 1.0001, 0.5); d1 = fma(d1, 1.0001, 0.5);
        x0 = (float)d0; x1 = (float)d1;
      }
    }
  }
  if (x0 + x1 + x2 + x3 == 123.456f) out[threadIdx.x] = x0;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  unsigned* out; CK(cudaMalloc(&out, 4096 * 4));
  const int iters = 4096;   // x8 unroll = 32768 atomics / thread
  const size_t smem = SMEM_WORDS * 4;
  const char* names[] = {"conflict_free_32lanes", "conflict_free_19lanes", "conflict_free_8lanes", "random_addr",
                         "same_addr_pairs", "bank_conflict_16way", "zpass_like_stride87", "all_same_addr", "packed_u16"};
#define RUN(M) { CK(cudaFuncSetAttribute(k_atoms<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int threads = 256; threads <= 1024; threads *= 2) { \
      float ms = time_ms([&] { k_atoms<M><<<sms, threads, smem>>>(iters, out); }, 5); \
      double lanes = (M == 1 ? 19.0 / 32 : M == 2 ? 8.0 / 32 : 1.0); \
      double ops = (double)sms * threads * iters * 8 * lanes; \
      double winstr = (double)sms * threads / 32 * iters * 8; \
      printf("{\"bench\": \"atoms\", \"mode\": \"%s\", \"threads\": %d, \"ms\": %.4f, \"Gatomic_lanes_per_s\": %.1f, \"Gwarp_instr_per_s\": %.2f}\n", \
             names[M], threads, ms, ops / ms * 1e-6, winstr / ms * 1e-6); } }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)

  const char* pnames[] = {"cxx_if_atomicAdd_59pct", "ptx_predicated_red_59pct", "uncond_add_hit_value_59pct", "uncond_ATOMS_ADD_reg_100pct"};
#define RUNP(M) { CK(cudaFuncSetAttribute(k_pred<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int threads = 512; threads <= 1024; threads *= 2) { \
      float ms = time_ms([&] { k_pred<M><<<sms, threads, smem>>>(iters, out, 1u); }, 5); \
      double winstr = (double)sms * threads / 32 * iters * 8; \
      printf("{\"bench\": \"pred\", \"mode\": \"%s\", \"threads\": %d, \"ms\": %.4f, \"Gwarp_instr_per_s\": %.2f, \"Gvotes_per_s\": %.1f}\n", \
             pnames[M], threads, ms, winstr / ms * 1e-6, winstr * 32 * (M == 3 ? 1.0 : 0.59) / ms * 1e-6); } }
  RUNP(0) RUNP(1) RUNP(2) RUNP(3)
  CK(cudaFuncSetAttribute(k_ldssts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { float ms = time_ms([&] { k_ldssts<<<sms, 1024, smem>>>(iters, out); }, 5);
    printf("{\"bench\": \"lds_add_sts\", \"threads\": 1024, \"ms\": %.4f, \"Grmw_lanes_per_s\": %.1f}\n", ms, (double)sms * 1024 * iters * 8 / ms * 1e-6); }
  const char* anames[] = {"sqrt_approx", "f2i_floor", "ffma", "ffma2_packed", "dfma"};
  double per_iter[] = {4, 4, 4, 4, 4};
#define RUNA(M) { float ms = time_ms([&] { k_alu<M><<<sms * 2, 1024>>>(iters, (float*)out); }, 5); \
    printf("{\"bench\": \"alu\", \"mode\": \"%s\", \"ms\": %.4f, \"Gops_lanes_per_s\": %.1f}\n", anames[M], ms, (double)sms * 2 * 1024 * iters * 8 * per_iter[M] / ms * 1e-6); }
  RUNA(0) RUNA(1) RUNA(2) RUNA(3) RUNA(4)
  return 0;
}
