// head1x1.cu -- K5: the producer's final 1x1 convolution (models/fcnresnet.py:118,187-189 of the reference:
// conv8 = Conv2d(32 -> 2, 1x1, bias); seg = out[:,0:1], radial = out[:,1:2]) as a tcgen05 tensor-core kernel.
//
//   out[b][n][p] = bias[n] + sum_k bf16(w[n][k]) * up[b][k][p]        n = 0,1   k = 0..31   p = 0..H*W-1
//
// GEMM view per tile: D[128 pixels x 16] = A[128 x 32] * B[16 x 32]^T, bf16 operands, fp32 accumulation in TMEM
// (N is padded from 2 to 16, the smallest N of an M = 128 UMMA).  The input is NCHW, so a tile of A is 32 channel
// rows of 128 contiguous pixels: exactly the canonical MN-major no-swizzle UMMA layout when each 16-byte chunk
// (8 pixels of one channel) is placed at  (pixel group) * 128 B + (channel % 8) * 16 B + (channel / 8) * 2048 B.
// Two forms of the copy exist.  The default (k_head1x1_tma, further down) lets TMA boxes land as the 128-byte-swizzled MN-major
// layout: 6.7-7.0 TB/s.  The first form (k_head1x1, RCV_HEAD_TMA=0) is described here:
// cp.async moves the chunks global -> shared without touching registers (2-stage ring of 4 tiles = 32 KB per stage, 3 CTAs per SM); one thread
// issues the two K = 16 MMAs of a tile, tcgen05.commit signals an mbarrier, and the four warps read their 32 TMEM
// lanes back with tcgen05.ld, add the bias and store both planes with fully coalesced 128-pixel rows.
// The kernel is HBM-bound: 64 B read + 8 B written per pixel, 39 MFLOP per 640x480 map.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rcvvote.h"

namespace rcv_head {

constexpr int kThreads = 128;
constexpr int kTileM = 128;                 // pixels per UMMA tile = UMMA M
constexpr int kC = 32;                      // input channels = 2 x UMMA K
constexpr int kN = 16;                      // UMMA N (2 real outputs)
constexpr int kATile = kC * kTileM * 2;     // 8192 bytes
constexpr int kBTile = kN * kC * 2;         // 1024 bytes
constexpr int kMaxKp = 8;                   // fused mode: weight tiles of up to 8 keypoint networks live in shared memory

// Fused mode: "image" b is the item (frame, keypoint) = (b / n_kpts, b % n_kpts); the epilogue applies the
// evaluator's mask rule (AccumulatorSpace.py:603-605, :837-840, :1049-1053) to the seg / radius values it has just read from
// tensor memory, writes one survival bit per pixel + the item's count (the outputs of k_frame_mask, rcvvote.cu) and the
// radius plane; the seg plane never reaches HBM and nothing reads the maps back.
struct FusedArgs {
  const void* depth; int depth_dtype; int n_kpts;
  const double* max_radii; int max_radii_stride; int flags; float sem_threshold;
  unsigned* bits; int words_per_item; int* cnt;
};

template <int kStages, int kTpi, bool kFused = false> constexpr int smem_bytes() { return kStages * kTpi * kATile + (kFused ? kMaxKp : 1) * kBTile + 64; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  // UMMA shared-memory descriptor, SWIZZLE_NONE: start address, leading / stride byte offsets (all >> 4), version 1
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// instruction descriptor, kind::f16: D = f32, A = B = bf16, A MN-major, B K-major, N = 16, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// Epilogue of one group: r[tl] = the thread's pixel of tile tl, accumulators of the seg and the radius output.
template <int kTpi, bool kFused>
__device__ __forceinline__ void head_epilogue(const uint32_t (&r)[kTpi][2], long long grp, int groups_per_image, long long HW, int tid,
                                            const float* __restrict__ bias, float* __restrict__ out, const FusedArgs& fz) {
  const long long b = grp / groups_per_image;
  const long long pix0 = (grp - b * groups_per_image) * (kTpi * kTileM) + tid;
  if constexpr (!kFused) {
    const float bias0 = bias[0], bias1 = bias[1];
    float* o = out + b * 2 * HW;
#pragma unroll
    for (int tl = 0; tl < kTpi; ++tl) {
      const long long pix = pix0 + tl * kTileM;
      if (pix < HW) {
        o[pix] = __uint_as_float(r[tl][0]) + bias0;
        o[HW + pix] = __uint_as_float(r[tl][1]) + bias1;
      }
    }
  } else {
    const int kp = (int)(b % fz.n_kpts);
    const long long frame = b / fz.n_kpts;
    const float bias0 = bias[2 * kp], bias1 = bias[2 * kp + 1];
    const double max_r = fz.max_radii ? fz.max_radii[frame * fz.max_radii_stride + kp] : 0.0;
    float* o = out + b * HW;                      // radius plane only
    unsigned* bits = fz.bits + b * (long long)fz.words_per_item;
    int n = 0;
#pragma unroll
    for (int tl = 0; tl < kTpi; ++tl) {
      const long long pix = pix0 + tl * kTileM;
      bool ok = false;
      if (pix < HW) {
        const float sv = __uint_as_float(r[tl][0]) + bias0, rad = __uint_as_float(r[tl][1]) + bias1;
        o[pix] = rad;
        const long long di = frame * HW + pix;
        const double z = fz.depth_dtype == RCV_U16 ? (double)((const unsigned short*)fz.depth)[di]
                         : fz.depth_dtype == RCV_F32 ? (double)((const float*)fz.depth)[di] : ((const double*)fz.depth)[di];
        ok = z != 0.0;
        if (fz.flags & RCV_MASK_MAX_RADIUS) ok = ok && ((double)rad <= max_r);
        if (fz.flags & RCV_MASK_RADIUS_NONZERO) ok = ok && (rad != 0.f);
        if (fz.flags & RCV_MASK_RADIUS_POSITIVE) ok = ok && (rad > 0.f);
        if (fz.flags & RCV_MASK_SEM_GT) ok = ok && (sv > fz.sem_threshold);
        if (fz.flags & RCV_MASK_SEM_GE) ok = ok && (sv >= fz.sem_threshold);
      }
      const unsigned word = __ballot_sync(0xffffffffu, ok);    // a warp owns 32 consecutive pixels: one word of the bit mask
      const long long wi = (pix - (tid & 31)) >> 5;
      if ((tid & 31) == 0 && wi < fz.words_per_item) bits[wi] = word;
      n += __popc(word);
    }
    if ((tid & 31) == 0 && n) atomicAdd(fz.cnt + b, n);
  }
}

// One iteration of a CTA handles a group of kTpi consecutive 128-pixel tiles of one image (one ring stage), so the
// serial chain  copy-wait -> barrier -> MMA -> commit -> mbarrier -> tcgen05.ld -> store  is paid once per
// kTpi * 8 KB of input.
template <int kStages, int kTpi, bool kFused>
__global__ void __launch_bounds__(kThreads) k_head1x1(const __nv_bfloat16* __restrict__ up, const float* __restrict__ w, const float* __restrict__ bias,
                                                     float* __restrict__ out, long long HW, int groups_per_image, long long n_groups, FusedArgs fz) {
  constexpr int kTmemCols = kTpi * kN < 32 ? 32 : kTpi * kN;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTpi * kATile;
  constexpr int kWTiles = kFused ? kMaxKp : 1;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sB + kWTiles * kBTile);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(sB + kWTiles * kBTile + 16);
  const int tid = threadIdx.x, warp = tid >> 5;

  // weights -> canonical K-major no-swizzle tile: (n % 8) * 16 + (n / 8) * 512 + (k % 8) * 2 + (k / 8) * 128
  const int n_wtiles = kFused ? fz.n_kpts : 1;   // fused: one weight tile per keypoint network
  for (int e = tid; e < n_wtiles * kN * kC; e += kThreads) {
    const int t = e / (kN * kC), n = (e / kC) % kN, k = e % kC;
    const float v = n < 2 ? w[t * 2 * kC + n * kC + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(sB + t * kBTile + (n % 8) * 16 + (n / 8) * 512 + (k % 8) * 2 + (k / 8) * 128) = __float2bfloat16_rn(v);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tile was written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;

  const long long step = gridDim.x;
  auto issue = [&](long long grp, int stage) {
    if (grp < n_groups) {
      const long long b = grp / groups_per_image;
      const long long m0 = (grp - b * groups_per_image) * (kTpi * kTileM);
      const __nv_bfloat16* src0 = up + b * kC * HW;
      const uint32_t dst0 = smem_u32(sA + stage * kTpi * kATile);
#pragma unroll
      for (int j = 0; j < kTpi * (kC * kTileM / 8) / kThreads; ++j) {
        // chunk = 8 pixels (16 bytes) of one channel.  A quarter-warp (8 lanes, one shared-memory wavefront) takes the 8
        // channels of one pixel group = 128 contiguous bytes of the canonical layout, so the shared-memory side of the copy
        // is conflict-free; the four quarters take 4 consecutive pixel groups (64 contiguous bytes per channel in global
        // memory).  The first version gave consecutive lanes consecutive pixel groups of ONE channel: 128-byte stride in
        // shared memory, every lane on the same four banks, 8x the wavefronts -- 16 B/clk/SM, the 3.87 TB/s plateau.
        const int chunk = tid + j * kThreads, lane = chunk & 31, wi = chunk >> 5;
        const int k = (wi / (4 * kTpi)) * 8 + (lane & 7), mg = (wi % (4 * kTpi)) * 4 + (lane >> 3);
        const long long pix = m0 + mg * 8;
        const bool valid = pix < HW;
        const __nv_bfloat16* src = src0 + (long long)k * HW + (valid ? pix : 0);
        const uint32_t dst = dst0 + (mg >> 4) * kATile + (mg & 15) * 128 + (k & 7) * 16 + (k >> 3) * 2048;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const long long grp0 = blockIdx.x;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(grp0 + s * step, s);
  uint32_t phase = 0;
  int it = 0;
  for (long long grp = grp0; grp < n_groups; grp += step, ++it) {
    issue(grp + (kStages - 1) * step, (it + kStages - 1) % kStages);
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA + (it % kStages) * kTpi * kATile);
      const uint32_t b0 = smem_u32(sB) + (kFused ? (uint32_t)((grp / groups_per_image) % fz.n_kpts) * kBTile : 0u);
#pragma unroll
      for (int tl = 0; tl < kTpi; ++tl)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = make_desc(a0 + tl * kATile + kb * 4096, 2048, 128);   // MN-major: 8-channel groups 2048 B apart, 8-pixel groups 128 B apart
          const uint64_t db = make_desc(b0 + kb * 256, 128, 512);                   // K-major:  8-element K chunks 128 B apart, 8-row groups 512 B apart
          const uint32_t accumulate = kb;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem + tl * kN), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
    } while (!done);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[kTpi][2];
#pragma unroll
    for (int tl = 0; tl < kTpi; ++tl)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[tl][0]), "=r"(r[tl][1])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + tl * kN) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    head_epilogue<kTpi, kFused>(r, grp, groups_per_image, HW, tid, bias, out, fz);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // TMEM and the stage buffer are free again
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}


// ---- TMA form of the copy (RCV_HEAD_TMA=1) ------------------------------------------------------------------------------
// cp.async fetches a whole 32-byte sector for every 16-byte lane (twice the L2 -> SM traffic).  Here one
// thread issues cp.async.bulk.tensor.2d boxes of 64 pixels x 32 channels (SWIZZLE_128B) that land as the canonical MN-major
// 128-byte-swizzled UMMA layout: a row = 64 pixels of one channel = 128 contiguous bytes in global AND in shared memory, rows
// 128 bytes apart, 8-row groups 1024 bytes apart (SBO), the second 64-pixel block of a tile 4096 bytes further (LBO); the
// bytes arrive on an mbarrier (expect_tx), no thread touches them.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) |
         (2ull << 61);   // layout type SWIZZLE_128B
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

template <int kStages, int kTpi, bool kFused>
__global__ void __launch_bounds__(kThreads) k_head1x1_tma(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, long long HW,
                                                         int groups_per_image, long long n_groups, FusedArgs fz) {
  constexpr int kTmemCols = kTpi * kN < 32 ? 32 : kTpi * kN;
  constexpr int kWTiles = kFused ? kMaxKp : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTpi * kATile;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kWTiles * kBTile);             // [kStages] bytes-landed barriers
  uint64_t* mbar = full + kStages;                                                 // MMA-complete barrier
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  const int n_wtiles = kFused ? fz.n_kpts : 1;
  for (int e = tid; e < n_wtiles * kN * kC; e += kThreads) {
    const int t = e / (kN * kC), n = (e / kC) % kN, k = e % kC;
    const float v = n < 2 ? w[t * 2 * kC + n * kC + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(sB + t * kBTile + (n % 8) * 16 + (n / 8) * 512 + (k % 8) * 2 + (k / 8) * 128) = __float2bfloat16_rn(v);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full + s)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;

  const long long step = gridDim.x;
  auto issue = [&](long long grp, int stage) {      // thread 0 only
    if (grp >= n_groups) return;
    const long long b = grp / groups_per_image;
    const int m0 = (int)((grp - b * groups_per_image) * (kTpi * kTileM));
    const uint32_t bar = smem_u32(full + stage);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTpi * kATile) : "memory");
#pragma unroll
    for (int h = 0; h < 2 * kTpi; ++h) {            // boxes of 64 pixels x 32 channels = 4096 bytes; pixels beyond HW arrive as zeros
      const uint32_t dst = smem_u32(sA + stage * kTpi * kATile + h * 4096);
      const int c0 = m0 + h * 64, c1 = (int)(b * kC);
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(bar) : "memory");
    }
  };

  const long long grp0 = blockIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) issue(grp0 + s * step, s);
  }
  uint32_t phase = 0;
  int it = 0;
  for (long long grp = grp0; grp < n_groups; grp += step, ++it) {
    if (tid == 0) {
      issue(grp + (kStages - 1) * step, (it + kStages - 1) % kStages);   // that stage was released by the barrier that ended the last iteration
      mbar_wait(smem_u32(full + it % kStages), (uint32_t)((it / kStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA + (it % kStages) * kTpi * kATile);
      const uint32_t b0 = smem_u32(sB) + (kFused ? (uint32_t)((grp / groups_per_image) % fz.n_kpts) * kBTile : 0u);
#pragma unroll
      for (int tl = 0; tl < kTpi; ++tl)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = make_desc_sw128(a0 + tl * kATile + kb * 2048, 4096, 1024);   // 16 channel rows per K step
          const uint64_t db = make_desc(b0 + kb * 256, 128, 512);
          const uint32_t accumulate = kb;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem + tl * kN), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    __syncwarp();
    mbar_wait(smem_u32(mbar), phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[kTpi][2];
#pragma unroll
    for (int tl = 0; tl < kTpi; ++tl)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[tl][0]), "=r"(r[tl][1])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + tl * kN) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    head_epilogue<kTpi, kFused>(r, grp, groups_per_image, HW, tid, bias, out, fz);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // TMEM and the stage buffer are free again
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace rcv_head

template <int kStages, int kTpi, bool kFused = false>
static int launch(const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw, int sms, int ctas_per_sm,
                  void* stream, rcv_head::FusedArgs fz = rcv_head::FusedArgs{}) {
  using namespace rcv_head;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_head1x1<kStages, kTpi, kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes<kStages, kTpi, kFused>());
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int groups_per_image = (int)((hw + kTpi * kTileM - 1) / (kTpi * kTileM));
  const long long n_groups = (long long)n_images * groups_per_image;
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > n_groups) grid = n_groups;
  k_head1x1<kStages, kTpi, kFused><<<(int)grid, kThreads, smem_bytes<kStages, kTpi, kFused>(), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)up_bf16, weight, bias, out, hw, groups_per_image, n_groups, fz);
  return (int)cudaGetLastError();
}


// Tensor map of the bf16 input seen as a 2-D array [n_images * 32 channel rows][hw pixels]; box = 64 pixels x 32 rows, SWIZZLE_128B.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_head_tensor_map(CUtensorMap* tm, const void* up_bf16, int n_images, long long hw) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return (int)cudaErrorNotSupported;
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)n_images * rcv_head::kC};
  const cuuint64_t strides[1] = {(cuuint64_t)hw * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)rcv_head::kC};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(up_bf16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int kStages, int kTpi, bool kFused = false>
static int launch_tma(const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw, int sms, int ctas_per_sm,
                      void* stream, rcv_head::FusedArgs fz = rcv_head::FusedArgs{}) {
  using namespace rcv_head;
  constexpr int kSmem = smem_bytes<kStages, kTpi, kFused>() + 1024 + 64;   // + alignment slack + the per-stage barriers
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_head1x1_tma<kStages, kTpi, kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  CUtensorMap tm;
  int rc = make_head_tensor_map(&tm, up_bf16, n_images, hw);
  if (rc) return rc;
  const int groups_per_image = (int)((hw + kTpi * kTileM - 1) / (kTpi * kTileM));
  const long long n_groups = (long long)n_images * groups_per_image;
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > n_groups) grid = n_groups;
  k_head1x1_tma<kStages, kTpi, kFused><<<(int)grid, kThreads, kSmem, (cudaStream_t)stream>>>(tm, weight, bias, out, hw, groups_per_image, n_groups, fz);
  return (int)cudaGetLastError();
}

static bool head_use_tma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RCV_HEAD_TMA"); v = (e && !atoi(e)) ? 0 : 1; }   // default: the TMA kernel; RCV_HEAD_TMA=0 selects the cp.async one
  return v == 1;
}

// Called by the C ABI (rcvvote.cu: rcv_head_1x1).  Returns a cudaError_t as int.
extern "C" int rcv_head1x1_launch(const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw, int sms,
                                  void* stream) {
  // tuning knobs (experiments only): ring stages x tiles per iteration, and resident CTAs per SM
  static int cfg = 0, ctas = 0;
  if (!cfg) {
    const char* e1 = getenv("RCV_HEAD_CFG"); const char* e2 = getenv("RCV_HEAD_CTAS");
    // measured best on B200 (profiles/r01_head1x1_sweep.txt): TMA 2 stages x 4 tiles, 2 CTAs/SM: 6.97 TB/s; cp.async 2 x 4, 3 CTAs/SM: 4.3 TB/s
    cfg = e1 ? atoi(e1) : 24; ctas = e2 ? atoi(e2) : (head_use_tma() ? 2 : 3);
  }
  if (head_use_tma()) {
    switch (cfg) {
      case 22: return launch_tma<2, 2>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
      case 34: return launch_tma<3, 4>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
      case 44: return launch_tma<4, 4>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
      default: return launch_tma<2, 4>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    }
  }
  switch (cfg) {   // cfg = 10 * stages + tiles per iteration
    case 31: return launch<3, 1>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    case 21: return launch<2, 1>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    case 22: return launch<2, 2>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    case 32: return launch<3, 2>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    case 34: return launch<3, 4>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    case 28: return launch<2, 8>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
    default: return launch<2, 4>(up_bf16, weight, bias, out, n_images, hw, sms, ctas, stream);
  }
}

// Fused head + mask rule (rcvvote.cu: rcv_head_vote_frames).  up [n_items][32][hw], weight [n_kpts][2][32], bias [n_kpts][2];
// writes radius_out [n_items][hw], bits [n_items][words_per_item] and adds the survivors to cnt [n_items] (zeroed by the caller).
extern "C" int rcv_head1x1_fused_launch(const void* up_bf16, const float* weight, const float* bias, float* radius_out, int n_items, long long hw,
                                        int sms, const void* depth, int depth_dtype, int n_kpts, const double* max_radii, int max_radii_stride,
                                        int flags, float sem_threshold, unsigned* bits, int words_per_item, int* cnt, void* stream) {
  if (n_kpts < 1 || n_kpts > rcv_head::kMaxKp) return (int)cudaErrorInvalidValue;
  rcv_head::FusedArgs fz{depth, depth_dtype, n_kpts, max_radii, max_radii_stride, flags, sem_threshold, bits, words_per_item, cnt};
  if (head_use_tma()) return launch_tma<2, 4, true>(up_bf16, weight, bias, radius_out, n_items, hw, sms, 2, stream, fz);
  return launch<2, 4, true>(up_bf16, weight, bias, radius_out, n_items, hw, sms, 3, stream, fz);   // 512-pixel groups: whole 256-pixel mask steps
}
