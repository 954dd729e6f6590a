// conv7head.cu -- K6: the tail of the radius-map producer as ONE tcgen05 kernel:
//   conv7 = Conv2d(64 -> 32, 3x3, padding 1) + BatchNorm2d(32) + ReLU      models/fcnresnet.py:114-116, :183-185
//   conv8 = Conv2d(32 -> 2, 1x1)                                            models/fcnresnet.py:118, :187-189
//   (fused mode) the evaluator's mask rule on seg / radius / depth          AccumulatorSpace.py:603-605, :837-840, :1049-1053
// so that neither the 32-channel activation (20 MB per 640x480 map) nor the seg plane ever reaches HBM.
//
// conv7 is an implicit GEMM: for a tile of 128 consecutive pixels of one image row,
//   D[128 px x 32] = sum over the 9 taps (dy, dx) of  A_tap[128 px x 64 ch] * W_tap[32 x 64]^T,
// bf16 operands, fp32 accumulation in tensor memory: 36 UMMAs of M = 128, N = 32, K = 16 per tile.
// The input is the NHWC (channels_last) bf16 output of up1: a pixel is 64 contiguous channels = 8 chunks of 16 bytes.  A tile's
// halo block (3 rows x 130 pixels) is staged ONCE in the canonical K-major no-swizzle UMMA layout with the 8-pixel groups packed:
//      byte offset of (row r, channel c, pixel p) = ((r * 8 + c / 8) * 130 + p) * 16 + (c % 8) * 2
// i.e. core matrices (8 pixels x 8 channels = 128 contiguous bytes) follow each other along the pixels (SBO = 128) and sit
// 130 * 16 bytes apart along the channels (LBO).  In this layout the operand of tap (dy, dx) is the SAME block read from a start
// address 16 * dx bytes further in row dy: the nine taps cost no extra copies, and the zero padding of the convolution is the
// zero fill of the out-of-image chunks (cp.async with src-size 0).  cp.async (16-byte chunks, all 128 threads, two stages) moves
// the block; TMA cannot produce this layout from NHWC without 16-byte boxes.
// Epilogue: each thread owns one pixel = one TMEM lane; it reads the 32 accumulators (tcgen05.ld 32x32b.x32), applies the folded
// BatchNorm (scale, shift) and ReLU, rounds to bf16 (the activation dtype of the unfused pipeline), takes the two conv8 dot
// products on the CUDA cores (64 FMAs) and hands seg / radial to the same epilogue as the 1x1 head (head1x1.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/rcvvote.h"

namespace rcv_c7 {

constexpr int kThreads = 128;
constexpr int kTileM = 128;          // pixels per tile = UMMA M
constexpr int kCin = 64, kCout = 32;
constexpr int kPx = kTileM + 2;      // pixels of a staged row (halo of one pixel each side)
constexpr int kStageBytes = 3 * (kCin / 8) * kPx * 16;   // 49,920
constexpr int kWTapBytes = kCout * kCin * 2;             // 4,096 per tap
constexpr int kStages = 2;
constexpr int kSmem = kStages * kStageBytes + 9 * kWTapBytes + (2 * kCout + 2 * kCout) * 4 + 64 + 128;

struct FusedArgs {            // the mask-rule epilogue (same meaning as rcv_head::FusedArgs); kp = the keypoint network of this launch
  const void* depth; int depth_dtype; int n_kpts; int kp;
  const double* max_radii; int max_radii_stride; int flags; float sem_threshold;
  unsigned* bits; int words_per_item; int* cnt;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_NONE, K-major: LBO = K direction, SBO = M/N direction
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16: D = f32, A = B = bf16, both K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kCout >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

// x [n_images][H][W][64] bf16; w7 [32][64][3][3], bn_scale / bn_shift [32], w8 [2][32], b8 [2] float32.
// !kFused: out [n_images][2][H*W] float32 (seg, radial).  kFused: out = radius planes [item][H*W], item = image * n_kpts + kp.
template <bool kFused>
__global__ void __launch_bounds__(kThreads) k_conv7_head(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w7,
                                                        const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                        const float* __restrict__ w8, const float* __restrict__ b8, float* __restrict__ out,
                                                        int H, int W, long long n_tiles, FusedArgs fz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kStageBytes;
  float* s_scale = reinterpret_cast<float*>(sB + 9 * kWTapBytes);
  float* s_shift = s_scale + kCout;
  float* s_w8 = s_shift + kCout;                                   // [2][32], rounded to bf16 like the 1x1 head's weights
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_w8 + 2 * kCout);  // MMA-complete barrier
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles_per_row = W / kTileM;
  const long long tiles_per_image = (long long)tiles_per_row * H;
  const long long HW = (long long)H * W;

  // weights: tap t = dy * 3 + dx, K-major no-swizzle core matrices: (n % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (n / 8) * 1024
  for (int e = tid; e < 9 * kCout * kCin; e += kThreads) {
    const int t = e / (kCout * kCin), n = (e / kCin) % kCout, k = e % kCin;
    const float v = w7[(n * kCin + k) * 9 + t];
    *reinterpret_cast<__nv_bfloat16*>(sB + t * kWTapBytes + (n / 8) * 1024 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = __float2bfloat16_rn(v);
  }
  if (tid < kCout) { s_scale[tid] = bn_scale[tid]; s_shift[tid] = bn_shift[tid]; }
  if (tid < 2 * kCout) s_w8[tid] = __bfloat162float(__float2bfloat16_rn(w8[tid]));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;

  // stage the halo block of a tile: chunk q = ((r * 130 + p) * 8 + kg): consecutive threads read consecutive 16-byte chunks of a pixel
  auto load = [&](long long tile, int stage) {
    const long long img = tile / tiles_per_image;
    const int rem = (int)(tile - img * tiles_per_image);
    const int y = rem / tiles_per_row, x0 = (rem - y * tiles_per_row) * kTileM;
    const uint32_t dst0 = smem_u32(sA + stage * kStageBytes);
    const __nv_bfloat16* xi = x + img * HW * kCin;
    for (int q = tid; q < 3 * kPx * 8; q += kThreads) {
      const int kg = q & 7, rp = q >> 3, r = rp / kPx, p = rp - r * kPx;
      const int yy = y + r - 1, xx = x0 + p - 1;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const __nv_bfloat16* src = xi + ((long long)(in ? yy : 0) * W + (in ? xx : 0)) * kCin + kg * 8;
      const uint32_t dst = dst0 + (uint32_t)(((r * 8 + kg) * kPx + p) * 16);
      const int nbytes = in ? 16 : 0;                               // out of the image: zero fill = the convolution's padding
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const long long step = gridDim.x;
  long long tile = blockIdx.x;
  if (tile < n_tiles) load(tile, 0);
  uint32_t phase = 0;
  for (int it = 0; tile < n_tiles; tile += step, ++it) {
    const int stage = it & 1;
    if (tile + step < n_tiles) {
      load(tile + step, stage ^ 1);                                 // that stage was released by the barrier that ended the last iteration
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async wrote through the generic proxy; the UMMA reads through the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA + stage * kStageBytes), b0 = smem_u32(sB);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t - dy * 3;
#pragma unroll
        for (int ks = 0; ks < kCin / 16; ++ks) {
          const uint64_t da = make_desc(a0 + (uint32_t)(((dy * 8 + 2 * ks) * kPx + dx) * 16), kPx * 16, 128);
          const uint64_t db = make_desc(b0 + t * kWTapBytes + ks * 256, 128, 1024);
          const uint32_t accumulate = (t | ks) != 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    __syncwarp();
    mbar_wait(smem_u32(mbar), phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[kCout];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // BatchNorm (folded) + ReLU + bf16 rounding, then conv8 on the CUDA cores
    float sv = b8[0], rad = b8[1];
#pragma unroll
    for (int ch = 0; ch < kCout; ++ch) {
      float a = fmaf(__uint_as_float(r[ch]), s_scale[ch], s_shift[ch]);
      a = __bfloat162float(__float2bfloat16_rn(fmaxf(a, 0.f)));
      sv = fmaf(s_w8[ch], a, sv);
      rad = fmaf(s_w8[kCout + ch], a, rad);
    }
    const long long img = tile / tiles_per_image;
    const long long pix = (tile - img * tiles_per_image) * kTileM + tid;      // = y * W + x0 + tid (W is a multiple of 128)
    if constexpr (!kFused) {
      out[(img * 2) * HW + pix] = sv;
      out[(img * 2 + 1) * HW + pix] = rad;
    } else {
      const long long item = img * fz.n_kpts + fz.kp;
      out[item * HW + pix] = rad;
      const long long di = img * HW + pix;
      const double z = fz.depth_dtype == RCV_U16 ? (double)((const unsigned short*)fz.depth)[di]
                       : fz.depth_dtype == RCV_F32 ? (double)((const float*)fz.depth)[di] : ((const double*)fz.depth)[di];
      bool ok = z != 0.0;
      if (fz.flags & RCV_MASK_MAX_RADIUS) ok = ok && ((double)rad <= fz.max_radii[img * fz.max_radii_stride + fz.kp]);
      if (fz.flags & RCV_MASK_RADIUS_NONZERO) ok = ok && (rad != 0.f);
      if (fz.flags & RCV_MASK_RADIUS_POSITIVE) ok = ok && (rad > 0.f);
      if (fz.flags & RCV_MASK_SEM_GT) ok = ok && (sv > fz.sem_threshold);
      if (fz.flags & RCV_MASK_SEM_GE) ok = ok && (sv >= fz.sem_threshold);
      const unsigned word = __ballot_sync(0xffffffffu, ok);      // a warp owns 32 consecutive pixels: one word of the bit mask
      if ((tid & 31) == 0) {
        const long long wi = (pix - (tid & 31)) >> 5;
        if (wi < fz.words_per_item) fz.bits[item * (long long)fz.words_per_item + wi] = word;
        if (word) atomicAdd(fz.cnt + item, __popc(word));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // TMEM and the stage buffer are free again
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

template <bool kFused>
static int launch(const void* x, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8, const float* b8, float* out,
                  int n_images, int H, int W, int sms, void* stream, FusedArgs fz) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv7_head<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long n_tiles = (long long)n_images * H * (W / kTileM);
  long long grid = sms;
  if (grid > n_tiles) grid = n_tiles;
  k_conv7_head<kFused><<<(int)grid, kThreads, kSmem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w7, bn_scale, bn_shift, w8, b8, out, H, W, n_tiles, fz);
  return (int)cudaGetLastError();
}

}  // namespace rcv_c7

// Called by the C ABI (rcvvote.cu).  Return a cudaError_t as int.  W must be a multiple of 128.
extern "C" int rcv_conv7_head_launch(const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                                     const float* b8, float* out, int n_images, int H, int W, int sms, void* stream) {
  return rcv_c7::launch<false>(x_nhwc_bf16, w7, bn_scale, bn_shift, w8, b8, out, n_images, H, W, sms, stream, rcv_c7::FusedArgs{});
}

// One keypoint network of the fused tail: x [n_frames][H][W][64], w7 / bn / w8 / b8 of network kp (already offset); writes radius_out [item][H*W] (item = frame * n_kpts + kp), bits and cnt like rcv_head1x1_fused_launch.
extern "C" int rcv_conv7_head_fused_launch(const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                                           const float* b8, float* radius_out, int n_frames, int H, int W, int sms, const void* depth,
                                           int depth_dtype, int n_kpts, int kp, const double* max_radii, int max_radii_stride, int flags,
                                           float sem_threshold, unsigned* bits, int words_per_item, int* cnt, void* stream) {
  rcv_c7::FusedArgs fz{depth, depth_dtype, n_kpts, kp, max_radii, max_radii_stride, flags, sem_threshold, bits, words_per_item, cnt};
  return rcv_c7::launch<true>(x_nhwc_bf16, w7, bn_scale, bn_shift, w8, b8, radius_out, n_frames, H, W, sms, stream, fz);
}
