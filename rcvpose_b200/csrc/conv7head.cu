// conv7head.cu -- K6: the tail of the radius-map producer as ONE tcgen05 kernel:
//   conv7 = Conv2d(64 -> 32, 3x3, padding 1) + BatchNorm2d(32) + ReLU      models/fcnresnet.py:114-116, :183-185
//   conv8 = Conv2d(32 -> 2, 1x1)                                            models/fcnresnet.py:118, :187-189
//   (fused mode) the evaluator's mask rule on seg / radius / depth          AccumulatorSpace.py:603-605, :837-840, :1049-1053
// so that neither the 32-channel activation (20 MB per 640x480 map) nor the seg plane ever reaches HBM.
//
// conv7 is an implicit GEMM: for a tile of 128 consecutive pixels of one image row,
//   D[128 px x 32] = sum over the 9 taps (dy, dx) of  A_tap[128 px x 64 ch] * W_tap[32 x 64]^T,
// bf16 operands, fp32 accumulation in tensor memory: 36 UMMAs of M = 128, N = 32, K = 16 per tile.
// The input is the NHWC (channels_last) bf16 output of up1: a pixel is 64 contiguous channels = one 128-byte row of a K-major
// operand.  A tile's halo block (3 image rows x 130 pixels) is staged ONCE by three TMA boxes (cp.async.bulk.tensor.4d over
// [N][H][W][64], SWIZZLE_128B, out-of-image pixels arrive as zeros = the convolution's padding), each into a slab of 136 rows of
// 128 bytes (17 swizzle atoms of 8 rows).  The operand of tap (dy, dx) is the SAME data read from a start address dx rows
// further in slab dy: the nine taps cost no extra copies.  The 128-byte swizzle is a function of the absolute shared-memory
// address for TMA and tensor core alike, so a start address that is not 1024-byte aligned needs nothing else: the descriptor's
// base-offset field stays 0 (measured: with (address >> 7) & 7 there the results are wrong, RCV_C7_BASE_OFFSET=1).
// (A first version staged the block with 16-byte cp.async chunks into a no-swizzle layout: 25 dependent copies per thread,
// bound by the copies a warp can keep in flight -- 48 us per 640x480 image against 22 with TMA.)
// Epilogue: each thread owns one pixel = one TMEM lane; it reads the 32 accumulators (tcgen05.ld 32x32b.x32), applies the folded
// BatchNorm (scale, shift) and ReLU, rounds to bf16 (the activation dtype of the unfused pipeline), takes the two conv8 dot
// products on the CUDA cores (64 FMAs) and hands seg / radial to the same epilogue as the 1x1 head (head1x1.cu).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/rcvvote.h"

namespace rcv_c7 {

constexpr int kThreads = 128;
constexpr int kTileM = 128;          // pixels per tile = UMMA M
constexpr int kCin = 64, kCout = 32;
constexpr int kPx = kTileM + 2;      // pixels of a staged row (halo of one pixel each side)
constexpr int kSlabRows = 136;       // rows of 128 bytes per staged image row: a multiple of the 8-row swizzle atom
constexpr int kSlabBytes = kSlabRows * 128;              // 17,408
constexpr int kStageBytes = 3 * kSlabBytes;              // 52,224
constexpr int kBoxBytes = kPx * kCin * 2;                // 16,640 per TMA box
constexpr int kWTapBytes = kCout * kCin * 2;             // 4,096 per tap
// One stage per CTA, two CTAs per SM (2 x 91 KB): while one CTA's single thread feeds the tensor pipe (36 small UMMAs: an M = 128,
// N = 32 UMMA re-reads 4 KB of A from shared memory for 131 kFLOP, so the pipe is bound by operand reads) or runs its epilogue,
// the other one's halo block streams in.
constexpr int kSmem = kStageBytes + 9 * kWTapBytes + 4 * kCout * 4 + 64 + 1024;

struct FusedArgs {            // the mask-rule epilogue (same meaning as rcv_head::FusedArgs); kp = the keypoint network of this launch
  const void* depth; int depth_dtype; int n_kpts; int kp;
  const double* max_radii; int max_radii_stride; int flags; float sem_threshold;
  unsigned* bits; int words_per_item; int* cnt;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // SWIZZLE_NONE, K-major: LBO = K direction, SBO = M/N direction
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// SWIZZLE_128B, K-major (rows of 128 bytes, 8-row atoms 1024 bytes apart); base offset = phase of the swizzle pattern at the start address
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_offset) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7u) << 49) | (2ull << 61);
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

// x [n_images][H][W][64] bf16 (through the tensor map); w7 [32][64][3][3], bn_scale / bn_shift [32], w8 [2][32], b8 [2] float32.
// !kFused: out [n_images][2][H*W] float32 (seg, radial).  kFused: out = radius planes [item][H*W], item = image * n_kpts + kp.
// bo_mode: 0 = base offset 0 (correct, the default); 1 = descriptors carry (address >> 7) & 7 (experiment, wrong results).
template <bool kFused>
__global__ void __launch_bounds__(kThreads, 2) k_conv7_head(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w7,
                                                           const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                           const float* __restrict__ w8, const float* __restrict__ b8, float* __restrict__ out,
                                                           int H, int W, long long n_tiles, int bo_mode, FusedArgs fz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStageBytes;
  float4* s_par = reinterpret_cast<float4*>(sB + 9 * kWTapBytes);   // per channel: BN scale, BN shift, conv8 weights (rounded to bf16 like the 1x1 head's)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_par + kCout);      // bytes-landed barrier
  uint64_t* mbar = full + 1;                                        // MMA-complete barrier
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
  if (tid < kCout)
    s_par[tid] = make_float4(bn_scale[tid], bn_shift[tid], __bfloat162float(__float2bfloat16_rn(w8[tid])), __bfloat162float(__float2bfloat16_rn(w8[kCout + tid])));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(full)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;
  const float bias_seg = b8[0], bias_rad = b8[1];
  const uint32_t a0 = smem_u32(sA);
  const uint64_t db_base = make_desc(smem_u32(sB), 128, 1024);

  // thread 0: three boxes of 130 pixels x 64 channels (rows y-1, y, y+1 from pixel x0-1), one per slab
  auto load = [&](long long tile) {
    const long long img = tile / tiles_per_image;
    const int rem = (int)(tile - img * tiles_per_image);
    const int y = rem / tiles_per_row, x0 = (rem - y * tiles_per_row) * kTileM;
    const uint32_t bar = smem_u32(full);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3 * kBoxBytes) : "memory");
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(a0 + dy * kSlabBytes), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(x0 - 1), "r"(y + dy - 1), "r"((int)img), "r"(bar)
                   : "memory");
  };

  const long long step = gridDim.x;
  long long tile = blockIdx.x;
  if (tid == 0 && tile < n_tiles) load(tile);
  uint32_t phase = 0;
  for (; tile < n_tiles; tile += step) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                // every thread has read the last tile's accumulators
    if (tid == 0) {
      mbar_wait(smem_u32(full), phase);                             // the halo block has landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t - dy * 3;
#pragma unroll
        for (int ks = 0; ks < kCin / 16; ++ks) {
          const uint64_t da = make_desc_sw128(a0 + dy * kSlabBytes + dx * 128 + ks * 32, 1024, bo_mode ? (uint32_t)dx : 0u);
          const uint64_t db = db_base + (uint64_t)((t * kWTapBytes + ks * 256) >> 4);     // start-address field only (no carry: < 2^14)
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
    if (tid == 0 && tile + step < n_tiles) load(tile + step);       // the UMMAs have read the stage: the next block streams in under the epilogue
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
    float sv = bias_seg, rad = bias_rad;
#pragma unroll
    for (int ch = 0; ch < kCout; ++ch) {
      const float4 pr = s_par[ch];
      float a = fmaf(__uint_as_float(r[ch]), pr.x, pr.y);
      a = __bfloat162float(__float2bfloat16_rn(fmaxf(a, 0.f)));
      sv = fmaf(pr.z, a, sv);
      rad = fmaf(pr.w, a, rad);
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
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

// Tensor map of x seen as [n_images][H][W][64] bf16; box = 64 channels x 130 pixels of one image row, SWIZZLE_128B, zero fill outside.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tensor_map(CUtensorMap* tm, const void* x, int n_images, int H, int W) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return (int)cudaErrorNotSupported;
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)kCin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
  const cuuint64_t strides[3] = {(cuuint64_t)kCin * 2, (cuuint64_t)W * kCin * 2, (cuuint64_t)H * W * kCin * 2};
  const cuuint32_t box[4] = {(cuuint32_t)kCin, (cuuint32_t)kPx, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <bool kFused>
static int launch(const void* x, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8, const float* b8, float* out,
                  int n_images, int H, int W, int sms, void* stream, FusedArgs fz) {
  static bool attr_set = false;
  static int bo_mode = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv7_head<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    const char* ev = getenv("RCV_C7_BASE_OFFSET");            // experiment only
    if (ev) bo_mode = atoi(ev);
    attr_set = true;
  }
  CUtensorMap tm;
  const int rc = make_tensor_map(&tm, x, n_images, H, W);
  if (rc) return rc;
  const long long n_tiles = (long long)n_images * H * (W / kTileM);
  long long grid = 2LL * sms;            // two resident CTAs per SM
  if (grid > n_tiles) grid = n_tiles;
  k_conv7_head<kFused><<<(int)grid, kThreads, kSmem, (cudaStream_t)stream>>>(tm, w7, bn_scale, bn_shift, w8, b8, out, H, W, n_tiles, bo_mode, fz);
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
