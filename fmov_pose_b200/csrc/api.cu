// C-ABI plumbing: last-error string, version, weight-image packing, tile-image (un)packing and the
// tcgen05 self-test GEMMs used by tests/ to validate descriptors independently of the chain kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void note_launch() { ++g_launches; }
int cuda_fail(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return ERR_CUDA;
}
}  // namespace fmov
using namespace fmov;

extern "C" const char* fmov_last_error(void) { return g_err; }
extern "C" int fmov_version(void) { return 100; }
/* kernels of this library launched so far by this process (every launch site counts itself; launches recorded during a
 * CUDA-graph capture are counted once, at capture) */
extern "C" unsigned long long fmov_launch_count(void) { return fmov::g_launches; }

// ----------------------------------------------------------------------------------------
// Weight image packing: fp32 matrix -> [npad rows x 64*kblocks] fp16/bf16 no-swizzle operand image.
// dst(n, k) = scale * src[(n + row_off)*stride_n + kmap(k)*stride_k] for n < n_valid and k inside a
// segment, else 0.  Segments let the caller permute / split input columns (skip connection,
// colour-net input order) without materialising a permuted matrix.
// ----------------------------------------------------------------------------------------
struct PackArgs {
  const float* src;
  long long stride_n, stride_k;
  int n_valid, row_off;
  int nseg;
  int seg_dst[4], seg_src[4], seg_len[4];
  float scale;
  int bf16;
  uint8_t* dst;
  int npad, kblocks;
};

__global__ void pack_image_kernel(PackArgs a) {
  const int total = a.npad * a.kblocks * 8;   // 16-byte chunks
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kb = i / (a.npad * 8);
    const int rem = i - kb * a.npad * 8;
    const int n = rem >> 3, ch = rem & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb * 64 + ch * 8 + j;
      float x = 0.f;
      if (n < a.n_valid) {
        for (int sgi = 0; sgi < a.nseg; ++sgi) {
          if (k >= a.seg_dst[sgi] && k < a.seg_dst[sgi] + a.seg_len[sgi]) {
            const long long sk = a.seg_src[sgi] + (k - a.seg_dst[sgi]);
            x = a.scale * a.src[(long long)(n + a.row_off) * a.stride_n + sk * a.stride_k];
          }
        }
      }
      v[j] = x;
    }
    uint4 q;
    if (a.bf16 == 2) {      // residual image: what the fp16 image of the same matrix drops
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] -= __half2float(__float2half_rn(v[j]));
    }
    if (a.bf16 == 1) {
      q.x = pack_bf2(v[0], v[1]); q.y = pack_bf2(v[2], v[3]); q.z = pack_bf2(v[4], v[5]); q.w = pack_bf2(v[6], v[7]);
    } else {
      q.x = pack_h2(v[0], v[1]); q.y = pack_h2(v[2], v[3]); q.z = pack_h2(v[4], v[5]); q.w = pack_h2(v[6], v[7]);
    }
    uint8_t* blk = a.dst + (size_t)kb * a.npad * 128;
    *reinterpret_cast<uint4*>(blk + (size_t)ch * a.npad * 16 + (size_t)n * 16) = q;   // [chunk][row][16 B]
  }
}

extern "C" int fmov_pack_image(const float* src, long long stride_n, long long stride_k, int n_valid, int row_off, int nseg,
                               const int* seg_dst, const int* seg_src, const int* seg_len, float scale, int bf16,
                               void* dst, int npad, int kblocks, void* stream) {
  FMOV_REQUIRE(src && dst && nseg >= 1 && nseg <= 4 && npad > 0 && npad % 8 == 0 && kblocks > 0 && n_valid <= npad,
               "fmov_pack_image: bad arguments (npad=%d kblocks=%d nseg=%d n_valid=%d)", npad, kblocks, nseg, n_valid);
  PackArgs a;
  a.src = src; a.stride_n = stride_n; a.stride_k = stride_k; a.n_valid = n_valid; a.row_off = row_off; a.nseg = nseg;
  for (int i = 0; i < 4; ++i) {
    a.seg_dst[i] = i < nseg ? seg_dst[i] : 0;
    a.seg_src[i] = i < nseg ? seg_src[i] : 0;
    a.seg_len[i] = i < nseg ? seg_len[i] : 0;
  }
  a.scale = scale; a.bf16 = bf16; a.dst = reinterpret_cast<uint8_t*>(dst); a.npad = npad; a.kblocks = kblocks;
  const int total = npad * kblocks * 8;
  pack_image_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("pack_image_kernel");
  return OK;
}

// ----------------------------------------------------------------------------------------
// Tile image <-> row-major fp32 (test / debug utilities; also used to seed stash tensors in tests)
// ----------------------------------------------------------------------------------------
__global__ void ti_from_rowmajor_kernel(const float* src, long long P, int cols, int ld, int kblocks, int bf16,
                                        uint8_t* dst) {
  const long long n_tiles = (P + 127) / 128;
  const long long total = n_tiles * kblocks * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long tk = i >> 10;
    const int kb = (int)(tk % kblocks);
    const long long tile = tk / kblocks;
    const long long p = tile * 128 + r;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = kb * 64 + ch * 8 + j;
      v[j] = (p < P && c < cols) ? src[p * ld + c] : 0.f;
    }
    uint4 q;
    if (bf16) {
      q.x = pack_bf2(v[0], v[1]); q.y = pack_bf2(v[2], v[3]); q.z = pack_bf2(v[4], v[5]); q.w = pack_bf2(v[6], v[7]);
    } else {
      q.x = pack_h2(v[0], v[1]); q.y = pack_h2(v[2], v[3]); q.z = pack_h2(v[4], v[5]); q.w = pack_h2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)tk * BLK_BYTES + ti_chunk_off(r, ch)) = q;
  }
}
__global__ void ti_to_rowmajor_kernel(const uint8_t* src, long long P, int cols, int ld, int kblocks, int bf16,
                                      float* dst) {
  const long long n_tiles = (P + 127) / 128;
  const long long total = n_tiles * kblocks * 128 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long tk = i >> 10;
    const int kb = (int)(tk % kblocks);
    const long long tile = tk / kblocks;
    const long long p = tile * 128 + r;
    if (p >= P) continue;
    const uint4 q = *reinterpret_cast<const uint4*>(src + (size_t)tk * BLK_BYTES + ti_chunk_off(r, ch));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = bf16 ? unpack_bf2(w[j]) : unpack_h2(w[j]);
      const int c = kb * 64 + ch * 8 + 2 * j;
      if (c < cols) dst[p * ld + c] = f.x;
      if (c + 1 < cols) dst[p * ld + c + 1] = f.y;
    }
  }
}

extern "C" int fmov_ti_from_rowmajor(const float* src, long long P, int cols, int ld, int kblocks, int bf16, void* dst,
                                     void* stream) {
  FMOV_REQUIRE(src && dst && P > 0 && cols > 0 && cols <= kblocks * 64 && ld >= cols, "fmov_ti_from_rowmajor: bad arguments");
  ti_from_rowmajor_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(src, P, cols, ld, kblocks, bf16, (uint8_t*)dst);
  FMOV_LAUNCH_CHECK("ti_from_rowmajor_kernel");
  return OK;
}
extern "C" int fmov_ti_to_rowmajor(const void* src, long long P, int cols, int ld, int kblocks, int bf16, float* dst,
                                   void* stream) {
  FMOV_REQUIRE(src && dst && P > 0 && cols > 0 && cols <= kblocks * 64 && ld >= cols, "fmov_ti_to_rowmajor: bad arguments");
  ti_to_rowmajor_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, P, cols, ld, kblocks, bf16, dst);
  FMOV_LAUNCH_CHECK("ti_to_rowmajor_kernel");
  return OK;
}

// ----------------------------------------------------------------------------------------
// Self-test GEMMs (one CTA, one 128-row tile) exercising exactly the descriptors the product
// kernels use:  mode 0  D[128 x N] = A[128 x 64*KB] * W[N x 64*KB]^T   (K-major A and B)
//               mode 1  D[128 x N] = A^T B with A = TI[128 pts x 128 feats] (first 128 feats), B =
//                       TI[128 pts x N feats]  (MN-major A and B, K = 128 points) -> the dW form.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
selftest_gemm_kernel(const uint8_t* a_img, const uint8_t* b_img, int n, int kblocks, int a_fmt, int b_fmt, int mode,
                     float* out) {
  extern __shared__ uint8_t smem_raw[];
  uintptr_t pp = (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023);
  uint8_t* base = reinterpret_cast<uint8_t*>(pp);
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int a_blocks = (mode == 0) ? kblocks : 2;
  const int b_blocks = (mode == 0) ? kblocks : (n + 63) / 64;
  const uint32_t a_bytes = a_blocks * BLK_BYTES;
  const uint32_t b_bytes = (mode == 0) ? (uint32_t)kblocks * n * 128u : (uint32_t)b_blocks * BLK_BYTES;
  uint8_t* sa = base;
  uint8_t* sb = base + a_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, a_bytes + b_bytes);
    for (int i = 0; i < a_blocks; ++i) bulk_g2s(sa + i * BLK_BYTES, a_img + (size_t)i * BLK_BYTES, BLK_BYTES, &bar_load);
    if (mode == 0) {
      for (int i = 0; i < kblocks; ++i) bulk_g2s(sb + (size_t)i * n * 128, b_img + (size_t)i * n * 128, n * 128, &bar_load);
    } else {
      for (int i = 0; i < b_blocks; ++i) bulk_g2s(sb + i * BLK_BYTES, b_img + (size_t)i * BLK_BYTES, BLK_BYTES, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = umma_idesc(128, n, a_fmt, b_fmt, 0, 0);
      for (int kb = 0; kb < kblocks; ++kb)
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem, umma_desc_kmajor(smem_u32(sa + kb * BLK_BYTES) + ks * 2 * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                   umma_desc_kmajor(smem_u32(sb + (size_t)kb * n * 128) + ks * 2 * (n * 16), n * 16), idesc,
                   (kb | ks) ? 1u : 0u);
    } else {
      const uint32_t idesc = umma_idesc(128, n, a_fmt, b_fmt, 1, 1);
      for (int ks = 0; ks < 8; ++ks)   // K = 128 points, 16 per instruction = 2 groups of 8 rows = 256 bytes
        umma_f16(tmem, umma_desc_mnmajor(smem_u32(sa) + ks * 256, TI_CHUNK_STRIDE),
                 umma_desc_mnmajor(smem_u32(sb) + ks * 256, TI_CHUNK_STRIDE), idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t)row * n + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

extern "C" int fmov_selftest_gemm(const void* a_img, const void* b_img, int n, int kblocks, int a_bf16, int b_bf16,
                                  int mode, float* out, void* stream) {
  FMOV_REQUIRE(a_img && b_img && out && n >= 16 && n <= 256 && n % 16 == 0 && kblocks >= 1 && kblocks <= 5 &&
                   (mode == 0 || mode == 1),
               "fmov_selftest_gemm: bad arguments n=%d kblocks=%d mode=%d", n, kblocks, mode);
  const int a_blocks = (mode == 0) ? kblocks : 2;
  const int b_bytes = (mode == 0) ? kblocks * n * 128 : ((n + 63) / 64) * BLK_BYTES;
  const int smem = a_blocks * BLK_BYTES + b_bytes + 1024;
  FMOV_REQUIRE(smem <= 200 * 1024, "fmov_selftest_gemm: operands do not fit in shared memory (%d bytes)", smem);
  static bool attr_set = false;
  if (!attr_set) {
    FMOV_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  selftest_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)a_img, (const uint8_t*)b_img, n, kblocks,
                                                            a_bf16 ? FMT_BF16 : FMT_F16, b_bf16 ? FMT_BF16 : FMT_F16, mode, out);
  FMOV_LAUNCH_CHECK("selftest_gemm_kernel");
  return OK;
}
