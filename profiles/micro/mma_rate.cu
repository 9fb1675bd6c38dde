// Micro-benchmark: cycles per tcgen05.mma (M=128, N=256, K=16, kind::f16, cta_group::1) issued back to back on
// resident shared-memory operands, for the no-swizzle "interleaved" layout the chain engine uses and for SWIZZLE_128B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fmov_pose_b200/csrc profiles/micro/mma_rate.cu -o /tmp/mma_rate
#include "fmov_common.cuh"
#include <vector>
using namespace fmov;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // K-major, SWIZZLE_128B: SBO = 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1) << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) k(long long* out, int reps, int mode, int n, int with_ld) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint8_t* A = base;                 // 64 KB: 4 k-blocks [128 x 64]
  uint8_t* B = base + 64 * 1024;     // up to 4 x 32 KB would not fit: reuse 2 k-blocks (64 KB)
  const uint32_t idesc = umma_idesc(128, n, FMT_F16, FMT_F16, 0, 0);
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      for (int kb = 0; kb < 4; ++kb) {
        const uint32_t a_base = smem_u32(A + kb * BLK_BYTES);
        const uint32_t b_base = smem_u32(B + (kb & 1) * 32768);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (mode == 0)
            umma_f16(tmem, umma_desc_kmajor(a_base + ks * 2 * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                     umma_desc_kmajor(b_base + ks * 2 * (n * 16), n * 16), idesc, (kb | ks) != 0);
          else
            umma_f16(tmem, desc_sw128(a_base + ks * 32), desc_sw128(b_base + ks * 32), idesc, (kb | ks) != 0);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      tc_fence_after();
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (with_ld && threadIdx.x >= 32) {
    // concurrent shared-memory traffic from "epilogue" threads: st.shared of 16 B per thread in a loop
    uint4* p = reinterpret_cast<uint4*>(base + 128 * 1024) + threadIdx.x;
    for (int i = 0; i < reps * 200; ++i) { p[(i & 7) * 128] = make_uint4(i, i, i, i); }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 200;
  for (int n : {256, 128}) for (int mode : {0, 1}) for (int with_ld : {0, 1}) {
    k<<<148, 128, 200 * 1024>>>(d, reps, mode, n, with_ld);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60;
    for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    printf("N=%d layout=%s smem_traffic=%d: %.1f .. %.1f cycles per MMA (K=16)  [%s]\n", n, mode ? "SW128" : "interleaved", with_ld,
           (double)mn / (reps * 16), (double)mx / (reps * 16), cudaGetErrorString(e));
  }
  return 0;
}
