// Micro-benchmark 3: where do one thread's cp.async.bulk requests serialise — at issue or at completion — and do requests
// issued by several LANES of one warp instruction run in parallel?
//   burst   R requests of S bytes, each with its own stage and mbarrier: cycles to issue them all, cycles until all landed;
//           mode 0 = one lane issues them in a loop, mode 1 = R lanes issue one each in ONE warp instruction
//   ring    steady-state ring whose stages are filled by `lanes` lanes at once (each lane one piece of the stage)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fmov_pose_b200/csrc profiles/micro/l2_stream3.cu -o /tmp/l2_stream3
#include "fmov_common.cuh"
#include <cstdio>
#include <vector>
using namespace fmov;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_kind(uint64_t* bar, uint32_t parity, int spin) {
  if (spin) { while (!mbar_test_wait(bar, parity)) {} }
  else mbar_wait_poll(bar, parity);
}

__global__ void __launch_bounds__(1024, 1) burst(const uint8_t* __restrict__ src, int S, int R, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    for (int rep = 0; rep < 4; ++rep) {        // the last repetition is reported (L2 / TLB warm)
      __syncwarp();
      const long long t0 = clock64();
      if (mode == 0) {
        if (lane == 0)
          for (int i = 0; i < R; ++i) {
            mbar_expect_tx(&full[i], S);
            bulk_g2s(base + (size_t)i * S, src + (size_t)(i + rep * R) * S, S, &full[i]);
          }
      } else if (lane < R) {
        mbar_expect_tx(&full[lane], S);
        bulk_g2s(base + (size_t)lane * S, src + (size_t)(lane + rep * R) * S, S, &full[lane]);
      }
      __syncwarp();
      const long long t1 = clock64();
      long long tfirst = 0;
      if (lane == 0) {          // non-blocking polls over all barriers: when does each request land?
        unsigned int pending = (1u << R) - 1u;
        while (pending) {
          for (int i = 0; i < R; ++i)
            if ((pending >> i) & 1u) {
              if (mbar_test_wait(&full[i], rep & 1)) {
                pending &= ~(1u << i);
                if (tfirst == 0) tfirst = clock64();
                if (rep == 3) out[4 + i] = clock64() - t0;
              }
            }
        }
      }
      __syncwarp();
      const long long t2 = clock64();
      if (lane == 0) { out[0] = t1 - t0; out[1] = tfirst - t0; out[2] = t2 - t0; }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) ring_lanes(const uint8_t* __restrict__ src, long long src_bytes, int stage,
                                                      int depth, int lanes, int spin, long long total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const long long n = total / stage;
    const long long per = src_bytes / stage;
    long long off = (long long)blockIdx.x * 7 % per;
    const int psz = stage / lanes;
    const long long t0 = clock64();
    for (long long i = 0; i < n + depth; ++i) {
      if (i >= depth) wait_kind(&full[i % depth], ((i / depth) - 1) & 1, spin);      // all lanes poll (warp-uniform result)
      if (i < n) {
        if (lane == 0) mbar_expect_tx(&full[i % depth], stage);
        __syncwarp();
        if (lane < lanes)
          bulk_g2s(base + (i % depth) * stage + lane * psz, src + off * stage + lane * psz, psz, &full[i % depth]);
        off = (off + 1) % per;
      }
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const long long SRC = 3LL << 20;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, 64LL << 20);
  cudaMemset(src, 1, 64LL << 20);
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ring_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[148];
  printf("burst,mode,S_KB,R,issue_cycles,first_landed,all_landed\n");
  for (int mode : {0, 1})
    for (int S : {2048, 4096, 16384})
      for (int R : {1, 2, 4, 8, 12}) {
        if ((long long)S * R > 192 * 1024) continue;
        burst<<<1, 32, 200 * 1024>>>(src, S, R, mode, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(h, out, 20 * sizeof(long long), cudaMemcpyDeviceToHost);
        printf("burst,%d,%d,%d,%lld,%lld,%lld | landed:", mode, S >> 10, R, h[0], h[1], h[2]);
        for (int i = 0; i < R; ++i) printf(" %lld", h[4 + i]);
        printf("\n");
      }
  const long long total = 24LL << 20;
  printf("ring_lanes,grid,stage_KB,depth,lanes,spin,B_per_clk_per_SM(avg),slowest\n");
  for (int grid : {1, 148})
   for (int spin : {0, 1})
    for (int stage : {4096, 16384, 32768})
      for (int depth : {2, 4, 8})
        for (int lanes : {1, 4}) {
          if ((long long)stage * depth > 192 * 1024) continue;
          for (int rep = 0; rep < 2; ++rep) ring_lanes<<<grid, 32, 200 * 1024>>>(src, SRC, stage, depth, lanes, spin, total, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          double s = 0, mx = 0;
          for (int i = 0; i < grid; ++i) { s += (double)h[i]; mx = mx > h[i] ? mx : (double)h[i]; }
          printf("ring_lanes,%d,%d,%d,%d,%d,%.1f,%.1f\n", grid, stage >> 10, depth, lanes, spin, total / (s / grid), total / mx);
          fflush(stdout);
        }
  return 0;
}
