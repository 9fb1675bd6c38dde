// Micro-benchmark: how fast ONE SM can stream L2-resident data, (a) with cp.async.bulk into a shared-memory ring of
// `depth` stages of `stage` bytes (what the chain engine's weight producer does) and (b) with 16-byte ld.global by
// `warps` warps (what the epilogue threads' stash reads do), for 1 / 8 / 148 CTAs running at once.
// Answers: is the ~29 B/clk per SM the weight ring reaches a per-SM limit, a ring-depth limit or chip-level contention?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fmov_pose_b200/csrc profiles/micro/l2_stream.cu -o /tmp/l2_stream
#include "fmov_common.cuh"
#include <cstdio>
#include <vector>
using namespace fmov;

constexpr int MAXD = 16;

__global__ void __launch_bounds__(1024, 1) bulk_ring(const uint8_t* __restrict__ src, long long src_bytes, int stage,
                                                     int depth, long long total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[MAXD];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long n = total / stage;
    const long long per = src_bytes / stage;
    long long off = (long long)blockIdx.x * 7 % per;
    const long long t0 = clock64();
    for (long long i = 0; i < n + depth; ++i) {
      if (i >= depth) mbar_wait_poll(&full[i % depth], ((i / depth) - 1) & 1);      // oldest copy landed -> stage reusable
      if (i < n) {
        mbar_expect_tx(&full[i % depth], stage);
        bulk_g2s(base + (i % depth) * stage, src + off * stage, stage, &full[i % depth]);
        off = (off + 1) % per;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

__global__ void __launch_bounds__(1024, 1) ldg_stream(const uint4* __restrict__ src, long long src_vec, long long total_vec,
                                                      long long* out, uint4* sink) {
  // every warp reads 512 contiguous bytes per instruction, UNR independent loads in flight per thread
  constexpr int UNR = 8;
  const long long per_iter = (long long)blockDim.x * UNR;
  long long pos = ((long long)blockIdx.x * 9973 * per_iter) % src_vec;
  uint4 acc = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
  for (long long done = 0; done < total_vec; done += per_iter) {
    uint4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      long long idx = pos + (long long)u * blockDim.x + threadIdx.x;
      if (idx >= src_vec) idx -= src_vec;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                   : "l"(src + idx));
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    pos += per_iter;
    if (pos >= src_vec) pos -= src_vec;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (acc.x == 0x12345678u && acc.y == 77u) sink[threadIdx.x] = acc;
}

int main() {
  const long long SRC = 3LL << 20;                 // 3 MB: the size of one kernel's weight images, L2 resident
  uint8_t* src;
  long long* out;
  uint4* sink;
  cudaMalloc(&src, 64LL << 20);
  cudaMemset(src, 1, 64LL << 20);
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(uint4));
  cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<long long> h(148);
  const long long total = 32LL << 20;              // bytes streamed per CTA
  printf("mode,grid,src_MB,stage_KB,depth,inflight_KB,B_per_clk_per_SM(avg),slowest_SM\n");
  for (long long srcb : {SRC, 48LL << 20})
    for (int grid : {1, 8, 148})
      for (int stage : {4096, 8192, 16384, 32768})
        for (int depth : {2, 3, 4, 6, 8, 12}) {
          if ((long long)stage * depth > 192 * 1024) continue;
          bulk_ring<<<grid, 32, 200 * 1024>>>(src, srcb, stage, depth, total, out);
          bulk_ring<<<grid, 32, 200 * 1024>>>(src, srcb, stage, depth, total, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          double s = 0, mx = 0;
          for (int i = 0; i < grid; ++i) { s += (double)h[i]; mx = mx > h[i] ? mx : (double)h[i]; }
          printf("bulk,%d,%lld,%d,%d,%d,%.1f,%.1f\n", grid, srcb >> 20, stage >> 10, depth, (stage * depth) >> 10,
                 total / (s / grid), total / mx);
        }
  for (long long srcb : {SRC, 48LL << 20})
    for (int grid : {1, 8, 148})
      for (int warps : {4, 8, 16, 32}) {
        ldg_stream<<<grid, warps * 32>>>((const uint4*)src, srcb / 16, total / 16, out, sink);
        ldg_stream<<<grid, warps * 32>>>((const uint4*)src, srcb / 16, total / 16, out, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0, mx = 0;
        for (int i = 0; i < grid; ++i) { s += (double)h[i]; mx = mx > h[i] ? mx : (double)h[i]; }
        printf("ldg,%d,%lld,warps=%d,unroll=8,%d,%.1f,%.1f\n", grid, srcb >> 20, warps, warps * 32 * 8 * 16 >> 10,
               total / (s / grid), total / mx);
      }
  return 0;
}
