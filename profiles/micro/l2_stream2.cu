// Micro-benchmark 2 (follow-up of l2_stream.cu, which showed that ONE thread's cp.async.bulk requests complete one at a
// time, ~843 cycles each whatever their size: 16 KiB slices -> 19 B/clk per SM while plain ld.global reaches 90 B/clk).
// What lifts that limit?
//   parts    each ring stage filled by `parts` smaller requests on one mbarrier
//   issuers  several warps (one lane each), each with its own ring and barriers
//   tensor   cp.async.bulk.tensor.2d (tensor map over the flat buffer viewed as rows of 256 bytes) instead of cp.async.bulk
//   big      64 KiB requests
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fmov_pose_b200/csrc profiles/micro/l2_stream2.cu -o /tmp/l2_stream2
#include "fmov_common.cuh"
#include <cuda.h>
#include <cstdio>
#include <vector>
using namespace fmov;

constexpr int MAXB = 64;

__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(1024, 1) ring(const uint8_t* __restrict__ src, long long src_bytes, int stage, int depth,
                                                int parts, int issuers, int tensor, long long total,
                                                const __grid_constant__ CUtensorMap tm, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[MAXB];
  __shared__ long long cyc[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * issuers; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < issuers) {
    uint64_t* fb = full + warp * depth;
    uint8_t* rb = base + (size_t)warp * depth * stage;
    const long long n = total / stage / issuers;
    const long long per = src_bytes / stage;
    long long off = ((long long)blockIdx.x * 7 + warp * 3) % per;
    const int psz = stage / parts;
    const long long t0 = clock64();
    for (long long i = 0; i < n + depth; ++i) {
      if (i >= depth) mbar_wait_poll(&fb[i % depth], ((i / depth) - 1) & 1);
      if (i < n) {
        mbar_expect_tx(&fb[i % depth], stage);
        uint8_t* dst = rb + (i % depth) * stage;
        if (tensor) {
          for (int p = 0; p < parts; ++p) tma_2d(dst + p * psz, &tm, 0, (int)((off * stage + p * psz) / 256), &fb[i % depth]);
        } else {
          for (int p = 0; p < parts; ++p) bulk_g2s(dst + p * psz, src + off * stage + p * psz, psz, &fb[i % depth]);
        }
        off = (off + 1) % per;
      }
    }
    cyc[warp] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int w = 0; w < issuers; ++w) m = cyc[w] > m ? cyc[w] : m;
    out[blockIdx.x] = m;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long SRC = 3LL << 20;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, 64LL << 20);
  cudaMemset(src, 1, 64LL << 20);
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  std::vector<long long> h(148);
  const long long total = 24LL << 20;
  printf("kind,grid,stage_KB,depth,parts,issuers,inflight_KB,B_per_clk_per_SM(avg),slowest_SM\n");
  struct Cfg { int tensor, stage, depth, parts, issuers; };
  std::vector<Cfg> cfgs;
  for (int stage : {16384, 32768, 65536})
    for (int depth : {2, 3}) cfgs.push_back({0, stage, depth, 1, 1});
  for (int parts : {2, 4, 8}) cfgs.push_back({0, 16384, 4, parts, 1});
  for (int parts : {2, 4, 8}) cfgs.push_back({0, 32768, 3, parts, 1});
  for (int iss : {2, 4}) cfgs.push_back({0, 16384, 2, 1, iss});
  for (int iss : {2, 4, 8}) cfgs.push_back({0, 8192, 2, 1, iss});
  for (int stage : {8192, 16384, 32768, 65536})
    for (int depth : {2, 4}) {
      if ((long long)stage * depth > 192 * 1024) continue;
      cfgs.push_back({1, stage, depth, 1, 1});
    }
  for (int parts : {2, 4}) cfgs.push_back({1, 32768, 3, parts, 1});
  for (int iss : {2, 4}) cfgs.push_back({1, 16384, 2, 1, iss});
  for (int grid : {1, 148})
    for (const Cfg& c : cfgs) {
      if ((long long)c.stage * c.depth * c.issuers > 192 * 1024) continue;
      CUtensorMap tm;
      const int prow = c.stage / c.parts / 256;      // box rows of 256 bytes
      cuuint64_t gdim[2] = {256, (cuuint64_t)(SRC / 256)};
      cuuint64_t gstr[1] = {256};
      cuuint32_t box[2] = {256, (cuuint32_t)prow};
      cuuint32_t estr[2] = {1, 1};
      if (c.tensor) {
        if (prow > 256 || prow < 1) continue;
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      } else {
        memset(&tm, 0, sizeof(tm));
      }
      for (int rep = 0; rep < 2; ++rep)
        ring<<<grid, 32 * (c.issuers > 1 ? c.issuers : 1), 200 * 1024>>>(src, SRC, c.stage, c.depth, c.parts, c.issuers,
                                                                          c.tensor, total, tm, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0, mx = 0;
      for (int i = 0; i < grid; ++i) { s += (double)h[i]; mx = mx > h[i] ? mx : (double)h[i]; }
      printf("%s,%d,%d,%d,%d,%d,%d,%.1f,%.1f\n", c.tensor ? "tensor" : "bulk", grid, c.stage >> 10, c.depth, c.parts,
             c.issuers, (c.stage * c.depth * c.issuers) >> 10, total / (s / grid), total / mx);
      fflush(stdout);
    }
  return 0;
}
