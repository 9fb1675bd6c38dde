// Marching cubes on the dense u = -sdf grid (the step after extract_fields in validate_mesh: models/renderer.py:43,
// `mcubes.marching_cubes(u, threshold)` of PyMCubes 0.1.4 — CPU, third-party).  Produces the same INDEXED mesh: one
// vertex per crossed grid edge (shared by the cells around it), triangles from a 256-case table over the 12 cell edges.
//
// Three passes over the grid, all HBM-bound streaming kernels (thread per grid point, z fastest => coalesced rows;
// the +1 / +Z / +YZ neighbours come from L1/L2, so each pass reads u once from HBM: 4 B per point):
//   fmov_mc_count      per 256-point chunk: number of crossed edges starting at its points and of triangles of its cells
//   (host: exclusive prefix sum over the chunks — torch.cumsum; the totals size the outputs)
//   fmov_mc_vertices   vertex positions (index or world coordinates) + the vertex id of every crossed edge (vid3)
//   fmov_mc_triangles  case lookup, three vid3 reads per triangle corner
//   (both walk the list of non-empty chunks that fmov_mc_count appended to: ~2 % of the chunks at 512^3)
// Output order is deterministic: vertices by (grid point x-major, axis), triangles by (cell x-major, table order).
// Corner / edge numbering and the case table: fmov_pose_b200/mc_tables.py (uploaded once with fmov_mc_set_tables).
#include "fmov_common.cuh"
#include "mc_core.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

__device__ signed char d_mc_tri[256 * 3 * MC_MAX_TRIS];       // case table, read through L1 (mostly case 0 / 255)
__device__ unsigned char d_mc_ntri[256];
static bool g_mc_tables_set = false;

// exclusive prefix sum of v over the 256 threads of the block; *total = block sum
__device__ __forceinline__ int mc_block_exscan(int v, int* total) {
  __shared__ int warp_sum[MC_CHUNK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                       // previous use of warp_sum is over
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < MC_CHUNK / 32; ++w) {
    const int s = warp_sum[w];
    if (w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - v;
}

// A chunk without any crossing (the vast majority: the surface is a 2-D set in a 3-D grid) is recognised with one
// block-wide vote and skips both prefix sums; vertex and triangle counts share ONE scan (packed 16 + 16 bits: <= 3 * 256
// and <= 5 * 256 per chunk).
__device__ __forceinline__ int mc_pack(int nv, int nt) { return nv | (nt << 16); }

__global__ void __launch_bounds__(MC_CHUNK) mc_count_kernel(const McGrid g, int* __restrict__ chunk_nv,
                                                            int* __restrict__ chunk_nt, int* __restrict__ list,
                                                            int* __restrict__ n_list) {
  for (long long ch = blockIdx.x; ch < g.n_chunks; ch += gridDim.x) {
    const McPoint q = mc_point(g, ch * MC_CHUNK + threadIdx.x, true);
    const int v = mc_pack(mc_vertex_count(q), q.ntri);
    int tot = 0;
    if (__syncthreads_or(v)) mc_block_exscan(v, &tot);
    if (threadIdx.x == 0) {
      chunk_nv[ch] = tot & 0xFFFF;
      chunk_nt[ch] = tot >> 16;
      // the emit passes walk this list (about 2 % of the chunks at 512^3) instead of testing every chunk; its order is
      // whatever the atomics give, the OUTPUT positions come from the prefix sums and do not depend on it
      if (tot != 0) list[atomicAdd(n_list, 1)] = (int)ch;
    }
  }
}

__global__ void __launch_bounds__(MC_CHUNK) mc_vertices_kernel(const McGrid g, const long long* __restrict__ chunk_voff,
                                                               const int* __restrict__ list, const int* __restrict__ n_list,
                                                               const McXform xf, float* __restrict__ verts,
                                                               int* __restrict__ vid3) {
  const int n = *n_list;
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    const long long ch = list[li];
    if (chunk_voff[ch + 1] == chunk_voff[ch]) continue;          // triangles only (block-uniform)
    const long long p = ch * MC_CHUNK + threadIdx.x;
    const McPoint q = mc_point(g, p, false);
    const int nv = mc_vertex_count(q);
    int tv;
    const int local = mc_block_exscan(nv, &tv);
    if (nv) mc_emit_vertices(g, xf, p, q, chunk_voff[ch] + local, verts, vid3);
  }
}

__global__ void __launch_bounds__(MC_CHUNK) mc_triangles_kernel(const McGrid g, const long long* __restrict__ chunk_toff,
                                                                const int* __restrict__ list, const int* __restrict__ n_list,
                                                                const int* __restrict__ vid3, int* __restrict__ tris) {
  const int n = *n_list;
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    const long long ch = list[li];
    if (chunk_toff[ch + 1] == chunk_toff[ch]) continue;          // vertices only (block-uniform)
    const long long p = ch * MC_CHUNK + threadIdx.x;
    const McPoint q = mc_point(g, p, true);
    int tt;
    const int local = mc_block_exscan(q.ntri, &tt);
    if (q.ntri) mc_emit_triangles(g, p, q, chunk_toff[ch] + local, vid3, tris);
  }
}

static int mc_grid(McGrid& g, const float* u, int X, int Y, int Z, float iso) {
  FMOV_REQUIRE(u && X >= 2 && Y >= 2 && Z >= 2, "marching cubes: bad grid %d x %d x %d", X, Y, Z);
  FMOV_REQUIRE(g_mc_tables_set, "marching cubes: call fmov_mc_set_tables first");
  g.u = u; g.X = X; g.Y = Y; g.Z = Z; g.iso = iso;
  g.n = (long long)X * Y * Z;
  g.n_chunks = (g.n + MC_CHUNK - 1) / MC_CHUNK;
  mc_set_shifts(g);
  FMOV_REQUIRE(g.n < (1LL << 31), "marching cubes: grid too large (vertex ids are 32-bit)");
  void *pt = nullptr, *pn = nullptr;
  FMOV_CUDA(cudaGetSymbolAddress(&pt, d_mc_tri));
  FMOV_CUDA(cudaGetSymbolAddress(&pn, d_mc_ntri));
  g.tri = reinterpret_cast<const signed char*>(pt);
  g.ntri = reinterpret_cast<const unsigned char*>(pn);
  return OK;
}
static int mc_blocks(const McGrid& g) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (long long)sms * 8;          // 8 resident 256-thread blocks per SM
  return (int)(g.n_chunks < want ? g.n_chunks : want);
}

}  // namespace fmov
using namespace fmov;

/* tri_table: HOST [256][15] cell-edge numbers (-1 padded), n_tris: HOST [256] (fmov_pose_b200/mc_tables.py) */
extern "C" int fmov_mc_set_tables(const signed char* tri_table, const unsigned char* n_tris) {
  FMOV_REQUIRE(tri_table && n_tris, "fmov_mc_set_tables: null argument");
  for (int c = 0; c < 256; ++c) {
    FMOV_REQUIRE(n_tris[c] <= MC_MAX_TRIS, "fmov_mc_set_tables: case %d has %d triangles", c, (int)n_tris[c]);
    for (int i = 0; i < 3 * n_tris[c]; ++i)
      FMOV_REQUIRE(tri_table[c * 15 + i] >= 0 && tri_table[c * 15 + i] < 12, "fmov_mc_set_tables: bad edge in case %d", c);
  }
  FMOV_CUDA(cudaMemcpyToSymbol(d_mc_tri, tri_table, 256 * 3 * MC_MAX_TRIS));
  FMOV_CUDA(cudaMemcpyToSymbol(d_mc_ntri, n_tris, 256));
  g_mc_tables_set = true;
  return OK;
}

extern "C" long long fmov_mc_chunk_count(int X, int Y, int Z) {
  return ((long long)X * Y * Z + MC_CHUNK - 1) / MC_CHUNK;
}

extern "C" int fmov_mc_count(const float* u, int X, int Y, int Z, float iso, int* chunk_nv, int* chunk_nt, int* list,
                             int* n_list, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_nv && chunk_nt && list && n_list, "fmov_mc_count: null output");
  FMOV_CUDA(cudaMemsetAsync(n_list, 0, sizeof(int), (cudaStream_t)stream));
  mc_count_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_nv, chunk_nt, list, n_list);
  FMOV_LAUNCH_CHECK("mc_count_kernel");
  return OK;
}

extern "C" int fmov_mc_vertices(const float* u, int X, int Y, int Z, float iso, const long long* chunk_voff, const int* list,
                                const int* n_list, float sx, float sy, float sz, float ox, float oy, float oz, float* verts,
                                int* vid3, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_voff && list && n_list && verts && vid3, "fmov_mc_vertices: null argument");
  McXform xf;
  xf.s[0] = sx; xf.s[1] = sy; xf.s[2] = sz; xf.o[0] = ox; xf.o[1] = oy; xf.o[2] = oz;
  mc_vertices_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_voff, list, n_list, xf, verts, vid3);
  FMOV_LAUNCH_CHECK("mc_vertices_kernel");
  return OK;
}

extern "C" int fmov_mc_triangles(const float* u, int X, int Y, int Z, float iso, const long long* chunk_toff, const int* list,
                                 const int* n_list, const int* vid3, int* tris, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_toff && list && n_list && vid3 && tris, "fmov_mc_triangles: null argument");
  mc_triangles_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_toff, list, n_list, vid3, tris);
  FMOV_LAUNCH_CHECK("mc_triangles_kernel");
  return OK;
}
