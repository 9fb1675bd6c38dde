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
//   (both skip chunks whose prefix-sum entry shows nothing to emit: chunk_voff / chunk_toff have n_chunks + 1 entries)
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

// mc_point(g, p, true) for a warp of 32 consecutive points: every lane loads the four (x, y) corner rows at its own z and
// takes the z + 1 values from lane + 1 by shuffle (lane 31 loads them), i.e. 4 instead of 8 loads per point.  Same values,
// same marks, same case index as mc_core.cuh's mc_point (which the host emulation checks against the oracle).
__device__ __forceinline__ McPoint mc_point_warp(const McGrid& g, long long p) {
  McPoint q;
  q.valid = p < g.n;
  q.i = q.j = q.k = 0;
  q.f0 = 0.f; q.b0 = false; q.ntri = 0; q.cubecase = 0;
  q.cross[0] = q.cross[1] = q.cross[2] = false;
  q.f1[0] = q.f1[1] = q.f1[2] = 0.f;
  const int lane = threadIdx.x & 31;
  const int YZ = g.Y * g.Z;
  bool hx = false, hy = false, hz = false;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;          // u at (i,j,k) (i,j+1,k) (i+1,j,k) (i+1,j+1,k)
  if (q.valid) {
    const unsigned int pu = (unsigned int)p;
    q.i = (int)(pu / (unsigned int)YZ);
    const int r = (int)(pu - (unsigned int)q.i * (unsigned int)YZ);
    q.j = r / g.Z;
    q.k = r - q.j * g.Z;
    hx = q.i + 1 < g.X; hy = q.j + 1 < g.Y; hz = q.k + 1 < g.Z;
    a00 = g.u[p];
    if (hy) a01 = g.u[p + g.Z];
    if (hx) a10 = g.u[p + YZ];
    if (hx && hy) a11 = g.u[p + YZ + g.Z];
  }
  // z + 1 neighbours: lane + 1 holds point p + 1 = (i, j, k + 1) whenever hz (same row); lane 31 loads its own
  float b00 = __shfl_down_sync(0xffffffffu, a00, 1), b01 = __shfl_down_sync(0xffffffffu, a01, 1);
  float b10 = __shfl_down_sync(0xffffffffu, a10, 1), b11 = __shfl_down_sync(0xffffffffu, a11, 1);
  if (lane == 31 && q.valid && hz) {
    b00 = g.u[p + 1];
    if (hy) b01 = g.u[p + g.Z + 1];
    if (hx) b10 = g.u[p + YZ + 1];
    if (hx && hy) b11 = g.u[p + YZ + g.Z + 1];
  }
  if (!q.valid) return q;
  q.f0 = a00;
  q.b0 = a00 < g.iso;
  if (hx) { q.f1[0] = a10; q.cross[0] = (a10 < g.iso) != q.b0; }
  if (hy) { q.f1[1] = a01; q.cross[1] = (a01 < g.iso) != q.b0; }
  if (hz) { q.f1[2] = b00; q.cross[2] = (b00 < g.iso) != q.b0; }
  if (hx && hy && hz) {
    // corners v0..v7 (mc_tables.py): (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
    int c = q.b0 ? 1 : 0;
    c |= (a10 < g.iso) ? 2 : 0;
    c |= (a11 < g.iso) ? 4 : 0;
    c |= (a01 < g.iso) ? 8 : 0;
    c |= (b00 < g.iso) ? 16 : 0;
    c |= (b10 < g.iso) ? 32 : 0;
    c |= (b11 < g.iso) ? 64 : 0;
    c |= (b01 < g.iso) ? 128 : 0;
    q.cubecase = c;
    q.ntri = (c == 0 || c == 255) ? 0 : g.ntri[c];
  }
  return q;
}

__global__ void __launch_bounds__(MC_CHUNK) mc_count_kernel(const McGrid g, int* __restrict__ chunk_nv,
                                                            int* __restrict__ chunk_nt) {
  for (long long ch = blockIdx.x; ch < g.n_chunks; ch += gridDim.x) {
    const McPoint q = mc_point_warp(g, ch * MC_CHUNK + threadIdx.x);
    const int v = mc_pack(mc_vertex_count(q), q.ntri);
    int tot = 0;
    if (__syncthreads_or(v)) mc_block_exscan(v, &tot);
    if (threadIdx.x == 0) {
      chunk_nv[ch] = tot & 0xFFFF;
      chunk_nt[ch] = tot >> 16;
    }
  }
}

// The emit passes only touch chunks that emit something (about 2 % of them at 512^3): a block looks at MC_CHUNK chunk
// offsets at a time (one coalesced load per thread), compacts the non-empty ones into a shared list and walks that.
__device__ __forceinline__ int mc_nonempty_list(const long long* __restrict__ off, long long n_chunks, long long batch,
                                                int* list) {
  __shared__ int n_list;
  const long long ch = batch * MC_CHUNK + threadIdx.x;
  const bool ne = ch < n_chunks && off[ch + 1] != off[ch];
  if (threadIdx.x == 0) n_list = 0;
  __syncthreads();
  if (ne) list[atomicAdd(&n_list, 1)] = threadIdx.x;
  __syncthreads();
  const int n = n_list;
  __syncthreads();          // n_list is reset by the next call
  return n;
}

__global__ void __launch_bounds__(MC_CHUNK) mc_vertices_kernel(const McGrid g, const long long* __restrict__ chunk_voff,
                                                               const McXform xf, float* __restrict__ verts,
                                                               int* __restrict__ vid3) {
  __shared__ int list[MC_CHUNK];
  const long long n_batches = (g.n_chunks + MC_CHUNK - 1) / MC_CHUNK;
  for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int n = mc_nonempty_list(chunk_voff, g.n_chunks, batch, list);
    for (int k = 0; k < n; ++k) {
      const long long ch = batch * MC_CHUNK + list[k];
      const long long p = ch * MC_CHUNK + threadIdx.x;
      const McPoint q = mc_point(g, p, false);
      const int nv = mc_vertex_count(q);
      int tv;
      const int local = mc_block_exscan(nv, &tv);
      if (nv) mc_emit_vertices(g, xf, p, q, chunk_voff[ch] + local, verts, vid3);
    }
    __syncthreads();          // `list` is rewritten by the next batch
  }
}

__global__ void __launch_bounds__(MC_CHUNK) mc_triangles_kernel(const McGrid g, const long long* __restrict__ chunk_toff,
                                                                const int* __restrict__ vid3, int* __restrict__ tris) {
  __shared__ int list[MC_CHUNK];
  const long long n_batches = (g.n_chunks + MC_CHUNK - 1) / MC_CHUNK;
  for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int n = mc_nonempty_list(chunk_toff, g.n_chunks, batch, list);
    for (int k = 0; k < n; ++k) {
      const long long ch = batch * MC_CHUNK + list[k];
      const long long p = ch * MC_CHUNK + threadIdx.x;
      const McPoint q = mc_point_warp(g, p);
      int tt;
      const int local = mc_block_exscan(q.ntri, &tt);
      if (q.ntri) mc_emit_triangles(g, p, q, chunk_toff[ch] + local, vid3, tris);
    }
    __syncthreads();
  }
}

static int mc_grid(McGrid& g, const float* u, int X, int Y, int Z, float iso) {
  FMOV_REQUIRE(u && X >= 2 && Y >= 2 && Z >= 2, "marching cubes: bad grid %d x %d x %d", X, Y, Z);
  FMOV_REQUIRE(g_mc_tables_set, "marching cubes: call fmov_mc_set_tables first");
  g.u = u; g.X = X; g.Y = Y; g.Z = Z; g.iso = iso;
  g.n = (long long)X * Y * Z;
  g.n_chunks = (g.n + MC_CHUNK - 1) / MC_CHUNK;
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

extern "C" int fmov_mc_count(const float* u, int X, int Y, int Z, float iso, int* chunk_nv, int* chunk_nt, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_nv && chunk_nt, "fmov_mc_count: null output");
  mc_count_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_nv, chunk_nt);
  FMOV_LAUNCH_CHECK("mc_count_kernel");
  return OK;
}

extern "C" int fmov_mc_vertices(const float* u, int X, int Y, int Z, float iso, const long long* chunk_voff, float sx,
                                float sy, float sz, float ox, float oy, float oz, float* verts, int* vid3, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_voff && verts && vid3, "fmov_mc_vertices: null argument");
  McXform xf;
  xf.s[0] = sx; xf.s[1] = sy; xf.s[2] = sz; xf.o[0] = ox; xf.o[1] = oy; xf.o[2] = oz;
  mc_vertices_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_voff, xf, verts, vid3);
  FMOV_LAUNCH_CHECK("mc_vertices_kernel");
  return OK;
}

extern "C" int fmov_mc_triangles(const float* u, int X, int Y, int Z, float iso, const long long* chunk_toff,
                                 const int* vid3, int* tris, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_toff && vid3 && tris, "fmov_mc_triangles: null argument");
  mc_triangles_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_toff, vid3, tris);
  FMOV_LAUNCH_CHECK("mc_triangles_kernel");
  return OK;
}
