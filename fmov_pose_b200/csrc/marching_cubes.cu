// Marching cubes on the dense u = -sdf grid (the step after extract_fields in validate_mesh: models/renderer.py:43,
// `mcubes.marching_cubes(u, threshold)` of PyMCubes 0.1.4 — CPU, third-party).  Produces the same INDEXED mesh: one
// vertex per crossed grid edge (shared by the cells around it), triangles from a 256-case table over the 12 cell edges.
//
// Three passes over the grid (z fastest => coalesced rows; the +Z / +YZ neighbour rows come from L1/L2, so each pass reads u
// once from HBM: 4 B per point) and a scan between them:
//   fmov_mc_count      per 256-point chunk: number of crossed edges starting at its points and of triangles of its cells.
//                      Three kernels, the first that applies: mc_count_march_kernel (Z % 4 == 0 and Y * Z % 256 == 0, every
//                      validate_mesh grid: warp per chunk column marching along x on 16-byte quads and bit masks, each value
//                      loaded once), mc_count_quad_kernel (Z % 4 == 0: warp per chunk, quads), mc_count_kernel (thread per point)
//   fmov_mc_scan       exclusive 64-bit prefix sums over the chunks, totals appended; the totals size the outputs
//   fmov_mc_vertices   vertex positions (index or world coordinates) + the vertex id of every crossed edge (vid3)
//   fmov_mc_triangles  case lookup, three vid3 reads per triangle corner
//   (both walk the list of non-empty chunks that fmov_mc_count appended to)
// Output order is deterministic: vertices by (grid point x-major, axis), triangles by (cell x-major, table order).
// Corner / edge numbering and the case table: fmov_pose_b200/mc_tables.py (uploaded once with fmov_mc_set_tables).
#include "fmov_common.cuh"
#include "mc_core.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

__device__ __align__(16) signed char d_mc_tri[256 * MC_TRI_STRIDE];       // case table, one 16-byte row per case, read through L1
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
      // the emit passes walk this list (a quarter of the chunks at 512^3) instead of testing every chunk; its order is
      // whatever the atomics give, the OUTPUT positions come from the prefix sums and do not depend on it
      if (tot != 0) list[atomicAdd(n_list, 1)] = (int)ch;
    }
  }
}

// The same counts on quads (mc_core.cuh: Z % 4 == 0): one WARP per chunk, lane = two quads (16-byte loads of the four rows
// around them), chunk totals by one warp reduction — no shared memory, no block barrier, a quarter of the load
// instructions and ~1/6 of the issue slots of the per-point kernel (which was issue-bound at 0.7 TB/s).
__global__ void __launch_bounds__(MC_CHUNK) mc_count_quad_kernel(const McGrid g, int* __restrict__ chunk_nv,
                                                                 int* __restrict__ chunk_nt, int* __restrict__ list,
                                                                 int* __restrict__ n_list) {
  const int lane = threadIdx.x & 31;
  const long long n_warps = (long long)gridDim.x * (MC_CHUNK / 32);
  // non-empty chunks are appended to the list 32 at a time (lane k keeps the k-th one found): one atomic per 32 chunks
  // instead of ~10^5 atomics on one address
  int pending = 0, mine = 0;
  for (long long ch = (long long)blockIdx.x * (MC_CHUNK / 32) + (threadIdx.x >> 5); ch < g.n_chunks; ch += n_warps) {
    // both halves' loads (8 x 16 B + 8 x 4 B per lane) are in flight before the first comparison; n is a multiple of 4, so a
    // quad is inside the grid or not at all (then it re-reads quad 0 and counts nothing)
    const long long p0 = ch * MC_CHUNK + lane * 4, p1 = p0 + MC_CHUNK / 2;
    const bool in0 = p0 < g.n, in1 = p1 < g.n;
    const McQuadRows q0 = mc_quad_load(g, in0 ? (unsigned int)p0 : 0u), q1 = mc_quad_load(g, in1 ? (unsigned int)p1 : 0u);
    int packed = in0 ? mc_quad_eval(g, q0) : 0;
    packed += in1 ? mc_quad_eval(g, q1) : 0;
    const int tot = (int)__reduce_add_sync(0xffffffffu, (unsigned int)packed);          // <= 768 | 1280 << 16 per chunk
    if (lane == 0) {
      chunk_nv[ch] = tot & 0xFFFF;
      chunk_nt[ch] = tot >> 16;
    }
    if (tot != 0) {          // warp-uniform
      if (lane == pending) mine = (int)ch;
      if (++pending == 32) {
        int at = 0;
        if (lane == 0) at = atomicAdd(n_list, 32);
        at = __shfl_sync(0xffffffffu, at, 0);
        list[at + lane] = mine;
        pending = 0;
      }
    }
  }
  if (pending) {
    int at = 0;
    if (lane == 0) at = atomicAdd(n_list, pending);
    at = __shfl_sync(0xffffffffu, at, 0);
    if (lane < pending) list[at + lane] = mine;
  }
}

// The count pass marching along x (mc_core.cuh: Y * Z % 256 == 0, every validate_mesh grid): block = 8 adjacent chunk columns
// (their j + 1 rows are each other's rows: L1), a contiguous range of (column group, i) items per block, i fastest.  Per
// step a lane loads 2 rows x 2 quads of plane i + 1 and carries their four masks into the next step.
__global__ void __launch_bounds__(MC_CHUNK) mc_count_march_kernel(const McGrid g, const int cpp, int* __restrict__ chunk_nv,
                                                                  int* __restrict__ chunk_nt, int* __restrict__ list,
                                                                  int* __restrict__ n_list) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int groups = (cpp + MC_CHUNK / 32 - 1) / (MC_CHUNK / 32);
  const int total = groups * g.X;          // < 2^31 / 256
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int r0 = (int)blockIdx.x * per, r1 = r0 + per < total ? r0 + per : total;
  if (r0 >= r1) return;
  const unsigned int YZ = (unsigned int)(g.Y * g.Z);
  int cg = r0 / g.X, i = r0 - cg * g.X;
  bool fresh = true;          // the masks of plane i have to be loaded (first step of the block / of a column)
  unsigned int q[2] = {0u, 0u}, zoff[2] = {0u, 0u}, m00[2] = {0u, 0u}, m01[2] = {0u, 0u};
  bool hy[2] = {false, false}, hz4[2] = {false, false};
  int pending = 0, mine = 0;
  for (int r = r0; r < r1; ++r) {
    const int col = cg * (MC_CHUNK / 32) + wi;
    if (col < cpp) {          // warp-uniform
      if (fresh) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          q[h] = (unsigned int)col * MC_CHUNK + h * (MC_CHUNK / 2) + lane * 4;          // offset of the quad in its plane
          int pi, pj, pk;
          mc_split(g, q[h], pi, pj, pk);          // pi == 0
          hy[h] = pj + 1 < g.Y;
          hz4[h] = pk + 4 < g.Z;
          zoff[h] = hy[h] ? (unsigned int)g.Z : 0u;          // an absent row j + 1 repeats row j (never counted: hy)
        }
        const unsigned int base = (unsigned int)i * YZ;
        const McRow5 a0 = mc_row_load(g, base + q[0], hz4[0]), b0 = mc_row_load(g, base + q[0] + zoff[0], hz4[0]);
        const McRow5 a1 = mc_row_load(g, base + q[1], hz4[1]), b1 = mc_row_load(g, base + q[1] + zoff[1], hz4[1]);
        m00[0] = mc_row_mask(g, a0, hz4[0]); m01[0] = mc_row_mask(g, b0, hz4[0]);
        m00[1] = mc_row_mask(g, a1, hz4[1]); m01[1] = mc_row_mask(g, b1, hz4[1]);
        fresh = false;
      }
      const bool hx = i + 1 < g.X;
      const unsigned int next = (unsigned int)(hx ? i + 1 : i) * YZ;          // no plane i + 1: plane i again (never counted: hx)
      const McRow5 a0 = mc_row_load(g, next + q[0], hz4[0]), b0 = mc_row_load(g, next + q[0] + zoff[0], hz4[0]);
      const McRow5 a1 = mc_row_load(g, next + q[1], hz4[1]), b1 = mc_row_load(g, next + q[1] + zoff[1], hz4[1]);
      const unsigned int m10[2] = {mc_row_mask(g, a0, hz4[0]), mc_row_mask(g, a1, hz4[1])};
      const unsigned int m11[2] = {mc_row_mask(g, b0, hz4[0]), mc_row_mask(g, b1, hz4[1])};
      const int packed = mc_quad_eval_flags(g, m00[0], m01[0], m10[0], m11[0], hx, hy[0], hz4[0]) +
                         mc_quad_eval_flags(g, m00[1], m01[1], m10[1], m11[1], hx, hy[1], hz4[1]);
      const int tot = (int)__reduce_add_sync(0xffffffffu, (unsigned int)packed);
      const int ch = i * cpp + col;
      if (lane == 0) {
        chunk_nv[ch] = tot & 0xFFFF;
        chunk_nt[ch] = tot >> 16;
      }
      if (tot != 0) {          // warp-uniform; appended 32 at a time as in mc_count_quad_kernel
        if (lane == pending) mine = ch;
        if (++pending == 32) {
          int at = 0;
          if (lane == 0) at = atomicAdd(n_list, 32);
          at = __shfl_sync(0xffffffffu, at, 0);
          list[at + lane] = mine;
          pending = 0;
        }
      }
      m00[0] = m10[0]; m00[1] = m10[1]; m01[0] = m11[0]; m01[1] = m11[1];
    }
    if (++i == g.X) { i = 0; ++cg; fresh = true; }
  }
  if (pending) {
    int at = 0;
    if (lane == 0) at = atomicAdd(n_list, pending);
    at = __shfl_sync(0xffffffffu, at, 0);
    if (lane < pending) list[at + lane] = mine;
  }
}

// Sums of the chunk counts per group of 4096 chunks (vertices | triangles << 32): the starting points of the scan's blocks.
constexpr int MC_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(MC_SCAN_THREADS) mc_group_sums_kernel(const int* __restrict__ chunk_nv,
                                                                        const int* __restrict__ chunk_nt, long long n_chunks,
                                                                        unsigned long long* __restrict__ group_sums) {
  __shared__ unsigned long long warp_sum[MC_SCAN_THREADS / 32];
  const long long first = (long long)blockIdx.x << MC_GROUP_SHIFT;
  unsigned long long s = 0;
  for (int r = 0; r < (1 << MC_GROUP_SHIFT) / MC_SCAN_THREADS; ++r) {
    const long long c = first + r * MC_SCAN_THREADS + threadIdx.x;
    if (c < n_chunks) s += (unsigned long long)chunk_nv[c] | ((unsigned long long)chunk_nt[c] << 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < MC_SCAN_THREADS / 32; ++w) t += warp_sum[w];
    group_sums[blockIdx.x] = t;          // both fields < 2^23
  }
}

// Exclusive prefix sums of the chunk counts (64-bit, totals appended): block b adds up the groups before its own and scans
// its 4096 chunks in four rounds of 1024.
__global__ void __launch_bounds__(MC_SCAN_THREADS) mc_scan_kernel(const int* __restrict__ chunk_nv,
                                                                  const int* __restrict__ chunk_nt,
                                                                  const unsigned long long* __restrict__ group_sums,
                                                                  long long n_chunks, long long* __restrict__ voff,
                                                                  long long* __restrict__ toff, long long* __restrict__ totals) {
  __shared__ unsigned long long warp_sum[MC_SCAN_THREADS / 32];
  __shared__ long long base_v, base_t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // groups before this block (32 + 32 bits would overflow over many groups: two 64-bit sums)
  long long sv = 0, st = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += MC_SCAN_THREADS) {
    const unsigned long long s = group_sums[b];
    sv += (long long)(s & 0xFFFFFFFFull);
    st += (long long)(s >> 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
    st += __shfl_xor_sync(0xffffffffu, st, o);
  }
  if (threadIdx.x == 0) { base_v = 0; base_t = 0; }
  __syncthreads();
  if (lane == 0 && (sv | st)) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&base_v), (unsigned long long)sv);
    atomicAdd(reinterpret_cast<unsigned long long*>(&base_t), (unsigned long long)st);
  }
  __syncthreads();
  long long run_v = base_v, run_t = base_t;
  const long long first = (long long)blockIdx.x << MC_GROUP_SHIFT;
  for (int r = 0; r < (1 << MC_GROUP_SHIFT) / MC_SCAN_THREADS; ++r) {
    const long long c = first + r * MC_SCAN_THREADS + threadIdx.x;
    const bool in = c < n_chunks;
    // within a group both sums stay below 2^23: one packed 64-bit scan
    const unsigned long long v = in ? ((unsigned long long)chunk_nv[c] | ((unsigned long long)chunk_nt[c] << 32)) : 0ull;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();                       // warp_sum of the previous round is no longer read
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < MC_SCAN_THREADS / 32; ++w) {
      const unsigned long long s = warp_sum[w];
      if (w < warp) before += s;
      total += s;
    }
    const unsigned long long ex = before + inc - v;
    if (in) {
      voff[c] = run_v + (long long)(ex & 0xFFFFFFFFull);
      toff[c] = run_t + (long long)(ex >> 32);
      if (c == n_chunks - 1) {
        const long long tv = run_v + (long long)((ex + v) & 0xFFFFFFFFull), tt = run_t + (long long)((ex + v) >> 32);
        voff[n_chunks] = tv; toff[n_chunks] = tt;
        totals[0] = tv; totals[1] = tt;
      }
    }
    run_v += (long long)(total & 0xFFFFFFFFull);
    run_t += (long long)(total >> 32);
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

// The emit passes on quads (Z % 4 == 0): a WARP per listed chunk, the quad loads and masks of the count pass again, ids from
// one packed warp scan (first half | second half << 16), and only the lanes whose quads cross the surface emit.
struct McWarpChunk {
  McQuadRows q0, q1;
  unsigned int m0[4], m1[4];
  unsigned int p0, p1;
  int c0, c1;          // packed counts (vertices | triangles << 16) of this lane's two quads
};
__device__ __forceinline__ void mc_warp_chunk(const McGrid& g, long long ch, int lane, McWarpChunk& w) {
  const long long p0 = ch * MC_CHUNK + lane * 4, p1 = p0 + MC_CHUNK / 2;
  const bool in0 = p0 < g.n, in1 = p1 < g.n;
  w.p0 = in0 ? (unsigned int)p0 : 0u;
  w.p1 = in1 ? (unsigned int)p1 : 0u;
  w.q0 = mc_quad_load(g, w.p0);
  w.q1 = mc_quad_load(g, w.p1);
  mc_quad_masks(g, w.q0, w.m0);
  mc_quad_masks(g, w.q1, w.m1);
  w.c0 = in0 ? mc_quad_eval_masks(g, w.q0, w.m0) : 0;
  w.c1 = in1 ? mc_quad_eval_masks(g, w.q1, w.m1) : 0;
}
// exclusive offsets of this lane's two quads within the chunk (quads in point order: first half lanes 0..31, then second half)
__device__ __forceinline__ void mc_warp_offsets(int n0, int n1, int lane, int& ex0, int& ex1) {
  int inc = n0 | (n1 << 16);          // <= 640 per half: no carry between the fields
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const int tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xFFFF;
  ex0 = (inc & 0xFFFF) - n0;
  ex1 = tot0 + (inc >> 16) - n1;
}

__global__ void __launch_bounds__(MC_CHUNK) mc_vertices_quad_kernel(const McGrid g, const long long* __restrict__ chunk_voff,
                                                                    const int* __restrict__ list,
                                                                    const int* __restrict__ n_list, const McXform xf,
                                                                    float* __restrict__ verts, int* __restrict__ vid3) {
  const int n = *n_list, lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * (MC_CHUNK / 32);
  int li = blockIdx.x * (MC_CHUNK / 32) + (threadIdx.x >> 5);
  if (li >= n) return;
  long long ch = list[li];
  while (true) {
    // the next list entry, this chunk's offsets and its grid values are all requested before anything is used: one
    // memory latency per chunk in front of the emission instead of three dependent ones
    const int li_next = li + n_warps;
    const long long ch_next = li_next < n ? list[li_next] : 0;
    const long long base = chunk_voff[ch], end = chunk_voff[ch + 1];
    McWarpChunk w;
    mc_warp_chunk(g, ch, lane, w);
    if (end != base) {          // else triangles only (warp-uniform)
      const int n0 = w.c0 & 0xFFFF, n1 = w.c1 & 0xFFFF;
      int ex0, ex1;
      mc_warp_offsets(n0, n1, lane, ex0, ex1);
      if (n0) mc_quad_emit_vertices(g, xf, w.p0, w.q0, w.m0, base + ex0, verts, vid3);
      if (n1) mc_quad_emit_vertices(g, xf, w.p1, w.q1, w.m1, base + ex1, verts, vid3);
    }
    if (li_next >= n) break;
    li = li_next;
    ch = ch_next;
  }
}

// Triangles of ONE quad written by the whole warp: lane = (cell t = lane / 8, corner slot s = lane % 8 -> corners s and s + 8
// of the cell's <= 15), so the scattered vertex-id reads of a quad are one round of parallel loads instead of a divergent
// per-lane walk (the per-lane walk cost ~1100 issue slots per chunk, this ~100 per quad that has triangles).
// p = first point of the quad, mpack = its four 5-bit row masks (m00 | m01 << 5 | m10 << 10 | m11 << 15) | hz4 << 20,
// t0 = index of the quad's first triangle.  All arguments warp-uniform.
__device__ __forceinline__ void mc_warp_emit_triangles(const McGrid& g, int lane, unsigned int p, unsigned int mpack,
                                                       long long t0, const int* __restrict__ vid3, int* __restrict__ tris) {
  const unsigned int m00 = mpack & 31u, m01 = (mpack >> 5) & 31u, m10 = (mpack >> 10) & 31u, m11 = (mpack >> 15) & 31u;
  const unsigned int mixed = mc_quad_mixed(m00, m01, m10, m11, ((mpack >> 20) & 1u) ? 0xFu : 0x7u);
  const int t = lane >> 3, s = lane & 7;
  int c = 0, n = 0;
  if ((mixed >> t) & 1u) {
    c = mc_quad_case(m00, m01, m10, m11, t);
    n = g.ntri[c];
  }
  // triangles of the quad's cells before mine
  const int n0 = __shfl_sync(0xffffffffu, n, 0), n1 = __shfl_sync(0xffffffffu, n, 8), n2 = __shfl_sync(0xffffffffu, n, 16);
  const long long first = t0 + (t > 0 ? n0 : 0) + (t > 1 ? n1 : 0) + (t > 2 ? n2 : 0);
  const long long pc = (long long)p + t;
  int id0 = 0, id1 = 0;
  if (s < 3 * n) id0 = vid3[mc_corner_slot(g, pc, c, s)];
  if (s + 8 < 3 * n) id1 = vid3[mc_corner_slot(g, pc, c, s + 8)];
  if (s < 3 * n) tris[first * 3 + s] = id0;
  if (s + 8 < 3 * n) tris[first * 3 + s + 8] = id1;
}

__global__ void __launch_bounds__(MC_CHUNK, 3) mc_triangles_quad_kernel(const McGrid g, const long long* __restrict__ chunk_toff,
                                                                        const int* __restrict__ list,
                                                                        const int* __restrict__ n_list,
                                                                        const int* __restrict__ vid3, int* __restrict__ tris) {
  const int n = *n_list, lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * (MC_CHUNK / 32);
  int li = blockIdx.x * (MC_CHUNK / 32) + (threadIdx.x >> 5);
  if (li >= n) return;
  long long ch = list[li];
  while (true) {
    const int li_next = li + n_warps;          // as in mc_vertices_quad_kernel
    const long long ch_next = li_next < n ? list[li_next] : 0;
    const long long base = chunk_toff[ch], end = chunk_toff[ch + 1];
    McWarpChunk w;
    mc_warp_chunk(g, ch, lane, w);
    if (end != base) {          // else vertices only (warp-uniform)
      const int n0 = w.c0 >> 16, n1 = w.c1 >> 16;
      int ex0, ex1;
      mc_warp_offsets(n0, n1, lane, ex0, ex1);
      const unsigned int mp0 = w.m0[0] | (w.m0[1] << 5) | (w.m0[2] << 10) | (w.m0[3] << 15) | (w.q0.hz4 ? 1u << 20 : 0u);
      const unsigned int mp1 = w.m1[0] | (w.m1[1] << 5) | (w.m1[2] << 10) | (w.m1[3] << 15) | (w.q1.hz4 ? 1u << 20 : 0u);
      // quads with triangles, in point order (n > 0 implies hx && hy)
      for (unsigned int owners = __ballot_sync(0xffffffffu, n0 != 0); owners; owners &= owners - 1) {
        const int o = __ffs(owners) - 1;
        mc_warp_emit_triangles(g, lane, __shfl_sync(0xffffffffu, w.p0, o), __shfl_sync(0xffffffffu, mp0, o),
                               base + __shfl_sync(0xffffffffu, ex0, o), vid3, tris);
      }
      for (unsigned int owners = __ballot_sync(0xffffffffu, n1 != 0); owners; owners &= owners - 1) {
        const int o = __ffs(owners) - 1;
        mc_warp_emit_triangles(g, lane, __shfl_sync(0xffffffffu, w.p1, o), __shfl_sync(0xffffffffu, mp1, o),
                               base + __shfl_sync(0xffffffffu, ex1, o), vid3, tris);
      }
    }
    if (li_next >= n) break;
    li = li_next;
    ch = ch_next;
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
// grid of a grid-stride kernel = exactly the blocks that are resident at once: all warps then sweep the grid as ONE window of
// consecutive chunks, and the rows of plane i + 1 that a chunk reads are still in L2 when the window reaches them as plane i
// (with more blocks than fit, the late blocks re-read what the first wave had already fetched: 1.5x the DRAM traffic)
template <class K>
static int mc_resident_blocks(K kernel, long long max_useful) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MC_CHUNK, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long long want = (long long)sms * per_sm;
  return (int)(max_useful < want ? (max_useful < 1 ? 1 : max_useful) : want);
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
  signed char rows[256 * MC_TRI_STRIDE];
  for (int c = 0; c < 256; ++c)
    for (int k = 0; k < MC_TRI_STRIDE; ++k) rows[c * MC_TRI_STRIDE + k] = k < 15 ? tri_table[c * 15 + k] : (signed char)-1;
  FMOV_CUDA(cudaMemcpyToSymbol(d_mc_tri, rows, sizeof(rows)));
  FMOV_CUDA(cudaMemcpyToSymbol(d_mc_ntri, n_tris, 256));
  g_mc_tables_set = true;
  return OK;
}

extern "C" long long fmov_mc_chunk_count(int X, int Y, int Z) {
  return ((long long)X * Y * Z + MC_CHUNK - 1) / MC_CHUNK;
}

extern "C" long long fmov_mc_group_count(int X, int Y, int Z) {
  return (fmov_mc_chunk_count(X, Y, Z) + (1LL << MC_GROUP_SHIFT) - 1) >> MC_GROUP_SHIFT;
}

extern "C" int fmov_mc_count(const float* u, int X, int Y, int Z, float iso, int* chunk_nv, int* chunk_nt, int* list,
                             int* n_list, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_nv && chunk_nt && list && n_list, "fmov_mc_count: null output");
  FMOV_CUDA(cudaMemsetAsync(n_list, 0, sizeof(int), (cudaStream_t)stream));
  if (mc_march_ok(g)) {
    const int cpp = (int)(((long long)Y * Z) / MC_CHUNK);
    const long long items = (long long)((cpp + MC_CHUNK / 32 - 1) / (MC_CHUNK / 32)) * X;
    const int blocks = mc_resident_blocks(mc_count_march_kernel, items);
    mc_count_march_kernel<<<blocks, MC_CHUNK, 0, (cudaStream_t)stream>>>(g, cpp, chunk_nv, chunk_nt, list, n_list);
    FMOV_LAUNCH_CHECK("mc_count_march_kernel");
  } else if (mc_quads_ok(g)) {
    const int blocks = mc_resident_blocks(mc_count_quad_kernel, (g.n_chunks + MC_CHUNK / 32 - 1) / (MC_CHUNK / 32));
    mc_count_quad_kernel<<<blocks, MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_nv, chunk_nt, list, n_list);
    FMOV_LAUNCH_CHECK("mc_count_quad_kernel");
  } else {
    mc_count_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_nv, chunk_nt, list, n_list);
    FMOV_LAUNCH_CHECK("mc_count_kernel");
  }
  return OK;
}

extern "C" int fmov_mc_scan(const int* chunk_nv, const int* chunk_nt, long long n_chunks, unsigned long long* group_scratch,
                            long long* chunk_voff, long long* chunk_toff, long long* totals, void* stream) {
  FMOV_REQUIRE(chunk_nv && chunk_nt && group_scratch && chunk_voff && chunk_toff && totals, "fmov_mc_scan: null argument");
  FMOV_REQUIRE(n_chunks >= 1 && n_chunks <= (1LL << 31) / MC_CHUNK, "fmov_mc_scan: bad chunk count %lld", n_chunks);
  const int groups = (int)((n_chunks + (1LL << MC_GROUP_SHIFT) - 1) >> MC_GROUP_SHIFT);
  mc_group_sums_kernel<<<groups, MC_SCAN_THREADS, 0, (cudaStream_t)stream>>>(chunk_nv, chunk_nt, n_chunks, group_scratch);
  FMOV_LAUNCH_CHECK("mc_group_sums_kernel");
  mc_scan_kernel<<<groups, MC_SCAN_THREADS, 0, (cudaStream_t)stream>>>(chunk_nv, chunk_nt, group_scratch, n_chunks, chunk_voff,
                                                                        chunk_toff, totals);
  FMOV_LAUNCH_CHECK("mc_scan_kernel");
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
  if (mc_quads_ok(g)) {
    const int blocks = mc_resident_blocks(mc_vertices_quad_kernel, g.n_chunks);
    mc_vertices_quad_kernel<<<blocks, MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_voff, list, n_list, xf, verts, vid3);
    FMOV_LAUNCH_CHECK("mc_vertices_quad_kernel");
  } else {
    mc_vertices_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_voff, list, n_list, xf, verts, vid3);
    FMOV_LAUNCH_CHECK("mc_vertices_kernel");
  }
  return OK;
}

extern "C" int fmov_mc_triangles(const float* u, int X, int Y, int Z, float iso, const long long* chunk_toff, const int* list,
                                 const int* n_list, const int* vid3, int* tris, void* stream) {
  McGrid g;
  int st = mc_grid(g, u, X, Y, Z, iso);
  if (st) return st;
  FMOV_REQUIRE(chunk_toff && list && n_list && vid3 && tris, "fmov_mc_triangles: null argument");
  if (mc_quads_ok(g)) {
    const int blocks = mc_resident_blocks(mc_triangles_quad_kernel, g.n_chunks);
    mc_triangles_quad_kernel<<<blocks, MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_toff, list, n_list, vid3, tris);
    FMOV_LAUNCH_CHECK("mc_triangles_quad_kernel");
  } else {
    mc_triangles_kernel<<<mc_blocks(g), MC_CHUNK, 0, (cudaStream_t)stream>>>(g, chunk_toff, list, n_list, vid3, tris);
    FMOV_LAUNCH_CHECK("mc_triangles_kernel");
  }
  return OK;
}
