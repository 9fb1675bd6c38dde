// Per-grid-point logic of the marching-cubes kernels (csrc/marching_cubes.cu), written as plain host/device inline
// functions so that the very same code is also compiled by g++ and checked against the numpy oracle on the CPU
// (tests/test_marching_cubes_host_emulation.py).  No CUDA intrinsics in here.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace fmov {

constexpr int MC_CHUNK = 256;          // grid points per chunk = threads per block
constexpr int MC_MAX_TRIS = 5;
constexpr int MC_TRI_STRIDE = 16;       // bytes per case in McGrid::tri: 15 cell-edge numbers + 1 pad (one 16-byte load)
constexpr int MC_GROUP_SHIFT = 12;      // the offsets scan works on groups of 4096 chunks (fmov_mc_scan)

struct McGrid {
  const float* u;
  int X, Y, Z;
  float iso;
  long long n;          // X*Y*Z
  long long n_chunks;
  const signed char* tri;        // [256][MC_TRI_STRIDE] case table, 16-byte aligned (device memory in the kernels)
  const unsigned char* ntri;     // [256]
  int sh_z, sh_yz;               // log2(Z), log2(Y*Z) when those are powers of two, else -1 (mc_set_shifts)
};
MC_HD int mc_log2_or_neg(long long v) {
  int s = 0;
  while ((1LL << s) < v) ++s;
  return ((1LL << s) == v) ? s : -1;
}
MC_HD void mc_set_shifts(McGrid& g) {
  g.sh_z = mc_log2_or_neg(g.Z);
  g.sh_yz = mc_log2_or_neg((long long)g.Y * g.Z);
}
struct McXform { float s[3], o[3]; };     // output coordinate = index * s + o

struct McPoint {
  int i, j, k;
  bool valid;
  float f0;
  bool b0;
  bool cross[3];        // crossed grid edge starting here along x, y, z
  float f1[3];          // value at the other end of those edges
  int ntri;             // triangles of the cell whose lowest corner is this point
  int cubecase;
};

// linear index -> (i, j, k)
MC_HD void mc_split(const McGrid& g, unsigned int pu, int& i, int& j, int& k) {
  const int YZ = g.Y * g.Z;
  if (g.sh_z >= 0 && g.sh_yz >= 0) {          // power-of-two Z and Y*Z (the 512^3 validate_mesh grid): shifts and masks
    k = (int)(pu & (unsigned int)(g.Z - 1));
    j = (int)((pu & (unsigned int)(YZ - 1)) >> g.sh_z);
    i = (int)(pu >> g.sh_yz);
  } else {
    i = (int)(pu / (unsigned int)YZ);
    const int r = (int)(pu - (unsigned int)i * (unsigned int)YZ);
    j = r / g.Z;
    k = r - j * g.Z;
  }
}

// values and marks around point p (all loads are row-contiguous across the warp)
MC_HD McPoint mc_point(const McGrid& g, long long p, bool want_case) {
  McPoint q;
  q.valid = p < g.n;
  q.i = q.j = q.k = 0;
  q.f0 = 0.f; q.b0 = false; q.ntri = 0; q.cubecase = 0;
  q.cross[0] = q.cross[1] = q.cross[2] = false;
  q.f1[0] = q.f1[1] = q.f1[2] = 0.f;
  if (!q.valid) return q;
  // the grid has fewer than 2^31 points (checked by the launcher): 32-bit index arithmetic (64-bit division is ~5x the
  // instructions, and these kernels are issue-bound, not bandwidth-bound — ncu: 68 % issue-active at 0.7 TB/s)
  const int YZ = g.Y * g.Z;
  mc_split(g, (unsigned int)p, q.i, q.j, q.k);
  const bool hx = q.i + 1 < g.X, hy = q.j + 1 < g.Y, hz = q.k + 1 < g.Z;
  q.f0 = g.u[p];
  q.b0 = q.f0 < g.iso;
  if (hx) { q.f1[0] = g.u[p + YZ]; q.cross[0] = (q.f1[0] < g.iso) != q.b0; }
  if (hy) { q.f1[1] = g.u[p + g.Z]; q.cross[1] = (q.f1[1] < g.iso) != q.b0; }
  if (hz) { q.f1[2] = g.u[p + 1]; q.cross[2] = (q.f1[2] < g.iso) != q.b0; }
  if (want_case && hx && hy && hz) {
    // corners v0..v7 (mc_tables.py): (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
    int c = q.b0 ? 1 : 0;
    c |= (q.f1[0] < g.iso) ? 2 : 0;
    c |= (g.u[p + YZ + g.Z] < g.iso) ? 4 : 0;
    c |= (q.f1[1] < g.iso) ? 8 : 0;
    c |= (q.f1[2] < g.iso) ? 16 : 0;
    c |= (g.u[p + YZ + 1] < g.iso) ? 32 : 0;
    c |= (g.u[p + YZ + g.Z + 1] < g.iso) ? 64 : 0;
    c |= (g.u[p + g.Z + 1] < g.iso) ? 128 : 0;
    q.cubecase = c;
    q.ntri = g.ntri[c];
  }
  return q;
}
MC_HD int mc_vertex_count(const McPoint& q) { return (q.cross[0] ? 1 : 0) + (q.cross[1] ? 1 : 0) + (q.cross[2] ? 1 : 0); }

// ---- the count pass on QUADS: four z-consecutive points k0 .. k0+3 (k0 % 4 == 0, Z % 4 == 0: a quad never straddles a
// row and is one 16-byte load).  Everything is done on 5-bit "below the iso-value" masks of the four rows around the quad:
// bit t = u(row, k0 + t) < iso, t = 0..4 (bit 4 = first point of the next quad; absent at the row end).
MC_HD unsigned int mc_quad_mask(float a, float b, float c, float d, float e, float iso) {
  return (a < iso ? 1u : 0u) | (b < iso ? 2u : 0u) | (c < iso ? 4u : 0u) | (d < iso ? 8u : 0u) | (e < iso ? 16u : 0u);
}
MC_HD int mc_popc4(unsigned int m) { return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u)); }
// cells of the quad (bit t) whose eight corners do not all agree: bits t and t+1 neither all clear nor all set in the four rows
MC_HD unsigned int mc_quad_mixed(unsigned int m00, unsigned int m01, unsigned int m10, unsigned int m11, unsigned int zmask) {
  const unsigned int any = m00 | m01 | m10 | m11, all = m00 & m01 & m10 & m11;
  return ~((~any & ~(any >> 1)) | (all & (all >> 1))) & zmask;
}
// case of cell t; corners v0..v7 as in mc_point: (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
MC_HD int mc_quad_case(unsigned int m00, unsigned int m01, unsigned int m10, unsigned int m11, int t) {
  return (int)(((m00 >> t) & 1u) | (((m10 >> t) & 1u) << 1) | (((m11 >> t) & 1u) << 2) | (((m01 >> t) & 1u) << 3) |
               (((m00 >> (t + 1)) & 1u) << 4) | (((m10 >> (t + 1)) & 1u) << 5) | (((m11 >> (t + 1)) & 1u) << 6) |
               (((m01 >> (t + 1)) & 1u) << 7));
}
// crossed edges starting at the quad's points | triangles of their cells << 16: the sums of mc_vertex_count / ntri that
// mc_point(.., true) gives for the four points.  m00 m01 m10 m11 = masks of rows (i,j) (i,j+1) (i+1,j) (i+1,j+1);
// hx / hy: i + 1 < X, j + 1 < Y (masks of absent rows are ignored); hz4: k0 + 4 < Z (bit 4 present).
MC_HD int mc_quad_counts(const McGrid& g, unsigned int m00, unsigned int m01, unsigned int m10, unsigned int m11, bool hx,
                         bool hy, bool hz4) {
  const unsigned int zmask = hz4 ? 0xFu : 0x7u;          // points whose +z neighbour exists
  int nv = mc_popc4((m00 ^ (m00 >> 1)) & zmask);
  if (hy) nv += mc_popc4(m00 ^ m01);
  if (hx) nv += mc_popc4(m00 ^ m10);
  int nt = 0;
  if (hx && hy) {
    const unsigned int mixed = mc_quad_mixed(m00, m01, m10, m11, zmask);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < 4; ++t) {
      if (!((mixed >> t) & 1u)) continue;
      nt += g.ntri[mc_quad_case(m00, m01, m10, m11, t)];
    }
  }
  return nv | (nt << 16);
}
// the loads of one quad, all issued before anything looks at them: p = linear index of its first point (a multiple of 4)
struct McQuadRows {
  float v[4][5];          // rows (i,j) (i,j+1) (i+1,j) (i+1,j+1) at k0 .. k0+4; an absent row repeats row (i,j)
  bool hx, hy, hz4;
};
MC_HD McQuadRows mc_quad_load(const McGrid& g, unsigned int p) {
  McQuadRows q;
  int i, j, k0;
  mc_split(g, p, i, j, k0);
  const int YZ = g.Y * g.Z;
  q.hx = i + 1 < g.X; q.hy = j + 1 < g.Y; q.hz4 = k0 + 4 < g.Z;
  // an absent row never counts (hx / hy in mc_quad_counts): it reads row (i,j) again, so there is no branch between the loads
  const unsigned int off[4] = {0u, q.hy ? (unsigned int)g.Z : 0u, q.hx ? (unsigned int)YZ : 0u,
                               (q.hx && q.hy) ? (unsigned int)(YZ + g.Z) : 0u};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 4; ++r) {
    const float* row = g.u + (p + off[r]);
#if defined(__CUDA_ARCH__)
    const float4 f = *reinterpret_cast<const float4*>(row);
    q.v[r][0] = f.x; q.v[r][1] = f.y; q.v[r][2] = f.z; q.v[r][3] = f.w;
#else
    for (int t = 0; t < 4; ++t) q.v[r][t] = row[t];
#endif
    q.v[r][4] = q.hz4 ? row[4] : 0.f;
  }
  return q;
}
MC_HD void mc_quad_masks(const McGrid& g, const McQuadRows& q, unsigned int m[4]) {
  const unsigned int full = q.hz4 ? 31u : 15u;          // bit 4 only where that point exists
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 4; ++r) m[r] = mc_quad_mask(q.v[r][0], q.v[r][1], q.v[r][2], q.v[r][3], q.v[r][4], g.iso) & full;
}
MC_HD int mc_quad_eval_flags(const McGrid& g, unsigned int m00, unsigned int m01, unsigned int m10, unsigned int m11, bool hx,
                             bool hy, bool hz4) {
  // nearly every quad is far from the surface: all of its (up to) 20 values on one side
  const unsigned int full = hz4 ? 31u : 15u;
  if ((m00 | m01 | m10 | m11) == 0u || (m00 & m01 & m10 & m11) == full) return 0;
  return mc_quad_counts(g, m00, m01, m10, m11, hx, hy, hz4);
}
MC_HD int mc_quad_eval_masks(const McGrid& g, const McQuadRows& q, const unsigned int m[4]) {
  return mc_quad_eval_flags(g, m[0], m[1], m[2], m[3], q.hx, q.hy, q.hz4);
}
MC_HD int mc_quad_eval(const McGrid& g, const McQuadRows& q) {
  unsigned int m[4];
  mc_quad_masks(g, q, m);
  return mc_quad_eval_masks(g, q, m);
}
MC_HD int mc_quad(const McGrid& g, unsigned int p) { return mc_quad_eval(g, mc_quad_load(g, p)); }
// ---- the count pass MARCHING along x (planes are whole numbers of chunks: Y * Z % 256 == 0): a warp keeps one chunk column,
// the masks of plane i + 1 computed at step i are the masks of plane i at step i + 1, so every value is loaded and compared
// once (the quad kernel above loads and compares each row four times: as row (i,j), (i,j+1), (i+1,j) and (i+1,j+1)).
struct McRow5 { float v[5]; };
MC_HD McRow5 mc_row_load(const McGrid& g, unsigned int p, bool hz4) {          // p = linear index of the quad (multiple of 4)
  McRow5 r;
  const float* row = g.u + p;
#if defined(__CUDA_ARCH__)
  const float4 f = *reinterpret_cast<const float4*>(row);
  r.v[0] = f.x; r.v[1] = f.y; r.v[2] = f.z; r.v[3] = f.w;
#else
  for (int t = 0; t < 4; ++t) r.v[t] = row[t];
#endif
  r.v[4] = hz4 ? row[4] : 0.f;
  return r;
}
MC_HD unsigned int mc_row_mask(const McGrid& g, const McRow5& r, bool hz4) {
  return mc_quad_mask(r.v[0], r.v[1], r.v[2], r.v[3], r.v[4], g.iso) & (hz4 ? 31u : 15u);
}

// the quad path needs whole quads per row and 16-byte loads
MC_HD bool mc_quads_ok(const McGrid& g) { return (g.Z & 3) == 0 && (reinterpret_cast<unsigned long long>(g.u) & 15ull) == 0; }
MC_HD bool mc_march_ok(const McGrid& g) { return mc_quads_ok(g) && ((long long)g.Y * g.Z) % MC_CHUNK == 0; }

// vertices of the crossed edges that start at point p; `id` = id of the first one
MC_HD void mc_emit_vertices(const McGrid& g, const McXform& xf, long long p, const McPoint& q, long long id, float* verts,
                            int* vid3) {
  const float base[3] = {(float)q.i, (float)q.j, (float)q.k};
  for (int a = 0; a < 3; ++a) {
    if (!q.cross[a]) continue;
    // PyMCubes: x1 + (x2 - x1) * (iso - f1) / (f2 - f1) with x2 - x1 = 1 grid step
    const float t = (g.iso - q.f0) / (q.f1[a] - q.f0);
    float v[3] = {base[0], base[1], base[2]};
    v[a] += t;
    verts[id * 3 + 0] = fmaf(v[0], xf.s[0], xf.o[0]);
    verts[id * 3 + 1] = fmaf(v[1], xf.s[1], xf.o[1]);
    verts[id * 3 + 2] = fmaf(v[2], xf.s[2], xf.o[2]);
    vid3[p * 3 + a] = (int)id;
    ++id;
  }
}

// cell edge e starts at point + origin(e) and runs along axis(e) (mc_tables.py EDGE_ORIGIN / EDGE_AXIS):
//   e:      0   1   2   3   4   5   6   7   8   9   10  11
//   axis:   x   y   x   y   x   y   x   y   z   z   z   z
//   origin: 000 100 010 000 001 101 011 001 000 100 110 010
// packed 5 bits per edge: axis | ox << 2 | oy << 3 | oz << 4
MC_HD unsigned int mc_edge_code(int e) {
  const unsigned long long lut = (0ull << 0) | (5ull << 5) | (8ull << 10) | (1ull << 15) | (16ull << 20) | (21ull << 25) |
                                 (24ull << 30) | (17ull << 35) | (2ull << 40) | (6ull << 45) | (14ull << 50) | (10ull << 55);
  return (unsigned int)(lut >> (5 * e)) & 31u;
}
// where the vertex id of corner k (0 .. 3 * ntri - 1, case-table order) of the cell at point p lives in vid3
MC_HD long long mc_corner_slot(const McGrid& g, long long p, int cubecase, int k) {
  const unsigned int code = mc_edge_code(g.tri[cubecase * MC_TRI_STRIDE + k]);
  const long long pe = p + (long long)((code >> 2) & 1u) * ((long long)g.Y * g.Z) + (long long)((code >> 3) & 1u) * g.Z +
                       ((code >> 4) & 1u);
  return pe * 3 + (code & 3u);
}
// triangles of the cell whose lowest corner is point p; `t0` = index of the first one.  Batches of three triangles (most
// cells have no more), reads first and stores after: the vertex-id reads are scattered 4-byte reads, one DRAM latency each
MC_HD void mc_emit_triangles(const McGrid& g, long long p, const McPoint& q, long long t0, const int* vid3, int* tris) {
  const int nc = 3 * q.ntri;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k0 = 0; k0 < nc; k0 += 9) {
    int id[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 9; ++k)
      if (k0 + k < nc) id[k] = vid3[mc_corner_slot(g, p, q.cubecase, k0 + k)];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 9; ++k)
      if (k0 + k < nc) tris[t0 * 3 + k0 + k] = id[k];
  }
}

// ---- the emit passes on quads (same order and ids as the per-point functions above) ----
// vertices of the crossed edges starting at the quad's four points, in (point, axis) order; `id` = id of the first one
MC_HD void mc_quad_emit_vertices(const McGrid& g, const McXform& xf, unsigned int p, const McQuadRows& q,
                                 const unsigned int m[4], long long id, float* verts, int* vid3) {
  int i, j, k0;
  mc_split(g, p, i, j, k0);
  const unsigned int zmask = q.hz4 ? 0xFu : 0x7u;
  const unsigned int cx = q.hx ? ((m[0] ^ m[2]) & 0xFu) : 0u, cy = q.hy ? ((m[0] ^ m[1]) & 0xFu) : 0u,
                     cz = (m[0] ^ (m[0] >> 1)) & zmask;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int t = 0; t < 4; ++t) {
    if (!(((cx | cy | cz) >> t) & 1u)) continue;
    const float f0 = q.v[0][t];
    const float f1[3] = {q.v[2][t], q.v[1][t], q.v[0][t + 1]};          // other end along x, y, z
    const bool cr[3] = {((cx >> t) & 1u) != 0u, ((cy >> t) & 1u) != 0u, ((cz >> t) & 1u) != 0u};
    const float base[3] = {(float)i, (float)j, (float)(k0 + t)};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int a = 0; a < 3; ++a) {
      if (!cr[a]) continue;
      const float w = (g.iso - f0) / (f1[a] - f0);          // as mc_emit_vertices
      float v[3] = {base[0], base[1], base[2]};
      v[a] += w;
      verts[id * 3 + 0] = fmaf(v[0], xf.s[0], xf.o[0]);
      verts[id * 3 + 1] = fmaf(v[1], xf.s[1], xf.o[1]);
      verts[id * 3 + 2] = fmaf(v[2], xf.s[2], xf.o[2]);
      vid3[((long long)p + t) * 3 + a] = (int)id;
      ++id;
    }
  }
}
// triangles of the quad's four cells; `t0` = index of the first one.  This in-order walk is what the host stand-in runs; the
// device hands the same (cell, corner) items to the lanes of a warp (marching_cubes.cu: mc_warp_emit_triangles) — same
// mc_quad_mixed / mc_quad_case / mc_corner_slot, same output positions.
MC_HD void mc_quad_emit_triangles(const McGrid& g, unsigned int p, const McQuadRows& q, const unsigned int m[4], long long t0,
                                  const int* vid3, int* tris) {
  if (!(q.hx && q.hy)) return;
  const unsigned int mixed = mc_quad_mixed(m[0], m[1], m[2], m[3], q.hz4 ? 0xFu : 0x7u);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int t = 0; t < 4; ++t) {
    if (!((mixed >> t) & 1u)) continue;
    McPoint c;
    c.cubecase = mc_quad_case(m[0], m[1], m[2], m[3], t);
    c.ntri = g.ntri[c.cubecase];
    mc_emit_triangles(g, (long long)p + t, c, t0, vid3, tris);
    t0 += c.ntri;
  }
}

}  // namespace fmov
