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

struct McGrid {
  const float* u;
  int X, Y, Z;
  float iso;
  long long n;          // X*Y*Z
  long long n_chunks;
  const signed char* tri;        // [256][3*MC_MAX_TRIS] case table (device memory in the kernels)
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
  const unsigned int pu = (unsigned int)p;
  if (g.sh_z >= 0 && g.sh_yz >= 0) {          // power-of-two Z and Y*Z (the 512^3 validate_mesh grid): shifts and masks
    q.k = (int)(pu & (unsigned int)(g.Z - 1));
    q.j = (int)((pu & (unsigned int)(YZ - 1)) >> g.sh_z);
    q.i = (int)(pu >> g.sh_yz);
  } else {
    q.i = (int)(pu / (unsigned int)YZ);
    const int r = (int)(pu - (unsigned int)q.i * (unsigned int)YZ);
    q.j = r / g.Z;
    q.k = r - q.j * g.Z;
  }
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

// triangles of the cell whose lowest corner is point p; `t0` = index of the first one
MC_HD void mc_emit_triangles(const McGrid& g, long long p, const McPoint& q, long long t0, const int* vid3, int* tris) {
  const long long YZ = (long long)g.Y * g.Z;
  const signed char* row = g.tri + q.cubecase * (3 * MC_MAX_TRIS);
  for (int t = 0; t < q.ntri; ++t) {
    for (int c = 0; c < 3; ++c) {
      const int e = row[3 * t + c];
      // cell edge e starts at point + origin(e) and runs along axis(e) (mc_tables.py EDGE_ORIGIN / EDGE_AXIS):
      //   e:      0   1   2   3   4   5   6   7   8   9   10  11
      //   axis:   x   y   x   y   x   y   x   y   z   z   z   z
      //   origin: 000 100 010 000 001 101 011 001 000 100 110 010
      const int axis = e >= 8 ? 2 : (e & 1);
      const int ox = (e == 1 || e == 5 || e == 9 || e == 10) ? 1 : 0;
      const int oy = (e == 2 || e == 6 || e == 10 || e == 11) ? 1 : 0;
      const int oz = (e >= 4 && e <= 7) ? 1 : 0;
      const long long pe = p + ox * YZ + (long long)oy * g.Z + oz;
      tris[(t0 + t) * 3 + c] = vid3[pe * 3 + axis];
    }
  }
}

}  // namespace fmov
