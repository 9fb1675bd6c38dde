// Host emulation of csrc/marching_cubes.cu: the per-point device functions of csrc/mc_core.cuh compiled by g++ and
// driven chunk by chunk exactly as the three kernels drive them (sequential "threads", prefix sums on the host).
// Used by tests/test_marching_cubes_host_emulation.py; test infrastructure, never shipped.
//   mc_host_emul X Y Z iso u.bin tables.bin verts.bin tris.bin     (tables.bin = int8[256*15] + uint8[256])
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fmov_pose_b200/csrc/mc_core.cuh"

using namespace fmov;

template <class T>
static std::vector<T> slurp(const char* path, size_t n) {
  std::vector<T> v(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}
template <class T>
static void dump(const char* path, const std::vector<T>& v) {
  FILE* f = fopen(path, "wb");
  if (!f || (v.size() && fwrite(v.data(), sizeof(T), v.size(), f) != v.size())) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
  fclose(f);
}


// mc_count_march_kernel on the host: every chunk column marched along x with the masks of plane i carried; returns the
// packed counts (vertices | triangles << 16) of chunk (i, col) through `out[i * cpp + col]`
static void march_counts(const McGrid& g, std::vector<int>& out) {
  const int cpp = (int)(((long long)g.Y * g.Z) / MC_CHUNK);
  const unsigned int YZ = (unsigned int)(g.Y * g.Z);
  out.assign((size_t)cpp * g.X, 0);
  for (int col = 0; col < cpp; ++col)
    for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {          // lane x half
      const unsigned int q = (unsigned int)col * MC_CHUNK + q4 * 4;
      int pi, pj, pk;
      mc_split(g, q, pi, pj, pk);
      const bool hy = pj + 1 < g.Y, hz4 = pk + 4 < g.Z;
      const unsigned int zoff = hy ? (unsigned int)g.Z : 0u;
      unsigned int m00 = mc_row_mask(g, mc_row_load(g, q, hz4), hz4), m01 = mc_row_mask(g, mc_row_load(g, q + zoff, hz4), hz4);
      for (int i = 0; i < g.X; ++i) {
        const bool hx = i + 1 < g.X;
        const unsigned int next = (unsigned int)(hx ? i + 1 : i) * YZ;
        const unsigned int m10 = mc_row_mask(g, mc_row_load(g, next + q, hz4), hz4),
                           m11 = mc_row_mask(g, mc_row_load(g, next + q + zoff, hz4), hz4);
        out[(size_t)i * cpp + col] += mc_quad_eval_flags(g, m00, m01, m10, m11, hx, hy, hz4);
        m00 = m10; m01 = m11;
      }
    }
}

int main(int argc, char** argv) {
  if (argc != 9) return 2;
  McGrid g;
  g.X = atoi(argv[1]); g.Y = atoi(argv[2]); g.Z = atoi(argv[3]); g.iso = (float)atof(argv[4]);
  g.n = (long long)g.X * g.Y * g.Z;
  g.n_chunks = (g.n + MC_CHUNK - 1) / MC_CHUNK;
  mc_set_shifts(g);
  std::vector<float> u = slurp<float>(argv[5], (size_t)g.n);
  std::vector<signed char> tab = slurp<signed char>(argv[6], 256 * 15 + 256);
  g.u = u.data();
  std::vector<signed char> rows(256 * MC_TRI_STRIDE, -1);          // 16-byte rows like the device copy
  for (int c = 0; c < 256; ++c)
    for (int k = 0; k < 15; ++k) rows[c * MC_TRI_STRIDE + k] = tab[c * 15 + k];
  g.tri = rows.data();
  g.ntri = reinterpret_cast<const unsigned char*>(tab.data() + 256 * 15);
  std::vector<int> marched;
  if (mc_march_ok(g)) march_counts(g, marched);
  // pass 1: per-chunk counts (mc_count_kernel)
  std::vector<long long> voff(g.n_chunks + 1, 0), toff(g.n_chunks + 1, 0);
  for (long long ch = 0; ch < g.n_chunks; ++ch) {
    long long nv = 0, nt = 0;
    for (int tid = 0; tid < MC_CHUNK; ++tid) {
      const McPoint q = mc_point(g, ch * MC_CHUNK + tid, true);
      nv += mc_vertex_count(q);
      nt += q.ntri;
    }
    if (mc_quads_ok(g)) {          // mc_count_quad_kernel must count what the per-point code counts
      long long qv = 0, qt = 0;
      for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {
        const long long p = ch * MC_CHUNK + q4 * 4;
        if (p >= g.n) continue;
        const int packed = mc_quad(g, (unsigned int)p);
        qv += packed & 0xFFFF;
        qt += packed >> 16;
      }
      if (mc_march_ok(g) && ((marched[ch] & 0xFFFF) != nv || (marched[ch] >> 16) != nt)) {
        fprintf(stderr, "marched counts differ in chunk %lld\n", ch);
        return 4;
      }
      if (qv != nv || qt != nt) { fprintf(stderr, "quad counts differ in chunk %lld: %lld %lld vs %lld %lld\n", ch, qv, qt, nv, nt); return 3; }
    }
    voff[ch + 1] = voff[ch] + nv;
    toff[ch + 1] = toff[ch] + nt;
  }
  std::vector<float> verts((size_t)voff[g.n_chunks] * 3);
  std::vector<int> tris((size_t)toff[g.n_chunks] * 3, -7);
  std::vector<int> vid3((size_t)g.n * 3, -1);          // -1: an entry the triangle pass must never read
  McXform xf;
  for (int a = 0; a < 3; ++a) { xf.s[a] = 1.f; xf.o[a] = 0.f; }
  // pass 2: vertices (mc_vertices_kernel)
  for (long long ch = 0; ch < g.n_chunks; ++ch) {
    int local = 0;
    if (mc_quads_ok(g)) {          // mc_vertices_quad_kernel
      for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {
        const long long p = ch * MC_CHUNK + q4 * 4;
        if (p >= g.n) continue;
        const McQuadRows q = mc_quad_load(g, (unsigned int)p);
        unsigned int m[4];
        mc_quad_masks(g, q, m);
        const int nv = mc_quad_eval_masks(g, q, m) & 0xFFFF;
        if (nv) mc_quad_emit_vertices(g, xf, (unsigned int)p, q, m, voff[ch] + local, verts.data(), vid3.data());
        local += nv;
      }
      continue;
    }
    for (int tid = 0; tid < MC_CHUNK; ++tid) {
      const long long p = ch * MC_CHUNK + tid;
      const McPoint q = mc_point(g, p, false);
      const int nv = mc_vertex_count(q);
      if (nv) mc_emit_vertices(g, xf, p, q, voff[ch] + local, verts.data(), vid3.data());
      local += nv;
    }
  }
  // pass 3: triangles (mc_triangles_kernel)
  for (long long ch = 0; ch < g.n_chunks; ++ch) {
    int local = 0;
    if (mc_quads_ok(g)) {          // mc_triangles_quad_kernel
      for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {
        const long long p = ch * MC_CHUNK + q4 * 4;
        if (p >= g.n) continue;
        const McQuadRows q = mc_quad_load(g, (unsigned int)p);
        unsigned int m[4];
        mc_quad_masks(g, q, m);
        const int nt = mc_quad_eval_masks(g, q, m) >> 16;
        if (nt) mc_quad_emit_triangles(g, (unsigned int)p, q, m, toff[ch] + local, vid3.data(), tris.data());
        local += nt;
      }
      continue;
    }
    for (int tid = 0; tid < MC_CHUNK; ++tid) {
      const long long p = ch * MC_CHUNK + tid;
      const McPoint q = mc_point(g, p, true);
      if (q.ntri) mc_emit_triangles(g, p, q, toff[ch] + local, vid3.data(), tris.data());
      local += q.ntri;
    }
  }
  dump(argv[7], verts);
  dump(argv[8], tris);
  return 0;
}
