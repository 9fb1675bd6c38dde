// Host stand-in for the marching-cubes entry points of libfmov_b200.so (same C signatures, host pointers): the
// kernels' per-point code (csrc/mc_core.cuh) driven chunk by chunk.  Lets the Python glue of fmov_pose_b200/mcubes_gpu.py
// (argument marshalling, prefix sums, output sizing) run on the CPU in tests/test_marching_cubes_host_emulation.py.
// Test infrastructure, never shipped.
#include <cstring>
#include <vector>
#include "../../fmov_pose_b200/csrc/mc_core.cuh"

using namespace fmov;

static signed char g_tri[256 * MC_TRI_STRIDE];          // 16-byte rows like the device copy
static unsigned char g_ntri[256];
static bool g_set = false;

static bool grid(McGrid& g, const float* u, int X, int Y, int Z, float iso) {
  if (!u || X < 2 || Y < 2 || Z < 2 || !g_set) return false;
  g.u = u; g.X = X; g.Y = Y; g.Z = Z; g.iso = iso;
  g.n = (long long)X * Y * Z;
  g.n_chunks = (g.n + MC_CHUNK - 1) / MC_CHUNK;
  mc_set_shifts(g);
  g.tri = g_tri; g.ntri = g_ntri;
  return true;
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

extern "C" {
int fmov_mc_set_tables(const signed char* tri, const unsigned char* ntri) {
  if (!tri || !ntri) return -1;
  for (int c = 0; c < 256; ++c)
    for (int k = 0; k < MC_TRI_STRIDE; ++k) g_tri[c * MC_TRI_STRIDE + k] = k < 15 ? tri[c * 15 + k] : (signed char)-1;
  memcpy(g_ntri, ntri, sizeof(g_ntri));
  g_set = true;
  return 0;
}
long long fmov_mc_chunk_count(int X, int Y, int Z) { return ((long long)X * Y * Z + MC_CHUNK - 1) / MC_CHUNK; }
long long fmov_mc_group_count(int X, int Y, int Z) {
  return (fmov_mc_chunk_count(X, Y, Z) + (1LL << MC_GROUP_SHIFT) - 1) >> MC_GROUP_SHIFT;
}
int fmov_mc_count(const float* u, int X, int Y, int Z, float iso, int* chunk_nv, int* chunk_nt, int* list, int* n_list,
                  void*) {
  McGrid g;
  if (!grid(g, u, X, Y, Z, iso) || !chunk_nv || !chunk_nt || !list || !n_list) return -1;
  *n_list = 0;
  // the device appends in whatever order its atomics give: walk the chunks backwards here so that the glue and the emit
  // passes are exercised with a list that is NOT in chunk order
  std::vector<int> marched;
  if (mc_march_ok(g)) march_counts(g, marched);
  for (long long ch = g.n_chunks - 1; ch >= 0; --ch) {
    int nv = 0, nt = 0;
    if (mc_march_ok(g)) {          // mc_count_march_kernel
      nv = marched[ch] & 0xFFFF;
      nt = marched[ch] >> 16;
    } else if (mc_quads_ok(g)) {          // mc_count_quad_kernel: 32 lanes x 2 quads
      for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) {
          const long long p = ch * MC_CHUNK + h * (MC_CHUNK / 2) + lane * 4;
          if (p >= g.n) continue;
          const int packed = mc_quad(g, (unsigned int)p);
          nv += packed & 0xFFFF;
          nt += packed >> 16;
        }
    } else {
      for (int tid = 0; tid < MC_CHUNK; ++tid) {
        const McPoint q = mc_point(g, ch * MC_CHUNK + tid, true);
        nv += mc_vertex_count(q);
        nt += q.ntri;
      }
    }
    chunk_nv[ch] = nv;
    chunk_nt[ch] = nt;
    if (nv | nt) list[(*n_list)++] = (int)ch;
  }
  return 0;
}
int fmov_mc_scan(const int* chunk_nv, const int* chunk_nt, long long n_chunks, unsigned long long* group_sums,
                 long long* voff, long long* toff, long long* totals, void*) {
  if (!chunk_nv || !chunk_nt || !group_sums || !voff || !toff || !totals || n_chunks < 1) return -1;
  // like the device: sums per group of 4096 chunks first, then each group starts from the sums of the groups before it
  const long long groups = (n_chunks + (1LL << MC_GROUP_SHIFT) - 1) >> MC_GROUP_SHIFT;
  for (long long b = 0; b < groups; ++b) {
    group_sums[b] = 0;
    for (long long c = b << MC_GROUP_SHIFT; c < ((b + 1) << MC_GROUP_SHIFT) && c < n_chunks; ++c)
      group_sums[b] += (unsigned long long)chunk_nv[c] | ((unsigned long long)chunk_nt[c] << 32);
  }
  for (long long b = 0; b < groups; ++b) {
    long long v = 0, t = 0;
    for (long long a = 0; a < b; ++a) { v += (long long)(group_sums[a] & 0xFFFFFFFFull); t += (long long)(group_sums[a] >> 32); }
    const long long end = ((b + 1) << MC_GROUP_SHIFT) < n_chunks ? ((b + 1) << MC_GROUP_SHIFT) : n_chunks;
    for (long long c = b << MC_GROUP_SHIFT; c < end; ++c) {
      voff[c] = v; toff[c] = t;
      v += chunk_nv[c]; t += chunk_nt[c];
    }
    if (end == n_chunks) { voff[n_chunks] = v; toff[n_chunks] = t; totals[0] = v; totals[1] = t; }
  }
  return 0;
}
int fmov_mc_vertices(const float* u, int X, int Y, int Z, float iso, const long long* chunk_voff, const int* list,
                     const int* n_list, float sx, float sy, float sz, float ox, float oy, float oz, float* verts, int* vid3,
                     void*) {
  McGrid g;
  if (!grid(g, u, X, Y, Z, iso) || !chunk_voff || !list || !n_list || !verts || !vid3) return -1;
  McXform xf;
  xf.s[0] = sx; xf.s[1] = sy; xf.s[2] = sz; xf.o[0] = ox; xf.o[1] = oy; xf.o[2] = oz;
  for (int li = 0; li < *n_list; ++li) {
    const long long ch = list[li];
    int local = 0;
    if (mc_quads_ok(g)) {          // mc_vertices_quad_kernel: quads in point order, ids from the running sum
      for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {
        const long long p = ch * MC_CHUNK + q4 * 4;
        if (p >= g.n) continue;
        const McQuadRows q = mc_quad_load(g, (unsigned int)p);
        unsigned int m[4];
        mc_quad_masks(g, q, m);
        const int nv = mc_quad_eval_masks(g, q, m) & 0xFFFF;
        if (nv) mc_quad_emit_vertices(g, xf, (unsigned int)p, q, m, chunk_voff[ch] + local, verts, vid3);
        local += nv;
      }
      continue;
    }
    for (int tid = 0; tid < MC_CHUNK; ++tid) {
      const long long p = ch * MC_CHUNK + tid;
      const McPoint q = mc_point(g, p, false);
      const int nv = mc_vertex_count(q);
      if (nv) mc_emit_vertices(g, xf, p, q, chunk_voff[ch] + local, verts, vid3);
      local += nv;
    }
  }
  return 0;
}
int fmov_mc_triangles(const float* u, int X, int Y, int Z, float iso, const long long* chunk_toff, const int* list,
                      const int* n_list, const int* vid3, int* tris, void*) {
  McGrid g;
  if (!grid(g, u, X, Y, Z, iso) || !chunk_toff || !list || !n_list || !vid3 || !tris) return -1;
  for (int li = 0; li < *n_list; ++li) {
    const long long ch = list[li];
    int local = 0;
    if (mc_quads_ok(g)) {          // mc_triangles_quad_kernel
      for (int q4 = 0; q4 < MC_CHUNK / 4; ++q4) {
        const long long p = ch * MC_CHUNK + q4 * 4;
        if (p >= g.n) continue;
        const McQuadRows q = mc_quad_load(g, (unsigned int)p);
        unsigned int m[4];
        mc_quad_masks(g, q, m);
        const int nt = mc_quad_eval_masks(g, q, m) >> 16;
        if (nt) mc_quad_emit_triangles(g, (unsigned int)p, q, m, chunk_toff[ch] + local, vid3, tris);
        local += nt;
      }
      continue;
    }
    for (int tid = 0; tid < MC_CHUNK; ++tid) {
      const long long p = ch * MC_CHUNK + tid;
      const McPoint q = mc_point(g, p, true);
      if (q.ntri) mc_emit_triangles(g, p, q, chunk_toff[ch] + local, vid3, tris);
      local += q.ntri;
    }
  }
  return 0;
}
const char* fmov_last_error(void) { return "host stand-in"; }
}
