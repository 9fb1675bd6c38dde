// Compositing + losses, forward and backward (warp per ray, warp scans; HBM-bound).
//
//   fmov_composite_fwd / _bwd   NeuSRenderer.render_core tail: dists, mid-points, true_cos / iter_cos,
//                               prev/next CDF -> alpha -> transmittance cumprod -> weights, colour,
//                               depth, weight_sum/max, eikonal partial sums
//                               (models/renderer.py:261-272, 290-358, 477-498)
//   fmov_loss_fwd_bwd           masked L1 colour + BCE(mask) per-ray terms and their gradients
//                               (exp_runner.py:562-599, 772-779)
//   fmov_ray_reduce_bwd         per-sample point / view-dir gradients -> rays_o, rays_d, z gradients
//                               (autograd of pts = o + d*mid_z, renderer.py:269-272)
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

constexpr int CW = 4;          // warps (rays) per block
constexpr int CMAX = 8;        // samples per lane (S <= 256)

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wscan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float wscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

struct SampleTerms {   // everything alpha depends on, for one sample
  float dist, mid, tc, ic, prev_c, next_c, num, den, alpha_raw, alpha;
};
__device__ __forceinline__ SampleTerms sample_terms(float z0, float z1, bool last, float sample_dist, float sdf,
                                                    float nx, float ny, float nz, float dx, float dy, float dz,
                                                    float inv_s, float car) {
  SampleTerms t;
  t.dist = last ? sample_dist : (z1 - z0);
  t.mid = z0 + t.dist * 0.5f;
  t.tc = dx * nx + dy * ny + dz * nz;
  t.ic = -(fmaxf(-t.tc * 0.5f + 0.5f, 0.f) * (1.0f - car) + fmaxf(-t.tc, 0.f) * car);
  const float en = sdf + t.ic * t.dist * 0.5f, ep = sdf - t.ic * t.dist * 0.5f;
  t.prev_c = 1.0f / (1.0f + expf(-ep * inv_s));
  t.next_c = 1.0f / (1.0f + expf(-en * inv_s));
  t.num = t.prev_c - t.next_c + 1e-5f;
  t.den = t.prev_c + 1e-5f;
  t.alpha_raw = t.num / t.den;
  t.alpha = fminf(fmaxf(t.alpha_raw, 0.f), 1.f);
  return t;
}

struct CompArgs {
  long long B;
  int S;
  const float* rays_o; const float* rays_d; const float* z;          // [B,3],[B,3],[B,S]
  const float* sdf; const float* nrm; const float* rgb;              // [B*S],[B*S,3],[B*S,3]
  const float* inv_s;                                                 // device scalar (clipped)
  float sample_dist, cos_anneal;
  const float* bg;                                                    // [3] or null
  // outputs (fwd)
  float* color; float* wsum_; float* wmax_; float* depth;             // [B,3],[B],[B],[B]
  float* weights; float* cdf; float* inside; float* mid_z; float* pts;   // [B,S] x4, [B*S,3]
  float* eik;                                                         // [B,2] per-ray (num, den)
};

__global__ void __launch_bounds__(CW * 32) composite_fwd_kernel(CompArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * CW + warp;
  if (ray >= a.B) return;
  const int S = a.S;
  const int per = (S + 31) / 32;
  const int j0 = lane * per;
  const float ox = a.rays_o[ray * 3], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const float inv_s = *a.inv_s;
  const float* zr = a.z + ray * S;
  float alpha[CMAX], cr[CMAX], cg[CMAX], cb[CMAX], midz[CMAX];
  float prod = 1.f, e_num = 0.f, e_den = 0.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    const int j = j0 + k;
    alpha[k] = 0.f; cr[k] = cg[k] = cb[k] = 0.f; midz[k] = 0.f;
    if (k < per && j < S) {
      const long long p = ray * S + j;
      const float z0 = zr[j], z1 = (j + 1 < S) ? zr[j + 1] : 0.f;
      const float nx = a.nrm[p * 3], ny = a.nrm[p * 3 + 1], nz = a.nrm[p * 3 + 2];
      const SampleTerms t = sample_terms(z0, z1, j + 1 == S, a.sample_dist, a.sdf[p], nx, ny, nz, dx, dy, dz, inv_s,
                                         a.cos_anneal);
      alpha[k] = t.alpha;
      midz[k] = t.mid;
      cr[k] = a.rgb[p * 3]; cg[k] = a.rgb[p * 3 + 1]; cb[k] = a.rgb[p * 3 + 2];
      const float px = ox + dx * t.mid, py = oy + dy * t.mid, pz = oz + dz * t.mid;
      const float pn = sqrtf(px * px + py * py + pz * pz);
      const float gn = sqrtf(nx * nx + ny * ny + nz * nz);
      const float relax = pn < 1.2f ? 1.f : 0.f;
      e_num += relax * (gn - 1.f) * (gn - 1.f);
      e_den += relax;
      if (a.cdf) a.cdf[p] = t.prev_c;
      if (a.inside) a.inside[p] = pn < 1.0f ? 1.f : 0.f;
      if (a.mid_z) a.mid_z[p] = t.mid;
      if (a.pts) { a.pts[p * 3] = px; a.pts[p * 3 + 1] = py; a.pts[p * 3 + 2] = pz; }
      prod *= (1.f - t.alpha + 1e-7f);
    }
  }
  const float incl = wscan_mul(prod, lane);
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f, mw = 0.f, sdp = 0.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    const int j = j0 + k;
    if (k < per && j < S) {
      const float w = alpha[k] * T;
      T *= (1.f - alpha[k] + 1e-7f);
      sr += w * cr[k]; sg += w * cg[k]; sb += w * cb[k];
      sw += w; mw = fmaxf(mw, w); sdp += w * midz[k];
      if (a.weights) a.weights[ray * S + j] = w;
    }
  }
  sr = wsum(sr); sg = wsum(sg); sb = wsum(sb); sw = wsum(sw); sdp = wsum(sdp); mw = wmax(mw);
  e_num = wsum(e_num); e_den = wsum(e_den);
  if (lane == 0) {
    if (a.bg) { sr += a.bg[0] * (1.f - sw); sg += a.bg[1] * (1.f - sw); sb += a.bg[2] * (1.f - sw); }
    a.color[ray * 3] = sr; a.color[ray * 3 + 1] = sg; a.color[ray * 3 + 2] = sb;
    a.wsum_[ray] = sw; a.wmax_[ray] = mw; a.depth[ray] = sdp;
    a.eik[ray * 2] = e_num; a.eik[ray * 2 + 1] = e_den;
  }
}

struct CompBwdArgs {
  CompArgs f;
  // upstream gradients
  const float* g_color;    // [B,3] or null
  const float* g_wsum;     // [B] or null
  const float* g_depth;    // [B] or null
  const float* g_weights;  // [B,S] or null
  const float* g_eik;      // device scalar: dL/d(gradient_error) or null
  const float* eik_den;    // device scalar: sum(relax) over ALL rays of the (global) batch
  const float* g_nrm_ext;  // [B*S,3] or null: upstream gradient on the returned `gradients`
  // outputs
  float* d_sdf;            // [B*S]
  float* d_nrm;            // [B*S,3]
  float* d_rgb;            // [B*S,3]
  float* d_dir;            // [B,3]   (true_cos term only)
  float* d_dist;           // [B,S]   dL/d dists (alpha path only)
  float* d_mid;            // [B,S]   dL/d mid_z (depth term only)
  float* d_invs;           // [B]     per-ray partial of dL/d inv_s
};

__global__ void __launch_bounds__(CW * 32) composite_bwd_kernel(CompBwdArgs b) {
  const CompArgs& a = b.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * CW + warp;
  if (ray >= a.B) return;
  const int S = a.S;
  const int per = (S + 31) / 32;
  const int j0 = lane * per;
  const float ox = a.rays_o[ray * 3], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const float inv_s = *a.inv_s;
  const float* zr = a.z + ray * S;
  const float gc0 = b.g_color ? b.g_color[ray * 3] : 0.f, gc1 = b.g_color ? b.g_color[ray * 3 + 1] : 0.f,
              gc2 = b.g_color ? b.g_color[ray * 3 + 2] : 0.f;
  float gws = b.g_wsum ? b.g_wsum[ray] : 0.f;
  if (a.bg && b.g_color) gws -= gc0 * a.bg[0] + gc1 * a.bg[1] + gc2 * a.bg[2];
  const float gdp = b.g_depth ? b.g_depth[ray] : 0.f;
  const float geik = b.g_eik ? (*b.g_eik) / (*b.eik_den + 1e-5f) : 0.f;

  SampleTerms tt[CMAX];
  float nx[CMAX], ny[CMAX], nz[CMAX], gw[CMAX], sdfv[CMAX];
  float prod = 1.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    const int j = j0 + k;
    gw[k] = 0.f; nx[k] = ny[k] = nz[k] = 0.f; sdfv[k] = 0.f;
    tt[k].alpha = 0.f;
    if (k < per && j < S) {
      const long long p = ray * S + j;
      const float z0 = zr[j], z1 = (j + 1 < S) ? zr[j + 1] : 0.f;
      nx[k] = a.nrm[p * 3]; ny[k] = a.nrm[p * 3 + 1]; nz[k] = a.nrm[p * 3 + 2];
      sdfv[k] = a.sdf[p];
      tt[k] = sample_terms(z0, z1, j + 1 == S, a.sample_dist, sdfv[k], nx[k], ny[k], nz[k], dx, dy, dz, inv_s,
                           a.cos_anneal);
      // dL/dw_j from colour, weight_sum, depth and a direct upstream on weights
      gw[k] = gc0 * a.rgb[p * 3] + gc1 * a.rgb[p * 3 + 1] + gc2 * a.rgb[p * 3 + 2] + gws + gdp * tt[k].mid +
              (b.g_weights ? b.g_weights[p] : 0.f);
      prod *= (1.f - tt[k].alpha + 1e-7f);
    }
  }
  const float incl = wscan_mul(prod, lane);
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  // weights and the suffix sum  S_j = sum_{k>j} gw_k * w_k
  float w[CMAX], Tj[CMAX];
  float lsum = 0.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    const int j = j0 + k;
    w[k] = 0.f; Tj[k] = 0.f;
    if (k < per && j < S) {
      Tj[k] = T;
      w[k] = tt[k].alpha * T;
      T *= (1.f - tt[k].alpha + 1e-7f);
      lsum += gw[k] * w[k];
    }
  }
  const float total = wsum(lsum);
  const float incl_s = wscan_add(lsum, lane);
  float suffix = total - incl_s;             // sum over lanes > this lane
  float dd0 = 0.f, dd1 = 0.f, dd2 = 0.f, dinvs = 0.f;
#pragma unroll
  for (int k = CMAX - 1; k >= 0; --k) {
    const int j = j0 + k;
    if (k < per && j < S) {
      const long long p = ray * S + j;
      const SampleTerms& t = tt[k];
      float g_alpha = gw[k] * Tj[k] - suffix / (1.f - t.alpha + 1e-7f);
      suffix += gw[k] * w[k];
      if (!(t.alpha_raw >= 0.f && t.alpha_raw <= 1.f)) g_alpha = 0.f;      // clip(0,1) backward
      const float g_prev_c = g_alpha * t.next_c / (t.den * t.den);        // d(num/den)/d prev
      const float g_next_c = -g_alpha / t.den;
      const float g_ep = g_prev_c * t.prev_c * (1.f - t.prev_c) * inv_s;   // estimated_prev_sdf
      const float g_en = g_next_c * t.next_c * (1.f - t.next_c) * inv_s;
      const float ep = sdfv[k] - t.ic * t.dist * 0.5f, en = sdfv[k] + t.ic * t.dist * 0.5f;
      dinvs += g_prev_c * t.prev_c * (1.f - t.prev_c) * ep + g_next_c * t.next_c * (1.f - t.next_c) * en;
      const float g_sdf = g_ep + g_en;
      const float g_ic = (g_en - g_ep) * t.dist * 0.5f;
      const float g_dist = (g_en - g_ep) * t.ic * 0.5f;
      // iter_cos = -(relu(.5-.5tc)(1-r) + relu(-tc) r)
      const float g_tc = g_ic * ((t.tc < 1.f ? 0.5f * (1.f - a.cos_anneal) : 0.f) + (t.tc < 0.f ? a.cos_anneal : 0.f));
      // eikonal: relax * (|n|-1)^2 / (sum relax + 1e-5)
      const float px = ox + dx * t.mid, py = oy + dy * t.mid, pz = oz + dz * t.mid;
      const float relax = sqrtf(px * px + py * py + pz * pz) < 1.2f ? 1.f : 0.f;
      const float gn = sqrtf(nx[k] * nx[k] + ny[k] * ny[k] + nz[k] * nz[k]);
      const float ge = (gn > 0.f) ? geik * relax * 2.f * (gn - 1.f) / gn : 0.f;
      float gnx = g_tc * dx + ge * nx[k], gny = g_tc * dy + ge * ny[k], gnz = g_tc * dz + ge * nz[k];
      if (b.g_nrm_ext) { gnx += b.g_nrm_ext[p * 3]; gny += b.g_nrm_ext[p * 3 + 1]; gnz += b.g_nrm_ext[p * 3 + 2]; }
      b.d_sdf[p] = g_sdf;
      b.d_nrm[p * 3] = gnx; b.d_nrm[p * 3 + 1] = gny; b.d_nrm[p * 3 + 2] = gnz;
      b.d_rgb[p * 3] = w[k] * gc0; b.d_rgb[p * 3 + 1] = w[k] * gc1; b.d_rgb[p * 3 + 2] = w[k] * gc2;
      b.d_dist[p] = (j + 1 == S) ? 0.f : g_dist;
      b.d_mid[p] = gdp * w[k];
      dd0 += g_tc * nx[k]; dd1 += g_tc * ny[k]; dd2 += g_tc * nz[k];
    }
  }
  dd0 = wsum(dd0); dd1 = wsum(dd1); dd2 = wsum(dd2); dinvs = wsum(dinvs);
  if (lane == 0) {
    b.d_dir[ray * 3] = dd0; b.d_dir[ray * 3 + 1] = dd1; b.d_dir[ray * 3 + 2] = dd2;
    b.d_invs[ray] = dinvs;
  }
}

// ---------------------------------------------------------------------------------------------------------
// per-ray loss terms (exp_runner.py:562-599): masked L1 colour / mask_sum, BCE(clip(weight_sum), mask) / n_rays.
// Writes per-ray loss partials [B,2] (colour, bce) and the gradients w.r.t. colour and weight_sum for
// loss = colour + mask_weight * bce  (the eikonal term's gradient is passed separately as g_eik).
// ---------------------------------------------------------------------------------------------------------
__global__ void loss_fwd_bwd_kernel(const float* __restrict__ color, const float* __restrict__ wsum_,
                                    const float* __restrict__ true_rgb, const float* __restrict__ mask, long long B,
                                    const float* __restrict__ mask_sum, float inv_n_rays, float mask_weight,
                                    int use_mask, float* __restrict__ partial, float* __restrict__ g_color,
                                    float* __restrict__ g_wsum) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float m = use_mask ? (mask[r] > 0.5f ? 1.f : 0.f) : 1.f;
  const float ms = *mask_sum;                 // sum(mask) + 1e-5 over the GLOBAL batch
  float cl = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float e = (color[r * 3 + c] - true_rgb[r * 3 + c]) * m;
    cl += fabsf(e);
    g_color[r * 3 + c] = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * m / ms;
  }
  const float ws = wsum_[r];
  const float x = fminf(fmaxf(ws, 1e-3f), 1.f - 1e-3f);
  // F.binary_cross_entropy clamps log at -100 (irrelevant inside [1e-3, 1-1e-3])
  const float bce = -(m * logf(x) + (1.f - m) * logf(1.f - x));
  const float g_x = (-(m / x) + (1.f - m) / (1.f - x)) * inv_n_rays * mask_weight;
  g_wsum[r] = (ws >= 1e-3f && ws <= 1.f - 1e-3f) ? g_x : 0.f;
  partial[r * 2] = cl / ms;
  partial[r * 2 + 1] = bce * inv_n_rays;
}

// ---------------------------------------------------------------------------------------------------------
// pts = o + d * mid_z backward, plus view-dir gradients from the colour net; optional z gradients
// (only needed when n_importance == 0 so that z_vals carry grad through near/far, renderer.py:390).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CW * 32)
ray_reduce_bwd_kernel(const float* __restrict__ d_pts, const float* __restrict__ d_dirs, const float* __restrict__ d_dir_tc,
                      const float* __restrict__ d_dist, const float* __restrict__ d_mid, const float* __restrict__ rays_d,
                      const float* __restrict__ z, long long B, int S, float sample_dist, float* __restrict__ d_o,
                      float* __restrict__ d_d, float* __restrict__ d_z) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * CW + warp;
  if (ray >= B) return;
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
  for (int j = lane; j < S; j += 32) {
    const long long p = ray * S + j;
    const float z0 = z[p];
    const float dist = (j + 1 < S) ? z[p + 1] - z0 : sample_dist;
    const float mid = z0 + dist * 0.5f;
    const float g0 = d_pts[p * 3], g1 = d_pts[p * 3 + 1], g2 = d_pts[p * 3 + 2];
    o0 += g0; o1 += g1; o2 += g2;
    v0 += g0 * mid; v1 += g1 * mid; v2 += g2 * mid;
    if (d_dirs) { v0 += d_dirs[p * 3]; v1 += d_dirs[p * 3 + 1]; v2 += d_dirs[p * 3 + 2]; }
    if (d_z) {
      // total dL/dmid_j and dL/ddist_j, then z_j enters mid_j (+1), dist_j (-1), dist_{j-1} (+1)
      const float gmid = d_mid[p] + g0 * dx + g1 * dy + g2 * dz;
      const float gdist = ((j + 1 < S) ? d_dist[p] + 0.5f * gmid : 0.f);
      float gz = gmid - gdist;
      if (j > 0) {
        const long long q = p - 1;
        const float distq = z0 - z[q];
        const float midq = z[q] + distq * 0.5f;
        (void)midq;
        const float gmidq = d_mid[q] + d_pts[q * 3] * dx + d_pts[q * 3 + 1] * dy + d_pts[q * 3 + 2] * dz;
        gz += d_dist[q] + 0.5f * gmidq;
      }
      d_z[p] = gz;
    }
  }
  o0 = wsum(o0); o1 = wsum(o1); o2 = wsum(o2); v0 = wsum(v0); v1 = wsum(v1); v2 = wsum(v2);
  if (lane == 0) {
    d_o[ray * 3] = o0; d_o[ray * 3 + 1] = o1; d_o[ray * 3 + 2] = o2;
    d_d[ray * 3] = v0 + d_dir_tc[ray * 3]; d_d[ray * 3 + 1] = v1 + d_dir_tc[ray * 3 + 1];
    d_d[ray * 3 + 2] = v2 + d_dir_tc[ray * 3 + 2];
  }
}

}  // namespace fmov
using namespace fmov;

static int fill_comp(CompArgs& a, long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                     const float* sdf, const float* nrm, const float* rgb, const float* inv_s, float sample_dist,
                     float cos_anneal, const float* bg) {
  FMOV_REQUIRE(B >= 0 && S >= 1 && S <= 32 * CMAX, "composite: S=%d out of range (1..%d)", S, 32 * CMAX);
  FMOV_REQUIRE(B == 0 || (rays_o && rays_d && z && sdf && nrm && rgb && inv_s), "composite: null argument");
  memset(&a, 0, sizeof(a));
  a.B = B; a.S = S; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.sdf = sdf; a.nrm = nrm; a.rgb = rgb;
  a.inv_s = inv_s; a.sample_dist = sample_dist; a.cos_anneal = cos_anneal; a.bg = bg;
  return OK;
}

extern "C" int fmov_composite_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                                  const float* sdf, const float* nrm, const float* rgb, const float* inv_s,
                                  float sample_dist, float cos_anneal, const float* bg, float* color, float* weight_sum,
                                  float* weight_max, float* depth, float* weights, float* cdf, float* inside,
                                  float* mid_z, float* pts, float* eik_partial, void* stream) {
  CompArgs a;
  int st = fill_comp(a, B, S, rays_o, rays_d, z, sdf, nrm, rgb, inv_s, sample_dist, cos_anneal, bg);
  if (st) return st;
  if (B == 0) return OK;
  FMOV_REQUIRE(color && weight_sum && weight_max && depth && eik_partial, "fmov_composite_fwd: null output");
  a.color = color; a.wsum_ = weight_sum; a.wmax_ = weight_max; a.depth = depth; a.weights = weights; a.cdf = cdf;
  a.inside = inside; a.mid_z = mid_z; a.pts = pts; a.eik = eik_partial;
  composite_fwd_kernel<<<(unsigned)((B + CW - 1) / CW), CW * 32, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("composite_fwd_kernel");
  return OK;
}

extern "C" int fmov_composite_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                                  const float* sdf, const float* nrm, const float* rgb, const float* inv_s,
                                  float sample_dist, float cos_anneal, const float* bg, const float* g_color,
                                  const float* g_wsum, const float* g_depth, const float* g_weights, const float* g_eik,
                                  const float* eik_den, const float* g_nrm_ext, float* d_sdf, float* d_nrm, float* d_rgb,
                                  float* d_dir, float* d_dist, float* d_mid, float* d_invs, void* stream) {
  CompBwdArgs b;
  memset(&b, 0, sizeof(b));
  int st = fill_comp(b.f, B, S, rays_o, rays_d, z, sdf, nrm, rgb, inv_s, sample_dist, cos_anneal, bg);
  if (st) return st;
  if (B == 0) return OK;
  FMOV_REQUIRE(d_sdf && d_nrm && d_rgb && d_dir && d_dist && d_mid && d_invs, "fmov_composite_bwd: null output");
  FMOV_REQUIRE(!g_eik || eik_den, "fmov_composite_bwd: g_eik needs eik_den");
  b.g_color = g_color; b.g_wsum = g_wsum; b.g_depth = g_depth; b.g_weights = g_weights; b.g_eik = g_eik;
  b.eik_den = eik_den; b.g_nrm_ext = g_nrm_ext;
  b.d_sdf = d_sdf; b.d_nrm = d_nrm; b.d_rgb = d_rgb; b.d_dir = d_dir; b.d_dist = d_dist; b.d_mid = d_mid; b.d_invs = d_invs;
  composite_bwd_kernel<<<(unsigned)((B + CW - 1) / CW), CW * 32, 0, (cudaStream_t)stream>>>(b);
  FMOV_LAUNCH_CHECK("composite_bwd_kernel");
  return OK;
}

extern "C" int fmov_loss_fwd_bwd(const float* color, const float* weight_sum, const float* true_rgb, const float* mask,
                                 long long B, const float* mask_sum, long long n_rays_global, float mask_weight,
                                 float* partial, float* g_color, float* g_wsum, void* stream) {
  FMOV_REQUIRE(B >= 0 && n_rays_global > 0, "fmov_loss_fwd_bwd: bad sizes");
  if (B == 0) return OK;
  FMOV_REQUIRE(color && weight_sum && true_rgb && mask && mask_sum && partial && g_color && g_wsum,
               "fmov_loss_fwd_bwd: null argument");
  loss_fwd_bwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      color, weight_sum, true_rgb, mask, B, mask_sum, 1.0f / (float)n_rays_global, mask_weight, mask_weight > 0.f ? 1 : 0,
      partial, g_color, g_wsum);
  FMOV_LAUNCH_CHECK("loss_fwd_bwd_kernel");
  return OK;
}

extern "C" int fmov_ray_reduce_bwd(const float* d_pts, const float* d_dirs, const float* d_dir_tc, const float* d_dist,
                                   const float* d_mid, const float* rays_d, const float* z, long long B, int S,
                                   float sample_dist, float* d_o, float* d_d, float* d_z, void* stream) {
  FMOV_REQUIRE(B >= 0 && S >= 1, "fmov_ray_reduce_bwd: bad sizes");
  if (B == 0) return OK;
  FMOV_REQUIRE(d_pts && d_dir_tc && rays_d && z && d_o && d_d && (!d_z || (d_dist && d_mid)),
               "fmov_ray_reduce_bwd: null argument");
  ray_reduce_bwd_kernel<<<(unsigned)((B + CW - 1) / CW), CW * 32, 0, (cudaStream_t)stream>>>(
      d_pts, d_dirs, d_dir_tc, d_dist, d_mid, rays_d, z, B, S, sample_dist, d_o, d_d, d_z);
  FMOV_LAUNCH_CHECK("ray_reduce_bwd_kernel");
  return OK;
}
