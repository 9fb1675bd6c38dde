// Hierarchical sampling kernels (warp per ray, everything in shared memory / registers).
//
//   fmov_sample_coarse     NeuSRenderer.render coarse z + jitter         models/renderer.py:385-405
//   fmov_sample_round      cat_z_vals merge (sort of two sorted lists, SDF permuted alongside)
//                          models/renderer.py:222-242, followed by up_sample :168-220 and
//                          sample_pdf(det=True) :54-86 for the next round.
// The SDF queries between rounds are fmov_sdf_query_rays (mlp_query.cu); the host sequences
//   coarse -> query -> [round(i): (merge) + up_sample -> query]* -> final merge.
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

constexpr int SAMP_MAXS = 256;
constexpr int SAMP_WARPS = 4;

__device__ __forceinline__ float lin01(int n, int i) {   // torch.linspace(0,1,n)[i]
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.0f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

__global__ void sample_coarse_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                     const float* __restrict__ t_rand, long long B, int n, int z_stride,
                                     float* __restrict__ z) {
  const long long total = B * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int j = (int)(i - r * n);
    const float nr = near[r], fr = far[r];
    float v = __fadd_rn(nr, __fmul_rn(__fsub_rn(fr, nr), lin01(n, j)));          // renderer.py:389-390
    if (t_rand) v = __fadd_rn(v, __fdiv_rn(__fmul_rn(__fsub_rn(t_rand[r], 0.5f), 2.0f), (float)n));  // :404-405
    z[r * z_stride + j] = v;
  }
}

__device__ __forceinline__ float warp_incl_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sample_pdf(bins, weights, n_new, det=True) of models/renderer.py:54-86 for one ray, by one warp:
// Z[0..n) = bins, Q[0..n-1) = weights + 1e-5 (renderer.py:57), C[0..n) = scratch that receives the cdf, out[0..n_new).
__device__ __forceinline__ void warp_sample_pdf(const float* Z, const float* Q, float* C, int n, int n_new, int lane,
                                                float* out) {
  const int ns = n - 1;
  const int per = (ns + 31) / 32;             // contiguous sections per lane
  const int j0 = lane * per;
  float lsum = 0.f;
  for (int k = 0; k < per; ++k)
    if (j0 + k < ns) lsum += Q[j0 + k];
  const float total = warp_sum(lsum);
  const float incl_s = warp_incl_scan_add(lsum, lane);
  const float run = incl_s - lsum;             // exclusive prefix of this lane's chunk
  // cdf[0] = 0, cdf[j+1] = cumsum(pdf)[j]   (renderer.py:58-60)
  if (lane == 0) C[0] = 0.f;
  float cum = run / total;
  for (int k = 0; k < per; ++k) {
    const int j = j0 + k;
    if (j < ns) {
      cum += Q[j] / total;
      C[j + 1] = cum;
    }
  }
  __syncwarp();
  // ---- inverse CDF (renderer.py:62-84) ---------------------------------------------------
  for (int k = lane; k < n_new; k += 32) {
    // torch.linspace(0.5/m, 1-0.5/m, m)
    const float a = 0.5f / (float)n_new, b = 1.0f - 0.5f / (float)n_new;
    const float step = (b - a) / (float)(n_new - 1 > 0 ? n_new - 1 : 1);
    const float u = (k < n_new / 2) ? a + step * (float)k : b - step * (float)(n_new - 1 - k);
    int lo = 0, hi = n;                         // searchsorted(cdf, u, right=True) = #{cdf <= u}
    while (lo < hi) { int m = (lo + hi) >> 1; if (C[m] <= u) lo = m + 1; else hi = m; }
    const int below = max(lo - 1, 0), above = min(n - 1, lo);
    const float cb = C[below], ca = C[above];
    float denom = ca - cb;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = (u - cb) / denom;
    out[k] = Z[below] + t * (Z[above] - Z[below]);
  }
}

// sample_pdf on given bins / weights (the inverse-CDF stage of the round kernel on its own): warp per ray
__global__ void __launch_bounds__(SAMP_WARPS * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, long long B, int n, int n_new,
                  float* __restrict__ out) {
  __shared__ float sz[SAMP_WARPS][SAMP_MAXS];
  __shared__ float sc[SAMP_WARPS][SAMP_MAXS];
  __shared__ float sq[SAMP_WARPS][SAMP_MAXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  for (int i = lane; i < n; i += 32) {
    sz[warp][i] = bins[ray * n + i];
    if (i < n - 1) sq[warp][i] = weights[ray * (n - 1) + i] + 1e-5f;
  }
  __syncwarp();
  warp_sample_pdf(sz[warp], sq[warp], sc[warp], n, n_new, lane, out + ray * n_new);
}

// One warp per ray.
//   merge:    z[:, :n_sorted] (sorted) and z[:, n_sorted:n_sorted+n_tail] (sorted) -> z[:, :n_cur] sorted,
//             sdf permuted identically when with_sdf (not on the last round, renderer.py:229-240)
//   upsample: n_new > 0 -> new samples written to z[:, n_cur : n_cur+n_new]
__global__ void __launch_bounds__(SAMP_WARPS * 32)
sample_round_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ z,
                    float* __restrict__ sdf, long long B, int z_stride, int n_sorted, int n_tail, int with_sdf,
                    int n_new, float inv_s) {
  __shared__ float sz[SAMP_WARPS][SAMP_MAXS];
  __shared__ float ss[SAMP_WARPS][SAMP_MAXS];
  __shared__ float sc[SAMP_WARPS][SAMP_MAXS];   // merge scratch, then cdf
  __shared__ float sq[SAMP_WARPS][SAMP_MAXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  float* zr = z + ray * z_stride;
  float* sr = sdf ? sdf + ray * z_stride : nullptr;
  float* Z = sz[warp];
  float* Sd = ss[warp];
  float* C = sc[warp];
  float* Q = sq[warp];
  const int n = n_sorted + n_tail;

  // ---- load (and merge) ---------------------------------------------------------------
  if (n_tail > 0) {
    for (int i = lane; i < n; i += 32) {
      C[i] = zr[i];
      if (with_sdf) Q[i] = sr[i];
    }
    __syncwarp();
    const float* a = C;              // existing samples
    const float* b = C + n_sorted;   // new samples (monotone: inverse CDF of increasing u)
    for (int i = lane; i < n; i += 32) {
      int pos;
      if (i < n_sorted) {            // rank = i + #{b < a_i}
        const float v = a[i];
        int lo = 0, hi = n_tail;
        while (lo < hi) { int m = (lo + hi) >> 1; if (b[m] < v) lo = m + 1; else hi = m; }
        pos = i + lo;
      } else {                       // rank = j + #{a <= b_j}   (stable: existing first on ties)
        const int j = i - n_sorted;
        const float v = b[j];
        int lo = 0, hi = n_sorted;
        while (lo < hi) { int m = (lo + hi) >> 1; if (a[m] <= v) lo = m + 1; else hi = m; }
        pos = j + lo;
      }
      Z[pos] = C[i];
      if (with_sdf) Sd[pos] = Q[i];
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      zr[i] = Z[i];
      if (with_sdf) sr[i] = Sd[i];
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      Z[i] = zr[i];
      if (n_new > 0) Sd[i] = sr[i];
    }
  }
  __syncwarp();
  if (n_new <= 0) return;

  // ---- up_sample (renderer.py:168-217): section weights --------------------------------
  const float ox = rays_o[ray * 3 + 0], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
  const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const int ns = n - 1;                       // sections
  const int per = (ns + 31) / 32;             // contiguous sections per lane
  const int j0 = lane * per;
  // pass 1: alpha per section into C[], running product prefix via warp scan
  float prod = 1.0f;
  for (int k = 0; k < per; ++k) {
    const int j = j0 + k;
    float alpha = 0.f;
    if (j < ns) {
      const float za = Z[j], zb = Z[j + 1], sa = Sd[j], sb = Sd[j + 1];
      const float pax = ox + dx * za, pay = oy + dy * za, paz = oz + dz * za;
      const float pbx = ox + dx * zb, pby = oy + dy * zb, pbz = oz + dz * zb;
      const float ra = sqrtf(pax * pax + pay * pay + paz * paz), rb = sqrtf(pbx * pbx + pby * pby + pbz * pbz);
      const bool inside = (ra < 1.0f) || (rb < 1.0f);
      const float dist = zb - za;
      float cosv = (sb - sa) / (dist + 1e-5f);
      float prev_cos = 0.f;
      if (j > 0) prev_cos = (sa - Sd[j - 1]) / (za - Z[j - 1] + 1e-5f);
      cosv = fminf(prev_cos, cosv);
      cosv = fminf(fmaxf(cosv, -1e3f), 0.0f) * (inside ? 1.0f : 0.0f);
      const float mid = (sa + sb) * 0.5f;
      const float pe = mid - cosv * dist * 0.5f, ne = mid + cosv * dist * 0.5f;
      const float pc = 1.0f / (1.0f + expf(-pe * inv_s)), nc = 1.0f / (1.0f + expf(-ne * inv_s));
      alpha = (pc - nc + 1e-5f) / (pc + 1e-5f);
      C[j] = alpha;
      prod *= (1.0f - alpha + 1e-7f);
    }
  }
  const float incl = warp_incl_scan_mul(prod, lane);
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.0f;
  // pass 2: weights + 1e-5 (renderer.py:57)
  for (int k = 0; k < per; ++k) {
    const int j = j0 + k;
    if (j < ns) {
      const float alpha = C[j];
      Q[j] = alpha * T + 1e-5f;
      T *= (1.0f - alpha + 1e-7f);
    }
  }
  __syncwarp();          // every lane has finished reading alpha from C[] before the cdf overwrites it
  warp_sample_pdf(Z, Q, C, n, n_new, lane, zr + n);
}

}  // namespace fmov
using namespace fmov;

extern "C" int fmov_sample_coarse(const float* near, const float* far, const float* t_rand, long long B, int n_samples,
                                  int z_stride, float* z, void* stream) {
  FMOV_REQUIRE(B >= 0 && n_samples > 0 && z_stride >= n_samples, "fmov_sample_coarse: bad shape");
  if (B == 0) return OK;
  FMOV_REQUIRE(near && far && z, "fmov_sample_coarse: null argument");
  const long long total = B * n_samples;
  const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  sample_coarse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(near, far, t_rand, B, n_samples, z_stride, z);
  FMOV_LAUNCH_CHECK("sample_coarse_kernel");
  return OK;
}

extern "C" int fmov_sample_round(const float* rays_o, const float* rays_d, float* z, float* sdf, long long B,
                                 int z_stride, int n_sorted, int n_tail, int with_sdf, int n_new, float inv_s,
                                 void* stream) {
  const int n = n_sorted + n_tail;
  FMOV_REQUIRE(B >= 0 && n >= 2 && n + n_new <= z_stride && n + n_new <= SAMP_MAXS && n_tail >= 0 && n_new >= 0,
               "fmov_sample_round: bad shape (n_sorted=%d n_tail=%d n_new=%d stride=%d, max %d samples)", n_sorted, n_tail,
               n_new, z_stride, SAMP_MAXS);
  if (B == 0) return OK;
  FMOV_REQUIRE(rays_o && rays_d && z && (sdf || (!with_sdf && n_new == 0)), "fmov_sample_round: null argument");
  const int grid = (int)((B + SAMP_WARPS - 1) / SAMP_WARPS);
  sample_round_kernel<<<grid, SAMP_WARPS * 32, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, sdf, B, z_stride, n_sorted,
                                                                         n_tail, with_sdf, n_new, inv_s);
  FMOV_LAUNCH_CHECK("sample_round_kernel");
  return OK;
}

// sample_pdf(bins [B,n], weights [B,n-1], n_new, det=True) -> out [B,n_new]  (models/renderer.py:54-86): the inverse-CDF
// stage of fmov_sample_round on caller-given weights (known-answer tests; also usable as a stand-alone resampler)
extern "C" int fmov_sample_pdf(const float* bins, const float* weights, long long B, int n, int n_new, float* out,
                               void* stream) {
  FMOV_REQUIRE(B >= 0 && n >= 2 && n <= SAMP_MAXS && n_new >= 1, "fmov_sample_pdf: bad shape (n=%d n_new=%d, max %d bins)", n,
               n_new, SAMP_MAXS);
  if (B == 0) return OK;
  FMOV_REQUIRE(bins && weights && out, "fmov_sample_pdf: null argument");
  const int grid = (int)((B + SAMP_WARPS - 1) / SAMP_WARPS);
  sample_pdf_kernel<<<grid, SAMP_WARPS * 32, 0, (cudaStream_t)stream>>>(bins, weights, B, n, n_new, out);
  FMOV_LAUNCH_CHECK("sample_pdf_kernel");
  return OK;
}
