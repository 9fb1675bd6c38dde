// Flow / reprojection loss core and unit-sphere loss core, forward and backward (warp per ray; HBM-bound).
//
//   fmov_flow_fwd / _bwd   exp_runner.py:605-688: every sample point pts = o + d*mid_z of a ray is projected into
//                          the matched frame (w2c, K), compared with the matched pixel and the weighted error is
//                          summed over the ray:  err = sum_j w_j * (pi(K (R_w p_j + t_w)) - xy).
//                          The [P,3] `pts` tensor and its gradient never exist: the backward reduces straight to
//                          d rays_o, d rays_d, d z (n_importance == 0 only), d weights and d w2c.
//   fmov_unit_sphere_fwd_bwd  exp_runner.py:714-724: mean |w| over the samples with |pts| > 1 (mask detached).
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

constexpr int FW = 4;   // warps (rays) per block

__device__ __forceinline__ float fsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FlowArgs {
  long long B;
  int S;
  const float* rays_o; const float* rays_d; const float* z; const float* weights;
  const float* w2c;      // [3,4] row-major, device
  const float* K;        // [3,3] row-major (row stride k_stride), device
  int k_stride;
  const float* xy;       // [B,2] matched pixel
  float sample_dist;
};

struct Proj {   // projection of one sample
  float p[3], q[3], u, v;
};

__device__ __forceinline__ float mid_of(const float* __restrict__ zr, int j, int S, float sample_dist) {
  const float z0 = zr[j];
  const float dist = (j + 1 < S) ? zr[j + 1] - z0 : sample_dist;      // models/renderer.py:261-267
  return z0 + dist * 0.5f;
}

__device__ __forceinline__ Proj project(const float* o, const float* d, float mid, const float* W, const float* K) {
  Proj r;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.p[i] = o[i] + d[i] * mid;
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = W[i * 4] * r.p[0] + W[i * 4 + 1] * r.p[1] + W[i * 4 + 2] * r.p[2] + W[i * 4 + 3];
#pragma unroll
  for (int i = 0; i < 3; ++i) r.q[i] = K[i * 3] * c[0] + K[i * 3 + 1] * c[1] + K[i * 3 + 2] * c[2];
  r.u = r.q[0] / r.q[2];
  r.v = r.q[1] / r.q[2];
  return r;
}

__device__ __forceinline__ void load_cam(const FlowArgs& a, float* sW, float* sK) {
  if (threadIdx.x < 12) sW[threadIdx.x] = a.w2c[threadIdx.x];
  if (threadIdx.x < 9) sK[threadIdx.x] = a.K[(threadIdx.x / 3) * a.k_stride + threadIdx.x % 3];
  __syncthreads();
}

__global__ void __launch_bounds__(FW * 32) flow_fwd_kernel(FlowArgs a, float* __restrict__ err) {
  __shared__ float sW[12], sK[9];
  load_cam(a, sW, sK);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * FW + warp;
  if (ray >= a.B) return;
  const float o[3] = {a.rays_o[ray * 3], a.rays_o[ray * 3 + 1], a.rays_o[ray * 3 + 2]};
  const float d[3] = {a.rays_d[ray * 3], a.rays_d[ray * 3 + 1], a.rays_d[ray * 3 + 2]};
  const float tx = a.xy[ray * 2], ty = a.xy[ray * 2 + 1];
  const float* zr = a.z + ray * a.S;
  const float* wr = a.weights + ray * a.S;
  float eu = 0.f, ev = 0.f;
  for (int j = lane; j < a.S; j += 32) {
    const Proj pr = project(o, d, mid_of(zr, j, a.S, a.sample_dist), sW, sK);
    const float w = wr[j];
    eu += (pr.u - tx) * w;
    ev += (pr.v - ty) * w;
  }
  eu = fsum(eu); ev = fsum(ev);
  if (lane == 0) { err[ray * 2] = eu; err[ray * 2 + 1] = ev; }
}

// dL/dp of sample j given du = g0*w, dv = g1*w; also returns dc (camera-frame gradient) for the w2c gradient
__device__ __forceinline__ void proj_bwd(const Proj& pr, float du, float dv, const float* W, const float* K, float* dc,
                                         float* dp) {
  const float iq = 1.0f / pr.q[2];
  const float dq[3] = {du * iq, dv * iq, -(du * pr.q[0] + dv * pr.q[1]) * iq * iq};
#pragma unroll
  for (int i = 0; i < 3; ++i) dc[i] = K[i] * dq[0] + K[3 + i] * dq[1] + K[6 + i] * dq[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) dp[i] = W[i] * dc[0] + W[4 + i] * dc[1] + W[8 + i] * dc[2];
}

__global__ void __launch_bounds__(FW * 32)
flow_bwd_kernel(FlowArgs a, const float* __restrict__ g_err, float* __restrict__ d_w, float* __restrict__ d_o,
                float* __restrict__ d_d, float* __restrict__ d_z, float* __restrict__ d_w2c) {
  __shared__ float sW[12], sK[9], acc[12];
  if (threadIdx.x < 12) acc[threadIdx.x] = 0.f;
  load_cam(a, sW, sK);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * FW + warp;
  float G[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) G[i] = 0.f;
  if (ray < a.B) {
    const float o[3] = {a.rays_o[ray * 3], a.rays_o[ray * 3 + 1], a.rays_o[ray * 3 + 2]};
    const float d[3] = {a.rays_d[ray * 3], a.rays_d[ray * 3 + 1], a.rays_d[ray * 3 + 2]};
    const float tx = a.xy[ray * 2], ty = a.xy[ray * 2 + 1];
    const float g0 = g_err[ray * 2], g1 = g_err[ray * 2 + 1];
    const float* zr = a.z + ray * a.S;
    const float* wr = a.weights + ray * a.S;
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    for (int j = lane; j < a.S; j += 32) {
      const float mid = mid_of(zr, j, a.S, a.sample_dist);
      const Proj pr = project(o, d, mid, sW, sK);
      const float w = wr[j];
      if (d_w) d_w[ray * a.S + j] = g0 * (pr.u - tx) + g1 * (pr.v - ty);
      float dc[3], dp[3];
      proj_bwd(pr, g0 * w, g1 * w, sW, sK, dc, dp);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        go[i] += dp[i];
        gd[i] += dp[i] * mid;
        G[i * 4] += dc[i] * pr.p[0]; G[i * 4 + 1] += dc[i] * pr.p[1]; G[i * 4 + 2] += dc[i] * pr.p[2];
        G[i * 4 + 3] += dc[i];
      }
      if (d_z) {
        // mid_j = (z_j + z_{j+1})/2 (last: z_j + sample_dist/2): z_j enters mid_j and mid_{j-1}
        const float gmid = dp[0] * d[0] + dp[1] * d[1] + dp[2] * d[2];
        float gz = (j + 1 < a.S) ? 0.5f * gmid : gmid;
        if (j > 0) {
          const Proj pq = project(o, d, mid_of(zr, j - 1, a.S, a.sample_dist), sW, sK);
          const float wq = wr[j - 1];
          float dcq[3], dpq[3];
          proj_bwd(pq, g0 * wq, g1 * wq, sW, sK, dcq, dpq);
          gz += 0.5f * (dpq[0] * d[0] + dpq[1] * d[1] + dpq[2] * d[2]);
        }
        d_z[ray * a.S + j] = gz;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { go[i] = fsum(go[i]); gd[i] = fsum(gd[i]); }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { d_o[ray * 3 + i] = go[i]; d_d[ray * 3 + i] = gd[i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float s = fsum(G[i]);
    if (lane == 0) atomicAdd(&acc[i], s);
  }
  __syncthreads();
  if (threadIdx.x < 12) atomicAdd(&d_w2c[threadIdx.x], acc[threadIdx.x]);
}

// partial[0] += sum_{|p|>1} |w| , partial[1] += count ; d_w = g_scale[0] * sign(w) * [|p|>1]  (g_scale: device scalar)
__global__ void __launch_bounds__(FW * 32)
unit_sphere_kernel(long long B, int S, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const float* __restrict__ z, const float* __restrict__ weights, float sample_dist,
                   const float* __restrict__ g_scale, float* __restrict__ partial, float* __restrict__ d_w) {
  __shared__ float acc[2];
  if (threadIdx.x < 2) acc[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * FW + warp;
  float s = 0.f, c = 0.f;
  if (ray < B) {
    const float o[3] = {rays_o[ray * 3], rays_o[ray * 3 + 1], rays_o[ray * 3 + 2]};
    const float d[3] = {rays_d[ray * 3], rays_d[ray * 3 + 1], rays_d[ray * 3 + 2]};
    const float gs = g_scale ? g_scale[0] : 0.f;
    for (int j = lane; j < S; j += 32) {
      const float mid = mid_of(z + ray * S, j, S, sample_dist);
      const float p0 = o[0] + d[0] * mid, p1 = o[1] + d[1] * mid, p2 = o[2] + d[2] * mid;
      const bool out = sqrtf(p0 * p0 + p1 * p1 + p2 * p2) > 1.0f;
      const float w = weights[ray * S + j];
      if (out) { s += fabsf(w); c += 1.f; }
      if (d_w) d_w[ray * S + j] = out ? gs * ((w > 0.f) - (w < 0.f)) : 0.f;
    }
  }
  s = fsum(s); c = fsum(c);
  if (lane == 0) { atomicAdd(&acc[0], s); atomicAdd(&acc[1], c); }
  __syncthreads();
  if (partial && threadIdx.x < 2) atomicAdd(&partial[threadIdx.x], acc[threadIdx.x]);
}

}  // namespace fmov
using namespace fmov;

static int fill_flow(FlowArgs& a, long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                     float sample_dist, const float* weights, const float* w2c34, const float* K33, int k_stride,
                     const float* xy) {
  FMOV_REQUIRE(B >= 0 && S >= 1 && k_stride >= 3, "flow: bad sizes (B=%lld S=%d)", B, S);
  FMOV_REQUIRE(B == 0 || (rays_o && rays_d && z && weights && w2c34 && K33 && xy), "flow: null argument");
  memset(&a, 0, sizeof(a));
  a.B = B; a.S = S; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.weights = weights; a.w2c = w2c34; a.K = K33;
  a.k_stride = k_stride; a.xy = xy; a.sample_dist = sample_dist;
  return OK;
}

extern "C" int fmov_flow_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                             float sample_dist, const float* weights, const float* w2c34, const float* K33, int k_stride,
                             const float* xy, float* err, void* stream) {
  FlowArgs a;
  int st = fill_flow(a, B, S, rays_o, rays_d, z, sample_dist, weights, w2c34, K33, k_stride, xy);
  if (st) return st;
  if (B == 0) return OK;
  FMOV_REQUIRE(err, "fmov_flow_fwd: null output");
  flow_fwd_kernel<<<(unsigned)((B + FW - 1) / FW), FW * 32, 0, (cudaStream_t)stream>>>(a, err);
  FMOV_LAUNCH_CHECK("flow_fwd_kernel");
  return OK;
}

extern "C" int fmov_flow_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                             float sample_dist, const float* weights, const float* w2c34, const float* K33, int k_stride,
                             const float* xy, const float* g_err, float* d_weights, float* d_o, float* d_d, float* d_z,
                             float* d_w2c34, void* stream) {
  FlowArgs a;
  int st = fill_flow(a, B, S, rays_o, rays_d, z, sample_dist, weights, w2c34, K33, k_stride, xy);
  if (st) return st;
  FMOV_REQUIRE(d_w2c34, "fmov_flow_bwd: null output");
  FMOV_CUDA(cudaMemsetAsync(d_w2c34, 0, 12 * sizeof(float), (cudaStream_t)stream));
  if (B == 0) return OK;
  FMOV_REQUIRE(g_err && d_o && d_d, "fmov_flow_bwd: null argument");
  flow_bwd_kernel<<<(unsigned)((B + FW - 1) / FW), FW * 32, 0, (cudaStream_t)stream>>>(a, g_err, d_weights, d_o, d_d,
                                                                                         d_z, d_w2c34);
  FMOV_LAUNCH_CHECK("flow_bwd_kernel");
  return OK;
}

extern "C" int fmov_unit_sphere_fwd_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                                        float sample_dist, const float* weights, const float* g_scale, float* partial2,
                                        float* d_weights, void* stream) {
  FMOV_REQUIRE(B >= 0 && S >= 1, "fmov_unit_sphere_fwd_bwd: bad sizes");
  FMOV_REQUIRE(partial2 || d_weights, "fmov_unit_sphere_fwd_bwd: no output requested");
  FMOV_REQUIRE(!d_weights || g_scale, "fmov_unit_sphere_fwd_bwd: d_weights needs g_scale");
  if (partial2) FMOV_CUDA(cudaMemsetAsync(partial2, 0, 2 * sizeof(float), (cudaStream_t)stream));
  if (B == 0) return OK;
  FMOV_REQUIRE(rays_o && rays_d && z && weights, "fmov_unit_sphere_fwd_bwd: null argument");
  unit_sphere_kernel<<<(unsigned)((B + FW - 1) / FW), FW * 32, 0, (cudaStream_t)stream>>>(
      B, S, rays_o, rays_d, z, weights, sample_dist, g_scale, partial2, d_weights);
  FMOV_LAUNCH_CHECK("unit_sphere_kernel");
  return OK;
}
