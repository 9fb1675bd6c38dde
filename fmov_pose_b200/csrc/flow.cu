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

constexpr int FW = 4;     // warps (rays) per block
constexpr int FMAXK = 8;  // samples per lane (S <= 256): a lane's samples are loaded up front, fully unrolled, so that
                          // all of a ray's loads are in flight together (the kernels are HBM/latency bound)

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

// M = K [R_w | t_w] (3x4), kept in 12 registers per thread for the whole persistent loop: the per-sample projection is
// 12 FMAs from registers (measured: reading K and w2c from shared memory per sample made the kernels LDS-bound)
__device__ __forceinline__ void load_proj_matrix(const FlowArgs& a, float* M) {
  float W[12], K[9];
#pragma unroll
  for (int i = 0; i < 12; ++i) W[i] = __ldg(a.w2c + i);
#pragma unroll
  for (int i = 0; i < 9; ++i) K[i] = __ldg(a.K + (i / 3) * a.k_stride + i % 3);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) M[i * 4 + j] = K[i * 3] * W[j] + K[i * 3 + 1] * W[4 + j] + K[i * 3 + 2] * W[8 + j];
}

__device__ __forceinline__ Proj project(const float* o, const float* d, float mid, const float* M) {
  Proj r;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.p[i] = o[i] + d[i] * mid;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.q[i] = M[i * 4] * r.p[0] + M[i * 4 + 1] * r.p[1] + M[i * 4 + 2] * r.p[2] + M[i * 4 + 3];
  const float iq = __frcp_rn(r.q[2]);
  r.u = r.q[0] * iq;
  r.v = r.q[1] * iq;
  return r;
}

// One ray's samples of this lane: z_j and w_j for j = lane + 32k (coalesced: a warp reads 128 contiguous bytes per k).
template <int KM>
struct RaySamples {
  float z[KM], w[KM];
  __device__ __forceinline__ void load(const FlowArgs& a, long long ray, int lane) {
    const float* zr = a.z + ray * a.S;
    const float* wr = a.weights + ray * a.S;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int j = lane + 32 * k;
      z[k] = j < a.S ? __ldg(zr + j) : 0.f;
      w[k] = j < a.S ? __ldg(wr + j) : 0.f;
    }
  }
  // mid_z of sample j = lane + 32k (models/renderer.py:261-267); z_{j+1} comes from the next lane (lane 31: lane 0, k+1)
  __device__ __forceinline__ float mid(int k, int lane, int S, float sample_dist) const {
    float zn = __shfl_down_sync(0xffffffffu, z[k], 1);
    const float wrap = __shfl_sync(0xffffffffu, k + 1 < KM ? z[k + 1 < KM ? k + 1 : k] : 0.f, 0);
    if (lane == 31) zn = wrap;
    const int j = lane + 32 * k;
    return z[k] + 0.5f * ((j + 1 < S) ? zn - z[k] : sample_dist);
  }
};

// Persistent, warp per ray; the next ray's samples are loaded while the current ray is evaluated (the kernels are
// latency bound: bytes in flight per SM = registers holding outstanding loads).
template <int KM>
__global__ void __launch_bounds__(FW * 32) flow_fwd_kernel(FlowArgs a, float* __restrict__ err) {
  float M[12];
  load_proj_matrix(a, M);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * FW;
  long long ray = (long long)blockIdx.x * FW + warp;
  RaySamples<KM> cur, nxt;
  if (ray < a.B) cur.load(a, ray, lane);
  for (; ray < a.B; ray += stride) {
    if (ray + stride < a.B) nxt.load(a, ray + stride, lane);
    const float o[3] = {__ldg(a.rays_o + ray * 3), __ldg(a.rays_o + ray * 3 + 1), __ldg(a.rays_o + ray * 3 + 2)};
    const float d[3] = {__ldg(a.rays_d + ray * 3), __ldg(a.rays_d + ray * 3 + 1), __ldg(a.rays_d + ray * 3 + 2)};
    const float tx = __ldg(a.xy + ray * 2), ty = __ldg(a.xy + ray * 2 + 1);
    float eu = 0.f, ev = 0.f;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const float mid = cur.mid(k, lane, a.S, a.sample_dist);
      if (lane + 32 * k < a.S) {
        const Proj pr = project(o, d, mid, M);
        eu += (pr.u - tx) * cur.w[k];
        ev += (pr.v - ty) * cur.w[k];
      }
    }
    eu = fsum(eu); ev = fsum(ev);
    if (lane == 0) { err[ray * 2] = eu; err[ray * 2 + 1] = ev; }
    cur = nxt;
  }
}

// dq = dL/d(K cam) of one sample given du = g0*w, dv = g1*w;  dp = (M[:, :3])^T dq
__device__ __forceinline__ void proj_bwd(const Proj& pr, float du, float dv, const float* M, float* dq, float* dp) {
  const float iq = __frcp_rn(pr.q[2]);
  dq[0] = du * iq;
  dq[1] = dv * iq;
  dq[2] = -(du * pr.q[0] + dv * pr.q[1]) * iq * iq;
#pragma unroll
  for (int i = 0; i < 3; ++i) dp[i] = M[i] * dq[0] + M[4 + i] * dq[1] + M[8 + i] * dq[2];
}

template <int KM>
__global__ void __launch_bounds__(FW * 32)
flow_bwd_kernel(FlowArgs a, const float* __restrict__ g_err, float* __restrict__ d_w, float* __restrict__ d_o,
                float* __restrict__ d_d, float* __restrict__ d_z, float* __restrict__ d_w2c) {
  __shared__ float acc[12];
  if (threadIdx.x < 12) acc[threadIdx.x] = 0.f;
  __syncthreads();
  float M[12];
  load_proj_matrix(a, M);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float G[12];        // sum over this lane's samples of dq (x) [p; 1]; d w2c = K^T G, applied once per block
#pragma unroll
  for (int i = 0; i < 12; ++i) G[i] = 0.f;
  const long long stride = (long long)gridDim.x * FW;
  long long ray = (long long)blockIdx.x * FW + warp;
  RaySamples<KM> cur, nxt;
  if (ray < a.B) cur.load(a, ray, lane);
  for (; ray < a.B; ray += stride) {
    if (ray + stride < a.B) nxt.load(a, ray + stride, lane);
    const float o[3] = {__ldg(a.rays_o + ray * 3), __ldg(a.rays_o + ray * 3 + 1), __ldg(a.rays_o + ray * 3 + 2)};
    const float d[3] = {__ldg(a.rays_d + ray * 3), __ldg(a.rays_d + ray * 3 + 1), __ldg(a.rays_d + ray * 3 + 2)};
    const float tx = __ldg(a.xy + ray * 2), ty = __ldg(a.xy + ray * 2 + 1);
    const float g0 = __ldg(g_err + ray * 2), g1 = __ldg(g_err + ray * 2 + 1);
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    float carry = 0.f;      // dL/dmid of sample 32k-1 (lane 31 of the previous 32-sample group), consumed by lane 0
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int j = lane + 32 * k;
      const bool on = j < a.S;
      const float mid = cur.mid(k, lane, a.S, a.sample_dist);
      float gmid = 0.f;
      if (on) {
        const Proj pr = project(o, d, mid, M);
        const float w = cur.w[k];
        if (d_w) d_w[ray * a.S + j] = g0 * (pr.u - tx) + g1 * (pr.v - ty);
        float dq[3], dp[3];
        proj_bwd(pr, g0 * w, g1 * w, M, dq, dp);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          go[i] += dp[i];
          gd[i] += dp[i] * mid;
          G[i * 4] += dq[i] * pr.p[0]; G[i * 4 + 1] += dq[i] * pr.p[1]; G[i * 4 + 2] += dq[i] * pr.p[2];
          G[i * 4 + 3] += dq[i];
        }
        gmid = dp[0] * d[0] + dp[1] * d[1] + dp[2] * d[2];
      }
      if (d_z) {
        // mid_j = (z_j + z_{j+1})/2 (last: z_j + sample_dist/2): z_j enters mid_j and mid_{j-1}; the neighbour's
        // dL/dmid comes from the previous lane (lane 0: lane 31 of the previous 32-sample group)
        float prev = __shfl_up_sync(0xffffffffu, gmid, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, gmid, 31);
        if (on) d_z[ray * a.S + j] = ((j + 1 < a.S) ? 0.5f * gmid : gmid) + 0.5f * prev;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { go[i] = fsum(go[i]); gd[i] = fsum(gd[i]); }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { d_o[ray * 3 + i] = go[i]; d_d[ray * 3 + i] = gd[i]; }
    }
    cur = nxt;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float s = fsum(G[i]);
    if (lane == 0) atomicAdd(&acc[i], s);
  }
  __syncthreads();
  if (threadIdx.x < 12) {      // d w2c[i][j] = sum_k K[k][i] * G[k][j]
    const int i = threadIdx.x >> 2, j = threadIdx.x & 3;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v += __ldg(a.K + k * a.k_stride + i) * acc[k * 4 + j];
    atomicAdd(&d_w2c[threadIdx.x], v);
  }
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
  float s = 0.f, c = 0.f;
  for (long long ray = (long long)blockIdx.x * FW + warp; ray < B; ray += (long long)gridDim.x * FW) {
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

// persistent grids: SM count x resident blocks per SM (occupancy query, cached per kernel), capped by the ray groups
template <typename Kern>
static unsigned flow_grid(long long B, Kern kern, int* cache) {
  if (*cache == 0) {
    int dev = 0, sms = 148, per_sm = 4;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FW * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    *cache = sms * per_sm;
  }
  const long long want = (B + FW - 1) / FW;
  return (unsigned)(want < *cache ? want : *cache);
}
static int g_grid_fwd[2], g_grid_bwd[2], g_grid_unit;

static int fill_flow(FlowArgs& a, long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                     float sample_dist, const float* weights, const float* w2c34, const float* K33, int k_stride,
                     const float* xy) {
  FMOV_REQUIRE(B >= 0 && S >= 1 && S <= 32 * FMAXK && k_stride >= 3, "flow: bad sizes (B=%lld, S=%d of 1..%d)", B, S,
               32 * FMAXK);
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
  if (S <= 128)
    flow_fwd_kernel<4><<<flow_grid(B, flow_fwd_kernel<4>, &g_grid_fwd[0]), FW * 32, 0, (cudaStream_t)stream>>>(a, err);
  else
    flow_fwd_kernel<FMAXK><<<flow_grid(B, flow_fwd_kernel<FMAXK>, &g_grid_fwd[1]), FW * 32, 0, (cudaStream_t)stream>>>(a, err);
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
  if (S <= 128)
    flow_bwd_kernel<4><<<flow_grid(B, flow_bwd_kernel<4>, &g_grid_bwd[0]), FW * 32, 0, (cudaStream_t)stream>>>(
        a, g_err, d_weights, d_o, d_d, d_z, d_w2c34);
  else
    flow_bwd_kernel<FMAXK><<<flow_grid(B, flow_bwd_kernel<FMAXK>, &g_grid_bwd[1]), FW * 32, 0, (cudaStream_t)stream>>>(
        a, g_err, d_weights, d_o, d_d, d_z, d_w2c34);
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
  unit_sphere_kernel<<<flow_grid(B, unit_sphere_kernel, &g_grid_unit), FW * 32, 0, (cudaStream_t)stream>>>(
      B, S, rays_o, rays_d, z, weights, sample_dist, g_scale, partial2, d_weights);
  FMOV_LAUNCH_CHECK("unit_sphere_kernel");
  return OK;
}
