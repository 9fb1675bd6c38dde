// Ray generation fused with the per-frame SE(3) pose correction, forward and backward.
//
//   pose modes
//     0  c2w[3,4] given
//     1  LearnPoseGF tail: c2w = [Exp(rot) | trans] @ [R0 | scale*t0]     models/picture_pose.py:176-186,
//        Exp = Rodrigues with theta = |r| + 1e-15                          models/batch_lie_group_helper.py:19-47
//     2  BARF: c2w = compose_pair(se3_to_SE3(wu), noise_pose)              models/camera.py:89-102, 53-60;
//        call site exp_runner.py:419-424 (10-term Taylor A/B/C, camera.py:130-156)
//   rays   p = K^-1 [x,y,1]; v = p/|p|; rays_d = R v; rays_o = t              models/dataset.py:656-671
//   near/far = mid -/+ 1, mid = -(o.d)/(d.d)                                  models/dataset.py:835-842
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

struct PoseParams {
  int mode;
  const float* c2w;      // mode 0: [3,4] (row stride 4)
  const float* rot;      // mode 1: [3]
  const float* trans;    // mode 1: [3]
  const float* scale;    // mode 1: [1] or null
  const float* init;     // mode 1: init_c2w rows [3][4] (row stride 4);  mode 2: noise pose [3][4]
  const float* se3;      // mode 2: [6] (w, u)
};

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {   // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void skew(const float* r, float* K) {
  K[0] = 0.f; K[1] = -r[2]; K[2] = r[1];
  K[3] = r[2]; K[4] = 0.f; K[5] = -r[0];
  K[6] = -r[1]; K[7] = r[0]; K[8] = 0.f;
}
__device__ __forceinline__ void unskew_add(const float* Kb, float* rb) {
  rb[0] += Kb[7] - Kb[5];
  rb[1] += Kb[2] - Kb[6];
  rb[2] += Kb[3] - Kb[1];
}
__device__ __forceinline__ float dot9(const float* A, const float* B) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += A[i] * B[i];
  return s;
}
// Kbar += c * (G K^T + K^T G)   (backward of K @ K)
__device__ __forceinline__ void kk_bwd_add(const float* G, const float* K, float c, float* Kb) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < 3; ++m) s += G[i * 3 + m] * K[j * 3 + m] + K[m * 3 + i] * G[m * 3 + j];
      Kb[i * 3 + j] += c * s;
    }
}

// ---- Rodrigues ------------------------------------------------------------------------------
__device__ void rodrigues_fwd(const float* r, float* R) {
  float K[9], K2[9];
  skew(r, K);
  mat3_mul(K, K, K2);
  const float th = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) + 1e-15f;
  const float a = sinf(th) / th, b = (1.f - cosf(th)) / (th * th);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + a * K[i] + b * K2[i];
}
__device__ void rodrigues_bwd(const float* r, const float* G, float* rb) {
  float K[9], K2[9], Kb[9];
  skew(r, K);
  mat3_mul(K, K, K2);
  const float nr = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const float th = nr + 1e-15f;
  const float s = sinf(th), c = cosf(th);
  const float a = s / th, b = (1.f - c) / (th * th);
  const float da = (c * th - s) / (th * th);
  const float db = (s * th * th - (1.f - c) * 2.f * th) / (th * th * th * th);
  const float thb = dot9(G, K) * da + dot9(G, K2) * db;
#pragma unroll
  for (int i = 0; i < 9; ++i) Kb[i] = a * G[i];
  kk_bwd_add(G, K, b, Kb);
  rb[0] = rb[1] = rb[2] = 0.f;
  if (nr > 0.f) { rb[0] = thb * r[0] / nr; rb[1] = thb * r[1] / nr; rb[2] = thb * r[2] / nr; }
  unskew_add(Kb, rb);
}

// ---- se3 exponential with the reference's 10-term Taylor series ------------------------------
__device__ void taylor_abc(float x, float* A, float* B, float* C, float* dA, float* dB, float* dC) {
  float a = 0.f, b = 0.f, c = 0.f, da = 0.f, db = 0.f, dc = 0.f;
  float denA = 1.f, denB = 1.f, denC = 1.f;
  float xp = 1.f;           // x^(2i)
  float xpm = 0.f;          // x^(2i-1)
  for (int i = 0; i <= 10; ++i) {
    if (i > 0) denA *= (float)((2 * i) * (2 * i + 1));
    denB *= (float)((2 * i + 1) * (2 * i + 2));
    denC *= (float)((2 * i + 2) * (2 * i + 3));
    const float sgn = (i & 1) ? -1.f : 1.f;
    a += sgn * xp / denA; b += sgn * xp / denB; c += sgn * xp / denC;
    if (i > 0) {
      const float d = sgn * (float)(2 * i) * xpm;
      da += d / denA; db += d / denB; dc += d / denC;
    }
    xpm = xp * x;           // x^(2i+1)
    xp = xpm * x;           // x^(2i+2)
  }
  *A = a; *B = b; *C = c; *dA = da; *dB = db; *dC = dc;
}
__device__ void se3_fwd(const float* wu, float* R, float* t) {
  float K[9], K2[9], V[9];
  skew(wu, K);
  mat3_mul(K, K, K2);
  const float th = sqrtf(wu[0] * wu[0] + wu[1] * wu[1] + wu[2] * wu[2]);
  float A, B, C, dA, dB, dC;
  taylor_abc(th, &A, &B, &C, &dA, &dB, &dC);
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float I = (i % 4 == 0) ? 1.f : 0.f;
    R[i] = I + A * K[i] + B * K2[i];
    V[i] = I + B * K[i] + C * K2[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * wu[3] + V[i * 3 + 1] * wu[4] + V[i * 3 + 2] * wu[5];
}
__device__ void se3_bwd(const float* wu, const float* GR, const float* gt, float* wub) {
  float K[9], K2[9], V[9], Vb[9], Kb[9];
  skew(wu, K);
  mat3_mul(K, K, K2);
  const float th = sqrtf(wu[0] * wu[0] + wu[1] * wu[1] + wu[2] * wu[2]);
  float A, B, C, dA, dB, dC;
  taylor_abc(th, &A, &B, &C, &dA, &dB, &dC);
#pragma unroll
  for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.f : 0.f) + B * K[i] + C * K2[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Vb[i * 3 + j] = gt[i] * wu[3 + j];
#pragma unroll
  for (int j = 0; j < 3; ++j) wub[3 + j] = V[j] * gt[0] + V[3 + j] * gt[1] + V[6 + j] * gt[2];
  const float Ab = dot9(GR, K), Bb = dot9(GR, K2) + dot9(Vb, K), Cb = dot9(Vb, K2);
#pragma unroll
  for (int i = 0; i < 9; ++i) Kb[i] = A * GR[i] + B * Vb[i];
  kk_bwd_add(GR, K, B, Kb);
  kk_bwd_add(Vb, K, C, Kb);
  const float thb = Ab * dA + Bb * dB + Cb * dC;
  wub[0] = wub[1] = wub[2] = 0.f;
  if (th > 0.f) { wub[0] = thb * wu[0] / th; wub[1] = thb * wu[1] / th; wub[2] = thb * wu[2] / th; }
  unskew_add(Kb, wub);
}

// pose (R[9], t[3]) from the parameters
__device__ void pose_fwd(const PoseParams& pp, float* R, float* t) {
  if (pp.mode == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = pp.c2w[i * 4 + j];
      t[i] = pp.c2w[i * 4 + 3];
    }
  } else if (pp.mode == 1) {
    float r[3] = {pp.rot[0], pp.rot[1], pp.rot[2]}, Re[9], R0[9], t0[3];
    rodrigues_fwd(r, Re);
    const float s = pp.scale ? pp.scale[0] : 1.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) R0[i * 3 + j] = pp.init[i * 4 + j];
      t0[i] = pp.init[i * 4 + 3] * s;
    }
    mat3_mul(Re, R0, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = Re[i * 3] * t0[0] + Re[i * 3 + 1] * t0[1] + Re[i * 3 + 2] * t0[2] + pp.trans[i];
  } else {
    float wu[6], Ra[9], ta[3], Rb[9], tb[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) wu[i] = pp.se3[i];
    se3_fwd(wu, Ra, ta);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) Rb[i * 3 + j] = pp.init[i * 4 + j];
      tb[i] = pp.init[i * 4 + 3];
    }
    mat3_mul(Rb, Ra, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = Rb[i * 3] * ta[0] + Rb[i * 3 + 1] * ta[1] + Rb[i * 3 + 2] * ta[2] + tb[i];
  }
}

__global__ void pose_fwd_kernel(PoseParams pp, float* c2w34) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float R[9], t[3];
    pose_fwd(pp, R, t);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) c2w34[i * 4 + j] = R[i * 3 + j];
      c2w34[i * 4 + 3] = t[i];
    }
  }
}

// dL/dc2w[3,4] -> parameter gradients (single thread)
__global__ void pose_bwd_kernel(PoseParams pp, const float* g34, float* g_rot, float* g_trans, float* g_scale,
                                float* g_se3) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float GR[9], gt[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) GR[i * 3 + j] = g34[i * 4 + j];
    gt[i] = g34[i * 4 + 3];
  }
  if (pp.mode == 1) {
    float r[3] = {pp.rot[0], pp.rot[1], pp.rot[2]}, Re[9], R0[9], t0[3], Gre[9], rb[3];
    rodrigues_fwd(r, Re);
    const float s = pp.scale ? pp.scale[0] : 1.f;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) R0[i * 3 + j] = pp.init[i * 4 + j];
      t0[i] = pp.init[i * 4 + 3];
    }
    // Rn = Re R0 ; tn = Re (s t0) + trans
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float v = 0.f;
        for (int m = 0; m < 3; ++m) v += GR[i * 3 + m] * R0[j * 3 + m];
        Gre[i * 3 + j] = v + gt[i] * s * t0[j];
      }
    rodrigues_bwd(r, Gre, rb);
    for (int i = 0; i < 3; ++i) { g_rot[i] = rb[i]; g_trans[i] = gt[i]; }
    if (g_scale) {
      float v = 0.f;
      for (int j = 0; j < 3; ++j) v += (Re[j] * gt[0] + Re[3 + j] * gt[1] + Re[6 + j] * gt[2]) * t0[j];
      g_scale[0] = v;
    }
  } else if (pp.mode == 2) {
    float wu[6], Rb[9], Ga[9], gta[3], wub[6];
    for (int i = 0; i < 6; ++i) wu[i] = pp.se3[i];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rb[i * 3 + j] = pp.init[i * 4 + j];
    // R = Rb Ra ; t = Rb ta + tb   ->  Ga = Rb^T GR ; gta = Rb^T gt
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        float v = 0.f;
        for (int m = 0; m < 3; ++m) v += Rb[m * 3 + i] * GR[m * 3 + j];
        Ga[i * 3 + j] = v;
      }
      gta[i] = Rb[i] * gt[0] + Rb[3 + i] * gt[1] + Rb[6 + i] * gt[2];
    }
    se3_bwd(wu, Ga, gta, wub);
    for (int i = 0; i < 6; ++i) g_se3[i] = wub[i];
  }
}

struct RayArgs {
  PoseParams pp;
  const float* intr_inv;     // [3,3] (row stride given)
  int intr_stride;
  const long long* px; const long long* py;
  const float* pxf; const float* pyf;      // sub-pixel coordinates (flow matches); used when px == nullptr
  long long B;
  float* rays_o; float* rays_d; float* near; float* far; float* c2w_out;
};

__global__ void raygen_fwd_kernel(RayArgs a) {
  __shared__ float sR[9], st[3], sK[9];
  if (threadIdx.x == 0) {
    pose_fwd(a.pp, sR, st);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) sK[i * 3 + j] = a.intr_inv[i * a.intr_stride + j];
    if (blockIdx.x == 0 && a.c2w_out)
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) a.c2w_out[i * 4 + j] = sR[i * 3 + j];
        a.c2w_out[i * 4 + 3] = st[i];
      }
  }
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.B) return;
  const float x = a.px ? (float)a.px[r] : a.pxf[r], y = a.px ? (float)a.py[r] : a.pyf[r];
  float p[3], v[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = sK[i * 3] * x + sK[i * 3 + 1] * y + sK[i * 3 + 2];
  const float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = p[i] / pn;
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = sR[i * 3] * v[0] + sR[i * 3 + 1] * v[1] + sR[i * 3 + 2] * v[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) { a.rays_o[r * 3 + i] = st[i]; a.rays_d[r * 3 + i] = d[i]; }
  // Dataset.near_far_from_sphere (models/dataset.py:835-842) in torch's own operation order and rounding: a = sum(d**2),
  // b = 2 * sum(o*d), mid = 0.5 * (-b) / a with every product and sum rounded separately (no FMA contraction), so that the
  // fused path draws bit-identical samples to the torch expression evaluated on the same rays
  const float aa = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
  const float bb = __fmul_rn(2.f, __fadd_rn(__fadd_rn(__fmul_rn(st[0], d[0]), __fmul_rn(st[1], d[1])), __fmul_rn(st[2], d[2])));
  const float mid = __fdiv_rn(__fmul_rn(0.5f, -bb), aa);
  if (a.near) a.near[r] = mid - 1.f;
  if (a.far) a.far[r] = mid + 1.f;
}

// d(c2w)[3,4] += sum over rays; g_near/g_far optional.
__global__ void raygen_bwd_kernel(RayArgs a, const float* __restrict__ g_o, const float* __restrict__ g_d,
                                  const float* __restrict__ g_near, const float* __restrict__ g_far, float* g34) {
  __shared__ float sK[9];
  __shared__ float acc[12];
  if (threadIdx.x < 12) acc[threadIdx.x] = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) sK[i * 3 + j] = a.intr_inv[i * a.intr_stride + j];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float G[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) G[i] = 0.f;
  if (r < a.B) {
    const float x = a.px ? (float)a.px[r] : a.pxf[r], y = a.px ? (float)a.py[r] : a.pyf[r];
    float p[3], v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = sK[i * 3] * x + sK[i * 3 + 1] * y + sK[i * 3 + 2];
    const float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = p[i] / pn;
    float go[3] = {g_o ? g_o[r * 3] : 0.f, g_o ? g_o[r * 3 + 1] : 0.f, g_o ? g_o[r * 3 + 2] : 0.f};
    float gd[3] = {g_d ? g_d[r * 3] : 0.f, g_d ? g_d[r * 3 + 1] : 0.f, g_d ? g_d[r * 3 + 2] : 0.f};
    const float gm = (g_near ? g_near[r] : 0.f) + (g_far ? g_far[r] : 0.f);
    if (gm != 0.f) {
      // mid = -(o.d)/(d.d)
      const float* o = a.rays_o + r * 3;
      const float* d = a.rays_d + r * 3;
      const float aa = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const float od = o[0] * d[0] + o[1] * d[1] + o[2] * d[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        go[i] += gm * (-d[i] / aa);
        gd[i] += gm * (-o[i] / aa + 2.f * od * d[i] / (aa * aa));
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) G[i * 4 + j] = gd[i] * v[j];
      G[i * 4 + 3] = go[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float s = G[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[i], s);
  }
  __syncthreads();
  if (threadIdx.x < 12) atomicAdd(&g34[threadIdx.x], acc[threadIdx.x]);
}


// =====================================================================================================
// LearnPoseGF as ONE launch per direction (models/picture_pose.py:140-186): Gaussian-Fourier features of the frame
// index -> Linear 256->64, GELU -> Linear 64->64, GELU -> heads (lin3 [6] | lin3_rot [3], lin3_trans [3], lin3_scale [1])
// -> Rodrigues tail (mode 1 above).  21 k parameters: one block; the torch formulation is ~15 launches forward and ~30
// backward per frame, which dominates the shipped 512-ray iterations.
// =====================================================================================================
struct PoseGfArgs {
  const long long* cid;      // device int64 [1]: frame index
  const float* b;            // [128] Fourier frequencies
  const float* W1; const float* b1;      // [64,256], [64]
  const float* W2; const float* b2;      // [64,64], [64]
  const float* Wh[3]; const float* bh[3];   // heads, rows concatenated in order: rot(3) trans(3) [scale(1)]
  int rows[3];
  int n_heads;
  float rot_k;               // pi or pi/6 (small_rot)
  const float* init_all;     // [N,4,4] initial poses (or null: identity)
  float* save;               // [256 ff | 64 z1 | 64 z2 | 8 out(rot*k, trans, scale)] = 392 floats
};
constexpr int GF_SAVE = 392;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * expf(-0.5f * x * x) * 0.39894228040143267794f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) pose_gf_fwd_kernel(PoseGfArgs a, float* __restrict__ c2w34) {
  __shared__ float ff[256], z1[64], h1[64], z2[64], h2[64], out[8], init[12];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long cid = a.cid[0];
  if (t < 128) {
    const float x = 6.283185307179586f * (float)cid;      // (2*pi*cid) in fp32, as the reference's tensor expression
    const float ang = x * a.b[t];
    ff[t] = sinf(ang) / 11.313708498984761f;              // / sqrt(embedding_size = 128)
    ff[128 + t] = cosf(ang) / 11.313708498984761f;
  }
  if (t < 12) init[t] = a.init_all ? a.init_all[cid * 16 + t] : ((t % 5) == 0 ? 1.f : 0.f);
  if (t < 8) out[t] = 0.f;
  __syncthreads();
  for (int j = warp * 8; j < warp * 8 + 8; ++j) {          // lin1: 8 output rows per warp
    float s = 0.f;
#pragma unroll
    for (int i = lane; i < 256; i += 32) s = fmaf(a.W1[j * 256 + i], ff[i], s);
    s = warp_sum(s);
    if (lane == 0) { z1[j] = s + a.b1[j]; h1[j] = gelu_erf(z1[j]); }
  }
  __syncthreads();
  for (int j = warp * 8; j < warp * 8 + 8; ++j) {          // lin2
    float s = fmaf(a.W2[j * 64 + lane], h1[lane], a.W2[j * 64 + 32 + lane] * h1[32 + lane]);
    s = warp_sum(s);
    if (lane == 0) { z2[j] = s + a.b2[j]; h2[j] = gelu_erf(z2[j]); }
  }
  __syncthreads();
  if (warp < 7) {                                          // heads: one output row per warp
    int r = warp, hd = 0, base = 0;
    while (hd < a.n_heads && r >= a.rows[hd]) { r -= a.rows[hd]; base += a.rows[hd]; ++hd; }
    if (hd < a.n_heads) {
      float s = fmaf(a.Wh[hd][r * 64 + lane], h2[lane], a.Wh[hd][r * 64 + 32 + lane] * h2[32 + lane]);
      s = warp_sum(s);
      if (lane == 0) out[base + r] = (s + a.bh[hd][r]) * (base + r < 3 ? a.rot_k : 1.f);
    }
  }
  __syncthreads();
  if (t < 256) a.save[t] = ff[t];
  if (t < 64) { a.save[256 + t] = z1[t]; a.save[320 + t] = z2[t]; }
  if (t < 8) a.save[384 + t] = out[t];
  if (t == 0) {
    int total = 0;
    for (int h = 0; h < a.n_heads; ++h) total += a.rows[h];
    PoseParams pp;
    pp.mode = 1; pp.c2w = nullptr; pp.rot = out; pp.trans = out + 3; pp.scale = total > 6 ? out + 6 : nullptr;
    pp.init = init; pp.se3 = nullptr;
    float R[9], tt[3];
    pose_fwd(pp, R, tt);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) c2w34[i * 4 + j] = R[i * 3 + j];
      c2w34[i * 4 + 3] = tt[i];
    }
  }
}

struct PoseGfGrads {
  float* dW1; float* db1; float* dW2; float* db2;      // any may be null (frozen parameters)
  float* dWh[3]; float* dbh[3];
};

__global__ void __launch_bounds__(256)
pose_gf_bwd_kernel(PoseGfArgs a, const float* __restrict__ g34, PoseGfGrads g) {
  __shared__ float ff[256], z1[64], h1[64], z2[64], h2[64], gout[8], gz2[64], gz1[64], init[12], outv[8];
  const int t = threadIdx.x;
  const long long cid = a.cid[0];
  ff[t] = a.save[t];
  if (t < 64) {
    z1[t] = a.save[256 + t]; z2[t] = a.save[320 + t];
    h1[t] = gelu_erf(z1[t]); h2[t] = gelu_erf(z2[t]);
  }
  if (t < 8) { outv[t] = a.save[384 + t]; gout[t] = 0.f; }
  if (t < 12) init[t] = a.init_all ? a.init_all[cid * 16 + t] : ((t % 5) == 0 ? 1.f : 0.f);
  __syncthreads();
  int total = 0;
  for (int h = 0; h < a.n_heads; ++h) total += a.rows[h];
  if (t == 0) {      // tail backward (same maths as pose_bwd_kernel mode 1)
    float GR[9], gt[3], Re[9], R0[9], t0[3], Gre[9], rb[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) { GR[i * 3 + j] = g34[i * 4 + j]; R0[i * 3 + j] = init[i * 4 + j]; }
      gt[i] = g34[i * 4 + 3];
      t0[i] = init[i * 4 + 3];
    }
    const float r[3] = {outv[0], outv[1], outv[2]};
    rodrigues_fwd(r, Re);
    const float sc = total > 6 ? outv[6] : 1.f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float v = 0.f;
        for (int m = 0; m < 3; ++m) v += GR[i * 3 + m] * R0[j * 3 + m];
        Gre[i * 3 + j] = v + gt[i] * sc * t0[j];
      }
    rodrigues_bwd(r, Gre, rb);
    for (int i = 0; i < 3; ++i) { gout[i] = rb[i] * a.rot_k; gout[3 + i] = gt[i]; }
    if (total > 6) {
      float v = 0.f;
      for (int j = 0; j < 3; ++j) v += (Re[j] * gt[0] + Re[3 + j] * gt[1] + Re[6 + j] * gt[2]) * t0[j];
      gout[6] = v;
    }
  }
  __syncthreads();
  // heads: dWh[r][i] = gout[r] * h2[i], dbh[r] = gout[r];  g_h2[i] = sum_r Wh[r][i] gout[r]
  {
    int base = 0;
    for (int hd = 0; hd < a.n_heads; ++hd) {
      const int n = a.rows[hd] * 64;
      if (g.dWh[hd] && t < n) g.dWh[hd][t] = gout[base + t / 64] * h2[t % 64];
      if (n > 256 && g.dWh[hd] && t + 256 < n) g.dWh[hd][t + 256] = gout[base + (t + 256) / 64] * h2[(t + 256) % 64];
      if (g.dbh[hd] && t < a.rows[hd]) g.dbh[hd][t] = gout[base + t];
      base += a.rows[hd];
    }
  }
  if (t < 64) {
    float s = 0.f;
    int base = 0;
    for (int hd = 0; hd < a.n_heads; ++hd) {
      for (int r = 0; r < a.rows[hd]; ++r) s = fmaf(a.Wh[hd][r * 64 + t], gout[base + r], s);
      base += a.rows[hd];
    }
    gz2[t] = s * gelu_erf_grad(z2[t]);
    if (g.db2) g.db2[t] = gz2[t];
  }
  __syncthreads();
  if (g.dW2)
    for (int e = t; e < 64 * 64; e += 256) g.dW2[e] = gz2[e >> 6] * h1[e & 63];
  if (t < 64) {
    float s = 0.f;
    for (int j = 0; j < 64; ++j) s = fmaf(a.W2[j * 64 + t], gz2[j], s);
    gz1[t] = s * gelu_erf_grad(z1[t]);
    if (g.db1) g.db1[t] = gz1[t];
  }
  __syncthreads();
  if (g.dW1)
    for (int e = t; e < 64 * 256; e += 256) g.dW1[e] = gz1[e >> 8] * ff[e & 255];
}

}  // namespace fmov
using namespace fmov;

static int make_pose(PoseParams& pp, int mode, const float* c2w, const float* rot, const float* trans, const float* scale,
                     const float* init, const float* se3) {
  FMOV_REQUIRE(mode >= 0 && mode <= 2, "pose: unknown mode %d", mode);
  FMOV_REQUIRE((mode != 0 || c2w) && (mode != 1 || (rot && trans && init)) && (mode != 2 || (se3 && init)),
               "pose: missing parameters for mode %d", mode);
  pp.mode = mode; pp.c2w = c2w; pp.rot = rot; pp.trans = trans; pp.scale = scale; pp.init = init; pp.se3 = se3;
  return OK;
}

extern "C" int fmov_pose_fwd(int mode, const float* rot, const float* trans, const float* scale, const float* init34,
                             const float* se3, float* c2w34, void* stream) {
  PoseParams pp;
  FMOV_REQUIRE(mode == 1 || mode == 2, "fmov_pose_fwd: mode must be 1 (Rodrigues) or 2 (se3)");
  int st = make_pose(pp, mode, nullptr, rot, trans, scale, init34, se3);
  if (st) return st;
  FMOV_REQUIRE(c2w34, "fmov_pose_fwd: null output");
  pose_fwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, c2w34);
  FMOV_LAUNCH_CHECK("pose_fwd_kernel");
  return OK;
}

extern "C" int fmov_pose_bwd(int mode, const float* rot, const float* trans, const float* scale, const float* init34,
                             const float* se3, const float* g_c2w34, float* g_rot, float* g_trans, float* g_scale,
                             float* g_se3, void* stream) {
  PoseParams pp;
  FMOV_REQUIRE(mode == 1 || mode == 2, "fmov_pose_bwd: mode must be 1 (Rodrigues) or 2 (se3)");
  int st = make_pose(pp, mode, nullptr, rot, trans, scale, init34, se3);
  if (st) return st;
  FMOV_REQUIRE(g_c2w34 && ((mode == 1 && g_rot && g_trans) || (mode == 2 && g_se3)), "fmov_pose_bwd: null argument");
  pose_bwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, g_c2w34, g_rot, g_trans, g_scale, g_se3);
  FMOV_LAUNCH_CHECK("pose_bwd_kernel");
  return OK;
}

extern "C" int fmov_raygen_fwd(int mode, const float* c2w34, const float* rot, const float* trans, const float* scale,
                               const float* init34, const float* se3, const float* intr_inv, int intr_stride,
                               const long long* px, const long long* py, long long B, float* rays_o, float* rays_d,
                               float* near, float* far, float* c2w_out, void* stream) {
  RayArgs a;
  memset(&a, 0, sizeof(a));
  int st = make_pose(a.pp, mode, c2w34, rot, trans, scale, init34, se3);
  if (st) return st;
  FMOV_REQUIRE(B >= 0 && intr_stride >= 3, "fmov_raygen_fwd: bad sizes");
  if (B == 0) return OK;
  FMOV_REQUIRE(intr_inv && px && py && rays_o && rays_d, "fmov_raygen_fwd: null argument");
  a.intr_inv = intr_inv; a.intr_stride = intr_stride; a.px = px; a.py = py; a.B = B;
  a.rays_o = rays_o; a.rays_d = rays_d; a.near = near; a.far = far; a.c2w_out = c2w_out;
  raygen_fwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("raygen_fwd_kernel");
  return OK;
}

extern "C" int fmov_raygen_bwd(const float* intr_inv, int intr_stride, const long long* px, const long long* py,
                               long long B, const float* rays_o, const float* rays_d, const float* g_o, const float* g_d,
                               const float* g_near, const float* g_far, float* g_c2w34, void* stream) {
  RayArgs a;
  memset(&a, 0, sizeof(a));
  FMOV_REQUIRE(B >= 0 && intr_stride >= 3, "fmov_raygen_bwd: bad sizes");
  FMOV_REQUIRE(g_c2w34, "fmov_raygen_bwd: null output");
  FMOV_CUDA(cudaMemsetAsync(g_c2w34, 0, 12 * sizeof(float), (cudaStream_t)stream));
  if (B == 0) return OK;
  FMOV_REQUIRE(intr_inv && px && py && rays_o && rays_d, "fmov_raygen_bwd: null argument");
  a.intr_inv = intr_inv; a.intr_stride = intr_stride; a.px = px; a.py = py; a.B = B;
  a.rays_o = const_cast<float*>(rays_o); a.rays_d = const_cast<float*>(rays_d);
  raygen_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, g_o, g_d, g_near, g_far, g_c2w34);
  FMOV_LAUNCH_CHECK("raygen_bwd_kernel");
  return OK;
}

/* float-pixel variants (LoFTR flow matches are sub-pixel: models/dataset.py:712-715, 745-760) */
extern "C" int fmov_raygen_xy_fwd(int mode, const float* c2w34, const float* rot, const float* trans, const float* scale,
                                  const float* init34, const float* se3, const float* intr_inv, int intr_stride,
                                  const float* px, const float* py, long long B, float* rays_o, float* rays_d,
                                  float* near, float* far, float* c2w_out, void* stream) {
  RayArgs a;
  memset(&a, 0, sizeof(a));
  int st = make_pose(a.pp, mode, c2w34, rot, trans, scale, init34, se3);
  if (st) return st;
  FMOV_REQUIRE(B >= 0 && intr_stride >= 3, "fmov_raygen_xy_fwd: bad sizes");
  if (B == 0) return OK;
  FMOV_REQUIRE(intr_inv && px && py && rays_o && rays_d, "fmov_raygen_xy_fwd: null argument");
  a.intr_inv = intr_inv; a.intr_stride = intr_stride; a.pxf = px; a.pyf = py; a.B = B;
  a.rays_o = rays_o; a.rays_d = rays_d; a.near = near; a.far = far; a.c2w_out = c2w_out;
  raygen_fwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("raygen_fwd_kernel");
  return OK;
}

extern "C" int fmov_raygen_xy_bwd(const float* intr_inv, int intr_stride, const float* px, const float* py, long long B,
                                  const float* rays_o, const float* rays_d, const float* g_o, const float* g_d,
                                  const float* g_near, const float* g_far, float* g_c2w34, void* stream) {
  RayArgs a;
  memset(&a, 0, sizeof(a));
  FMOV_REQUIRE(B >= 0 && intr_stride >= 3, "fmov_raygen_xy_bwd: bad sizes");
  FMOV_REQUIRE(g_c2w34, "fmov_raygen_xy_bwd: null output");
  FMOV_CUDA(cudaMemsetAsync(g_c2w34, 0, 12 * sizeof(float), (cudaStream_t)stream));
  if (B == 0) return OK;
  FMOV_REQUIRE(intr_inv && px && py && rays_o && rays_d, "fmov_raygen_xy_bwd: null argument");
  a.intr_inv = intr_inv; a.intr_stride = intr_stride; a.pxf = px; a.pyf = py; a.B = B;
  a.rays_o = const_cast<float*>(rays_o); a.rays_d = const_cast<float*>(rays_d);
  raygen_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, g_o, g_d, g_near, g_far, g_c2w34);
  FMOV_LAUNCH_CHECK("raygen_bwd_kernel");
  return OK;
}

/* ---- LearnPoseGF in one launch per direction ---------------------------------------------------------- */
static int fill_gf(PoseGfArgs& a, const long long* cid, const float* b, const float* W1, const float* b1, const float* W2,
                   const float* b2, int n_heads, const float* const* Wh, const float* const* bh, const int* rows,
                   float rot_k, const float* init_all, float* save) {
  FMOV_REQUIRE(cid && b && W1 && b1 && W2 && b2 && Wh && bh && rows && save, "pose_gf: null argument");
  FMOV_REQUIRE(n_heads >= 1 && n_heads <= 3, "pose_gf: 1..3 heads (got %d)", n_heads);
  memset(&a, 0, sizeof(a));
  int total = 0;
  for (int h = 0; h < n_heads; ++h) {
    FMOV_REQUIRE(Wh[h] && bh[h] && rows[h] >= 1 && rows[h] <= 6, "pose_gf: bad head %d", h);
    a.Wh[h] = Wh[h]; a.bh[h] = bh[h]; a.rows[h] = rows[h];
    total += rows[h];
  }
  FMOV_REQUIRE(total == 6 || total == 7, "pose_gf: heads must give rot(3)+trans(3)[+scale(1)] rows (got %d)", total);
  a.cid = cid; a.b = b; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.n_heads = n_heads; a.rot_k = rot_k;
  a.init_all = init_all; a.save = save;
  return OK;
}

extern "C" int fmov_pose_gf_save_floats(void) { return GF_SAVE; }

extern "C" int fmov_pose_gf_fwd(const long long* cid, const float* b, const float* W1, const float* b1, const float* W2,
                                const float* b2, int n_heads, const float* const* Wh, const float* const* bh,
                                const int* rows, float rot_k, const float* init_all, float* save, float* c2w34,
                                void* stream) {
  PoseGfArgs a;
  int st = fill_gf(a, cid, b, W1, b1, W2, b2, n_heads, Wh, bh, rows, rot_k, init_all, save);
  if (st) return st;
  FMOV_REQUIRE(c2w34, "fmov_pose_gf_fwd: null output");
  pose_gf_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a, c2w34);
  FMOV_LAUNCH_CHECK("pose_gf_fwd_kernel");
  return OK;
}

extern "C" int fmov_pose_gf_bwd(const long long* cid, const float* b, const float* W1, const float* b1, const float* W2,
                                const float* b2, int n_heads, const float* const* Wh, const float* const* bh,
                                const int* rows, float rot_k, const float* init_all, const float* save,
                                const float* g_c2w34, float* dW1, float* db1, float* dW2, float* db2, float* const* dWh,
                                float* const* dbh, void* stream) {
  PoseGfArgs a;
  int st = fill_gf(a, cid, b, W1, b1, W2, b2, n_heads, Wh, bh, rows, rot_k, init_all, const_cast<float*>(save));
  if (st) return st;
  FMOV_REQUIRE(g_c2w34 && dWh && dbh, "fmov_pose_gf_bwd: null argument");
  PoseGfGrads g;
  memset(&g, 0, sizeof(g));
  g.dW1 = dW1; g.db1 = db1; g.dW2 = dW2; g.db2 = db2;
  for (int h = 0; h < n_heads; ++h) { g.dWh[h] = dWh[h]; g.dbh[h] = dbh[h]; }
  pose_gf_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a, g_c2w34, g);
  FMOV_LAUNCH_CHECK("pose_gf_bwd_kernel");
  return OK;
}
