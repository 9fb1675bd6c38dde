// Weight / bias gradients of both MLPs from the activation- and gradient-tile stash written by
// fmov_fine_fwd / fmov_fine_bwd  (the dW part of loss.backward(), exp_runner.py:802).
//
//   dW_l = Zbar_l^T U_l + Delta_l^T Vbar_l        (oracle/explicit_adjoint.py: sdf_backward)
// Both products are GEMMs whose reduction dimension is the POINT index, so the tile images are used as
// MN-major tcgen05 operands (no transposes): A = gradient tile (M = output features), B = input tile
// (N = input features), K = 64 points per pipeline stage, fp32 accumulation of the whole point range of a
// CTA in TMEM (2 x [128 x N] accumulators = all 512 columns for a 256 x 256 layer), one atomic flush at the end.
// Bias gradients, lin8 row 0 and the 3-row colour output layer are column sums done by a CUDA-core kernel.
#include "mlp_chain.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

// stash ids (must match mlp_fine.cu)
enum { S_PE = 0, S_H1 = 1, S_F = 9, S_D0 = 10, S_X = 18, S_C1 = 19, S_ZC0 = 23, S_FB = 27, S_GE = 28, S_V1 = 29,
       S_Q0 = 37, S_Z0 = 45, S_COUNT = 53 };
__host__ __device__ inline int s_kb(int id) { return (id == S_PE || id == S_X || id == S_GE) ? 1 : 4; }

constexpr int DW_MAX_JOBS = 20;
constexpr int DW_ENTRIES = 3;                     // ring of whole 128-point tiles (64 KiB each): A, B, A, B, ...
constexpr int DW_ENTRY_BYTES = 4 * BLK_BYTES;
constexpr int DW_THREADS = 192;

struct DwJob {
  uint8_t npairs;
  uint8_t a_id[2], b_id[2];        // stash tensors: A = gradient tile (4 blocks), B = input tile
  uint8_t a_bf16[2], b_bf16[2];
  uint8_t b_blocks;                // feature blocks of B used (N = 64*b_blocks)
  uint8_t pad0;
  int m_valid, n_valid;            // rows / cols of the fp32 output actually written
  int ld, col_off;                 // out[m*ld + col_off + n]
  long long out_off;               // offset (floats) into the flat gradient buffer
  float scale;
  int cta_first, cta_count;
  long long bias_off;              // float offset of the bias gradient fed by column sums of pair 0's A tile, or -1
  int bias_n;                      // number of valid columns
};
struct DwPlan {
  int n_jobs;
  DwJob job[DW_MAX_JOBS];
};
struct DwPtrs {
  const uint8_t* stash[S_COUNT];
};

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_kernel(const __grid_constant__ DwPlan plan, const __grid_constant__ DwPtrs ptrs, long long n_tiles, float* grads,
          const float* amax) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = chain_smem_base(smem_raw);
  __shared__ uint64_t full[DW_ENTRIES], ready[DW_ENTRIES], empty[DW_ENTRIES], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // which job / tile range
  int ji = 0;
  for (int j = 0; j < plan.n_jobs; ++j)
    if ((int)blockIdx.x >= plan.job[j].cta_first && (int)blockIdx.x < plan.job[j].cta_first + plan.job[j].cta_count) ji = j;
  const DwJob& jb = plan.job[ji];
  const int local = blockIdx.x - jb.cta_first;
  const long long per = (n_tiles + jb.cta_count - 1) / jb.cta_count;
  const long long t0 = local * per, t1 = (t0 + per < n_tiles) ? t0 + per : n_tiles;
  const int n_mma = jb.b_blocks * 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < DW_ENTRIES; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool has_work = t1 > t0;
  // ring entry sequence: (tile, pair) -> A entry then B entry

  if (warp == 0 && lane == 0 && has_work) {
    uint32_t it = 0;
    for (long long t = t0; t < t1; ++t)
      for (int pr = 0; pr < jb.npairs; ++pr)
        for (int ab = 0; ab < 2; ++ab, ++it) {
          const uint32_t e = it % DW_ENTRIES, n = it / DW_ENTRIES;
          const int id = ab ? jb.b_id[pr] : jb.a_id[pr];
          const uint32_t bytes = (uint32_t)(ab ? jb.b_blocks : 4) * BLK_BYTES;
          const uint8_t* src = ptrs.stash[id] + (size_t)t * s_kb(id) * BLK_BYTES;
          mbar_wait_backoff(&empty[e], (n & 1) ^ 1);
          mbar_expect_tx(&full[e], bytes);
          for (uint32_t off = 0; off < bytes; off += BLK_BYTES)
            bulk_g2s(base + e * DW_ENTRY_BYTES + off, src + off, BLK_BYTES, &full[e]);
        }
  } else if (warp == 1 && lane == 0 && has_work) {
    uint32_t it = 0;
    bool first = true;
    // tcgen05.mma kind::f16 cannot mix an fp16 with a bf16 operand (illegal instruction on sm_100a): in bf16
    // gradient mode the converter warps rewrite the fp16 operand of the pair as bf16 in shared memory.
    const uint32_t idesc = umma_idesc(128, n_mma, kGradFmt, kGradFmt, 1, 1);
    for (long long t = t0; t < t1; ++t)
      for (int pr = 0; pr < jb.npairs; ++pr, it += 2) {
        const uint32_t ea = it % DW_ENTRIES, na = it / DW_ENTRIES;
        const uint32_t eb = (it + 1) % DW_ENTRIES, nb = (it + 1) / DW_ENTRIES;
        const uint32_t sa = smem_u32(base + ea * DW_ENTRY_BYTES);
        const uint32_t sb = smem_u32(base + eb * DW_ENTRY_BYTES);
        mbar_wait_backoff(&ready[ea], na & 1);
        mbar_wait_backoff(&ready[eb], nb & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {        // 16 points per instruction = two 8-row groups = 256 bytes
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {      // 128 output features = 16 chunk columns
            umma_f16(tmem + mh * 256, umma_desc_mnmajor(sa + mh * 16 * TI_CHUNK_STRIDE + ks * 256, TI_CHUNK_STRIDE),
                     umma_desc_mnmajor(sb + ks * 256, TI_CHUNK_STRIDE), idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        first = false;
        umma_commit(&empty[ea]);
        umma_commit(&empty[eb]);
      }
    umma_commit(&done_bar);
  } else if (warp >= 2 && has_work) {
    // helper warps: (bf16 mode) fp16 operand -> bf16 in place; column sums of the gradient tile (pair 0's A
    // operand) = bias gradient, two columns per thread
    float bsum0 = 0.f, bsum1 = 0.f;
    {
      const int ctid = threadIdx.x - 64;
      const int bcol = 2 * ctid;                                   // columns bcol, bcol+1
      const int bcc = bcol >> 3;                                   // chunk column
      const uint32_t boff = (uint32_t)bcc * TI_CHUNK_STRIDE + (uint32_t)((bcol & 7) >> 1) * 4;
      uint32_t it = 0;
      for (long long t = t0; t < t1; ++t)
        for (int pr = 0; pr < jb.npairs; ++pr)
          for (int ab = 0; ab < 2; ++ab, ++it) {
            const uint32_t e = it % DW_ENTRIES, n = it / DW_ENTRIES;
            uint8_t* se = base + e * DW_ENTRY_BYTES;
            mbar_wait(&full[e], n & 1);
            if (kGradBf16 && !(ab ? jb.b_bf16[pr] : jb.a_bf16[pr])) {
              uint4* reg = reinterpret_cast<uint4*>(se);
              const int nchunks = (ab ? jb.b_blocks : 4) * BLK_BYTES / 16;
              for (int i = ctid; i < nchunks; i += 128) {
                uint4 q = reg[i];
                float2 f;
                f = unpack_h2(q.x); q.x = pack_bf2(f.x, f.y);
                f = unpack_h2(q.y); q.y = pack_bf2(f.x, f.y);
                f = unpack_h2(q.z); q.z = pack_bf2(f.x, f.y);
                f = unpack_h2(q.w); q.w = pack_bf2(f.x, f.y);
                reg[i] = q;
              }
              fence_proxy_async();
            }
            if (jb.bias_off >= 0 && pr == 0 && ab == 0) {
#pragma unroll 8
              for (int i = 0; i < TILE_M; ++i) {
                const int r = (i + (bcc & 7)) & (TILE_M - 1);      // staggered start row: conflict-free across chunk columns
                const uint32_t w = *reinterpret_cast<const uint32_t*>(se + boff + r * 16);
                const float2 f = kGradBf16 ? unpack_bf2(w) : unpack_h2(w);
                bsum0 += f.x;
                bsum1 += f.y;
              }
            }
            mbar_arrive(&ready[e]);
          }
      if (jb.bias_off >= 0) {
        const float ginv_b = 1.0f / grad_scale_from_amax(__ldg(amax));
        if (bcol < jb.bias_n) atomicAdd(grads + jb.bias_off + bcol, bsum0 * ginv_b);
        if (bcol + 1 < jb.bias_n) atomicAdd(grads + jb.bias_off + bcol + 1, bsum1 * ginv_b);
      }
    }
    // epilogue: wait for the whole range, then flush the two [128 x N] accumulators with atomics
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    float* out = grads + jb.out_off;
    const float oscale = jb.scale / grad_scale_from_amax(__ldg(amax));
    for (int mh = 0; mh < 2; ++mh) {
      const int m = mh * 128 + row;
      for (int c0 = 0; c0 < n_mma; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + mh * 256 + c0, v);
        tmem_ld_wait();
        if (m < jb.m_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < jb.n_valid) atomicAdd(out + (size_t)m * jb.ld + jb.col_off + c0 + j, v[j] * oscale);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- column sums: lin8 row 0, colour lin4, scalar biases ---------------------------------------------------------
struct ColsumArgs {
  DwPtrs ptrs;
  long long n_tiles, P;
  const float* d_sdf;     // [P]
  const float* zc4;       // [P,4]  (loss-scaled)
  const float* amax;
  float* grads;
  long long off_b_sdf[9];  // float offsets of lin{l}.bias grads
  long long off_b_col[5];
  long long off_w8;        // lin8 weight grad [257,256] (row 0 written here)
  long long off_wc4;       // colour lin4 weight grad [3,256]
};

// lin8 row 0 (sum_p sbar_p*H8[p,:] + Vbar8[p,:]), colour lin4 (sum_p zc4[p,j]*C4[p,:]), and the two scalar-ish
// biases.  grid = (tile groups, 4): a warp owns ONE 16-byte chunk column (8 features) of the three tiles and its
// lanes own consecutive rows, so every load instruction covers 512 contiguous bytes.
__global__ void __launch_bounds__(256) colsum_kernel(ColsumArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cc = blockIdx.y * 8 + warp;           // chunk column 0..31 -> features 8cc .. 8cc+7
  float w8[8], c4[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { w8[i] = 0.f; c4[0][i] = c4[1][i] = c4[2][i] = 0.f; }
  float extra[4] = {0.f, 0.f, 0.f, 0.f};
  const float ginv = 1.0f / grad_scale_from_amax(__ldg(a.amax));
  for (long long t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const size_t toff = (size_t)t * 4 * BLK_BYTES + (size_t)cc * TI_CHUNK_STRIDE;
    const uint8_t* th = a.ptrs.stash[S_H1 + 7] + toff;
    const uint8_t* tv = a.ptrs.stash[S_V1 + 7] + toff;
    const uint8_t* tc = a.ptrs.stash[S_C1 + 3] + toff;
#pragma unroll
    for (int rgp = 0; rgp < 4; ++rgp) {
      const int r = rgp * 32 + lane;
      const long long p = t * 128 + r;
      const bool ok = p < a.P;
      const float sb = ok ? a.d_sdf[p] : 0.f;
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) z4 = *reinterpret_cast<const float4*>(a.zc4 + p * 4);
      const float z0 = z4.x * ginv, z1 = z4.y * ginv, z2 = z4.z * ginv;
      const uint4 qh = *reinterpret_cast<const uint4*>(th + r * 16);
      const uint4 qv = *reinterpret_cast<const uint4*>(tv + r * 16);
      const uint4 qc = *reinterpret_cast<const uint4*>(tc + r * 16);
      const uint32_t hh[4] = {qh.x, qh.y, qh.z, qh.w}, vv[4] = {qv.x, qv.y, qv.z, qv.w}, cw[4] = {qc.x, qc.y, qc.z, qc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fh = unpack_h2(hh[j]), fv = kGradBf16 ? unpack_bf2(vv[j]) : unpack_h2(vv[j]), fc = unpack_h2(cw[j]);
        w8[2 * j] += sb * fh.x + fv.x * ginv; w8[2 * j + 1] += sb * fh.y + fv.y * ginv;
        c4[0][2 * j] += z0 * fc.x; c4[0][2 * j + 1] += z0 * fc.y;
        c4[1][2 * j] += z1 * fc.x; c4[1][2 * j + 1] += z1 * fc.y;
        c4[2][2 * j] += z2 * fc.x; c4[2][2 * j + 1] += z2 * fc.y;
      }
      if (cc == 0) { extra[0] += sb; extra[1] += z0; extra[2] += z1; extra[3] += z2; }
    }
  }
  float* g = a.grads;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s0 = w8[i], s1 = c4[0][i], s2 = c4[1][i], s3 = c4[2][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if (lane == 0) {
      const int col = cc * 8 + i;
      atomicAdd(g + a.off_w8 + col, s0);
      atomicAdd(g + a.off_wc4 + col, s1);
      atomicAdd(g + a.off_wc4 + 256 + col, s2);
      atomicAdd(g + a.off_wc4 + 512 + col, s3);
    }
  }
  if (cc == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e = extra[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if (lane == 0) atomicAdd(j == 0 ? g + a.off_b_sdf[8] : g + a.off_b_col[4] + (j - 1), e);
    }
  }
}

// max |x| over the three upstream per-sample gradient arrays -> device scalar (float bits, non-negative)
__global__ void grad_amax_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb,
                                 const float* __restrict__ c, long long nc, unsigned int* out) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) m = fmaxf(m, fabsf(a[i]));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride) m = fmaxf(m, fabsf(b[i]));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) m = fmaxf(m, fabsf(c[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(out, __float_as_uint(m));
}

}  // namespace fmov
using namespace fmov;

/* amax (device float) = max |d_sdf|, |d_nrm|, |d_rgb| : fixes the loss scale of the fp16 gradient tiles */
extern "C" int fmov_grad_amax(const float* d_sdf, const float* d_nrm, const float* d_rgb, long long P, float* amax,
                              void* stream) {
  FMOV_REQUIRE(P > 0 && d_sdf && d_nrm && d_rgb && amax, "fmov_grad_amax: bad arguments");
  FMOV_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), (cudaStream_t)stream));
  grad_amax_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(d_sdf, P, d_nrm, 3 * P, d_rgb, 3 * P,
                                                         reinterpret_cast<unsigned int*>(amax));
  FMOV_LAUNCH_CHECK("grad_amax_kernel");
  return OK;
}

// ---- flat gradient buffer layout (fp32), shapes = the reference parameters' effective weights -----------------
static const int SDF_OUT[9] = {256, 256, 256, 217, 256, 256, 256, 256, 257};
static const int SDF_IN[9] = {39, 256, 256, 256, 256, 256, 256, 256, 256};
static const int COL_OUT[5] = {256, 256, 256, 256, 3};
static const int COL_IN[5] = {289, 256, 256, 256, 256};

// kind: 0 sdf weight, 1 sdf bias, 2 colour weight, 3 colour bias
extern "C" long long fmov_grad_offset(int kind, int layer) {
  long long off = 0;
  for (int k = 0; k < 4; ++k) {
    const int nl = (k < 2) ? 9 : 5;
    for (int l = 0; l < nl; ++l) {
      if (k == kind && l == layer) return off;
      const int o = (k < 2) ? SDF_OUT[l] : COL_OUT[l];
      const int i = (k < 2) ? SDF_IN[l] : COL_IN[l];
      off += (k == 0 || k == 2) ? (long long)o * i : o;
    }
  }
  return (kind == 4) ? off : -1;
}
extern "C" long long fmov_grad_floats(void) { return fmov_grad_offset(4, 0); }

static void add_job(DwPlan& pl, int npairs, int a0, int b0, int a0bf, int b0bf, int a1, int b1, int a1bf, int b1bf,
                    int b_blocks, int m_valid, int n_valid, int ld, int col_off, long long out_off, float scale,
                    long long bias_off = -1, int bias_n = 0) {
  DwJob& j = pl.job[pl.n_jobs++];
  memset(&j, 0, sizeof(j));
  j.npairs = (uint8_t)npairs;
  j.a_id[0] = (uint8_t)a0; j.b_id[0] = (uint8_t)b0; j.a_bf16[0] = (uint8_t)a0bf; j.b_bf16[0] = (uint8_t)b0bf;
  j.a_id[1] = (uint8_t)a1; j.b_id[1] = (uint8_t)b1; j.a_bf16[1] = (uint8_t)a1bf; j.b_bf16[1] = (uint8_t)b1bf;
  j.b_blocks = (uint8_t)b_blocks; j.m_valid = m_valid; j.n_valid = n_valid; j.ld = ld; j.col_off = col_off;
  j.out_off = out_off; j.scale = scale; j.bias_off = bias_off; j.bias_n = bias_n;
}

static void build_plan(DwPlan& pl, int n_ctas) {
  memset(&pl, 0, sizeof(pl));
  const float rs2 = 0.70710678118654752f;
  const int G = kGradBf16 ? 1 : 0;     // format flag of gradient tiles
  // SDF layers 1..7: (Zbar_l, H_l) + (Delta_l, Vbar_l)
  for (int l = 1; l <= 7; ++l) {
    const int nv = (l == 4) ? 217 : 256;
    add_job(pl, 2, S_Z0 + l, S_H1 + (l - 1), G, 0, S_D0 + l, S_V1 + (l - 1), 0, G, 4, SDF_OUT[l], nv, 256, 0,
            fmov_grad_offset(0, l), l == 4 ? rs2 : 1.f, fmov_grad_offset(1, l), SDF_OUT[l]);
  }
  // layer 4, PE columns 217..255: (Zbar_4, PE) + (Delta_4, GE)
  add_job(pl, 2, S_Z0 + 4, S_PE, G, 0, S_D0 + 4, S_GE, 0, G, 1, 256, 39, 256, 217, fmov_grad_offset(0, 4), rs2);
  // layer 0: (Zbar_0, PE) + (Delta_0, GE)
  add_job(pl, 2, S_Z0 + 0, S_PE, G, 0, S_D0 + 0, S_GE, 0, G, 1, 256, 39, 39, 0, fmov_grad_offset(0, 0), 1.f,
          fmov_grad_offset(1, 0), 256);
  // layer 8 feature rows 1..256: (fbar, H8)
  add_job(pl, 1, S_FB, S_H1 + 7, G, 0, 0, 0, 0, 0, 4, 256, 256, 256, 0, fmov_grad_offset(0, 8) + 256, 1.f,
          fmov_grad_offset(1, 8) + 1, 256);
  // colour layers 1..3: (Zbar_cl, C_l); layer 0: (Zbar_c0, F) -> cols 33.., (Zbar_c0, X) -> cols 0..32
  for (int l = 1; l <= 3; ++l)
    add_job(pl, 1, S_ZC0 + l, S_C1 + (l - 1), G, 0, 0, 0, 0, 0, 4, 256, 256, 256, 0, fmov_grad_offset(2, l), 1.f,
            fmov_grad_offset(3, l), 256);
  add_job(pl, 1, S_ZC0 + 0, S_F, G, 0, 0, 0, 0, 0, 4, 256, 256, 289, 33, fmov_grad_offset(2, 0), 1.f,
          fmov_grad_offset(3, 0), 256);
  add_job(pl, 1, S_ZC0 + 0, S_X, G, 0, 0, 0, 0, 0, 1, 256, 33, 289, 0, fmov_grad_offset(2, 0), 1.f);
  // distribute CTAs proportionally to cost (bytes streamed per tile)
  float cost[DW_MAX_JOBS], total = 0.f;
  for (int j = 0; j < pl.n_jobs; ++j) {
    cost[j] = pl.job[j].npairs * (4.f + pl.job[j].b_blocks);   // 16 KiB blocks streamed per tile
    total += cost[j];
  }
  int used = 0;
  for (int j = 0; j < pl.n_jobs; ++j) {
    int c = (int)(cost[j] / total * n_ctas);
    if (c < 1) c = 1;
    pl.job[j].cta_count = c;
    used += c;
  }
  // hand leftovers to / take excess from the biggest jobs
  for (int j = 0; used != n_ctas; j = (j + 1) % pl.n_jobs) {
    if (used < n_ctas && pl.job[j].npairs == 2 && pl.job[j].b_blocks == 4) { ++pl.job[j].cta_count; ++used; }
    else if (used > n_ctas && pl.job[j].cta_count > 1) { --pl.job[j].cta_count; --used; }
  }
  int first = 0;
  for (int j = 0; j < pl.n_jobs; ++j) { pl.job[j].cta_first = first; first += pl.job[j].cta_count; }
}

/* grads: flat fp32 buffer of fmov_grad_floats() floats, zeroed by this call. stash: HOST array of device pointers. */
extern "C" int fmov_dw(long long P, void* const* stash, const float* d_sdf, const float* zc4, const float* amax, float* grads,
                       void* stream) {
  FMOV_REQUIRE(P > 0 && stash && d_sdf && zc4 && amax && grads, "fmov_dw: bad arguments");
  static DwPlan plan;
  static bool init = false;
  static int n_ctas = 148;
  if (!init) {
    int dev = 0, sms = 148;
    FMOV_CUDA(cudaGetDevice(&dev));
    FMOV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    n_ctas = sms;
    build_plan(plan, n_ctas);
    FMOV_CUDA(cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_ENTRIES * DW_ENTRY_BYTES + 1024));
    init = true;
  }
  DwPtrs ptrs;
  for (int i = 0; i < S_COUNT; ++i) {
    const bool is_q = i >= S_Q0 && i < S_Q0 + 8;          // q tiles feed no weight gradient (and are not stored:
    if (!is_q) FMOV_REQUIRE(stash[i], "fmov_dw: stash tensor %d is null", i);      // the backward rebuilds q)
    ptrs.stash[i] = reinterpret_cast<const uint8_t*>(stash[i]);
  }
  const long long n_tiles = (P + 127) / 128;
  FMOV_CUDA(cudaMemsetAsync(grads, 0, (size_t)fmov_grad_floats() * sizeof(float), (cudaStream_t)stream));
  dw_kernel<<<n_ctas, DW_THREADS, DW_ENTRIES * DW_ENTRY_BYTES + 1024, (cudaStream_t)stream>>>(plan, ptrs, n_tiles, grads, amax);
  FMOV_LAUNCH_CHECK("dw_kernel");
  ColsumArgs ca;
  ca.ptrs = ptrs; ca.n_tiles = n_tiles; ca.P = P; ca.d_sdf = d_sdf; ca.zc4 = zc4; ca.amax = amax; ca.grads = grads;
  for (int l = 0; l < 9; ++l) ca.off_b_sdf[l] = fmov_grad_offset(1, l);
  for (int l = 0; l < 5; ++l) ca.off_b_col[l] = fmov_grad_offset(3, l);
  ca.off_w8 = fmov_grad_offset(0, 8);
  ca.off_wc4 = fmov_grad_offset(2, 4);
  const int cs_grid = (int)(n_tiles < 2 * n_ctas ? n_tiles : 2 * n_ctas);
  colsum_kernel<<<dim3(cs_grid, 4), 256, 0, (cudaStream_t)stream>>>(ca);
  FMOV_LAUNCH_CHECK("colsum_kernel");
  return OK;
}
