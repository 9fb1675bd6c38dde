// fmov_sdf_query: fused PE(6) -> 8 x Softplus(beta=100) layers -> sdf  (inference, value only).
//
// Replaces SDFNetwork.sdf on the no-grad paths of the reference: coarse + up-sample queries
// (models/renderer.py:424-428, :230-232) and the dense grid query of extract_fields
// (models/renderer.py:9-37, query_func = -sdf at :506).   Network semantics: models/fields.py:88-107.
//
// One persistent CTA per SM walks 128-point tiles through the chain engine (mlp_chain.cuh): the
// activation tile never leaves shared memory, weights stream from L2 via TMA bulk copies, accumulators
// live in TMEM.  The last linear layer only needs row 0 (the SDF), which is evaluated in fp32 by the
// epilogue of layer 7 from the un-rounded activations.
// the value-only chain has no HBM traffic in its epilogue and fits 96 registers: two column warpgroups per tile slot
// (16 epilogue warps) measured 1.45 -> 1.33 ms per 1M points; the fine-stage kernels keep one (register bound)
#define FMOV_CH_WGS 2
#include "mlp_chain.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

using QL = ChainLayout;

struct QueryArgs {
  // input modes: 0 = points [P,3]; 1 = rays (o,d [B,3]) x z [B, z_stride], S samples per ray;
  //              2 = regular grid res^3 over [bmin,bmax] (torch.linspace semantics), x-major
  int mode;
  long long P;
  const float* pts;
  const float* rays_o;
  const float* rays_d;
  const float* z;
  int S;
  int z_stride;
  int z_off;
  int res;
  long long grid_off;   // first flat grid index handled by this call (grid partitioning across ranks)
  float bmin[3], bmax[3];
  float in_scale;       // SDFNetwork.scale
  float out_scale;      // sign / scale applied to the output (-1/scale for extract_fields)
  const float* bias;    // [8][256] biases of layers 0..7 (padded with zeros)
  const float* w8;      // [256] row 0 of lin8 (effective weight), fp32
  const float* b8;      // lin8.bias (device), element 0 used
  float* out;           // [P]
};

__device__ __forceinline__ float linspace_at(float a, float b, int n, int i) {
  // torch.linspace: step = (b-a)/(n-1); first half counts up from a, second half down from b
  if (n <= 1) return a;
  float step = (b - a) / (float)(n - 1);
  return (i < n / 2) ? a + step * (float)i : b - step * (float)(n - 1 - i);
}

__device__ __forceinline__ void load_point(const QueryArgs& a, long long p, float x[3]) {
  if (a.mode == 0) {
    x[0] = a.pts[p * 3 + 0]; x[1] = a.pts[p * 3 + 1]; x[2] = a.pts[p * 3 + 2];
  } else if (a.mode == 1) {
    long long r = p / a.S;
    int j = (int)(p - r * a.S);
    float zz = a.z[r * a.z_stride + a.z_off + j];
    x[0] = a.rays_o[r * 3 + 0] + a.rays_d[r * 3 + 0] * zz;
    x[1] = a.rays_o[r * 3 + 1] + a.rays_d[r * 3 + 1] * zz;
    x[2] = a.rays_o[r * 3 + 2] + a.rays_d[r * 3 + 2] * zz;
  } else {
    long long g = p + a.grid_off;
    long long rr = (long long)a.res * a.res;
    int ix = (int)(g / rr);
    int iy = (int)((g - ix * rr) / a.res);
    int iz = (int)(g - ix * rr - (long long)iy * a.res);
    x[0] = linspace_at(a.bmin[0], a.bmax[0], a.res, ix);
    x[1] = linspace_at(a.bmin[1], a.bmax[1], a.res, iy);
    x[2] = linspace_at(a.bmin[2], a.bmax[2], a.res, iz);
  }
}

// Positional encoding of one point into its row of a [128 x 64] fp16 block: 39 channels
// [x, sin(2^k x), cos(2^k x)]_{k<6} (models/embedder.py:28-37), columns 39..63 zero.
__device__ __forceinline__ void pe6_to_block(uint8_t* blk, int row, const float x[3]) {
  float e[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(x[c] * f, &s, &co);
      e[3 + 6 * k + c] = s;
      e[6 + 6 * k + c] = co;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) e[PE_RES_COL + c] = x[c] - __half2float(__float2half_rn(x[c]));      // see pe_with_residual
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 q[4];
    pack4(e + 32 * h, false, q);
    row_half_store(blk + row * 16, h, q);
  }
}
__device__ __forceinline__ void pe6_half_to_block(uint8_t* blk, int row, const float x[3], int h, bool bias_cols = false) {
  float e[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(x[c] * f, &s, &co);
      e[3 + 6 * k + c] = s;
      e[6 + 6 * k + c] = co;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) e[PE_RES_COL + c] = x[c] - __half2float(__float2half_rn(x[c]));      // see pe_with_residual
  if (bias_cols) { e[AUX_ONE_COL] = 1.0f; e[AUX_ONE_COL + 1] = 1.0f / BIAS_LO_SCALE; }              // x the bias slices (bias16 steps)
  uint4 q[4];
  if (h == 0) pack4(e, false, q); else pack4(e + 32, false, q);
  row_half_store(blk + row * 16, h, q);
}

// split-precision variants: value = hi + lo with hi = fp16(v), lo = fp16(v - hi) (relative error 2^-22)
__device__ __forceinline__ void split16(const float* v, float* lo) {
#pragma unroll
  for (int i = 0; i < 16; ++i) lo[i] = v[i] - __half2float(__float2half_rn(v[i]));
}
__device__ __forceinline__ void pe6_half_to_block_hilo(uint8_t* blk_hi, uint8_t* blk_lo, int row, const float x[3], int h) {
  float e[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(x[c] * f, &s, &co);      // abs error ~1e-6: far below the 1e-4 the split-precision chain is held to
      e[3 + 6 * k + c] = s;
      e[6 + 6 * k + c] = co;
    }
  }
  const float* eh = e + 32 * h;
  uint4 q[4];
  pack4(eh, false, q);
  row_half_store(blk_hi + row * 16, h, q);
  float lo[32];
  split16(eh, lo);
  split16(eh + 16, lo + 16);
  pack4(lo, false, q);
  row_half_store(blk_lo + row * 16, h, q);
}
// MODE 0: one-CTA engine, row-interleaved weight images, biases added in the epilogue (any caller's stand-alone blob)
// MODE 1: split-precision chain (one-CTA engine)
// MODE 2: CTA-pair engine (clusters of two, tcgen05.mma.cta_group::2, half-major images with bias slices: the FP0..FP7 images
//         of the fine-stage blob) — the hierarchical-sampling queries of the train step
constexpr int QM_SINGLE = 0, QM_PRECISE = 1, QM_PAIR = 2;
template <int MODE>
__global__ void __launch_bounds__(CH_THREADS, 1)
sdf_query_kernel(const __grid_constant__ ChainTable tb, const __grid_constant__ ChainPtrs ptrs,
                 const __grid_constant__ QueryArgs a, const __grid_constant__ PairMaps maps) {
  constexpr bool PRECISE = MODE == QM_PRECISE, PAIR = MODE == QM_PAIR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = chain_smem_base(smem_raw);
  ChainSmem* s = reinterpret_cast<ChainSmem*>(base);
  uint8_t* act0 = base + QL::ACT;
  uint8_t* aux0 = base + QL::AUX;
  uint8_t* wst = base + QL::WST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const long long n_tiles = (a.P + TILE_M - 1) / TILE_M;
  // pair mode: both CTAs of a pair walk as many tiles as the even one has (the odd CTA pads with an all-invalid tile)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const long long first_tile = PAIR ? (long long)(blockIdx.x & ~1u) : (long long)blockIdx.x;
  const int n_my = (int)((n_tiles - first_tile + gridDim.x - 1) / gridDim.x);

  if (threadIdx.x == 0) chain_init_barriers(s, PRECISE ? 2 * EPI_THREADS : EPI_THREADS, PAIR);
  if (warp == ISSUER_WARP) {
    if (PAIR) tmem_alloc_pair(&s->tmem_base, 512);
    else tmem_alloc(&s->tmem_base, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s->tmem_base;

  if (warp >= CTRL_WARP0) {
    if (warp == PRODUCER_WARP) {
      if (lane == 0) {
        if (PAIR) chain_weight_producer_pair<true>(tb, maps, s, wst, n_my, rank);
        else chain_weight_producer(tb, ptrs, s, wst, n_my, blockIdx.x, gridDim.x);
      }
    } else if (warp == ISSUER_WARP) {
      if (lane == 0) {
        if (!PAIR) chain_mma_issuer<PRECISE>(tb, s, act0, aux0, wst, tmem, n_my);
        else if (rank == 0) chain_mma_issuer_pair<true>(tb, s, act0, aux0, wst, tmem, n_my);
      }
    }
  } else if (PRECISE) {
    // Split-precision chain: ONE tile in flight.  Slot 1's ACT / AUX hold the fp16 residuals of slot 0's operands, and all
    // 16 epilogue warps work on the one accumulator: column group cg = warp / 4 owns columns [64 cg, 64 cg + 64).
    static_assert(CH_WGS == 2, "the split-precision epilogue is written for 16 epilogue warps");
    EpiCtx c;
    epi_init(c, s, act0, aux0, tmem);
    const int cg = (warp - EPI_WARP0) >> 2;                     // 0..3
    c.slot = 0;
    c.act = act0;
    c.aux = aux0;
    c.tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* const act_lo = act0 + 4 * BLK_BYTES;
    uint8_t* const aux_lo = aux0 + BLK_BYTES;
    float* const part = reinterpret_cast<float*>(aux0);          // [4][128] partial dot products (AUX is free after layer 4)
    const float b8 = __ldg(a.b8);
    for (int k = 0; k < n_my; ++k) {
      const long long tile = (long long)blockIdx.x + (long long)k * gridDim.x;
      const long long p = tile * TILE_M + c.row;
      const bool valid = p < a.P;
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) load_point(a, p, x);
      x[0] *= a.in_scale; x[1] *= a.in_scale; x[2] *= a.in_scale;
      if (cg < 2) pe6_half_to_block_hilo(c.aux, aux_lo, c.row, x, cg);
      epi_signal_act(c);
      float sdf = 0.f;
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const float* bias = a.bias + l * 256;
        const int n_mma = (l == 3) ? 224 : 256;
        uint8_t* actp = c.act + c.row * 16;
        epi_wait_acc(c);
#pragma unroll 2
        for (int i = 0; i < 4; ++i) {
          const int ck = cg * 4 + i;
          if (ck * 16 >= n_mma) break;
          float v[16];
          acc_load16(c, ck * 16, v);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + ck * 16) + j4);
            v[j4 * 4 + 0] = softplus100(v[j4 * 4 + 0] + b4.x);
            v[j4 * 4 + 1] = softplus100(v[j4 * 4 + 1] + b4.y);
            v[j4 * 4 + 2] = softplus100(v[j4 * 4 + 2] + b4.z);
            v[j4 * 4 + 3] = softplus100(v[j4 * 4 + 3] + b4.w);
          }
          if (l == 7) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w8 + ck * 16) + j4);
              sdf = fmaf(v[j4 * 4 + 0], w4.x, sdf); sdf = fmaf(v[j4 * 4 + 1], w4.y, sdf);
              sdf = fmaf(v[j4 * 4 + 2], w4.z, sdf); sdf = fmaf(v[j4 * 4 + 3], w4.w, sdf);
            }
          } else {
            uint4 q2[2];
            pack2(v, false, q2);
            chunk_store(actp, ck, q2);
            float lo[16];
            split16(v, lo);
            pack2(lo, false, q2);
            chunk_store(act_lo + c.row * 16, ck, q2);
          }
        }
        if (l < 7) epi_signal_act(c);
        else tc_fence_before();
      }
      // combine the four column groups of the lin8-row-0 dot product
      part[cg * 128 + c.row] = sdf;
      named_bar_sync(3, 2 * EPI_THREADS);
      if (cg == 0 && valid)
        a.out[p] = (part[c.row] + part[128 + c.row] + part[256 + c.row] + part[384 + c.row] + b8) * a.out_scale;
      named_bar_sync(3, 2 * EPI_THREADS);          // `part` aliases AUX, which the next tile's encoding overwrites
    }
  } else {
    EpiCtx c;
    epi_init(c, s, act0, aux0, tmem);
    const float b8 = __ldg(a.b8);
    for (int k = c.slot; k < n_my; k += CH_SLOTS) {
      const long long tile = (long long)blockIdx.x + (long long)k * gridDim.x;
      const long long p = tile * TILE_M + c.row;
      const bool valid = p < a.P;
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) load_point(a, p, x);
      x[0] *= a.in_scale; x[1] *= a.in_scale; x[2] *= a.in_scale;
      if (CH_WGS == 2) pe6_half_to_block(c.aux, c.row, x, c.wg, PAIR);   // warpgroup wg writes PE columns [32wg, 32wg+32)
      else pe6_to_block(c.aux, c.row, x);
      epi_signal_act<PAIR>(c);
      float sdf = 0.f;
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const float* bias = a.bias + l * 256;
        const int n_mma = (l == 3) ? 224 : 256;       // lin3 has 217 outputs; columns 217.. meet zero weights next
        uint8_t* actp = c.act + c.row * 16;
        // the chunk's 16 biases are fetched one chunk ahead (the first before the accumulator wait), see fine_fwd_kernel
        float4 bb[4];
        if (!PAIR) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) bb[j4] = __ldg(reinterpret_cast<const float4*>(bias + c.wg * CH_CHUNKS * 16) + j4);
        }
        epi_wait_acc(c);
#pragma unroll 2
        for (int i = 0; i < CH_CHUNKS; ++i) {
          const int ck = c.wg * CH_CHUNKS + i;
          if (ck * 16 >= n_mma) break;
          float v[16];
          acc_load16(c, ck * 16, v);
          if (!PAIR) {          // (pair mode: bias16 steps, the accumulator already holds W u + b)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              v[j4 * 4 + 0] += bb[j4].x; v[j4 * 4 + 1] += bb[j4].y; v[j4 * 4 + 2] += bb[j4].z; v[j4 * 4 + 3] += bb[j4].w;
            }
            if (i + 1 < CH_CHUNKS) {          // (the bias rows are padded to 256 floats: the read past n_mma is in bounds)
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) bb[j4] = ldg_f4_volatile(reinterpret_cast<const float4*>(bias + (ck + 1) * 16) + j4);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = softplus100(v[j]);
          if (l == 7) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w8 + ck * 16) + j4);
              sdf = fmaf(v[j4 * 4 + 0], w4.x, sdf); sdf = fmaf(v[j4 * 4 + 1], w4.y, sdf);
              sdf = fmaf(v[j4 * 4 + 2], w4.z, sdf); sdf = fmaf(v[j4 * 4 + 3], w4.w, sdf);
            }
          } else {
            uint4 q2[2];
            pack2(v, false, q2);
            chunk_store(actp, ck, q2);
          }
        }
        if (l < 7) epi_signal_act<PAIR>(c);
        else tc_fence_before();
      }
      // combine the two column halves of the lin8-row-0 dot product
      if (CH_WGS == 2) {
        if (c.wg == 0) s->scratch[c.slot][c.row] = sdf;
        slot_sync(c);
        if (c.wg == 1) sdf += s->scratch[c.slot][c.row];
        slot_sync(c);      // scratch may be rewritten by the next tile
      }
      if (c.wg == CH_WGS - 1 && valid) a.out[p] = (sdf + b8) * a.out_scale;
    }
  }
  if (PAIR) cluster_sync_all();      // neither CTA may leave while its partner still signals its barriers / reads its smem
  else __syncthreads();
  if (warp == ISSUER_WARP) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

}  // namespace fmov

using namespace fmov;

// Step table of the query chain. Weight blob layout (built by fmov_pack_sdf_weights):
// layers 0..7 forward images, k-blocks consecutive.
static void build_query_table(ChainTable& tb) {
  memset(&tb, 0, sizeof(tb));
  tb.n_steps = 8;
  uint32_t off = 0;
  for (int l = 0; l < 8; ++l) {
    ChainStep& st = tb.step[l];
    st.w_off = off;
    st.n = (l == 3) ? 224 : 256;
    st.nkb_a = (l == 0) ? 0 : 4;
    st.nkb_aux = (l == 0 || l == 4) ? 1 : 0;
    st.a_fmt = FMT_F16;
    st.b_fmt = FMT_F16;
    st.pf[0] = st.pf[1] = 0xFF;
    off += (uint32_t)st.n * 128u * (st.nkb_a + st.nkb_aux);
  }
}

// CTA-pair chain: the half-major images FP0..FP7 of the fine-stage blob, each followed by its bias slice (mlp_fine.cu,
// img_bytes); offsets relative to FP0.
static void build_query_table_pair(ChainTable& tb) {
  build_query_table(tb);
  uint32_t off = 0;
  for (int l = 0; l < 8; ++l) {
    ChainStep& st = tb.step[l];
    st.w_off = off;
    st.bias16 = 1;
    off += (uint32_t)st.n * 128u * (st.nkb_a + st.nkb_aux) + (uint32_t)st.n * 32u;
  }
}
extern "C" long long fmov_sdf_pair_blob_bytes(void) {
  if (!FMOV_FINE_PAIR) return 0;          // built without the pair engine: FP0..FP7 are plain copies of F0..F7
  ChainTable tb;
  build_query_table_pair(tb);
  const ChainStep& st = tb.step[7];
  return (long long)st.w_off + (long long)st.n * 128 * (st.nkb_a + st.nkb_aux) + (long long)st.n * 32;
}

// Split-precision chain: per layer three products into one accumulator, hi*W_hi + lo*W_hi + hi*W_lo (the residual images
// have the layout of the forward blob), one tile in flight.  The first two share one pass over W_hi (CHF_DUAL_A), so a
// layer streams two weight images instead of three.  Result error ~1e-5 instead of the fp16 chain's ~1e-3.
static void build_query_table_precise(ChainTable& tb) {
  ChainTable f;
  build_query_table(f);
  memset(&tb, 0, sizeof(tb));
  tb.n_steps = 16;
  tb.slots = 1;
  for (int l = 0; l < 8; ++l) {
    ChainStep st = f.step[l];
    st.flags = CHF_DUAL_A | CHF_NO_COMMIT;          // (hi + lo) * W_hi
    tb.step[l * 2] = st;
    st.flags = CHF_ACCUM | CHF_W2;                  // + hi * W_lo, then signal the epilogue
    tb.step[l * 2 + 1] = st;
  }
}

// Activation-split chain: only the activations and encoded inputs are carried as fp16 hi + lo, the weights stay plain fp16:
// per layer (hi + lo) * W in ONE pass over W (both A operands against each weight slice).  Half the weight stream and two
// thirds of the MMAs of the full split; SDF error ~3e-4 on the +-1.01 box (full split: 1e-4, plain chain: 1.3e-3).
static void build_query_table_split_act(ChainTable& tb) {
  ChainTable f;
  build_query_table(f);
  memset(&tb, 0, sizeof(tb));
  tb.n_steps = 8;
  tb.slots = 1;
  for (int l = 0; l < 8; ++l) {
    tb.step[l] = f.step[l];
    tb.step[l].flags = CHF_DUAL_A;
  }
}

extern "C" long long fmov_sdf_fwd_blob_bytes(void) {
  ChainTable tb;
  build_query_table(tb);
  const ChainStep& st = tb.step[7];
  return (long long)st.w_off + (long long)st.n * 128 * (st.nkb_a + st.nkb_aux);
}

extern "C" long long fmov_sdf_fwd_blob_offset(int layer) {
  ChainTable tb;
  build_query_table(tb);
  if (layer < 0 || layer > 7) return -1;
  return tb.step[layer].w_off;
}

static int launch_query(const QueryArgs& a, const void* wblob, int max_ctas, cudaStream_t stream, const void* wblob_lo = nullptr,
                        bool pair = false, bool split = false) {
  static ChainTable tb, tbp, tba, tb2;
  static bool init = false;
  if (!init) {
    build_query_table(tb); build_query_table_precise(tbp); build_query_table_split_act(tba); build_query_table_pair(tb2);
    init = true;
  }
  ChainPtrs ptrs;
  memset(&ptrs, 0, sizeof(ptrs));
  ptrs.weights = reinterpret_cast<const uint8_t*>(wblob);
  ptrs.weights2 = reinterpret_cast<const uint8_t*>(wblob_lo);
  int dev = 0, sms = 0;
  FMOV_CUDA(cudaGetDevice(&dev));
  FMOV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_set = false;
  if (!attr_set) {
    FMOV_CUDA(cudaFuncSetAttribute(sdf_query_kernel<QM_SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, QL::DYN_BYTES));
    FMOV_CUDA(cudaFuncSetAttribute(sdf_query_kernel<QM_PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, QL::DYN_BYTES));
    FMOV_CUDA(cudaFuncSetAttribute(sdf_query_kernel<QM_PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, QL::DYN_BYTES));
    attr_set = true;
  }
  long long n_tiles = (a.P + TILE_M - 1) / TILE_M;
  if (n_tiles == 0) return OK;
  int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  PairMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (pair) {
    int st = chain_pair_maps(wblob, fmov_sdf_pair_blob_bytes(), maps);
    if (st) return st;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(CH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = QL::DYN_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs = 0;
    if (max_pairs == 0) {
      cfg.gridDim = dim3(2 * 74, 1, 1);
      int n = 0;
      FMOV_CUDA(cudaOccupancyMaxActiveClusters(&n, sdf_query_kernel<QM_PAIR>, &cfg));
      FMOV_REQUIRE(n > 0, "sdf_query_kernel: no CTA pair fits on this device");
      max_pairs = n;
    }
    grid = (grid + 1) & ~1;
    if (grid > 2 * max_pairs) grid = 2 * max_pairs;
    cfg.gridDim = dim3(grid, 1, 1);
    FMOV_CUDA(cudaLaunchKernelEx(&cfg, sdf_query_kernel<QM_PAIR>, tb2, ptrs, a, maps));
  } else if (split) {          // split-precision kernel: full split with the residual weight images, else activations only
    sdf_query_kernel<QM_PRECISE><<<grid, CH_THREADS, QL::DYN_BYTES, stream>>>(wblob_lo ? tbp : tba, ptrs, a, maps);
  } else {
    sdf_query_kernel<QM_SINGLE><<<grid, CH_THREADS, QL::DYN_BYTES, stream>>>(tb, ptrs, a, maps);
  }
  FMOV_LAUNCH_CHECK("sdf_query_kernel");
  return OK;
}

extern "C" int fmov_sdf_query_points(const float* pts, long long P, const void* wblob, const float* bias8x256,
                                     const float* w8_row0, const float* b8, float in_scale, float out_scale, float* out,
                                     void* stream) {
  FMOV_REQUIRE(P >= 0 && (P == 0 || (pts && out && wblob && bias8x256 && w8_row0)), "fmov_sdf_query_points: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 0; a.P = P; a.pts = pts; a.in_scale = in_scale; a.out_scale = out_scale;
  a.bias = bias8x256; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob, 0, (cudaStream_t)stream);
}

extern "C" int fmov_sdf_query_rays(const float* rays_o, const float* rays_d, const float* z, long long B, int S,
                                   int z_stride, int z_off, const void* wblob, const float* bias8x256,
                                   const float* w8_row0, const float* b8, float in_scale, float out_scale, float* out,
                                   void* stream) {
  FMOV_REQUIRE(B >= 0 && S > 0 && z_stride >= S + z_off, "fmov_sdf_query_rays: bad shape B=%lld S=%d stride=%d off=%d", B, S,
               z_stride, z_off);
  FMOV_REQUIRE(B == 0 || (rays_o && rays_d && z && out && wblob && bias8x256 && w8_row0), "fmov_sdf_query_rays: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 1; a.P = B * S; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.S = S; a.z_stride = z_stride; a.z_off = z_off;
  a.in_scale = in_scale; a.out_scale = out_scale; a.bias = bias8x256; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob, 0, (cudaStream_t)stream);
}

/* same query on the CTA-pair engine: `wblob_pair` = the half-major images FP0..FP7 with their bias slices
 * (fmov_sdf_pair_blob_bytes() bytes: the FP0.. region of the blob fmov_pack_all writes); biases come from the slices */
extern "C" int fmov_sdf_query_rays_pair(const float* rays_o, const float* rays_d, const float* z, long long B, int S,
                                        int z_stride, int z_off, const void* wblob_pair, const float* w8_row0, const float* b8,
                                        float in_scale, float out_scale, float* out, void* stream) {
  FMOV_REQUIRE(B >= 0 && S > 0 && z_stride >= S + z_off, "fmov_sdf_query_rays_pair: bad shape B=%lld S=%d stride=%d off=%d", B,
               S, z_stride, z_off);
  FMOV_REQUIRE(B == 0 || (rays_o && rays_d && z && out && wblob_pair && w8_row0), "fmov_sdf_query_rays_pair: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 1; a.P = B * S; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.S = S; a.z_stride = z_stride; a.z_off = z_off;
  a.in_scale = in_scale; a.out_scale = out_scale; a.bias = nullptr; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob_pair, 0, (cudaStream_t)stream, nullptr, true);
}

extern "C" int fmov_sdf_query_grid(const float* bmin3, const float* bmax3, int res, long long first, long long count,
                                   const void* wblob, const float* bias8x256, const float* w8_row0, const float* b8,
                                   float in_scale, float out_scale, float* out, void* stream) {
  FMOV_REQUIRE(res > 0 && first >= 0 && count >= 0 && first + count <= (long long)res * res * res,
               "fmov_sdf_query_grid: bad range first=%lld count=%lld res=%d", first, count, res);
  FMOV_REQUIRE(bmin3 && bmax3 && (count == 0 || (out && wblob && bias8x256 && w8_row0)), "fmov_sdf_query_grid: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 2; a.P = count; a.res = res; a.grid_off = first;
  for (int i = 0; i < 3; ++i) { a.bmin[i] = bmin3[i]; a.bmax[i] = bmax3[i]; }
  a.in_scale = in_scale; a.out_scale = out_scale; a.bias = bias8x256; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob, 0, (cudaStream_t)stream);
}

// ---- split-precision ("precise") entry points: SDF to ~1e-5 of the fp32 network everywhere (north_star: SDF <= 1e-3), for
// SDFNetwork.sdf and extract_fields (models/fields.py:106-107, models/renderer.py:9-37, :506).  `wblob_lo` = the images of
// the fp16 residuals W - fp16(W), same layout as `wblob` (fmov_pack_image with fmt = 2); NULL selects the activation-split
// chain (weights plain fp16, activations hi + lo: ~3e-4, 1.8x the throughput).
extern "C" int fmov_sdf_query_points_precise(const float* pts, long long P, const void* wblob, const void* wblob_lo,
                                             const float* bias8x256, const float* w8_row0, const float* b8, float in_scale,
                                             float out_scale, float* out, void* stream) {
  FMOV_REQUIRE(P >= 0 && (P == 0 || (pts && out && wblob && bias8x256 && w8_row0)),
               "fmov_sdf_query_points_precise: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 0; a.P = P; a.pts = pts; a.in_scale = in_scale; a.out_scale = out_scale;
  a.bias = bias8x256; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob, 0, (cudaStream_t)stream, wblob_lo, false, true);
}

extern "C" int fmov_sdf_query_grid_precise(const float* bmin3, const float* bmax3, int res, long long first, long long count,
                                           const void* wblob, const void* wblob_lo, const float* bias8x256,
                                           const float* w8_row0, const float* b8, float in_scale, float out_scale, float* out,
                                           void* stream) {
  FMOV_REQUIRE(res > 0 && first >= 0 && count >= 0 && first + count <= (long long)res * res * res,
               "fmov_sdf_query_grid_precise: bad range first=%lld count=%lld res=%d", first, count, res);
  FMOV_REQUIRE(bmin3 && bmax3 && (count == 0 || (out && wblob && bias8x256 && w8_row0)),
               "fmov_sdf_query_grid_precise: null argument");
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 2; a.P = count; a.res = res; a.grid_off = first;
  for (int i = 0; i < 3; ++i) { a.bmin[i] = bmin3[i]; a.bmax[i] = bmax3[i]; }
  a.in_scale = in_scale; a.out_scale = out_scale; a.bias = bias8x256; a.w8 = w8_row0; a.b8 = b8; a.out = out;
  return launch_query(a, wblob, 0, (cudaStream_t)stream, wblob_lo, false, true);
}
