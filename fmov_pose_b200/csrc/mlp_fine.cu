// Fine stage of NeuSRenderer.render_core as two fused chain kernels (forward, backward).
//
//   fmov_fine_fwd   per sample: SDF value + 256-d feature (SDFNetwork.forward, models/fields.py:88-104),
//                   analytic normal d sdf/dx (replaces the second forward + autograd.grad of
//                   SDFNetwork.gradient, models/fields.py:112-124), colour MLP (RenderingNetwork.forward,
//                   models/fields.py:166-193); call sites models/renderer.py:277-288.
//   fmov_fine_bwd   the matching part of loss.backward() (exp_runner.py:802): given dL/d{sdf, normal, rgb} per
//                   sample it produces dL/d{point, view dir} and every activation-gradient tile the weight
//                   gradient GEMMs (mlp_dw.cu) need, including the second-order path through the normal.
//
// Math: oracle/explicit_adjoint.py (checked against autograd) — same symbols:
//   value pass   z_l = W_l u_l + b_l, h_{l+1} = softplus_100(z_l), sigma_l = 1 - exp(-100 h_{l+1})
//   reverse sweep delta_7 = W_8[0,:]*sigma_7;  v_l = W_l^T delta_l;  delta_{l-1} = v_l*sigma_{l-1};  n = J_e^T g_e
//   adjoint pass dbar_l = W_l vbar_l;  vbar_{l+1} = dbar_l*sigma_l;  q_l = 100*dbar_l*delta_l*(1-sigma_l)
//   backward     zbar_{l-1} = (W_l^T zbar_l)*sigma_{l-1} + q_{l-1}
// fp16 operands in the forward (value pass, reverse sweep, colour net); gradient tiles are fp16 with a per-step
// power-of-two loss scale derived from max|upstream gradient| (fmov_grad_amax; bf16 tiles missed the 1e-2 gradient
// bar on cancellation-heavy sums); fp32 accumulation in TMEM; fp32 epilogue math.
// Epilogue warpgroups per tile slot (FMOV_FINE_WGS, default 2): with two, warpgroup w owns columns [128w, 128w+128) of
// every 256-wide tile (32-column chunks hb = 4w..4w+3); warpgroup WGS-1 additionally owns the per-row state (normal,
// x-bar, PE adjoints) and the narrow N=48 / N=16 steps.  A thread still only re-reads stash rows/chunks it wrote itself.
#ifndef FMOV_FINE_WGS
#define FMOV_FINE_WGS 2
#endif
#define FMOV_CH_WGS FMOV_FINE_WGS
// CTA-pair mode of the chain engine (mlp_chain.cuh): clusters of two CTAs, tcgen05.mma.cta_group::2, each CTA stages half of
// every weight slice.  -DFMOV_FINE_PAIR=0 builds the one-CTA engine of rounds 1-2 (kept for A/B measurements; the default
// lives in mlp_chain.cuh).
#define FMOV_CH_PAIR FMOV_FINE_PAIR
#include "mlp_chain.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

using FL = ChainLayout;
constexpr int HPW = 8 / CH_WGS;          // 32-column chunks per warpgroup
constexpr int NCK = CH_CHUNKS;           // 16-column chunks per warpgroup
constexpr int PFD = 1;                   // stash-read prefetch distance (chunks) in the backward hot loops
constexpr int RQ_PH = 2;                 // prefetch distance (8-column pieces) of the three operand streams that rebuild q
// Measured on B200 and decided (round 2, profiles/r2_fine_variants.txt; alternating builds on one box, 8192 rays x 128):
//   kept     ReLU signs of C1..C4 as one 32-bit word per row and 32 columns (16 -> 1 block per tile read by the colour
//            backward); q_l rebuilt in the ordinary backward from V-bar, delta and sigma instead of stored (-32 blocks);
//            L2 eviction hints in fine_bwd (first read of H evict_last, single-use traffic evict_first):
//            fine_bwd 7.72 -> 7.16 ms, stash 203 -> 172 blocks per tile (27 -> 23 GB per 8192 rays)
//   dropped  eviction hints in fine_fwd (5.17 -> 5.4 ms), bulk L2 prefetch from the producer warp, thread-issued prefetch
//            one step ahead, PFD = 2 (spills at 96 registers), double-buffered TMEM loads and peeled special-case layers,
//            setmaxnreg re-balancing (re-checked in round 2 with ptxas 12.9 -v: a setmaxnreg.dec anywhere in the kernel caps
//            the WHOLE kernel at the dec value — 3.6 KB of spills at dec 40 with or without the epilogue's inc, with or
//            without separate kernel tails per role — and an inc alone frees nothing: 'Used 96 registers' either way)
// -18 % fine_bwd traffic bought -7 % time: the epilogues are latency bound (16 warps, long-scoreboard 9.9 cycles per
// issue), not bandwidth bound — see DESIGN.md §3.
#define FINE_BOUNDS __launch_bounds__(CH_THREADS, 1)
// fine_bwd reads H twice (adjoint pass, then ordinary backward): KEEP / STREAM name the L2 eviction intent at each site
#define LD_KEEP(tp, ck, q) chunk_load_pol(tp, ck, q, pol_keep)
#define LD_STREAM(tp, ck, q) chunk_load_pol(tp, ck, q, pol_stream)
#define ST_STREAM(tp, ck, q) chunk_store_pol(tp, ck, q, pol_stream)
#define PF_KEEP(tp, a, b) tile_prefetch_l2_keep(tp, a, b)
// Register budget: 640 threads put 5 warps on every SM sub-partition (16 K registers each), so ptxas caps the kernels
// at 96 registers/thread.  The hot loops are written in 16-column chunks so that they fit 96 registers without spills;
// the remaining spills (ptxas -v) sit in the once-per-tile positional-encoding steps.

// ---- weight image directory ---------------------------------------------------------------------
enum ImgId {
  IMG_F0 = 0,          // F0..F8   forward images (fp16): lin0..lin7, lin8 feature rows
  IMG_T0 = 9,          // T0..T7   transposed images (fp16) for the reverse sweep
  IMG_C0 = 17,         // C0..C4   colour net forward images (fp16)
  IMG_CT0A = 22, IMG_CT0B = 23, IMG_CT1 = 24, IMG_CT2 = 25, IMG_CT3 = 26,   // colour transposed (bf16)
  IMG_FB0 = 27,        // FB0..FB7 forward images in bf16 (adjoint pass)
  IMG_TB0 = 35,        // TB0..TB8 transposed images in bf16 (backward pass; TB8 = lin8 feature rows^T)
  IMG_FP0 = 44,        // FP0..FP7 copies of F0..F7 for the fine kernels (F0..F7 themselves are the sampling queries' blob)
  IMG_COUNT = 52
};
struct ImgInfo { int npad, kblocks; };
__host__ __device__ inline ImgInfo img_info(int id) {
  if (id >= IMG_F0 && id < IMG_F0 + 9) {
    const int l = id - IMG_F0;
    return l == 0 ? ImgInfo{256, 1} : l == 3 ? ImgInfo{224, 4} : l == 4 ? ImgInfo{256, 5} : ImgInfo{256, 4};
  }
  if (id >= IMG_T0 && id < IMG_T0 + 8) return (id - IMG_T0) == 0 ? ImgInfo{48, 4} : ImgInfo{256, 4};
  if (id >= IMG_C0 && id < IMG_C0 + 5) {
    const int l = id - IMG_C0;
    return l == 0 ? ImgInfo{256, 5} : l == 4 ? ImgInfo{16, 4} : ImgInfo{256, 4};
  }
  if (id == IMG_CT0B) return ImgInfo{48, 4};
  if (id >= IMG_CT0A && id <= IMG_CT3) return ImgInfo{256, 4};
  if (id >= IMG_FB0 && id < IMG_FB0 + 8) return img_info(IMG_F0 + (id - IMG_FB0));
  if (id >= IMG_TB0 && id < IMG_TB0 + 9) return (id - IMG_TB0) == 0 ? ImgInfo{48, 4} : ImgInfo{256, 4};
  if (id >= IMG_FP0 && id < IMG_FP0 + 8) return img_info(IMG_F0 + (id - IMG_FP0));
  return ImgInfo{0, 0};
}
// Layout of an image's 64-wide k-block: row-interleaved [8 chunk columns][N rows][16 B] (F0..F7: shared with the sampling
// queries), or — every image only the fine kernels read, in CTA-pair mode — half-major [2 halves][8 chunk columns][N/2 rows]
// [16 B], so that the N/2 rows one CTA of a pair stages are contiguous.
__host__ __device__ inline bool img_half_major(int id) { return CH_PAIR && !(id >= IMG_F0 && id < IMG_F0 + 8); }
// Images of the biased forward layers carry, behind their k-blocks, a [N x 16] slice (half-major like the k-blocks) whose
// K columns 0 / 1 hold the layer's bias as fp16 hi + fp16 residual (x 2^12, so that it is a normal fp16 number): the chain's bias16 step multiplies it with the 1.0
// columns of the AUX block (pair mode only; the one-CTA engine adds the fp32 bias in the epilogue).
constexpr bool FINE_BIAS_MMA = CH_PAIR;
__host__ __device__ inline bool img_has_bias(int id) {
  return FINE_BIAS_MMA && (id == IMG_F0 + 8 || (id >= IMG_C0 && id < IMG_C0 + 4) || (id >= IMG_FP0 && id < IMG_FP0 + 8));
}
__host__ __device__ inline long long img_bytes(int id) {
  const ImgInfo ii = img_info(id);
  return (long long)ii.npad * 128 * ii.kblocks + (img_has_bias(id) ? (long long)ii.npad * 32 : 0);
}
__constant__ long long c_img_offset[IMG_COUNT + 1];
__device__ __forceinline__ long long img_offset_dev(int id) { return c_img_offset[id]; }
static long long img_offset(int id) {
  long long off = 0;
  for (int i = 0; i < id; ++i) off += img_bytes(i);
  return off;
}


// ---- stash tensors --------------------------------------------------------------------------------
enum StashId {
  ST_PE = 0, ST_H1 = 1 /*..H8=8*/, ST_F = 9, ST_D0 = 10 /*..D7=17*/, ST_X = 18, ST_C1 = 19 /*..C4=22*/,
  ST_ZC0 = 23 /*..ZC3=26*/, ST_FB = 27, ST_GE = 28, ST_V1 = 29 /*..V8=36*/, ST_Q0 = 37 /*..Q7=44*/,
  ST_Z0 = 45 /*..Z7=52*/,
  ST_CM = 53, ST_COUNT = 54     // experiment: ReLU sign bits of C1..C4, [layer][32-column chunk][row] uint32 = one block per tile
};
__host__ __device__ inline int stash_kb(int id) {
  if (id == ST_CM) return 1;
  if (id >= ST_Q0 && id < ST_Q0 + 8) return 0;          // q is never materialised: 32 blocks per tile less stash memory
  return (id == ST_PE || id == ST_X || id == ST_GE) ? 1 : 4;
}
// written by fmov_fine_fwd (a forward-only stash holds exactly these)
__host__ __device__ inline bool stash_is_forward(int id) {
  if (id == ST_CM) return true;
  return id <= ST_C1 + 3;
}

struct FineArgs {
  long long B;
  int S;
  const float* rays_o; const float* rays_d; const float* z;   // [B,3],[B,3],[B,S]
  float sample_dist;
  // fp32 side parameters
  const float* bias_sdf;   // [8][256]
  const float* b8;         // [257]
  const float* w8row;      // [256]  lin8 row 0 (effective)
  const float* bias_col;   // [4][256]
  const float* bc4;        // [3]
  const float* wc4;        // [3][256]  colour lin4 (effective), fp32 (backward only)
  // per-sample fp32 tensors
  float* sdf; float* nrm; float* rgb; float* ge;              // fwd outputs: [P],[P,3],[P,3], g_e scratch (row40)
  const float* d_sdf; const float* d_nrm; const float* d_rgb; // bwd inputs
  float* d_pts; float* d_dirs; float* zc4;                    // bwd outputs: [P,3],[P,3],[P,4]
  float* eb;                                                  // bwd scratch (row40)
  const float* amax;                                          // bwd: device scalar, max |upstream gradient|
};

__device__ __forceinline__ uint8_t* stash_tile(const ChainPtrs& ptrs, int id, long long tile) {
  return ptrs.stash[id] + (size_t)tile * stash_kb(id) * BLK_BYTES;
}

struct PointCtx { bool valid; long long p; float x[3]; float d[3]; };
// per-sample scratch rows of 40 floats (g_e of the forward, e-bar of the backward) live tile-major and column-major inside a
// tile, [tile][40][128]: entry i of row r of tile t is at (t*40 + i)*128 + r, so the 32 lanes of a warp (consecutive rows)
// touch 128 contiguous bytes per access instead of 32 sectors (the row-major [P,40] layout cost fine_fwd's normal step
// 27 K cycles per tile).  The buffers hold 40*128*ceil(P/128) floats and are opaque to the caller.
__device__ __forceinline__ long long row40(long long tile, int row, int i) { return (tile * 40 + i) * TILE_M + row; }
__device__ __forceinline__ PointCtx load_sample(const FineArgs& a, long long tile, int row) {
  PointCtx c;
  c.p = tile * TILE_M + row;
  c.valid = c.p < a.B * a.S;
  c.x[0] = c.x[1] = c.x[2] = 0.f;
  c.d[0] = c.d[1] = c.d[2] = 0.f;
  if (c.valid) {
    const long long r = c.p / a.S;
    const int j = (int)(c.p - r * a.S);
    const float z0 = a.z[c.p];
    const float dist = (j + 1 < a.S) ? a.z[c.p + 1] - z0 : a.sample_dist;
    const float mid = z0 + dist * 0.5f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c.d[i] = a.rays_d[r * 3 + i];
      c.x[i] = a.rays_o[r * 3 + i] + c.d[i] * mid;
    }
  }
  return c;
}

// PE(L) of a 3-vector into e[3+6L]: [x, sin(2^k x), cos(2^k x)]  (models/embedder.py:28-37)
template <int L>
__device__ __forceinline__ void pe_eval(const float x[3], float* e) {
  e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(x[c] * f, &s, &co);
      e[3 + 6 * k + c] = s;
      e[6 + 6 * k + c] = co;
    }
  }
}
// columns PE_RES_COL..+2 of the 64-wide encoded-input block: what fp16 drops of the raw coordinates (|x| up to ~1.75 has an
// fp16 spacing of 9.8e-4, and d sdf / d x ~ 1: the largest single contribution to the SDF error of the fp16 chain)
__device__ __forceinline__ void pe_with_residual(const float x[3], float* e64) {
  pe_eval<6>(x, e64);
#pragma unroll
  for (int c = 0; c < 3; ++c) e64[PE_RES_COL + c] = x[c] - __half2float(__float2half_rn(x[c]));
}
// n = J_e^T g   (g has 3+6L entries)
template <int L>
__device__ __forceinline__ void pe_jt(const float x[3], const float* g, float n[3]) {
  n[0] = g[0]; n[1] = g[1]; n[2] = g[2];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      fast_sincos(x[c] * f, &s, &co);
      n[c] += g[3 + 6 * k + c] * f * co - g[6 + 6 * k + c] * f * s;
    }
  }
}

__device__ __forceinline__ uint8_t* stash_row(const ChainPtrs& ptrs, int id, long long tile, int kb, int row) {
  return ptrs.stash[id] + ((size_t)tile * stash_kb(id) + kb) * BLK_BYTES + (size_t)row * 16;
}
// store one 32-column chunk (index hb = 0..7 within a 256-wide tile) to the ACT operand and/or a stash tensor
__device__ __forceinline__ void put_chunk(const EpiCtx& c, const ChainPtrs& ptrs, bool to_act, int stash_id,
                                          long long tile, int hb, bool bf16, const float* v) {
  uint4 q[4];
  pack4(v, bf16, q);
  if (to_act) row_half_store(c.act + (hb >> 1) * BLK_BYTES + c.row * 16, hb & 1, q);
  if (stash_id >= 0) row_half_store(stash_row(ptrs, stash_id, tile, hb >> 1, c.row), hb & 1, q);
}
__device__ __forceinline__ void put_chunk_grad(const EpiCtx& c, const ChainPtrs& ptrs, bool to_act, int stash_id,
                                               long long tile, int hb, const float* v) {
  uint4 q[4];
  pack4_grad(v, q);
  if (to_act) row_half_store(c.act + (hb >> 1) * BLK_BYTES + c.row * 16, hb & 1, q);
  if (stash_id >= 0) row_half_store(stash_row(ptrs, stash_id, tile, hb >> 1, c.row), hb & 1, q);
}
// hoisted-base variants: `tp` = stash tile base + row*16 (or ACT base + row*16); chunk hb = 32 columns
__device__ __forceinline__ uint8_t* tile_base(const ChainPtrs& ptrs, int id, long long tile, int row) {
  return ptrs.stash[id] + (size_t)tile * stash_kb(id) * BLK_BYTES + (size_t)row * 16;
}
__device__ __forceinline__ void ld_half(const uint8_t* tp, int hb, uint4* q) {
  row_half_load(tp + (hb >> 1) * BLK_BYTES, hb & 1, q);
}
__device__ __forceinline__ void st_half(uint8_t* tp, int hb, const uint4* q) {
  row_half_store(tp + (hb >> 1) * BLK_BYTES, hb & 1, q);
}
__device__ __forceinline__ void get_chunk_raw(const EpiCtx& c, const ChainPtrs& ptrs, int stash_id, long long tile,
                                              int hb, uint4* q) {
  row_half_load(stash_row(ptrs, stash_id, tile, hb >> 1, c.row), hb & 1, q);
}

// =====================================================================================================
// forward
// =====================================================================================================
__global__ void FINE_BOUNDS
fine_fwd_kernel(const __grid_constant__ ChainTable tb, const __grid_constant__ ChainPtrs ptrs,
                const __grid_constant__ FineArgs a, const __grid_constant__ PairMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = chain_smem_base(smem_raw);
  ChainSmem* s = reinterpret_cast<ChainSmem*>(base);
  uint8_t* act0 = base + FL::ACT;
  uint8_t* aux0 = base + FL::AUX;
  uint8_t* wst = base + FL::WST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long P = a.B * a.S;
  const long long n_tiles = (P + TILE_M - 1) / TILE_M;
  // pair mode: both CTAs of a pair walk as many tiles as the even one has (the odd CTA pads with a dummy tile)
  const uint32_t rank = CH_PAIR ? cluster_ctarank() : 0u;
  const long long first_tile = CH_PAIR ? (long long)(blockIdx.x & ~1u) : (long long)blockIdx.x;
  const int n_my = (int)((n_tiles - first_tile + gridDim.x - 1) / gridDim.x);

  if (threadIdx.x == 0) chain_init_barriers(s);
  if (warp == ISSUER_WARP) {
    if (CH_PAIR) tmem_alloc_pair(&s->tmem_base, 512);
    else tmem_alloc(&s->tmem_base, 512);
  }
  tc_fence_before();
  if (CH_PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s->tmem_base;

  if (warp >= CTRL_WARP0) {
    if (warp == PRODUCER_WARP) {
      if (lane == 0) {
        if (CH_PAIR) chain_weight_producer_pair<FINE_BIAS_MMA>(tb, maps, s, wst, n_my, rank);
        else chain_weight_producer(tb, ptrs, s, wst, n_my, blockIdx.x, gridDim.x);
      }
    } else if (warp == ISSUER_WARP) {
      if (lane == 0) {
        if (!CH_PAIR) chain_mma_issuer(tb, s, act0, aux0, wst, tmem, n_my);
        else if (rank == 0) chain_mma_issuer_pair<FINE_BIAS_MMA>(tb, s, act0, aux0, wst, tmem, n_my);
      }
    }
  } else {
    EpiCtx c;
    epi_init(c, s, act0, aux0, tmem);
    const int hb0 = c.wg * HPW;
    const int ck0 = c.wg * NCK;
    const bool owner = c.wg == CH_WGS - 1;      // warpgroup that owns the per-row state and the narrow steps
    for (int k = c.slot; k < n_my; k += CH_SLOTS) {
      const long long tile_raw = (long long)blockIdx.x + (long long)k * gridDim.x;
      // dummy tile of the odd CTA (pair mode): every point invalid, stash traffic goes to the padding tile n_tiles
      const long long tile = tile_raw < n_tiles ? tile_raw : n_tiles;
      const PointCtx pc = load_sample(a, tile_raw, c.row);
      // ---- input: PE6(x) -> AUX + stash ----------------------------------------------------------------
      {
        float e[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) e[i] = 0.f;
        pe_with_residual(pc.x, e);
        if (FINE_BIAS_MMA) { e[AUX_ONE_COL] = 1.0f; e[AUX_ONE_COL + 1] = 1.0f / BIAS_LO_SCALE; }      // x the bias slices of lin0..lin8
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (CH_WGS == 2 && h != c.wg) continue;        // each warpgroup stores one 32-column half
          uint4 q[4];
          pack4(e + 32 * h, false, q);
          row_half_store(c.aux + c.row * 16, h, q);
          row_half_store(stash_row(ptrs, ST_PE, tile, 0, c.row), h, q);
        }
        epi_signal_act(c);
      }
      // ---- value pass: layers 0..7 -----------------------------------------------------------------------
      float sdf = owner ? __ldg(a.b8) : 0.f;
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const int n_valid = (l == 3) ? 217 : 256;
        const int n_mma = (l == 3) ? 224 : 256;
        const float* bias = a.bias_sdf + l * 256;
        // the chunk's 16 biases are fetched one chunk ahead (the first before the accumulator wait): the broadcast loads are
        // L1 hits, but issued next to their use they still cost ~25 % of this loop in long-scoreboard stalls (ncu source view)
        float4 bb[4];
        if (!FINE_BIAS_MMA) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) bb[j4] = __ldg(reinterpret_cast<const float4*>(bias + ck0 * 16) + j4);
        }
        epi_wait_acc(c);
        // 16-column chunks; the TMEM load of chunk i+1 is in flight while chunk i is evaluated (two register buffers)
        uint8_t* actp = c.act + c.row * 16;
        uint8_t* hsp = tile_base(ptrs, ST_H1 + l, tile, c.row);
        float vbuf[2][16];
        tmem_ld16(c.tmem + ck0 * 16, vbuf[0]);
#pragma unroll
        for (int i = 0; i < NCK; ++i) {
          const int ck = ck0 + i;
          float* v = vbuf[i & 1];
          tmem_ld_wait();
          if (i + 1 < NCK && (ck + 1) * 16 < n_mma) tmem_ld16(c.tmem + (ck + 1) * 16, vbuf[(i + 1) & 1]);
          if (ck * 16 < n_mma) {
            if (!FINE_BIAS_MMA) {          // (pair mode: the accumulator already holds W u + b)
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                v[j4 * 4 + 0] += bb[j4].x; v[j4 * 4 + 1] += bb[j4].y; v[j4 * 4 + 2] += bb[j4].z; v[j4 * 4 + 3] += bb[j4].w;
              }
              if (i + 1 < NCK) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) bb[j4] = ldg_f4_volatile(reinterpret_cast<const float4*>(bias + (ck + 1) * 16) + j4);
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = softplus100(v[j]);
            if (ck * 16 + 16 > n_valid) {           // lin3: columns 217..223 are padding
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) v[jj] = (ck * 16 + jj < n_valid) ? v[jj] : 0.f;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] = 0.f;
          }
          if (l == 7) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w8row + ck * 16) + j4);
              sdf = fmaf(v[j4 * 4 + 0], w4.x, sdf); sdf = fmaf(v[j4 * 4 + 1], w4.y, sdf);
              sdf = fmaf(v[j4 * 4 + 2], w4.z, sdf); sdf = fmaf(v[j4 * 4 + 3], w4.w, sdf);
            }
          }
          uint4 q2[2];
          pack2(v, false, q2);
          chunk_store(actp, ck, q2);
          chunk_store(hsp, ck, q2);
        }
        epi_signal_act(c);
      }
      if (CH_WGS == 2) {       // combine the two column halves of the lin8-row-0 dot product
        if (!owner) s->scratch[c.slot][c.row] = sdf;
        slot_sync(c);
        if (owner) sdf += s->scratch[c.slot][c.row];
      }
      if (owner && pc.valid) a.sdf[pc.p] = sdf;
      // ---- lin8 feature rows -> stash F ; delta_7 = W8[0,:]*sigma_7 in place -------------------------------------
      epi_wait_acc(c);
#pragma unroll 2
      for (int hi = 0; hi < HPW; ++hi) {
        const int hb = hb0 + hi;
        float v[32], h[32];
        uint4 q[4];
        acc_load32(c, hb * 32, v);
        row_half_load(c.act + (hb >> 1) * BLK_BYTES + c.row * 16, hb & 1, q);
        unpack4(q, false, h);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = hb * 32 + j;
          if (!FINE_BIAS_MMA) v[j] += __ldg(a.b8 + 1 + col);
          h[j] = __ldg(a.w8row + col) * sigma_from_h(h[j]);
        }
        put_chunk(c, ptrs, false, ST_F, tile, hb, false, v);
        put_chunk(c, ptrs, true, ST_D0 + 7, tile, hb, false, h);
      }
      epi_signal_act(c);
      // ---- reverse sweep l = 7..1 -------------------------------------------------------------------------------
#pragma unroll 1
      for (int l = 7; l >= 1; --l) {
        const uint8_t* hp = tile_base(ptrs, ST_H1 + (l - 1), tile, c.row);      // H_l -> sigma_{l-1} (own rows)
        uint8_t* dp = tile_base(ptrs, ST_D0 + (l - 1), tile, c.row);
        uint8_t* ap = c.act + c.row * 16;
        uint4 sb[2][4];
        ld_half(hp, hb0, sb[0]);
        tile_prefetch_l2(hp, ck0 + 2, NCK - 2);
        epi_wait_acc(c);
#pragma unroll
        for (int hi = 0; hi < HPW; ++hi) {
          const int hb = hb0 + hi;
          float v[32], h[32];
          if (hi < HPW - 1) ld_half(hp, hb + 1, sb[(hi + 1) & 1]);
          acc_load32(c, hb * 32, v);
          unpack4(sb[hi & 1], false, h);
          if (l == 4 && hb == 6 && pc.valid) {
            // columns 217..255 of v_4 are the PE part of the skip input (1/sqrt2 folded into the image): parked in
            // the g_e output row until the W_0^T delta_0 term arrives (keeps 39 registers free across the sweep)
#pragma unroll
            for (int j = 25; j < 32; ++j) a.ge[row40(tile, c.row, j - 25)] = v[j];
          }
          if (l == 4 && hb == 7 && pc.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a.ge[row40(tile, c.row, 7 + j)] = v[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= sigma_from_h(h[j]);    // H_4 is zero beyond col 216 -> delta_3 too
          uint4 q[4];
          pack4(v, false, q);
          st_half(ap, hb, q);
          st_half(dp, hb, q);
        }
        epi_signal_act(c);
      }
      // ---- g_e += W_0^T delta_0 ; normal ; colour-net extras ; feature tile back into ACT -----------------------------
      epi_wait_acc(c);
      if (owner) {
        float v[48];
        float nrm[3];
        float ge[40];
        acc_load32(c, 0, v);
        acc_load16(c, 32, v + 32);
        ge[39] = 0.f;
#pragma unroll
        for (int i = 0; i < 39; ++i) ge[i] = v[i] + (pc.valid ? a.ge[row40(tile, c.row, i)] : 0.f);
        pe_jt<6>(pc.x, ge, nrm);
        if (pc.valid) {
#pragma unroll
          for (int i = 0; i < 3; ++i) a.nrm[pc.p * 3 + i] = nrm[i];
#pragma unroll
          for (int i = 0; i < 40; ++i) a.ge[row40(tile, c.row, i)] = ge[i];
        }
        // extras = [pts(3), PE4(dirs)(27), normals(3)]   (models/fields.py:172-175)
        float e[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) e[i] = 0.f;
        e[0] = pc.x[0]; e[1] = pc.x[1]; e[2] = pc.x[2];
        pe_eval<4>(pc.d, e + 3);
        e[30] = nrm[0]; e[31] = nrm[1]; e[32] = nrm[2];
        if (FINE_BIAS_MMA) { e[AUX_ONE_COL] = 1.0f; e[AUX_ONE_COL + 1] = 1.0f / BIAS_LO_SCALE; }      // x the bias slices of colour lin0..lin3
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 q[4];
          pack4(e + 32 * h, false, q);
          row_half_store(c.aux + c.row * 16, h, q);
          row_half_store(stash_row(ptrs, ST_X, tile, 0, c.row), h, q);
        }
      }
      {
#pragma unroll
        for (int hi = 0; hi < HPW; ++hi) {        // own chunks of F: written above by this thread
          const int hb = hb0 + hi;
          uint4 q[4];
          get_chunk_raw(c, ptrs, ST_F, tile, hb, q);
          row_half_store(c.act + (hb >> 1) * BLK_BYTES + c.row * 16, hb & 1, q);
        }
        epi_signal_act(c);
      }
      // ---- colour net: 4 ReLU layers ------------------------------------------------------------------------------
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
        const float* bias = a.bias_col + l * 256;
        epi_wait_acc(c);
#pragma unroll 2
        for (int hi = 0; hi < HPW; ++hi) {
          const int hb = hb0 + hi;
          float v[32];
          acc_load32(c, hb * 32, v);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!FINE_BIAS_MMA) b4 = __ldg(reinterpret_cast<const float4*>(bias + hb * 32) + j4);
            v[j4 * 4 + 0] = fmaxf(v[j4 * 4 + 0] + b4.x, 0.f);
            v[j4 * 4 + 1] = fmaxf(v[j4 * 4 + 1] + b4.y, 0.f);
            v[j4 * 4 + 2] = fmaxf(v[j4 * 4 + 2] + b4.z, 0.f);
            v[j4 * 4 + 3] = fmaxf(v[j4 * 4 + 3] + b4.w, 0.f);
          }
          put_chunk(c, ptrs, true, ST_C1 + l, tile, hb, false, v);
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
          reinterpret_cast<uint32_t*>(stash_tile(ptrs, ST_CM, tile))[(l * 8 + hb) * TILE_M + c.row] = m;
        }
        epi_signal_act(c);
      }
      // ---- colour output: sigmoid ------------------------------------------------------------------------------------
      {
        epi_wait_acc(c);
        if (owner) {
          float v[16];
          acc_load16(c, 0, v);
          if (pc.valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) a.rgb[pc.p * 3 + i] = sigmoidf_(v[i] + __ldg(a.bc4 + i));
          }
        }
        tc_fence_before();
      }
    }
  }
  if (CH_PAIR) cluster_sync_all();      // neither CTA may leave while its partner still signals its barriers / reads its smem
  else __syncthreads();
  if (warp == ISSUER_WARP) {
    tc_fence_after();
    if (CH_PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

// =====================================================================================================
// backward
// =====================================================================================================
__global__ void FINE_BOUNDS
fine_bwd_kernel(const __grid_constant__ ChainTable tb, const __grid_constant__ ChainPtrs ptrs,
                const __grid_constant__ FineArgs a, const __grid_constant__ PairMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = chain_smem_base(smem_raw);
  ChainSmem* s = reinterpret_cast<ChainSmem*>(base);
  uint8_t* act0 = base + FL::ACT;
  uint8_t* aux0 = base + FL::AUX;
  uint8_t* wst = base + FL::WST;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long P = a.B * a.S;
  const long long n_tiles = (P + TILE_M - 1) / TILE_M;
  // pair mode: both CTAs of a pair walk as many tiles as the even one has (the odd CTA pads with a dummy tile)
  const uint32_t rank = CH_PAIR ? cluster_ctarank() : 0u;
  const long long first_tile = CH_PAIR ? (long long)(blockIdx.x & ~1u) : (long long)blockIdx.x;
  const int n_my = (int)((n_tiles - first_tile + gridDim.x - 1) / gridDim.x);

  if (threadIdx.x == 0) chain_init_barriers(s);
  if (warp == ISSUER_WARP) {
    if (CH_PAIR) tmem_alloc_pair(&s->tmem_base, 512);
    else tmem_alloc(&s->tmem_base, 512);
  }
  tc_fence_before();
  if (CH_PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s->tmem_base;

  if (warp >= CTRL_WARP0) {
    if (warp == PRODUCER_WARP) {
      if (lane == 0) {
        if (CH_PAIR) chain_weight_producer_pair<false>(tb, maps, s, wst, n_my, rank);
        else chain_weight_producer(tb, ptrs, s, wst, n_my, blockIdx.x, gridDim.x);
      }
    } else if (warp == ISSUER_WARP) {
      if (lane == 0) {
        if (!CH_PAIR) chain_mma_issuer(tb, s, act0, aux0, wst, tmem, n_my);
        else if (rank == 0) chain_mma_issuer_pair<false>(tb, s, act0, aux0, wst, tmem, n_my);
      }
    }
  } else {
    EpiCtx c;
    epi_init(c, s, act0, aux0, tmem);
    const int hb0 = c.wg * HPW;
    const int ck0 = c.wg * NCK;
    const bool owner = c.wg == CH_WGS - 1;
    const float gscale = grad_scale_from_amax(__ldg(a.amax));
    const float ginv = 1.0f / gscale;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    for (int k = c.slot; k < n_my; k += CH_SLOTS) {
      const long long tile_raw = (long long)blockIdx.x + (long long)k * gridDim.x;
      // dummy tile of the odd CTA (pair mode): every point invalid, stash traffic goes to the padding tile n_tiles
      const long long tile = tile_raw < n_tiles ? tile_raw : n_tiles;
      const PointCtx pc = load_sample(a, tile_raw, c.row);
      float sbar = 0.f, nbar[3] = {0.f, 0.f, 0.f}, zc4[3] = {0.f, 0.f, 0.f}, xbar[3] = {0.f, 0.f, 0.f};
      if (pc.valid) {
        sbar = a.d_sdf[pc.p] * gscale;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          nbar[i] = a.d_nrm[pc.p * 3 + i] * gscale;
          const float r = a.rgb[pc.p * 3 + i];
          zc4[i] = a.d_rgb[pc.p * 3 + i] * gscale * r * (1.f - r);     // sigmoid backward (loss-scaled)
        }
        if (owner) {
#pragma unroll
          for (int i = 0; i < 3; ++i) a.zc4[pc.p * 4 + i] = zc4[i];
          a.zc4[pc.p * 4 + 3] = 0.f;
        }
      }
      // ---- colour lin4 backward on CUDA cores (3 x 256) + ReLU mask of C4 -> zbar_c3 ---------------------
#pragma unroll 2
      for (int hi = 0; hi < HPW; ++hi) {
        const int hb = hb0 + hi;
        float v[32];
        // experiment (off by default): the backward pass reads the ReLU signs as one word per 32 columns instead of the
        // fp16 C tiles (16 blocks -> 1 block per tile; DESIGN.md §3 "HBM budget")
        const uint32_t m = reinterpret_cast<const uint32_t*>(stash_tile(ptrs, ST_CM, tile))[(3 * 8 + hb) * TILE_M + c.row];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = hb * 32 + j;
          const float ab = zc4[0] * __ldg(a.wc4 + col) + zc4[1] * __ldg(a.wc4 + 256 + col) + zc4[2] * __ldg(a.wc4 + 512 + col);
          v[j] = ((m >> j) & 1u) ? ab : 0.f;
        }
        put_chunk_grad(c, ptrs, true, ST_ZC0 + 3, tile, hb, v);
      }
      epi_signal_act(c);
      // ---- colour layers 3..1:  zbar_c{l-1} = (zbar_cl W_cl) * [C_l > 0] -----------------------------------
#pragma unroll 1
      for (int l = 3; l >= 1; --l) {
        const uint8_t* cp = tile_base(ptrs, ST_C1 + (l - 1), tile, c.row);
        uint8_t* zp = tile_base(ptrs, ST_ZC0 + (l - 1), tile, c.row);
        uint8_t* ap = c.act + c.row * 16;
        (void)cp;
        uint32_t mw[HPW];
#pragma unroll
        for (int hi = 0; hi < HPW; ++hi)
          mw[hi] = reinterpret_cast<const uint32_t*>(stash_tile(ptrs, ST_CM, tile))[((l - 1) * 8 + hb0 + hi) * TILE_M + c.row];
        epi_wait_acc(c);
#pragma unroll
        for (int hi = 0; hi < HPW; ++hi) {
          const int hb = hb0 + hi;
          float v[32];
          acc_load32(c, hb * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((mw[hi] >> j) & 1u) ? v[j] : 0.f;
          uint4 q[4];
          pack4_grad(v, q);
          st_half(ap, hb, q);
          st_half(zp, hb, q);
        }
        epi_signal_act(c);
      }
      // ---- colour lin0 backward, extras part: pts-bar, PE4(dirs)-bar, normals-bar -------------------------
      epi_wait_acc(c);
      if (owner) {
        float v[48];
        acc_load32(c, 0, v);
        acc_load16(c, 32, v + 32);
        xbar[0] = v[0]; xbar[1] = v[1]; xbar[2] = v[2];
        nbar[0] += v[30]; nbar[1] += v[31]; nbar[2] += v[32];
        float dd[3];
        pe_jt<4>(pc.d, v + 3, dd);
        if (pc.valid) {
#pragma unroll
          for (int i = 0; i < 3; ++i) a.d_dirs[pc.p * 3 + i] = dd[i] * ginv;
        }
      }
      epi_signal_act(c);                 // ACT (zbar_c0) untouched: next step re-uses it
      // ---- colour lin0 backward, feature part -> fbar (bf16) -> stash only -----------------------------------
      epi_wait_acc(c);
#pragma unroll 2
      for (int hi = 0; hi < HPW; ++hi) {
        const int hb = hb0 + hi;
        float v[32];
        acc_load32(c, hb * 32, v);
        put_chunk_grad(c, ptrs, false, ST_FB, tile, hb, v);
      }
      // ---- adjoint of n = J_e^T g_e: gbar_e = J_e nbar -> AUX ; PE-Hessian term into xbar -------------------
      if (owner) {
        float e[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) e[i] = 0.f;
        e[0] = nbar[0]; e[1] = nbar[1]; e[2] = nbar[2];
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) {
          const float f = (float)(1 << kk);
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            float sn, co;
            fast_sincos(pc.x[ci] * f, &sn, &co);
            e[3 + 6 * kk + ci] = nbar[ci] * f * co;
            e[6 + 6 * kk + ci] = -nbar[ci] * f * sn;
            if (pc.valid) {
              const float gs = a.ge[row40(tile, c.row, 3 + 6 * kk + ci)], gc = a.ge[row40(tile, c.row, 6 + 6 * kk + ci)];
              xbar[ci] -= (gs * sn + gc * co) * f * f * nbar[ci];
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 q[4];
          pack4_grad(e + 32 * h, q);
          row_half_store(c.aux + c.row * 16, h, q);
          row_half_store(stash_row(ptrs, ST_GE, tile, 0, c.row), h, q);
        }
      }
      epi_signal_act(c);
      // ---- adjoint pass l = 0..7:  dbar_l = W_l vbar_l ; vbar_{l+1} = dbar_l*sigma_l ; q_l -------------------
      // 16-column chunks, H/delta chunks prefetched one chunk ahead (register budget: 96, see above)
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const int n_mma = (l == 3) ? 224 : 256;
        const uint8_t* hp = tile_base(ptrs, ST_H1 + l, tile, c.row);      // H_{l+1} -> sigma_l
        const uint8_t* dp = tile_base(ptrs, ST_D0 + l, tile, c.row);      // delta_l
        uint8_t* vp = tile_base(ptrs, ST_V1 + l, tile, c.row);
        uint8_t* ap = c.act + c.row * 16;
        // experiment (off by default): q_l is not spilled; the backward pass rebuilds it from V-bar, delta and sigma, which
        // removes the Q write here and the delta read ("HBM budget")
        uint4 sb[PFD + 1][2];
#pragma unroll
        for (int i = 0; i < PFD; ++i) LD_KEEP(hp, ck0 + i, sb[i]);
        PF_KEEP(hp, ck0 + PFD, NCK - PFD);
        (void)dp;
        epi_wait_acc(c);
#pragma unroll
        for (int i = 0; i < NCK; ++i) {
          const int ck = ck0 + i;
          if (i + PFD < NCK) LD_KEEP(hp, ck + PFD, sb[(i + PFD) % (PFD + 1)]);
          float v[16];
          if (ck * 16 < n_mma) {
            acc_load16(c, ck * 16, v);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] = 0.f;
          }
#pragma unroll
          for (int jp = 0; jp < 8; ++jp) {
            const float2 hh = chunk_pair(sb[i % (PFD + 1)], jp, false);
            v[2 * jp] *= sigma_from_h(hh.x);                                    // vbar_{l+1}
            v[2 * jp + 1] *= sigma_from_h(hh.y);
          }
          uint4 q2[2];
          pack2_grad(v, q2);
          chunk_store(ap, ck, q2);
          chunk_store(vp, ck, q2);
        }
        if (l < 7) epi_signal_act(c);
      }
      // ---- fbar back into ACT (own rows) -----------------------------------------------------------------------
#pragma unroll
      for (int hi = 0; hi < HPW; ++hi) {
        const int hb = hb0 + hi;
        uint4 q[4];
        get_chunk_raw(c, ptrs, ST_FB, tile, hb, q);
        row_half_store(c.act + (hb >> 1) * BLK_BYTES + c.row * 16, hb & 1, q);
      }
      epi_signal_act(c);
      // ---- ordinary backward l = 8..1: zbar_{l-1} = (zbar_l W_l)*sigma_{l-1} + q_{l-1} ------------------------
      // q_{l-1} = 100 * dbar_{l-1} * delta_{l-1} * (1 - sigma_{l-1}) with dbar_{l-1} = vbar_l / sigma_{l-1}: rebuilt from the
      // V-bar tile the adjoint pass wrote (own chunks) and the forward's delta tile.  sigma = 0 only where h = 0 (padding
      // columns, fp16 underflow): vbar and delta are 0 there as well and so is q.  Three operand streams: the loads are
      // pipelined in 8-column pieces (one uint4 per stream and stage) to stay inside the 96-register budget.
#pragma unroll 1
      for (int l = 8; l >= 1; --l) {
        const uint8_t* hp = tile_base(ptrs, ST_H1 + (l - 1), tile, c.row);    // H_l -> sigma_{l-1}
        const uint8_t* vp = tile_base(ptrs, ST_V1 + (l - 1), tile, c.row);    // vbar_l (gradient format)
        const uint8_t* dp = tile_base(ptrs, ST_D0 + (l - 1), tile, c.row);    // delta_{l-1} (fp16)
        uint8_t* zp = tile_base(ptrs, ST_Z0 + (l - 1), tile, c.row);
        uint8_t* ap = c.act + c.row * 16;
        constexpr int PH = RQ_PH;          // prefetch distance in 8-column pieces
        constexpr int NP = 2 * NCK;             // pieces per warpgroup
        const int pc0 = 2 * ck0;                // first piece (= 16-byte chunk column of the 256-wide tile)
        auto piece = [pol_stream](const uint8_t* tp, int pi) {          // last use of all three streams
          return ldg_pol(tp + (pi >> 3) * BLK_BYTES + (pi & 7) * TI_CHUNK_STRIDE, pol_stream);
        };
        uint4 sb[PH + 1], vb[PH + 1], db[PH + 1];
#pragma unroll
        for (int i = 0; i < PH; ++i) {
          sb[i] = piece(hp, pc0 + i);
          vb[i] = piece(vp, pc0 + i);
          db[i] = piece(dp, pc0 + i);
        }
        tile_prefetch_l2(hp, ck0 + (PH + 1) / 2, NCK - (PH + 1) / 2);
        tile_prefetch_l2(vp, ck0 + (PH + 1) / 2, NCK - (PH + 1) / 2);
        tile_prefetch_l2(dp, ck0 + (PH + 1) / 2, NCK - (PH + 1) / 2);
        epi_wait_acc(c);
        float v[16];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int ck = ck0 + (i >> 1);
          if (i + PH < NP) {
            sb[(i + PH) % (PH + 1)] = piece(hp, pc0 + i + PH);
            vb[(i + PH) % (PH + 1)] = piece(vp, pc0 + i + PH);
            db[(i + PH) % (PH + 1)] = piece(dp, pc0 + i + PH);
          }
          if ((i & 1) == 0) {
            acc_load16(c, ck * 16, v);
            if (l == 8) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w8row + ck * 16) + j4);
                v[j4 * 4 + 0] = fmaf(sbar, w4.x, v[j4 * 4 + 0]); v[j4 * 4 + 1] = fmaf(sbar, w4.y, v[j4 * 4 + 1]);
                v[j4 * 4 + 2] = fmaf(sbar, w4.z, v[j4 * 4 + 2]); v[j4 * 4 + 3] = fmaf(sbar, w4.w, v[j4 * 4 + 3]);
              }
            }
            if (l == 4 && ck >= 13 && pc.valid) {        // columns 217..255: PE part of the skip input -> scratch row
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const int col = ck * 16 + jj;
                if (col >= 217) a.eb[row40(tile, c.row, col - 217)] = v[jj];
              }
            }
          }
          const uint4 hq = sb[i % (PH + 1)], vq = vb[i % (PH + 1)], dq = db[i % (PH + 1)];
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            const float2 hh = chunk_pair(&hq, jp, false);
            const float2 vv = chunk_pair(&vq, jp, kGradBf16);
            const float2 dd = chunk_pair(&dq, jp, false);
            const float e0 = ex2_approx(hh.x * (-SP_BETA * 1.4426950408889634f));      // 1 - sigma
            const float e1 = ex2_approx(hh.y * (-SP_BETA * 1.4426950408889634f));
            const float s0 = 1.0f - e0, s1 = 1.0f - e1;
            const float r0 = s0 > 0.f ? __fdividef(SP_BETA * e0, s0) : 0.f;
            const float r1 = s1 > 0.f ? __fdividef(SP_BETA * e1, s1) : 0.f;
            const int j = (i & 1) * 8 + 2 * jp;
            v[j] = fmaf(v[j], s0, vv.x * dd.x * r0);
            v[j + 1] = fmaf(v[j + 1], s1, vv.y * dd.y * r1);
          }
          if (i & 1) {
            uint4 q2[2];
            pack2_grad(v, q2);
            chunk_store(ap, ck, q2);
            ST_STREAM(zp, ck, q2);
          }
        }
        epi_signal_act(c);
      }
      // ---- e-bar += W_0^T zbar_0 ; xbar += J_e^T e-bar -----------------------------------------------------------
      epi_wait_acc(c);
      if (owner) {
        float v[48];
        float eb[40];
        acc_load32(c, 0, v);
        acc_load16(c, 32, v + 32);
        eb[39] = 0.f;
#pragma unroll
        for (int i = 0; i < 39; ++i) eb[i] = v[i] + (pc.valid ? a.eb[row40(tile, c.row, i)] : 0.f);
        float xe[3];
        pe_jt<6>(pc.x, eb, xe);
        if (pc.valid) {
#pragma unroll
          for (int i = 0; i < 3; ++i) a.d_pts[pc.p * 3 + i] = (xbar[i] + xe[i]) * ginv;
        }
      }
      tc_fence_before();
    }
  }
  if (CH_PAIR) cluster_sync_all();      // neither CTA may leave while its partner still signals its barriers / reads its smem
  else __syncthreads();
  if (warp == ISSUER_WARP) {
    tc_fence_after();
    if (CH_PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}


// =====================================================================================================
// one-launch packing of every weight image + the fp32 side arrays from the 28 effective-weight tensors
// =====================================================================================================
struct PackSpec {
  int img;             // image id
  int src;             // 0..8 W_sdf[l], 14..18 -> W_col[l]  (index into PackAllArgs::src)
  int src_rows, src_cols;
  int transpose;       // image row n <-> source column
  int row_off, n_valid;
  int nseg, seg_dst[3], seg_src[3], seg_len[3];
  float scale;
  int bf16;
  int chunk0;          // first 16-byte chunk of this image in the global chunk numbering
};
constexpr int PACK_MAX = 48;
struct PackAllArgs {
  const float* src[28];        // W_sdf[9], b_sdf[9], W_col[5], b_col[5]
  uint8_t* blob;
  float* side;                 // [bias_sdf 8x256 | b8 257(+pad to 264) | w8row 256 | bias_col 4x256 | bc4 4 | wc4 3x256]
  int n_spec;
  int total_chunks;
  int n_bias;                  // bias slices (FINE_BIAS_MMA): image, source tensor (index into src), first source row, valid rows
  int bias_img[13], bias_src[13], bias_row_off[13], bias_nvalid[13];
  PackSpec spec[PACK_MAX];
};
constexpr int SIDE_BIAS_SDF = 0, SIDE_B8 = 2048, SIDE_W8ROW = 2048 + 264, SIDE_BIAS_COL = SIDE_W8ROW + 256,
              SIDE_BC4 = SIDE_BIAS_COL + 1024, SIDE_WC4 = SIDE_BC4 + 4, SIDE_FLOATS = SIDE_WC4 + 768;

__global__ void pack_all_kernel(const __grid_constant__ PackAllArgs a) {
  // ---- fp32 side arrays (block 0 .. few) ---------------------------------------------------------------
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SIDE_FLOATS; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < SIDE_B8) { const int l = i >> 8, c = i & 255; v = (l != 3 || c < 217) ? a.src[9 + l][c] : 0.f; }
    else if (i < SIDE_W8ROW) { const int c = i - SIDE_B8; v = c < 257 ? a.src[17][c] : 0.f; }
    else if (i < SIDE_BIAS_COL) v = a.src[8][i - SIDE_W8ROW];                       // lin8 row 0
    else if (i < SIDE_BC4) { const int j = i - SIDE_BIAS_COL; v = a.src[23 + (j >> 8)][j & 255]; }
    else if (i < SIDE_WC4) { const int c = i - SIDE_BC4; v = c < 3 ? a.src[27][c] : 0.f; }
    else v = a.src[22][i - SIDE_WC4];                                               // colour lin4 [3,256]
    a.side[i] = v;
  }
  // ---- operand images ----------------------------------------------------------------------------------
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < a.total_chunks; g += gridDim.x * blockDim.x) {
    int si = 0;
    while (si + 1 < a.n_spec && g >= a.spec[si + 1].chunk0) ++si;
    const PackSpec& sp = a.spec[si];
    const ImgInfo ii = img_info(sp.img);
    const int i = g - sp.chunk0;
    const int kb = i / (ii.npad * 8);
    const int rem = i - kb * ii.npad * 8;
    const int ch = rem / ii.npad, n = rem - ch * ii.npad;       // [k-block][chunk column][row]
    const float* src = a.src[sp.src];
    const long long sn = sp.transpose ? 1 : sp.src_cols, sk = sp.transpose ? sp.src_cols : 1;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb * 64 + ch * 8 + j;
      float x = 0.f;
      if (n < sp.n_valid) {
        for (int q = 0; q < sp.nseg; ++q)
          if (k >= sp.seg_dst[q] && k < sp.seg_dst[q] + sp.seg_len[q])
            x = sp.scale * src[(long long)(n + sp.row_off) * sn + (long long)(sp.seg_src[q] + k - sp.seg_dst[q]) * sk];
      }
      v[j] = x;
    }
    uint4 q4;
    if (sp.bf16) { q4.x = pack_bf2(v[0], v[1]); q4.y = pack_bf2(v[2], v[3]); q4.z = pack_bf2(v[4], v[5]); q4.w = pack_bf2(v[6], v[7]); }
    else { q4.x = pack_h2(v[0], v[1]); q4.y = pack_h2(v[2], v[3]); q4.z = pack_h2(v[4], v[5]); q4.w = pack_h2(v[6], v[7]); }
    size_t in_kb = (size_t)ch * ii.npad * 16 + (size_t)n * 16;
    if (img_half_major(sp.img)) {
      const int hn = ii.npad / 2, h = n / hn;
      in_kb = (size_t)h * hn * 128 + (size_t)ch * hn * 16 + (size_t)(n - h * hn) * 16;
    }
    *reinterpret_cast<uint4*>(a.blob + img_offset_dev(sp.img) + (size_t)kb * ii.npad * 128 + in_kb) = q4;
  }
  // ---- bias slices: [half][2 chunk columns][N/2 rows][16 B]; K column 0 = fp16(b), 1 = fp16(b - fp16(b)), rest 0 -------
  for (int bi = 0; bi < a.n_bias; ++bi) {
    const ImgInfo ii = img_info(a.bias_img[bi]);
    const int hn = ii.npad / 2;
    uint8_t* dst = a.blob + img_offset_dev(a.bias_img[bi]) + (size_t)ii.kblocks * ii.npad * 128;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * ii.npad; i += gridDim.x * blockDim.x) {
      const int ch = i / ii.npad, n = i - ch * ii.npad, h = n / hn;
      uint4 q4 = make_uint4(0u, 0u, 0u, 0u);
      if (ch == 0 && n < a.bias_nvalid[bi]) {
        const float b = a.src[a.bias_src[bi]][n + a.bias_row_off[bi]];
        const float hi = __half2float(__float2half_rn(b));
        q4.x = pack_h2(hi, (b - hi) * BIAS_LO_SCALE);          // the residual scaled into fp16's normal range
      }
      *reinterpret_cast<uint4*>(dst + (size_t)h * hn * 32 + (size_t)ch * hn * 16 + (size_t)(n - h * hn) * 16) = q4;
    }
  }
}
}  // namespace fmov
using namespace fmov;

// ---- tables ---------------------------------------------------------------------------------------------
static void set_step(ChainStep& st, int img, int nkb_a, int nkb_aux, uint32_t a_fmt, uint32_t b_fmt) {
  const ImgInfo ii = img_info(img);
  st.w_off = (uint32_t)img_offset(img);
  st.n = (uint16_t)ii.npad;
  st.nkb_a = (uint8_t)nkb_a;
  st.nkb_aux = (uint8_t)nkb_aux;
  st.a_fmt = (uint8_t)a_fmt;
  st.b_fmt = (uint8_t)b_fmt;
}
static void init_table(ChainTable& tb) {
  memset(&tb, 0, sizeof(tb));
  for (int i = 0; i < MAX_STEPS; ++i) tb.step[i].pf[0] = tb.step[i].pf[1] = 0xFF;
}
static void set_pf(ChainStep& st, int t0, int t1 = 0xFF) { st.pf[0] = (uint8_t)t0; st.pf[1] = (uint8_t)t1; }

static void build_fwd_table(ChainTable& tb) {
  init_table(tb);
  int n = 0;
  for (int l = 0; l < 8; ++l) set_step(tb.step[n++], IMG_FP0 + l, l == 0 ? 0 : 4, (l == 0 || l == 4) ? 1 : 0, FMT_F16, FMT_F16);
  set_step(tb.step[n++], IMG_F0 + 8, 4, 0, FMT_F16, FMT_F16);                 // lin8 feature rows
  for (int i = 0; i < 9; ++i) tb.step[i].bias16 = FINE_BIAS_MMA ? 1 : 0;       // biases of lin0..lin8 by the tensor core
  for (int l = 7; l >= 1; --l) {                                                // reverse sweep: reads H_l
    set_pf(tb.step[n], ST_H1 + (l - 1));
    set_step(tb.step[n++], IMG_T0 + l, 4, 0, FMT_F16, FMT_F16);
  }
  set_pf(tb.step[n], ST_F);                                                     // F goes back into ACT here
  set_step(tb.step[n++], IMG_T0 + 0, 4, 0, FMT_F16, FMT_F16);                  // W_0^T delta_0 (N = 48)
  for (int l = 0; l < 4; ++l) tb.step[n + l].bias16 = FINE_BIAS_MMA ? 1 : 0;   // colour lin0..lin3
  set_step(tb.step[n++], IMG_C0 + 0, 4, 1, FMT_F16, FMT_F16);                  // colour lin0: feat + extras
  for (int l = 1; l < 4; ++l) set_step(tb.step[n++], IMG_C0 + l, 4, 0, FMT_F16, FMT_F16);
  set_step(tb.step[n++], IMG_C0 + 4, 4, 0, FMT_F16, FMT_F16);                  // colour lin4 (N = 16)
  tb.n_steps = n;
}

static void build_bwd_table(ChainTable& tb) {
  init_table(tb);
  int n = 0;
  set_pf(tb.step[n], ST_C1 + 3);
  tb.step[n++].no_mma = 1;          // colour lin4 backward: epilogue only (CUDA cores, 3 x 256)
  for (int l = 3; l >= 1; --l) {
    set_pf(tb.step[n], ST_C1 + (l - 1));
    set_step(tb.step[n++], l == 3 ? IMG_CT3 : l == 2 ? IMG_CT2 : IMG_CT1, 4, 0, kGradFmt, kGradFmt);
  }
  set_step(tb.step[n++], IMG_CT0B, 4, 0, kGradFmt, kGradFmt);
  set_step(tb.step[n++], IMG_CT0A, 4, 0, kGradFmt, kGradFmt);
  for (int l = 0; l < 8; ++l) {                                                 // adjoint pass: reads H_{l+1}, delta_l
    set_pf(tb.step[n], ST_H1 + l, ST_D0 + l);
    set_step(tb.step[n++], (kGradBf16 ? IMG_FB0 : IMG_FP0) + l, l == 0 ? 0 : 4, (l == 0 || l == 4) ? 1 : 0, kGradFmt, kGradFmt);
  }
  for (int l = 8; l >= 1; --l) {                                                // ordinary backward: reads H_l, q_{l-1}
    set_pf(tb.step[n], ST_H1 + (l - 1), ST_Q0 + (l - 1));
    set_step(tb.step[n++], (kGradBf16 || l == 8 ? IMG_TB0 : IMG_T0) + l, 4, 0, kGradFmt, kGradFmt);   // (T8 only exists in the TB range)
  }
  set_step(tb.step[n++], (kGradBf16 ? IMG_TB0 : IMG_T0) + 0, 4, 0, kGradFmt, kGradFmt);
  tb.n_steps = n;
}


// ---- fused packing -----------------------------------------------------------------------------------------
static void add_spec(PackAllArgs& a, int img, int src, int rows, int cols, bool tr, int row_off, int n_valid, int nseg,
                     int d0, int s0, int l0, int d1, int s1, int l1, float scale, bool bf16, int d2 = 0, int s2 = 0,
                     int l2 = 0) {
  PackSpec& sp = a.spec[a.n_spec++];
  sp.img = img; sp.src = src; sp.src_rows = rows; sp.src_cols = cols; sp.transpose = tr ? 1 : 0; sp.row_off = row_off;
  sp.n_valid = n_valid; sp.nseg = nseg; sp.seg_dst[0] = d0; sp.seg_src[0] = s0; sp.seg_len[0] = l0;
  sp.seg_dst[1] = d1; sp.seg_src[1] = s1; sp.seg_len[1] = l1; sp.seg_dst[2] = d2; sp.seg_src[2] = s2; sp.seg_len[2] = l2;
  sp.scale = scale; sp.bf16 = bf16 ? 1 : 0;
  const ImgInfo ii = img_info(img);
  sp.chunk0 = a.total_chunks;
  a.total_chunks += ii.npad * ii.kblocks * 8;
}
static void build_pack_specs(PackAllArgs& a, bool backward) {
  a.n_spec = 0; a.total_chunks = 0; a.n_bias = 0;
  auto add_bias = [&](int img, int src, int row_off, int n_valid) {
    if (!img_has_bias(img)) return;
    const int i = a.n_bias++;
    a.bias_img[i] = img; a.bias_src[i] = src; a.bias_row_off[i] = row_off; a.bias_nvalid[i] = n_valid;
  };
  static const int so_b[8] = {256, 256, 256, 217, 256, 256, 256, 256};
  for (int l = 0; l < 8; ++l) add_bias(IMG_FP0 + l, 9 + l, 0, so_b[l]);
  add_bias(IMG_F0 + 8, 17, 1, 256);                      // lin8 feature rows 1..256
  for (int l = 0; l < 4; ++l) add_bias(IMG_C0 + l, 23 + l, 0, 256);
  const float rs2 = 0.70710678118654752f;
  static const int so[9] = {256, 256, 256, 217, 256, 256, 256, 256, 257}, si[9] = {39, 256, 256, 256, 256, 256, 256, 256, 256};
  auto fwd_img = [&](int img, int l, bool bf) {
    // the encoded-input k-block carries the fp16 residuals of the raw coordinates in columns 39..41 (pe_with_residual): the
    // images repeat the weight columns of x, y, z there, so x enters the first layer and the skip layer to ~2^-22
    if (l == 0) add_spec(a, img, 0, so[0], si[0], false, 0, 256, 2, 0, 0, 39, PE_RES_COL, 0, 3, 1.f, bf);
    else if (l == 4) add_spec(a, img, 4, 256, 256, false, 0, 256, 3, 0, 0, 217, 256, 217, 39, rs2, bf, 256 + PE_RES_COL, 217, 3);
    else if (l == 8) add_spec(a, img, 8, 257, 256, false, 1, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, bf);
    else add_spec(a, img, l, so[l], si[l], false, 0, so[l], 1, 0, 0, 256, 0, 0, 0, 1.f, bf);
  };
  auto tr_img = [&](int img, int l, bool bf) {      // image rows = input index, K = output index
    if (l == 0) add_spec(a, img, 0, 256, 39, true, 0, 39, 1, 0, 0, 256, 0, 0, 0, 1.f, bf);
    else if (l == 3) add_spec(a, img, 3, 217, 256, true, 0, 256, 1, 0, 0, 217, 0, 0, 0, 1.f, bf);
    else if (l == 4) add_spec(a, img, 4, 256, 256, true, 0, 256, 1, 0, 0, 256, 0, 0, 0, rs2, bf);
    else if (l == 8) add_spec(a, img, 8, 257, 256, true, 0, 256, 1, 0, 1, 256, 0, 0, 0, 1.f, bf);
    else add_spec(a, img, l, 256, 256, true, 0, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, bf);
  };
  for (int l = 0; l < 9; ++l) fwd_img(IMG_F0 + l, l, false);
  for (int l = 0; l < 8; ++l) fwd_img(IMG_FP0 + l, l, false);
  for (int l = 0; l < 8; ++l) tr_img(IMG_T0 + l, l, false);
  // colour net: kernel K order is [feat(256) | extras(33)], reference order is [extras(33) | feat(256)]
  add_spec(a, IMG_C0 + 0, 18, 256, 289, false, 0, 256, 2, 0, 33, 256, 256, 0, 33, 1.f, false);
  for (int l = 1; l <= 3; ++l) add_spec(a, IMG_C0 + l, 18 + l, 256, 256, false, 0, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, false);
  add_spec(a, IMG_C0 + 4, 22, 3, 256, false, 0, 3, 1, 0, 0, 256, 0, 0, 0, 1.f, false);
  if (backward) {
    const bool g = kGradBf16;
    add_spec(a, IMG_CT0A, 18, 256, 289, true, 33, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, g);
    add_spec(a, IMG_CT0B, 18, 256, 289, true, 0, 33, 1, 0, 0, 256, 0, 0, 0, 1.f, g);
    add_spec(a, IMG_CT1, 19, 256, 256, true, 0, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, g);
    add_spec(a, IMG_CT2, 20, 256, 256, true, 0, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, g);
    add_spec(a, IMG_CT3, 21, 256, 256, true, 0, 256, 1, 0, 0, 256, 0, 0, 0, 1.f, g);
    if (g) {
      for (int l = 0; l < 8; ++l) fwd_img(IMG_FB0 + l, l, true);
      for (int l = 0; l < 9; ++l) tr_img(IMG_TB0 + l, l, true);
    } else {
      tr_img(IMG_TB0 + 8, 8, false);
    }
  }
}

extern "C" int fmov_side_floats(void) { return SIDE_FLOATS; }
extern "C" int fmov_side_offset(int which) {
  const int offs[6] = {SIDE_BIAS_SDF, SIDE_B8, SIDE_W8ROW, SIDE_BIAS_COL, SIDE_BC4, SIDE_WC4};
  return (which >= 0 && which < 6) ? offs[which] : -1;
}
/* srcs: HOST array of 28 device pointers: W_sdf[0..8], b_sdf[0..8], W_col[0..4], b_col[0..4] (contiguous fp32,
 * reference shapes).  One launch writes every operand image into `blob` (fmov_fine_blob_bytes()) and the fp32 side
 * arrays into `side` (fmov_side_floats() floats; offsets from fmov_side_offset()). */
extern "C" int fmov_pack_all(const float* const* srcs, void* blob, float* side, int need_backward, void* stream) {
  FMOV_REQUIRE(srcs && blob && side, "fmov_pack_all: null argument");
  static PackAllArgs tmpl[2];
  static bool init = false;
  if (!init) {
    build_pack_specs(tmpl[0], false);
    build_pack_specs(tmpl[1], true);
    long long offs[IMG_COUNT + 1];
    for (int i = 0; i <= IMG_COUNT; ++i) offs[i] = img_offset(i);
    FMOV_CUDA(cudaMemcpyToSymbol(c_img_offset, offs, sizeof(offs)));
    init = true;
  }
  PackAllArgs a = tmpl[need_backward ? 1 : 0];
  for (int i = 0; i < 28; ++i) {
    FMOV_REQUIRE(srcs[i], "fmov_pack_all: source tensor %d is null", i);
    a.src[i] = srcs[i];
  }
  a.blob = reinterpret_cast<uint8_t*>(blob);
  a.side = side;
  pack_all_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("pack_all_kernel");
  return OK;
}

#ifdef FMOV_TRACE
/* debug builds only: copies the CTA-0 timeline of the last fine kernel to the host ([event][tag, clock], warp regions
 * concatenated) and resets it */
extern "C" int fmov_debug_trace(long long* host, int max_events) {
  unsigned int cnt[TR_WARPS];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(cnt, g_trace_cnt, sizeof(cnt));
  int n = 0;
  for (int w = 0; w < TR_WARPS; ++w) {
    int c = (int)cnt[w];
    if (c > TR_CAP) c = TR_CAP;
    if (n + c > max_events) c = max_events - n;
    if (c > 0) cudaMemcpyFromSymbol(host + 2 * (size_t)n, g_trace, (size_t)c * 2 * sizeof(long long), (size_t)w * TR_CAP * 2 * sizeof(long long));
    n += c > 0 ? c : 0;
  }
  memset(cnt, 0, sizeof(cnt));
  cudaMemcpyToSymbol(g_trace_cnt, cnt, sizeof(cnt));
  return n;
}
#endif

// ---- C ABI -----------------------------------------------------------------------------------------------
extern "C" int fmov_fine_image_count(void) { return IMG_COUNT; }
extern "C" int fmov_fine_image_info(int id, long long* offset, int* npad, int* kblocks) {
  FMOV_REQUIRE(id >= 0 && id < IMG_COUNT && offset && npad && kblocks, "fmov_fine_image_info: bad id %d", id);
  const ImgInfo ii = img_info(id);
  *offset = img_offset(id);
  *npad = ii.npad;
  *kblocks = ii.kblocks;
  return OK;
}
extern "C" long long fmov_fine_blob_bytes(void) { return img_offset(IMG_COUNT); }
/* byte offset, inside the fine blob, of the images the pair-engine queries read (FP0..FP7, fmov_sdf_pair_blob_bytes() bytes) */
extern "C" long long fmov_sdf_pair_blob_offset(void) { return img_offset(IMG_FP0); }
extern "C" int fmov_grad_is_bf16(void) { return kGradBf16 ? 1 : 0; }
extern "C" int fmov_fine_stash_count(void) { return ST_COUNT; }
extern "C" int fmov_fine_stash_blocks(int id) { return (id >= 0 && id < ST_COUNT) ? stash_kb(id) : -1; }
extern "C" int fmov_fine_stash_is_forward(int id) { return (id >= 0 && id < ST_COUNT) ? (stash_is_forward(id) ? 1 : 0) : -1; }

static int fill_args(FineArgs& a, ChainPtrs& ptrs, long long B, int S, const float* rays_o, const float* rays_d,
                     const float* z, float sample_dist, const void* wblob, void* const* stash, const float* bias_sdf,
                     const float* b8, const float* w8row, const float* bias_col, const float* bc4, const float* wc4) {
  FMOV_REQUIRE(B > 0 && S > 0, "fine: bad sizes B=%lld S=%d", B, S);
  FMOV_REQUIRE(rays_o && rays_d && z && wblob && stash && bias_sdf && b8 && w8row && bias_col && bc4, "fine: null argument");
  memset(&a, 0, sizeof(a));
  memset(&ptrs, 0, sizeof(ptrs));
  a.B = B; a.S = S; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.sample_dist = sample_dist;
  a.bias_sdf = bias_sdf; a.b8 = b8; a.w8row = w8row; a.bias_col = bias_col; a.bc4 = bc4; a.wc4 = wc4;
  ptrs.weights = reinterpret_cast<const uint8_t*>(wblob);
  for (int i = 0; i < ST_COUNT; ++i) ptrs.stash[i] = reinterpret_cast<uint8_t*>(stash[i]);
  return OK;
}
static int grid_for(long long P, int max_ctas) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (P + TILE_M - 1) / TILE_M;
  int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  return grid;
}
// One persistent CTA per SM; in CTA-pair mode clusters of two CTAs (as many pairs as can be resident at once: a pair that
// had to wait for another to finish would double the kernel time, the tile assignment is static).
static int pair_maps_for(const void* blob, PairMaps& out) {
  memset(&out, 0, sizeof(out));
  if (!CH_PAIR) return OK;
  return chain_pair_maps(blob, img_offset(IMG_COUNT), out);
}
template <typename Kernel>
static int launch_fine(Kernel kernel, long long P, const ChainTable& tb, const ChainPtrs& ptrs, const FineArgs& a,
                       cudaStream_t stream, const char* name) {
  int grid = grid_for(P, 0);
  PairMaps maps;
  int st = pair_maps_for(ptrs.weights, maps);
  if (st) return st;
  if (!CH_PAIR) {
    kernel<<<grid, CH_THREADS, FL::DYN_BYTES, stream>>>(tb, ptrs, a, maps);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(CH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = FL::DYN_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int max_pairs[2] = {0, 0};          // per kernel (fwd / bwd)
    int& mp = max_pairs[name[5] == 'f' ? 0 : 1];
    if (mp == 0) {
      cfg.gridDim = dim3(2 * 74, 1, 1);
      int n = 0;
      FMOV_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
      FMOV_REQUIRE(n > 0, "%s: no CTA pair fits on this device", name);
      mp = n;
    }
    grid = (grid + 1) & ~1;                    // pairs; the odd CTA of the last pair may only have a dummy tile
    if (grid > 2 * mp) grid = 2 * mp;
    cfg.gridDim = dim3(grid, 1, 1);
    FMOV_CUDA(cudaLaunchKernelEx(&cfg, kernel, tb, ptrs, a, maps));
  }
  FMOV_LAUNCH_CHECK(name);
  return OK;
}

/* stash: HOST array of ST_COUNT device pointers (tile-image tensors, fmov_fine_stash_blocks(id) blocks per tile) */
extern "C" int fmov_fine_fwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                             float sample_dist, const void* wblob, void* const* stash, const float* bias_sdf,
                             const float* b8, const float* w8row, const float* bias_col, const float* bc4, float* sdf,
                             float* nrm, float* rgb, float* ge, void* stream) {
  static ChainTable tb;
  static bool init = false;
  if (!init) {
    build_fwd_table(tb);
    FMOV_CUDA(cudaFuncSetAttribute(fine_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL::DYN_BYTES));
    init = true;
  }
  FineArgs a;
  ChainPtrs ptrs;
  int st = fill_args(a, ptrs, B, S, rays_o, rays_d, z, sample_dist, wblob, stash, bias_sdf, b8, w8row, bias_col, bc4, nullptr);
  if (st) return st;
  FMOV_REQUIRE(sdf && nrm && rgb && ge, "fmov_fine_fwd: null output");
  for (int i = 0; i < ST_COUNT; ++i)
    if (stash_is_forward(i)) FMOV_REQUIRE(stash[i], "fmov_fine_fwd: stash tensor %d is null", i);
  a.sdf = sdf; a.nrm = nrm; a.rgb = rgb; a.ge = ge;
  return launch_fine(fine_fwd_kernel, B * S, tb, ptrs, a, (cudaStream_t)stream, "fine_fwd_kernel");
}

extern "C" int fmov_fine_bwd(long long B, int S, const float* rays_o, const float* rays_d, const float* z,
                             float sample_dist, const void* wblob, void* const* stash, const float* bias_sdf,
                             const float* b8, const float* w8row, const float* bias_col, const float* bc4,
                             const float* wc4, const float* rgb, const float* ge, const float* d_sdf, const float* d_nrm,
                             const float* d_rgb, const float* amax, float* d_pts, float* d_dirs, float* zc4, float* eb_scratch,
                             void* stream) {
  static ChainTable tb;
  static bool init = false;
  if (!init) {
    build_bwd_table(tb);
    FMOV_CUDA(cudaFuncSetAttribute(fine_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL::DYN_BYTES));
    init = true;
  }
  FineArgs a;
  ChainPtrs ptrs;
  int st = fill_args(a, ptrs, B, S, rays_o, rays_d, z, sample_dist, wblob, stash, bias_sdf, b8, w8row, bias_col, bc4, wc4);
  if (st) return st;
  FMOV_REQUIRE(wc4 && rgb && ge && d_sdf && d_nrm && d_rgb && amax && d_pts && d_dirs && zc4 && eb_scratch, "fmov_fine_bwd: null argument");
  for (int i = 0; i < ST_COUNT; ++i)
    if (stash_kb(i) > 0) FMOV_REQUIRE(stash[i], "fmov_fine_bwd: stash tensor %d is null", i);      // 0 blocks: never touched
  a.rgb = const_cast<float*>(rgb); a.ge = const_cast<float*>(ge);
  a.d_sdf = d_sdf; a.d_nrm = d_nrm; a.d_rgb = d_rgb; a.d_pts = d_pts; a.d_dirs = d_dirs; a.zc4 = zc4; a.eb = eb_scratch; a.amax = amax;
  return launch_fine(fine_bwd_kernel, B * S, tb, ptrs, a, (cudaStream_t)stream, "fine_bwd_kernel");
}
