// Fused MLP "chain" engine: a persistent CTA walks 128-point tiles through a table of GEMM steps with the activation
// tile resident in shared memory, weights streamed from L2 into a ring, accumulators in TMEM and a per-kernel epilogue
// between steps.  TWO tiles are in flight per CTA so that the tensor core works on one tile while the other tile's
// epilogue runs.  Two forms:
//   one-CTA engine   tcgen05.mma.cta_group::1 (M = 128); weight k-slices by cp.async.bulk into a 4 x 16 KiB ring
//   CTA-pair engine  clusters of two CTAs; ONE tcgen05.mma.cta_group::2 stream (M = 256) issued by the even CTA for both
//                    CTAs' tiles; each CTA stages half of every weight k-block by cta_group::2 tensor-map TMA; commits
//                    multicast to both CTAs; optional bias16 steps add the layer bias by one extra K = 16 MMA
//
// Warp roles (CH_WGS = 2: 20 warps, 640 threads):
//   warps 0..7    epilogue of tile slot 0: two warpgroups, columns 0..127 and 128..255 of the tile
//   warps 8..15   epilogue of tile slot 1        (tcgen05.ld -> fp32 math -> st.shared next A operand
//                                                 + st.global / ld.global of the stash rows)
//   warp 16       weight producer   bulk / tensor-map copies of weight k-blocks -> WST ring          (1 lane)
//   warp 17       MMA issuer        tcgen05.mma    A = ACT[slot]/AUX[slot], B = WST stage            (1 lane) + TMEM alloc
//                 (pair engine: only the even CTA's issuer issues; the odd CTA's is idle)
//   warps 18,19   idle
// What bounds the kernels built on it was measured in round 2b (knock-out builds): with the pair engine the
// weight stream and the tensor core are off the critical path; the epilogue warps' instruction and memory work is it.
//
// Stash traffic is done by the row-owning thread: in the tile-image layout a tile row is 128 contiguous
// bytes per 64-feature block, so a warp reads/writes 4 KiB contiguous per block (fully coalesced), and a
// thread only ever re-reads rows it wrote itself (program-order coherence, no fences).  Cross-thread
// visibility is only needed for operands of the MMA (shared memory: fence.proxy.async + mbarrier).
//
// mbarriers (per tile slot): act_ready (A operand written + TMEM drained: one arrival per epilogue thread, or one per
// epilogue warp of both CTAs on the even CTA's barrier in pair mode) -> MMA issuer; acc_ready (tcgen05.commit) ->
// epilogue; weight ring full/empty.
#pragma once
#include "fmov_common.cuh"
#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace fmov {

constexpr int CH_SLOTS = 2;
#ifndef FMOV_CH_WGS
#define FMOV_CH_WGS 1
#endif
constexpr int CH_WGS = FMOV_CH_WGS;                // epilogue warpgroups (column ranges) per tile slot: 1 or 2
constexpr int CH_CHUNKS = 16 / CH_WGS;             // 16-column chunks handled by one warpgroup
// Warp layout: epilogue warps first (whole warpgroups), then one control warpgroup whose warps 0/1 are the weight producer /
// MMA issuer (warps 2/3 idle).  Registers are allocated per SM sub-partition (16 K each, warp w lives in partition w % 4):
// with 5 warps per partition the launch allocation is <= 96 registers/thread (setmaxnreg cannot re-balance it with
// ptxas 12.9: a .dec anywhere caps the whole kernel).
constexpr int EPI_WARP0 = 0;
constexpr int CTRL_WARP0 = CH_SLOTS * CH_WGS * 4;
constexpr int PRODUCER_WARP = CTRL_WARP0;
constexpr int ISSUER_WARP = CTRL_WARP0 + 1;
constexpr int CH_THREADS = (CTRL_WARP0 + 4) * 32;           // 640 (two warpgroups per slot) / 384
constexpr int EPI_THREADS = CH_WGS * 128;          // threads arriving per slot
// Weight ring: 64 KiB of shared memory in CH_WSTAGES slots of 1/CH_WSPLIT k-block each (a K = 64/CH_WSPLIT slice of
// all N rows is contiguous in the weight image).  One-CTA engine: the ring is latency bound — 48 KiB in flight behind the
// stage being read, per ~2-2.5 K-cycle L2 round trip under load = 23-29 B/clk per SM, although the SM itself can take
// 60-90 B/clk (profiles/micro/l2_stream*.cu); 2 x 32 KiB was worse (one copy in flight), 8 x 8 KiB too (per-request cost).
#ifndef FMOV_CH_WSPLIT
#define FMOV_CH_WSPLIT 2
#endif
constexpr int CH_WSPLIT = FMOV_CH_WSPLIT;          // ring slots per 64-wide k-block: 1, 2 or 4
constexpr int WSLOT_BYTES = 256 * 128 / CH_WSPLIT; // [256 rows x 64/CH_WSPLIT] 16-bit
constexpr int CH_WSTAGES = 2 * CH_WSPLIT;
// FMOV_CH_PAIR = 1 (set by the including kernel file): CTA-PAIR mode.  The kernel is launched in clusters of two CTAs
// (one per SM of a TPC); every GEMM step of a tile slot is ONE tcgen05.mma.cta_group::2 stream issued by the even CTA
// for both CTAs' 128-point tiles (M = 256), and each CTA stages only HALF of every weight slice (N/2 rows of B) in its
// ring.  Why: the weight ring is latency bound (64 KiB in flight per ~2.5 K-cycle L2 round trip = 23-29 B/clk per SM,
// profiles/micro/l2_stream*.cu, r2s2 traces), so a step's 128 KiB of weights took ~5.5 K cycles against 2 K of tensor
// time and the single issuing thread was busy 11 K of every 16 K cycles; with half the bytes per SM the same ring feeds a
// step in ~2.8 K.  Weight images used in this mode are stored half-major ([half][chunk column][N/2 rows][16 B] per
// 64-wide k-block) so that a CTA's half of a k-block is one contiguous bulk copy.
#ifndef FMOV_FINE_PAIR
#define FMOV_FINE_PAIR 1          // the fine-stage kernels and the train step's sampling queries use the pair engine
#endif
#ifndef FMOV_CH_PAIR
#define FMOV_CH_PAIR 0
#endif
constexpr bool CH_PAIR = FMOV_CH_PAIR != 0;
constexpr int CH_PW_STAGES = 4;                     // pair mode: ring stages of one half k-block (N/2 rows x 64 K) each
constexpr int CH_PW_STAGE_BYTES = 128 * 128;        // 16 KiB (N = 256)
constexpr uint32_t CH_PEER_MASK = 0xFEFFFFFFu;      // shared::cluster address of the pair's EVEN CTA (cute::Sm100MmaPeerBitMask)
constexpr int MAX_STEPS = 40;
constexpr int MAX_STASH = 56;

// ---- optional timeline trace (-DFMOV_TRACE): the tracing lanes of CTA 0 record (tag, clock64) pairs into per-warp regions
// of a device buffer (counters in shared memory, plain stores: ~50 cycles per event — the first tracer took its slot with a
// global atomic, ~850 cycles per event, which inflated every interval it measured); read back with fmov_debug_trace ----
#ifdef FMOV_TRACE
constexpr int TR_WARPS = 24, TR_CAP = 16384;
__device__ long long g_trace[TR_WARPS][TR_CAP][2];
__device__ unsigned int g_trace_cnt[TR_WARPS];
#define FMOV_TR(S, kind, who, idx)                                                                     \
  do {                                                                                                 \
    if (blockIdx.x == 0) {                                                                             \
      const unsigned int w__ = threadIdx.x >> 5, n__ = (S)->tr_cnt[w__]++;                             \
      if (n__ < (unsigned int)TR_CAP) {                                                                \
        g_trace[w__][n__][0] = ((long long)(kind) << 32) | ((long long)(who) << 16) | (long long)((idx) & 0xFFFF); \
        g_trace[w__][n__][1] = clock64();                                                              \
        g_trace_cnt[w__] = n__ + 1;                                                                    \
      }                                                                                                \
    }                                                                                                  \
  } while (0)
#else
#define FMOV_TR(S, kind, who, idx)
#endif

struct ChainStep {
  uint32_t w_off;       // byte offset of the first k-block of this step in the weight blob
  uint16_t n;           // MMA N (multiple of 16, <= 256); k-block bytes = n*128
  uint8_t nkb_a;        // k-blocks read from ACT ...
  uint8_t nkb_aux;      // ... followed by k-blocks read from AUX
  uint8_t a_fmt;        // FMT_F16 / FMT_BF16 (A and B must match: tcgen05 kind::f16 cannot mix them)
  uint8_t b_fmt;
  uint8_t no_mma;       // epilogue-only step (no weights / MMA / accumulator)
  uint8_t pf[2];        // stash tensors (4-block tiles) this step's epilogue reads from HBM; 0xFF = none (documentation of the
                        // tile program: the epilogue threads prefetch their own rows, see tile_prefetch_l2)
  uint8_t flags;        // CHF_* bits (0 for an ordinary step)
  uint8_t bias16;       // pair mode: one more K = 16 MMA after the k-blocks, A = columns 48..63 of the slot's AUX block (1.0 in
                        // columns 48 / 49), B = the [N x 16] slice appended to the step's weight image (bias as fp16 hi + lo):
                        // the layer's bias is added by the tensor core instead of 4 broadcast loads + 16 adds per chunk
  uint8_t pad[1];
};
// Step flags: let several GEMM passes share one accumulator (split-precision chains: hi*W_hi + lo*W_hi + hi*W_lo).
constexpr uint8_t CHF_A_OTHER = 1;     // A operand = the OTHER tile slot's ACT / AUX region (holds the fp16 residuals)
constexpr uint8_t CHF_ACCUM = 2;       // continue on the previous step's accumulator: no operand wait, never zero-initialise
constexpr uint8_t CHF_NO_COMMIT = 4;   // more passes follow: do not signal the epilogue after this step
constexpr uint8_t CHF_W2 = 8;          // weights come from ChainPtrs::weights2 (residual images)
constexpr uint8_t CHF_DUAL_A = 16;     // every weight slice multiplies BOTH slots' A operands (hi, then the residuals) before
                                       // it is released: two passes for one weight stream
struct ChainTable {
  int n_steps;
  int slots;            // tiles in flight per CTA: 0 -> CH_SLOTS; 1 for chains that use slot 1's buffers for residuals
  ChainStep step[MAX_STEPS];
};
struct ChainPtrs {
  const uint8_t* weights;        // packed weight images
  const uint8_t* weights2;       // second blob (CHF_W2 steps), may be null
  uint8_t* stash[MAX_STASH];     // tile-image tensors in HBM
};

struct ChainSmem {
  uint64_t act_ready[CH_SLOTS];
  uint64_t acc_ready[CH_SLOTS];
  uint64_t w_full[CH_WSTAGES];
  uint64_t w_empty[CH_WSTAGES];
  uint32_t tmem_base;
  uint32_t pad_;
  uint32_t tr_cnt[24];            // FMOV_TRACE: events recorded per warp
  float scratch[CH_SLOTS][128];   // per-row partial sums exchanged between the two column warpgroups of a slot
};

// Dynamic smem: [ChainSmem | pad to 1024][ACT0 64K][ACT1 64K][AUX0 16K][AUX1 16K][WST 2 x 32K]
struct ChainLayout {
  static constexpr int HDR = 2048;
  static constexpr int ACT = HDR;
  static constexpr int AUX = ACT + CH_SLOTS * 4 * BLK_BYTES;
  static constexpr int WST = AUX + CH_SLOTS * BLK_BYTES;
  static constexpr int TOTAL = WST + CH_WSTAGES * WSLOT_BYTES;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024-alignment (dynamic smem base is >= 16 B aligned)
};
static_assert(ChainLayout::DYN_BYTES <= 227 * 1024, "chain engine exceeds the 227 KiB shared-memory limit");

__device__ __forceinline__ uint8_t* chain_smem_base(uint8_t* raw) {
  uintptr_t p = reinterpret_cast<uintptr_t>(raw);
  p = (p + 1023) & ~uintptr_t(1023);
  return reinterpret_cast<uint8_t*>(p);
}

__device__ __forceinline__ void chain_init_barriers(ChainSmem* s, int epi_threads = EPI_THREADS, bool pair = CH_PAIR) {
  for (int i = 0; i < CH_SLOTS; ++i) {
    // pair mode: one elected arrival per epilogue warp of BOTH CTAs, on the even CTA's barrier
    mbar_init(&s->act_ready[i], pair ? 2 * (epi_threads / 32) : epi_threads);
    mbar_init(&s->acc_ready[i], 1);
  }
  for (int i = 0; i < 24; ++i) s->tr_cnt[i] = 0;
  for (int i = 0; i < CH_WSTAGES; ++i) {
    mbar_init(&s->w_full[i], 1);
    mbar_init(&s->w_empty[i], 1);
  }
  fence_mbar_init();
}

// ---- cluster / CTA-pair primitives ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {          // every thread of both CTAs (warps may arrive diverged)
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in the pair's EVEN CTA.  Default semantics (release at CTA scope),
// as cutlass::arch::umma_arrive_2x1SM_sm0 does: what the even CTA's issuer must see before it issues the pair MMA are this
// CTA's shared-memory operand writes, which every writing thread has already pushed to the async proxy (fence.proxy.async)
// before its warp elects the arriving lane.  (`.release.cluster` here compiled to a MEMBAR that also waited for the thread's
// outstanding global stash stores: 9 % of fine_fwd's and 5 % of fine_bwd's warp samples were `membar` stalls on it.)
__device__ __forceinline__ void mbar_arrive_even_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & CH_PEER_MASK) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_poll_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0xFFF) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) mbar_timeout_trap(bar, parity);
    }
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= [A_even; A_odd] * [B_even; B_odd]^T : M = 256 (128 rows per CTA), each CTA holds N/2 rows of B
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs when all previously issued pair MMAs have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(
          smem_u32(bar))
      : "memory");
}

// Tiles of this CTA: global tile index = blockIdx.x + k*gridDim.x, k = 0..n_my-1; slot = k & 1.
// Producer and issuer walk (pair, step, slot, k-block) in the same order.

// ---- producer warp -------------------------------------------------------------------------
// (measured on B200 and removed: a bulk L2 prefetch of the next step's stash tiles from this warp — fine_bwd 12.3 -> 14.1 ms
//  at 8192 rays; the stash reads are prefetched by the epilogue threads instead, see tile_prefetch_l2)
__device__ __forceinline__ void chain_weight_producer(const ChainTable& tb, const ChainPtrs& ptrs, ChainSmem* s,
                                                      uint8_t* wst, int n_my_tiles, long long tile0, long long tile_stride) {
  const int SL = tb.slots > 0 ? tb.slots : CH_SLOTS;
  uint32_t it = 0;
  for (int k0 = 0; k0 < n_my_tiles; k0 += SL) {
    const int nslot = (n_my_tiles - k0 < SL) ? (n_my_tiles - k0) : SL;
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint8_t* __restrict__ wblob = (st.flags & CHF_W2) ? ptrs.weights2 : ptrs.weights;
      const uint32_t bytes = (uint32_t)st.n * (128u / CH_WSPLIT);
      const int nsl = (st.nkb_a + st.nkb_aux) * CH_WSPLIT;      // consecutive slices of this step's weight image
      for (int slot = 0; slot < nslot; ++slot)
        for (int sl = 0; sl < nsl; ++sl, ++it) {
          const uint32_t stage = it % CH_WSTAGES, n = it / CH_WSTAGES;
          mbar_wait_poll(&s->w_empty[stage], (n & 1) ^ 1);
          FMOV_TR(s, 7, stage, it);
          mbar_expect_tx(&s->w_full[stage], bytes);
          bulk_g2s(wst + stage * WSLOT_BYTES, wblob + st.w_off + (size_t)sl * bytes, bytes, &s->w_full[stage]);
        }
    }
  }
}

// ---- warp 1 -----------------------------------------------------------------------------
// SPLIT = false: the plain chains (no step flags are looked at: the single issuing thread's loop is on the critical path of
// the MMA-bound value chain — the flag tests cost the query kernel 10 % when they were unconditional); SPLIT = true: the
// split-precision chain (CHF_* flags honoured).
template <bool SPLIT = false>
__device__ __forceinline__ void chain_mma_issuer(const ChainTable& tb, ChainSmem* s, uint8_t* act0, uint8_t* aux0,
                                                 uint8_t* wst, uint32_t tmem, int n_my_tiles) {
  uint32_t it = 0;
  uint32_t nstep[CH_SLOTS] = {0, 0};
  const int SL = tb.slots > 0 ? tb.slots : CH_SLOTS;
  for (int k0 = 0; k0 < n_my_tiles; k0 += SL) {
    const int nslot = (n_my_tiles - k0 < SL) ? (n_my_tiles - k0) : SL;
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint32_t idesc = umma_idesc(128, st.n, st.a_fmt, st.b_fmt, 0, 0);
      const int nkb = st.nkb_a + st.nkb_aux;
      for (int slot = 0; slot < nslot; ++slot) {
        const uint8_t fl = SPLIT ? st.flags : (uint8_t)0;
        const int aslot = slot ^ (fl & CHF_A_OTHER);
        uint8_t* act = act0 + aslot * 4 * BLK_BYTES;
        uint8_t* aux = aux0 + aslot * BLK_BYTES;
        const uint32_t accum = (fl & CHF_ACCUM) ? 1u : 0u;
        if (!accum) {
          FMOV_TR(s, 1, slot, nstep[slot]);            // issuer starts waiting for this slot's operand
          mbar_wait_poll(&s->act_ready[slot], nstep[slot] & 1);
          FMOV_TR(s, 2, slot, nstep[slot]);            // operand ready seen
          ++nstep[slot];
          tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t a_base = smem_u32(kb < st.nkb_a ? act + kb * BLK_BYTES : aux + (kb - st.nkb_a) * BLK_BYTES);
          const uint32_t a_base2 = smem_u32(kb < st.nkb_a ? act0 + (aslot ^ 1) * 4 * BLK_BYTES + kb * BLK_BYTES
                                                          : aux0 + (aslot ^ 1) * BLK_BYTES + (kb - st.nkb_a) * BLK_BYTES);
#pragma unroll
          for (int part = 0; part < CH_WSPLIT; ++part, ++it) {
            const uint32_t stage = it % CH_WSTAGES, n = it / CH_WSTAGES;
            FMOV_TR(s, 8, stage, it);
            mbar_wait_poll(&s->w_full[stage], n & 1);
            FMOV_TR(s, 9, stage, it);
            tc_fence_after();
            const uint32_t b_base = smem_u32(wst + stage * WSLOT_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4 / CH_WSPLIT; ++kk) {
              const int ks = part * (4 / CH_WSPLIT) + kk;            // K = 16 slice of the k-block
              umma_f16(tmem + slot * 256, umma_desc_kmajor(a_base + ks * 2 * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                       umma_desc_kmajor(b_base + kk * 2 * ((uint32_t)st.n * 16), (uint32_t)st.n * 16), idesc,
                       ((kb | ks) != 0 ? 1u : 0u) | accum);
              if (SPLIT && (fl & CHF_DUAL_A))
                umma_f16(tmem + slot * 256, umma_desc_kmajor(a_base2 + ks * 2 * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                         umma_desc_kmajor(b_base + kk * 2 * ((uint32_t)st.n * 16), (uint32_t)st.n * 16), idesc, 1u);
            }
            umma_commit(&s->w_empty[stage]);   // slot reusable once these MMAs have read it
          }
        }
        if (!(fl & CHF_NO_COMMIT)) umma_commit(&s->acc_ready[slot]);
        FMOV_TR(s, 3, slot, nstep[slot] - 1);        // all MMAs of the step issued
      }
    }
  }
}

// Tensor maps over the weight blob viewed as rows of 256 bytes, one per half-k-block size (N = 256 / 224 / 48 / 16 ->
// boxes of 64 / 56 / 12 / 4 rows): in pair mode the weight copies are cp.async.bulk.tensor.2d.cta_group::2, the only bulk
// copy whose completion may be signalled on the PARTNER CTA's mbarrier (a plain cp.async.bulk with a remote mbarrier
// never completes: measured, r2s2) — so both CTAs' halves count into the even CTA's w_full and no hop is needed.
struct PairMaps { CUtensorMap m[6]; };      // half k-blocks of N = 256 / 224 / 48 / 16, half bias slices of N = 256 / 224
__host__ __device__ inline int pair_map_index(int n) { return n == 256 ? 0 : n == 224 ? 1 : n == 48 ? 2 : n == 16 ? 3 : -1; }
__host__ __device__ inline int pair_bias_map_index(int n) { return n == 256 ? 4 : n == 224 ? 5 : -1; }
// bias16 steps: column 48 of the AUX block holds 1.0 and column 49 holds 2^-12; the bias slice holds fp16(b) and
// 2^12 (b - fp16(b)) in its K columns 0 / 1 — the residual is scaled into fp16's NORMAL range (an unscaled residual is a
// subnormal fp16 number for |b| < 0.12, and with those the tensor core lost it: colour-net gradient parity 8e-3 -> 1.6e-2)
constexpr int AUX_ONE_COL = 48;
constexpr float BIAS_LO_SCALE = 4096.0f;
__device__ __forceinline__ void tma2_load_rows(void* smem_dst, const CUtensorMap* map, int row, uint32_t bar_even) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(smem_dst)), "l"(map), "r"(0), "r"(row), "r"(bar_even)
      : "memory");
}

// Host side: the six tensor maps of a weight blob of `bytes` bytes; the last few blobs are cached per translation unit
// (a train loop re-packs into the same buffer, and the maps only encode address and extent).
static int chain_pair_maps(const void* blob, long long bytes, PairMaps& out) {
  struct Entry { const void* blob; long long bytes; PairMaps maps; };
  static Entry cache[8];
  static int n_cached = 0, next = 0;
  for (int i = 0; i < n_cached; ++i)
    if (cache[i].blob == blob && cache[i].bytes == bytes) { out = cache[i].maps; return OK; }
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FMOV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FMOV_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    encode = (EncodeFn)fn;
  }
  FMOV_REQUIRE(bytes > 0 && bytes % 256 == 0 && (reinterpret_cast<uintptr_t>(blob) & 15) == 0,
               "weight blob must be 16-byte aligned and a multiple of 256 bytes (%lld)", bytes);
  static const int box_rows[6] = {256 * 64 / 256, 224 * 64 / 256, 48 * 64 / 256, 16 * 64 / 256,      // half k-blocks
                                 256 * 16 / 256, 224 * 16 / 256};                                     // half bias slices
  Entry e;
  e.blob = blob;
  e.bytes = bytes;
  for (int i = 0; i < 6; ++i) {
    cuuint64_t gdim[2] = {256, (cuuint64_t)(bytes / 256)};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {256, (cuuint32_t)box_rows[i]};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&e.maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(blob), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FMOV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for the %d-row weight box", (int)r, box_rows[i]);
  }
  cache[next] = e;
  next = (next + 1) % 8;
  if (n_cached < 8) ++n_cached;
  out = e.maps;
  return OK;
}

// ---- CTA-pair mode (FMOV_CH_PAIR): producer / issuer ------------------------------------------------
// Both walk (tile pair, step, slot, k-block) in the same order; ring stage = it % CH_PW_STAGES holds this CTA's half
// (rows [rank*N/2, (rank+1)*N/2)) of one 64-wide k-block.  n_pair_tiles = tiles per CTA of the pair (the odd CTA pads
// with a dummy tile when it has one less).
// BIAS: the table may contain bias16 steps (the backward kernel's does not: its instantiation carries no code for them)
template <bool BIAS>
__device__ __forceinline__ void chain_weight_producer_pair(const ChainTable& tb, const PairMaps& maps, ChainSmem* s,
                                                           uint8_t* wst, int n_pair_tiles, uint32_t rank) {
  uint32_t it = 0;
  for (int k0 = 0; k0 < n_pair_tiles; k0 += CH_SLOTS) {
    const int nslot = (n_pair_tiles - k0 < CH_SLOTS) ? (n_pair_tiles - k0) : CH_SLOTS;
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint32_t half = (uint32_t)st.n * 64u;                  // bytes of N/2 rows x 64 K
      const CUtensorMap* map = &maps.m[pair_map_index(st.n)];
      const int row0 = (int)((st.w_off + rank * half) >> 8);       // this CTA's half of k-block 0, in rows of 256 bytes
      const int nkb = st.nkb_a + st.nkb_aux;
      for (int slot = 0; slot < nslot; ++slot) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t stage = it % CH_PW_STAGES, n = it / CH_PW_STAGES;
          mbar_wait_poll(&s->w_empty[stage], (n & 1) ^ 1);
          FMOV_TR(s, 7, stage, it);                       // ring stage seen free: the copy is issued now
          // both CTAs' copies complete on the EVEN CTA's barrier (the issuer lives there); it expects both halves
          if (rank == 0) mbar_expect_tx(&s->w_full[stage], 2u * half);
          tma2_load_rows(wst + stage * CH_PW_STAGE_BYTES, map, row0 + kb * (int)((2u * half) >> 8),
                         smem_u32(&s->w_full[stage]) & CH_PEER_MASK);
        }
        if (BIAS && st.bias16) {          // this CTA's N/2 rows of the [N x 16] bias slice behind the k-blocks
          const uint32_t bhalf = (uint32_t)st.n * 16u;
          const uint32_t stage = it % CH_PW_STAGES, n = it / CH_PW_STAGES;
          mbar_wait_poll(&s->w_empty[stage], (n & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&s->w_full[stage], 2u * bhalf);
          tma2_load_rows(wst + stage * CH_PW_STAGE_BYTES, &maps.m[pair_bias_map_index(st.n)],
                         (int)((st.w_off + (uint32_t)nkb * 2u * half + rank * bhalf) >> 8),
                         smem_u32(&s->w_full[stage]) & CH_PEER_MASK);
          ++it;
        }
      }
    }
  }
}
// even CTA: one thread issues every MMA of the pair
template <bool BIAS>
__device__ __forceinline__ void chain_mma_issuer_pair(const ChainTable& tb, ChainSmem* s, uint8_t* act0, uint8_t* aux0,
                                                      uint8_t* wst, uint32_t tmem, int n_pair_tiles) {
  uint32_t it = 0;
  uint32_t nstep[CH_SLOTS] = {0, 0};
  for (int k0 = 0; k0 < n_pair_tiles; k0 += CH_SLOTS) {
    const int nslot = (n_pair_tiles - k0 < CH_SLOTS) ? (n_pair_tiles - k0) : CH_SLOTS;
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint32_t idesc = umma_idesc(256, st.n, st.a_fmt, st.b_fmt, 0, 0);
      const uint32_t b_chunk = (uint32_t)st.n * 8u;                // (N/2 rows) x 16 B: distance between chunk columns of B
      const int nkb = st.nkb_a + st.nkb_aux;
      for (int slot = 0; slot < nslot; ++slot) {
        uint8_t* act = act0 + slot * 4 * BLK_BYTES;
        uint8_t* aux = aux0 + slot * BLK_BYTES;
        FMOV_TR(s, 1, slot, nstep[slot]);
        mbar_wait_poll_cluster(&s->act_ready[slot], nstep[slot] & 1);     // both CTAs' operands written, accumulators drained
        FMOV_TR(s, 2, slot, nstep[slot]);
        ++nstep[slot];
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t a_base = smem_u32(kb < st.nkb_a ? act + kb * BLK_BYTES : aux + (kb - st.nkb_a) * BLK_BYTES);
          const uint32_t stage = it % CH_PW_STAGES, n = it / CH_PW_STAGES;
          FMOV_TR(s, 8, stage, it);                       // issuer starts waiting for the stage
          mbar_wait_poll_cluster(&s->w_full[stage], n & 1);      // both halves landed (each in its own CTA's ring)
          FMOV_TR(s, 9, stage, it);                       // stage seen full
          tc_fence_after();
          const uint32_t b_base = smem_u32(wst + stage * CH_PW_STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma2_f16(tmem + slot * 256, umma_desc_kmajor(a_base + ks * 2 * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                      umma_desc_kmajor(b_base + ks * 2 * b_chunk, b_chunk), idesc, (kb | ks) != 0 ? 1u : 0u);
          umma2_commit_both(&s->w_empty[stage]);       // both CTAs' ring stages reusable once these MMAs have read them
        }
        if (BIAS && st.bias16) {          // + 1.0 * bias_hi + 2^-12 * (2^12 bias_lo): K = 16 slice of the AUX block x the bias slice
          const uint32_t stage = it % CH_PW_STAGES, n = it / CH_PW_STAGES;
          mbar_wait_poll_cluster(&s->w_full[stage], n & 1);
          tc_fence_after();
          umma2_f16(tmem + slot * 256, umma_desc_kmajor(smem_u32(aux) + (AUX_ONE_COL / 8) * TI_CHUNK_STRIDE, TI_CHUNK_STRIDE),
                    umma_desc_kmajor(smem_u32(wst + stage * CH_PW_STAGE_BYTES), b_chunk), idesc, 1u);
          umma2_commit_both(&s->w_empty[stage]);
          ++it;
        }
        umma2_commit_both(&s->acc_ready[slot]);
        FMOV_TR(s, 3, slot, nstep[slot] - 1);
      }
    }
  }
}

// ---- epilogue-side helpers (thread <-> tile row / TMEM lane) -------------------------------------------
struct EpiCtx {
  ChainSmem* s;
  uint8_t* act;        // this slot's ACT (4 blocks)
  uint8_t* aux;        // this slot's AUX (1 block)
  uint32_t tmem;       // TMEM address of this slot's accumulator with the warp's lane quarter folded in
  int slot;
  int wg;              // column half handled by this warpgroup: chunks of columns [128*wg, 128*wg + 128)
  int row;             // 0..127 (tile row == TMEM lane)
  uint32_t acc_n;      // accumulator phases consumed
};

__device__ __forceinline__ void epi_init(EpiCtx& c, ChainSmem* s, uint8_t* act0, uint8_t* aux0, uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;                  // TMEM lanes this warp may access
  c.slot = (warp - EPI_WARP0) / (4 * CH_WGS);
  c.wg = ((warp - EPI_WARP0) >> 2) % CH_WGS;
  c.s = s;
  c.act = act0 + c.slot * 4 * BLK_BYTES;
  c.aux = aux0 + c.slot * BLK_BYTES;
  c.row = quarter * 32 + lane;
  c.tmem = tmem_base + c.slot * 256 + ((uint32_t)(quarter * 32) << 16);
  c.acc_n = 0;
}
__device__ __forceinline__ void epi_wait_acc(EpiCtx& c) {
  if ((threadIdx.x & 31) == 0) FMOV_TR(c.s, 4, threadIdx.x >> 5, c.acc_n);      // epilogue warp starts waiting
  mbar_wait(&c.s->acc_ready[c.slot], c.acc_n & 1);
  if ((threadIdx.x & 31) == 0) FMOV_TR(c.s, 5, threadIdx.x >> 5, c.acc_n);      // accumulator ready
  ++c.acc_n;
  tc_fence_after();
}
// A operand for the next step is written (or nothing to write) and this slot's accumulator is drained.
template <bool PAIR = CH_PAIR>
__device__ __forceinline__ void epi_signal_act(EpiCtx& c) {
  tc_fence_before();
  fence_proxy_async();
  if (PAIR) {          // one arrival per warp, on the even CTA's barrier (its issuer serves both CTAs)
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_even_cta(&c.s->act_ready[c.slot]);
  } else {
    mbar_arrive(&c.s->act_ready[c.slot]);
  }
  if ((threadIdx.x & 31) == 0) FMOV_TR(c.s, 6, threadIdx.x >> 5, c.acc_n);      // epilogue warp done with the step
}
// 32 accumulator columns [col0, col0+32) of this thread's row
__device__ __forceinline__ void acc_load32(const EpiCtx& c, int col0, float* v) {
  tmem_ld32(c.tmem + col0, v);
  tmem_ld_wait();
}
__device__ __forceinline__ void acc_load16(const EpiCtx& c, int col0, float* v) {
  tmem_ld16(c.tmem + col0, v);
  tmem_ld_wait();
}
// all 256 threads of a tile slot (both column warpgroups); orders their shared/global writes
__device__ __forceinline__ void slot_sync(const EpiCtx& c) { named_bar_sync(1 + c.slot, EPI_THREADS); }

// ---- 16-column chunks (index ck = 0..15 within a 256-wide tile): 2 x 16-byte pieces of a row -----------------
__device__ __forceinline__ void pack2(const float* v, bool bf16, uint4* q) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (bf16) {
      q[i].x = pack_bf2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_bf2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_bf2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_bf2(v[i * 8 + 6], v[i * 8 + 7]);
    } else {
      q[i].x = pack_h2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_h2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_h2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_h2(v[i * 8 + 6], v[i * 8 + 7]);
    }
  }
}
__device__ __forceinline__ void pack2_grad(const float* v, uint4* q) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (kGradBf16) {
      q[i].x = pack_bf2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_bf2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_bf2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_bf2(v[i * 8 + 6], v[i * 8 + 7]);
    } else {
      q[i].x = pack_h2_sat(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_h2_sat(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_h2_sat(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_h2_sat(v[i * 8 + 6], v[i * 8 + 7]);
    }
  }
}
// element j (0..15) of a packed 16-column chunk
__device__ __forceinline__ float2 chunk_pair(const uint4* q, int jp, bool bf16) {   // jp = pair index 0..7
  const uint4 w = q[jp >> 2];
  const uint32_t u = (jp & 3) == 0 ? w.x : (jp & 3) == 1 ? w.y : (jp & 3) == 2 ? w.z : w.w;
  return bf16 ? unpack_bf2(u) : unpack_h2(u);
}
// `tilep` = base of a 256-wide tile (4 blocks) + row*16, shared or global; chunk ck -> block ck/4, pieces 2*(ck%4)..
__device__ __forceinline__ void chunk_store(uint8_t* tilep, int ck, const uint4* q) {
  uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
  *reinterpret_cast<uint4*>(p) = q[0];
  *reinterpret_cast<uint4*>(p + TI_CHUNK_STRIDE) = q[1];
}
__device__ __forceinline__ void chunk_load(const uint8_t* tilep, int ck, uint4* q) {
  const uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
  q[0] = *reinterpret_cast<const uint4*>(p);
  q[1] = *reinterpret_cast<const uint4*>(p + TI_CHUNK_STRIDE);
}

// L2 prefetch of this warp's rows of chunk columns [ck_first*2, (ck_first+n_ck)*2) of a 256-wide tile: a warp's 32 rows
// of one 16-byte chunk column are 512 contiguous bytes = 4 lines, so every 8th lane prefetches one line.  Issued by the
// epilogue threads when they start waiting for the accumulator: the MMA time (4-5 K cycles) is the lead that turns the
// HBM latency of the step's stash reads into L2 hits without holding registers.
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tile_prefetch_l2(const uint8_t* tilep, int ck_first, int n_ck) {
  if ((threadIdx.x & 7) == 0) {
    for (int ck = ck_first; ck < ck_first + n_ck; ++ck) {
      const uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
      prefetch_l2(p);
      prefetch_l2(p + TI_CHUNK_STRIDE);
    }
  }
}

// read-only 16-byte load that the compiler may not move (software-pipelined bias fetch)
__device__ __forceinline__ float4 ldg_f4_volatile(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ---- L2 eviction-priority hints: a kernel that reads a stash tensor twice marks the first
// read evict_last and everything it streams evict_first, so that the second read has a chance to hit the 126 MB L2.
// Hints cannot change results, only where lines live.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ldg_pol(const void* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_pol(void* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void chunk_load_pol(const uint8_t* tilep, int ck, uint4* q, uint64_t pol) {
  const uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
  q[0] = ldg_pol(p, pol);
  q[1] = ldg_pol(p + TI_CHUNK_STRIDE, pol);
}
__device__ __forceinline__ void chunk_store_pol(uint8_t* tilep, int ck, const uint4* q, uint64_t pol) {
  uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
  stg_pol(p, q[0], pol);
  stg_pol(p + TI_CHUNK_STRIDE, q[1], pol);
}
__device__ __forceinline__ void tile_prefetch_l2_keep(const uint8_t* tilep, int ck_first, int n_ck) {
  if ((threadIdx.x & 7) == 0) {
    for (int ck = ck_first; ck < ck_first + n_ck; ++ck) {
      const uint8_t* p = tilep + (ck >> 2) * BLK_BYTES + (2 * (ck & 3)) * TI_CHUNK_STRIDE;
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p + TI_CHUNK_STRIDE));
    }
  }
}

__device__ __forceinline__ void row_half_store_pol(uint8_t* rowp, int h, const uint4* q, uint64_t pol) {
#pragma unroll
  for (int i = 0; i < 4; ++i) stg_pol(rowp + (4 * h + i) * TI_CHUNK_STRIDE, q[i], pol);
}
__device__ __forceinline__ void row_half_load_pol(const uint8_t* rowp, int h, uint4* q, uint64_t pol) {
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ldg_pol(rowp + (4 * h + i) * TI_CHUNK_STRIDE, pol);
}

// ---- rows of tile images: a half block = 32 columns = 4 x 16-byte chunks --------------------------------
// chunk ch (0..7) of row r sits at blk + ch*2048 + r*16: the 32 lanes of a warp touch 512 contiguous bytes.
__device__ __forceinline__ void pack4_grad(const float* v, uint4* q) {   // gradient tile format
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (kGradBf16) {
      q[i].x = pack_bf2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_bf2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_bf2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_bf2(v[i * 8 + 6], v[i * 8 + 7]);
    } else {
      q[i].x = pack_h2_sat(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_h2_sat(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_h2_sat(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_h2_sat(v[i * 8 + 6], v[i * 8 + 7]);
    }
  }
}
__device__ __forceinline__ void pack4(const float* v, bool bf16, uint4* q) {   // 32 floats -> 4 chunks
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bf16) {
      q[i].x = pack_bf2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_bf2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_bf2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_bf2(v[i * 8 + 6], v[i * 8 + 7]);
    } else {
      q[i].x = pack_h2(v[i * 8 + 0], v[i * 8 + 1]); q[i].y = pack_h2(v[i * 8 + 2], v[i * 8 + 3]);
      q[i].z = pack_h2(v[i * 8 + 4], v[i * 8 + 5]); q[i].w = pack_h2(v[i * 8 + 6], v[i * 8 + 7]);
    }
  }
}
__device__ __forceinline__ void unpack4(const uint4* q, bool bf16, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = bf16 ? unpack_bf2(w[j]) : unpack_h2(w[j]);
      v[i * 8 + 2 * j] = f.x;
      v[i * 8 + 2 * j + 1] = f.y;
    }
  }
}
// half h (0/1) of a block row: chunks 4h..4h+3.  `rowp` = block base + row*16 (shared or global).
__device__ __forceinline__ void row_half_store(uint8_t* rowp, int h, const uint4* q) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(rowp + (4 * h + i) * TI_CHUNK_STRIDE) = q[i];
}
__device__ __forceinline__ void row_half_load(const uint8_t* rowp, int h, uint4* q) {
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const uint4*>(rowp + (4 * h + i) * TI_CHUNK_STRIDE);
}

}  // namespace fmov
