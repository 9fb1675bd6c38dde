// Fused MLP "chain" engine: one CTA walks a 128-point tile through a table of GEMM steps with the
// activation tile resident in shared memory, weights streamed from L2 by TMA bulk copies, the
// accumulator in TMEM and a per-kernel epilogue between steps.
//
// Warp roles (7 warps, 224 threads):
//   warp 0  weight producer   cp.async.bulk  global weight images -> WST ring     (1 lane)
//   warp 1  MMA issuer        tcgen05.mma    A = ACT/AUX blocks, B = WST slot      (1 lane) + TMEM alloc
//   warp 2  side producer     cp.async.bulk  stash blocks -> SIDE ring             (1 lane)
//   warps 3..6 epilogue       tcgen05.ld -> math -> st.shared (next A operand) -> bulk store to the stash
//
// Synchronisation (all mbarriers, one phase per step):
//   act_ready   (count 128)  epilogue/input stage -> MMA issuer : A operand written, TMEM drained
//   acc_ready   (count 1)    MMA issuer (tcgen05.commit) -> epilogue : accumulator complete
//   w_full/w_empty[NS]       weight ring;  side_full/side_empty[4] side ring (empty count 128)
//   stash_bar   (count 1)    epilogue store thread -> side producer : earlier bulk stores have landed
#pragma once
#include "fmov_common.cuh"

namespace fmov {

constexpr int CH_THREADS = 224;
constexpr int EPI_WARP0 = 3;
constexpr int EPI_THREADS = 128;
constexpr int WSLOT_BYTES = 256 * 128;     // [256 rows x 64] fp16
constexpr int SIDE_SLOTS = 4;
constexpr int MAX_STEPS = 40;
constexpr int MAX_SIDE = 160;
constexpr int MAX_STASH = 56;

struct ChainStep {
  uint32_t w_off;       // byte offset of the first k-block of this step in the weight blob
  uint16_t n;           // MMA N (multiple of 16, <= 256); k-block bytes = n*128
  uint8_t nkb_a;        // k-blocks read from ACT ...
  uint8_t nkb_aux;      // ... followed by k-blocks read from AUX
  uint8_t a_fmt;        // FMT_F16 / FMT_BF16 of the A operand (activations / gradients)
  uint8_t b_fmt;        // format of the weight image
  uint8_t side_first;   // first entry in side[] consumed by this step's epilogue
  uint8_t side_cnt;     // number of side blocks
  uint8_t wait_stash;   // side producer waits for stash_bar before this step's side loads
  uint8_t no_mma;       // epilogue-only step (side loads but no weights / MMA / accumulator)
  uint8_t pad[2];
};
struct SideRef {
  uint8_t tensor;       // index into stash[]
  uint8_t kb;           // block inside the tile
};
struct ChainTable {
  int n_steps;
  int n_side;
  ChainStep step[MAX_STEPS];
  SideRef side[MAX_SIDE];
  uint8_t stash_kb[MAX_STASH];   // blocks per tile of each stash tensor
};
struct ChainPtrs {
  const uint8_t* weights;        // packed weight images
  uint8_t* stash[MAX_STASH];     // tile-image tensors in HBM
};

struct ChainSmem {
  // barriers live at the front of dynamic smem (after 1024-alignment)
  uint64_t act_ready;
  uint64_t acc_ready;
  uint64_t stash_bar;
  uint64_t misc_bar;    // epilogue-issued bulk loads (tile re-loads into ACT)
  uint64_t w_full[4];
  uint64_t w_empty[4];
  uint64_t side_full[SIDE_SLOTS];
  uint64_t side_empty[SIDE_SLOTS];
  uint32_t tmem_base;
  uint32_t pad_;
};

// Dynamic smem carve-up: [ChainSmem | pad to 1024][ACT 64K][AUX 16K][SIDE 4x16K (optional)][WST NSx32K]
template <int NS, bool HAS_SIDE>
struct ChainLayout {
  static constexpr int HDR = 1024;
  static constexpr int ACT = HDR;
  static constexpr int AUX = ACT + 4 * BLK_BYTES;
  static constexpr int SIDE = AUX + BLK_BYTES;
  static constexpr int WST = SIDE + (HAS_SIDE ? SIDE_SLOTS * BLK_BYTES : 0);
  static constexpr int TOTAL = WST + NS * WSLOT_BYTES;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024-alignment
};

__device__ __forceinline__ uint8_t* chain_smem_base(uint8_t* raw) {
  uintptr_t p = reinterpret_cast<uintptr_t>(raw);
  p = (p + 1023) & ~uintptr_t(1023);
  return reinterpret_cast<uint8_t*>(p);
}

template <int NS>
__device__ __forceinline__ void chain_init_barriers(ChainSmem* s) {
  mbar_init(&s->act_ready, EPI_THREADS);
  mbar_init(&s->acc_ready, 1);
  mbar_init(&s->stash_bar, 1);
  mbar_init(&s->misc_bar, 1);
  for (int i = 0; i < NS; ++i) {
    mbar_init(&s->w_full[i], 1);
    mbar_init(&s->w_empty[i], 1);
  }
  for (int i = 0; i < SIDE_SLOTS; ++i) {
    mbar_init(&s->side_full[i], 1);
    mbar_init(&s->side_empty[i], EPI_THREADS);
  }
  fence_mbar_init();
}

// ---- warp 0 -----------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void chain_weight_producer(const ChainTable& tb, const uint8_t* __restrict__ wblob,
                                                      ChainSmem* s, uint8_t* wst, int n_my_tiles) {
  uint32_t it = 0;
  for (int t = 0; t < n_my_tiles; ++t) {
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint32_t bytes = (uint32_t)st.n * 128u;
      const int nkb = st.nkb_a + st.nkb_aux;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t slot = it % NS, n = it / NS;
        mbar_wait(&s->w_empty[slot], (n & 1) ^ 1);
        mbar_expect_tx(&s->w_full[slot], bytes);
        bulk_g2s(wst + slot * WSLOT_BYTES, wblob + st.w_off + (size_t)kb * bytes, bytes, &s->w_full[slot]);
      }
    }
  }
}

// ---- warp 1 -----------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void chain_mma_issuer(const ChainTable& tb, ChainSmem* s, uint8_t* act, uint8_t* aux,
                                                 uint8_t* wst, uint32_t tmem, int n_my_tiles) {
  uint32_t it = 0, nstep = 0;
  for (int t = 0; t < n_my_tiles; ++t) {
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.no_mma) continue;
      const uint32_t idesc = umma_idesc(128, st.n, st.a_fmt, st.b_fmt, 0, 0);
      mbar_wait(&s->act_ready, nstep & 1);
      ++nstep;
      tc_fence_after();
      const int nkb = st.nkb_a + st.nkb_aux;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t slot = it % NS, n = it / NS;
        mbar_wait(&s->w_full[slot], n & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(kb < st.nkb_a ? act + kb * BLK_BYTES : aux + (kb - st.nkb_a) * BLK_BYTES);
        const uint32_t b_base = smem_u32(wst + slot * WSLOT_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_f16(tmem, umma_desc_kmajor(a_base + ks * 32), umma_desc_kmajor(b_base + ks * 32), idesc,
                   (kb | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&s->w_empty[slot]);   // slot reusable once these MMAs have read it
      }
      umma_commit(&s->acc_ready);
    }
  }
}

// ---- warp 2 -----------------------------------------------------------------------------
__device__ __forceinline__ void chain_side_producer(const ChainTable& tb, const ChainPtrs& ptrs, ChainSmem* s,
                                                    uint8_t* side, int tile0, int tile_stride, int n_my_tiles) {
  uint32_t it = 0, nstash = 0;
  for (int t = 0; t < n_my_tiles; ++t) {
    const size_t tile = (size_t)tile0 + (size_t)t * tile_stride;
    for (int si = 0; si < tb.n_steps; ++si) {
      const ChainStep st = tb.step[si];
      if (st.wait_stash) {
        mbar_wait(&s->stash_bar, nstash & 1);
        ++nstash;
      }
      for (int j = 0; j < st.side_cnt; ++j, ++it) {
        const SideRef r = tb.side[st.side_first + j];
        const uint32_t slot = it % SIDE_SLOTS, n = it / SIDE_SLOTS;
        mbar_wait(&s->side_empty[slot], (n & 1) ^ 1);
        mbar_expect_tx(&s->side_full[slot], BLK_BYTES);
        const uint8_t* src = ptrs.stash[r.tensor] + (tile * tb.stash_kb[r.tensor] + r.kb) * (size_t)BLK_BYTES;
        bulk_g2s(side + slot * BLK_BYTES, src, BLK_BYTES, &s->side_full[slot]);
      }
    }
  }
}

// ---- epilogue-side helpers (warps 3..6, 128 threads; thread <-> tile row / TMEM lane) ------
struct EpiCtx {
  ChainSmem* s;
  uint8_t* act;
  uint8_t* aux;
  uint8_t* side;
  uint32_t tmem;       // TMEM base with this warp's lane quarter folded in
  int row;             // 0..127 (tile row == TMEM lane)
  int etid;            // 0..127 thread index inside the epilogue group
  uint32_t acc_n;      // accumulator phases consumed
  uint32_t side_n;     // side blocks consumed
  uint32_t stash_n;    // stash_bar signals issued (store thread only)
  uint32_t misc_n;     // misc_bar phases consumed
  bool store_pending;  // store thread: a bulk store may still be reading ACT/AUX
};

__device__ __forceinline__ void epi_init(EpiCtx& c, ChainSmem* s, uint8_t* act, uint8_t* aux, uint8_t* side,
                                         uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;                  // TMEM lanes this warp may access
  c.s = s; c.act = act; c.aux = aux; c.side = side;
  c.row = quarter * 32 + lane;
  c.etid = (warp - EPI_WARP0) * 32 + lane;
  c.tmem = tmem_base + ((uint32_t)(quarter * 32) << 16);
  c.acc_n = 0; c.side_n = 0; c.stash_n = 0; c.misc_n = 0; c.store_pending = false;
}
__device__ __forceinline__ void epi_wait_acc(EpiCtx& c) {
  mbar_wait(&c.s->acc_ready, c.acc_n & 1);
  ++c.acc_n;
  tc_fence_after();
}
// A operand for the next step is written (or nothing to write) and TMEM is drained.
__device__ __forceinline__ void epi_signal_act(EpiCtx& c) {
  tc_fence_before();
  fence_proxy_async();
  mbar_arrive(&c.s->act_ready);
}
// Call before overwriting ACT/AUX if a bulk store of it may be in flight.
__device__ __forceinline__ void epi_before_write(EpiCtx& c) {
  if (c.etid == 0 && c.store_pending) {
    bulk_wait_read0();
    c.store_pending = false;
  }
  named_bar_sync(1, EPI_THREADS);
}
// Store `nkb` blocks starting at smem `src` to the stash tensor image. All 128 threads call it after
// their st.shared writes.
__device__ __forceinline__ void epi_store_blocks(EpiCtx& c, const uint8_t* src, uint8_t* dst_tile, int nkb) {
  fence_proxy_async();
  named_bar_sync(1, EPI_THREADS);
  if (c.etid == 0) {
    for (int kb = 0; kb < nkb; ++kb) bulk_s2g(dst_tile + (size_t)kb * BLK_BYTES, src + kb * BLK_BYTES, BLK_BYTES);
    bulk_commit();
    c.store_pending = true;
  }
}
// All earlier stash stores of this CTA have landed in global memory -> release the side producer.
__device__ __forceinline__ void epi_publish_stash(EpiCtx& c) {
  if (c.etid == 0) {
    bulk_wait_all0();
    c.store_pending = false;
    mbar_arrive(&c.s->stash_bar);
  }
}
__device__ __forceinline__ const uint8_t* epi_side_wait(EpiCtx& c) {
  const uint32_t slot = c.side_n % SIDE_SLOTS, n = c.side_n / SIDE_SLOTS;
  mbar_wait(&c.s->side_full[slot], n & 1);
  return c.side + slot * BLK_BYTES;
}
__device__ __forceinline__ void epi_side_release(EpiCtx& c) {
  const uint32_t slot = c.side_n % SIDE_SLOTS;
  mbar_arrive(&c.s->side_empty[slot]);
  ++c.side_n;
}

// read / write one row-chunk (8 x 16-bit) of a block
__device__ __forceinline__ uint4 blk_ld_chunk(const uint8_t* blk, int row, int chunk) {
  return *reinterpret_cast<const uint4*>(blk + ti_chunk_off(row, chunk));
}
__device__ __forceinline__ void blk_st_chunk(uint8_t* blk, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(blk + ti_chunk_off(row, chunk)) = v;
}


// Publish variant that also covers plain st.global writes made by all epilogue threads (e.g. tile rows
// written straight to HBM): every thread fences its writes towards the async proxy first.
__device__ __forceinline__ void epi_publish_stash_all(EpiCtx& c) {
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  named_bar_sync(1, EPI_THREADS);
  epi_publish_stash(c);
}
// Re-load `nkb` blocks of a stash tile into smem `dst` (ACT); all 128 threads call and wait.
__device__ __forceinline__ void epi_reload_blocks(EpiCtx& c, uint8_t* dst, const uint8_t* src_tile, int nkb) {
  if (c.etid == 0) {
    mbar_expect_tx(&c.s->misc_bar, (uint32_t)nkb * BLK_BYTES);
    for (int kb = 0; kb < nkb; ++kb) bulk_g2s(dst + kb * BLK_BYTES, src_tile + (size_t)kb * BLK_BYTES, BLK_BYTES, &c.s->misc_bar);
  }
  mbar_wait(&c.s->misc_bar, c.misc_n & 1);
  ++c.misc_n;
}
// 64 accumulator columns [col0, col0+64) of this thread's row; columns >= n_mma read as zero.
__device__ __forceinline__ void acc_load64(const EpiCtx& c, int col0, int n_mma, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (col0 + q * 16 < n_mma) {
      tmem_ld16(c.tmem + col0 + q * 16, v + q * 16);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[q * 16 + j] = 0.f;
    }
  }
  tmem_ld_wait();
}
// one 64-wide row of a block <-> 64 floats
__device__ __forceinline__ void row_load64(const uint8_t* blk, int row, bool bf16, float* v) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const uint4 q = blk_ld_chunk(blk, row, ch);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = bf16 ? unpack_bf2(w[j]) : unpack_h2(w[j]);
      v[ch * 8 + 2 * j] = f.x;
      v[ch * 8 + 2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void row_store64(uint8_t* blk, int row, bool bf16, const float* v) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 q;
    if (bf16) {
      q.x = pack_bf2(v[ch * 8 + 0], v[ch * 8 + 1]); q.y = pack_bf2(v[ch * 8 + 2], v[ch * 8 + 3]);
      q.z = pack_bf2(v[ch * 8 + 4], v[ch * 8 + 5]); q.w = pack_bf2(v[ch * 8 + 6], v[ch * 8 + 7]);
    } else {
      q.x = pack_h2(v[ch * 8 + 0], v[ch * 8 + 1]); q.y = pack_h2(v[ch * 8 + 2], v[ch * 8 + 3]);
      q.z = pack_h2(v[ch * 8 + 4], v[ch * 8 + 5]); q.w = pack_h2(v[ch * 8 + 6], v[ch * 8 + 7]);
    }
    blk_st_chunk(blk, row, ch, q);
  }
}

}  // namespace fmov
