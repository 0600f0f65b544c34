// tcgen05 self-attention over unpadded sequences.  PERSISTENT: 3 CTAs per SM, each loops over work items
// (sequence, 128-query tile, head); TMEM / barriers are set up once per CTA and the TMA producer runs ahead into the
// next item's Q / K / V while the current item is still in its softmax.
//
//   warp 0      : TMA producer  (Q tile; K blocks of 64 keys in a 3-stage ring, V blocks in a 2-stage ring, K issued one
//                               block ahead of V; SWIZZLE_128B boxes)
//   warp 1      : MMA issuer    S  = Q K^T  (tcgen05.mma 128x64x16, both operands K-major)        -> TMEM S
//                               O += P V    (tcgen05.mma 128x64x16, A = P from smem, B = V MN-major) -> TMEM O
//   warps 2..5  : softmax       one warp per TMEM lane quarter, thread == query row (all 64 columns of the block, so
//                               the row max / sum never leave the thread): tcgen05.ld S (which frees S for the next
//                               block's QK^T at once), mask, exponentials in base 2 with packed fp32x2 arithmetic
//                               (FFMA2 / FADD2), P (fp16) -> swizzled smem
// The output accumulates in TMEM across key blocks (the MMA's accumulate flag), not in registers: P is scaled by a
// per-row reference max m_used that is only raised when the running max exceeds it by more than 2^8 ("lazy
// rescaling": p <= 256 stays far inside fp16 / fp32 range and O / l is exact for any reference), and only then
// the warp multiplies its O rows in TMEM by 2^(m_old - m_new) (tcgen05.ld / st between PV_{G-1} and PV_G).  That
// removes the per-block read-modify-write of O from the softmax warps and brings them to ~100 registers, so three
// CTAs fit per SM.
// Sizing (measured, profiles/README.md): the softmax warps are bound by dependent-latency hops per key block (mbarrier
// probes, TMEM loads, proxy fence) with the MUFU pipe at one third; more resident CTAs per SM is what fills it.
// All ring / buffer indices and mbarrier parities derive from per-role running counters (item count `it_n`, key-block
// count `g`), which every role advances identically; every parity wait is placed so that the barrier can be at most
// one phase ahead of the waiter.  q/k already carry RoPE (Wqkv GEMM epilogue).  On local layers only key blocks
// intersecting |i - j| <= window are visited; blocks fully outside a warp's window skip TMEM.
// Producer / MMA loops are warp-convergent with elect.sync around the TMA / MMA instructions (uniform datapath).
#include <stdlib.h>

#include <atomic>

#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int AQ = 128, AK = 64, AD = 64;
constexpr int SOFT_WARPS = 4;
// "lean" = split precision at 2 CTAs per SM: K ring 2 deep, V single-buffered -> 112 KB of tiles + 1 KB of barriers, so
// two CTAs fill the SM's 228 KB exactly and overlap each other's S -> softmax -> PV chains, which one CTA alone runs
// back to back (measured at 1024 x 512 tokens: 4.13 -> 3.02 ms global, 2.39 -> 1.78 ms local layers).
constexpr bool att_lean(bool split) { return split; }
constexpr int att_ks(bool lean) { return lean ? 2 : 3; }   // K ring depth: K of block g+1 / g+2 loads while block g is in its softmax, so
                                         // S = Q K^T of the next block never waits for an L2 round trip
constexpr int att_vs(bool lean) { return lean ? 1 : 2; }   // V ring depth (V is needed one softmax later than K)
constexpr int ATT_CTAS_PER_SM = 3;       // fp16 kernel; the split kernel runs 2 (att_lean)
constexpr int att_ctas_per_sm(bool split) { return split ? 2 : ATT_CTAS_PER_SM; }
constexpr int ATT_THREADS = 32 * (2 + SOFT_WARPS);
constexpr uint32_t ATT_TMEM_COLS = 128;  // S [0,64)  O [64,128)
constexpr int SQ_BYTES = AQ * AD * 2;    // 16384
constexpr int SKV_BYTES = AK * AD * 2;   // 8192
constexpr int SP_BYTES = AQ * AK * 2;    // 16384
// SPLIT (split-precision / "precise" mode, gemm.cuh): every tile exists twice, hi plane then lo plane
template <bool SPLIT>
constexpr int att_smem() {
  constexpr bool lean = att_lean(SPLIT);
  constexpr int tiles = (SPLIT ? 2 : 1) * (SQ_BYTES + (att_ks(lean) + att_vs(lean)) * SKV_BYTES + SP_BYTES);
  return lean ? tiles + 1024 : tiles + 1024 + 256;   // lean: no slack for an align-up, the window must be 1 KB aligned
}
constexpr float RESCALE_THRESHOLD = 8.f; // log2 units

// split-precision planes of two values: hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split_half2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16-byte store of chunk c of a SWIZZLE_128B row: address = row_base_with_swizzle ^ (c << 4).  The XOR sits inside the
// asm block so that the compiler cannot hoist eight loop-invariant addresses out of the item loop (it did, and spilled
// them).
template <int C>
__device__ __forceinline__ void st_p_chunk(uint32_t row_sw, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile(
      "{\n\t.reg .b32 addr;\n\txor.b32 addr, %0, %1;\n\tst.shared.v4.b32 [addr], {%2, %3, %4, %5};\n\t}\n" ::"r"(row_sw),
      "n"(C << 4), "r"(a), "r"(b), "r"(c), "r"(d)
      : "memory");
}

template <int V>
struct IntC {
  static constexpr int value = V;
};

struct Item {  // one (sequence, 128-query tile, head)
  int s0, L, q0, head, j_lo, nb;
};

// Item w = pair * heads + head of this CTA's stream (w = blockIdx.x, += gridDim.x).  Only the producer warp walks the
// work list (16-byte entries {first token of the sequence, sequence length, q0, -}, next entry prefetched) and
// publishes {s0, L, q0, head} in shared memory (two slots, item parity) before it arms q_full; the MMA and softmax
// warps read it after a barrier wait that is causally later, so they carry no decode state in registers.
template <bool LOCAL>
__device__ __forceinline__ Item make_item(int4 e, int window) {
  Item it;
  it.s0 = e.x;
  it.L = e.y;
  it.q0 = e.z;
  it.head = e.w;
  int kv_lo = 0, kv_hi = it.L;
  if (LOCAL) {
    kv_lo = max(0, it.q0 - window);
    kv_hi = min(it.L, it.q0 + AQ + window);
  }
  it.j_lo = kv_lo / AK;
  it.nb = (kv_hi + AK - 1) / AK - it.j_lo;
  return it;
}
template <bool LOCAL>
__device__ __forceinline__ Item read_item(const int4* info, uint32_t it_n, int window) {
  int4 e;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
               : "r"(smem_u32(info + (it_n & 1)))
               : "memory");
  return make_item<LOCAL>(e, window);
}

// SPLIT: q | k | v, P and the output are pairs of fp16 planes (x = hi + lo); S = Q_hi K_hi + Q_lo K_hi + Q_hi K_lo and
// O += P_hi V_hi + P_lo V_hi + P_hi V_lo (three MMAs per k-step); 113 KB of shared memory -> two CTAs per SM.
template <bool LOCAL, bool SPLIT>
__global__ void __launch_bounds__(ATT_THREADS, att_ctas_per_sm(SPLIT))
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmQlo, const __grid_constant__ CUtensorMap tmKVlo,
                    __half* __restrict__ out, __half* __restrict__ out_lo,
                    const int4* __restrict__ work, int n_pairs, int heads, int hidden, float scale_log2e,
                    int window) {
  constexpr int PL = SPLIT ? 2 : 1;            // planes per tile
  constexpr bool LEAN = att_lean(SPLIT);
  constexpr int KS = att_ks(LEAN), VS = att_vs(LEAN);
  constexpr int SQ_SLOT = PL * SQ_BYTES, SKV_SLOT = PL * SKV_BYTES, SP_SLOT = PL * SP_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if constexpr (LEAN) {   // two CTAs x (113 KB + 1 KB reserved) = the SM's 228 KB: nothing left for the align-up
    if (smem != smem_raw) __trap();
  }
  uint8_t* sQ = smem;                          // 16 KB (per plane)
  uint8_t* sK = sQ + SQ_SLOT;                  // K ring: slot s at sK + s * SKV_SLOT (hi plane, then lo plane)
  uint8_t* sV = sK + KS * SKV_SLOT;            // V ring: slot s at sV + s * SKV_SLOT
  uint8_t* sP = sV + VS * SKV_SLOT;            // 16 KB (per plane)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + SP_SLOT);
  uint64_t* q_full = bars;                   // 1
  uint64_t* q_empty = q_full + 1;            // 1
  uint64_t* k_full = q_empty + 1;            // [KS]
  uint64_t* k_empty = k_full + KS;           // [KS]
  uint64_t* v_full = k_empty + KS;           // [VS]
  uint64_t* v_empty = v_full + VS;           // [VS]
  uint64_t* s_full = v_empty + VS;           // 1
  uint64_t* s_empty = s_full + 1;            // 1
  uint64_t* p_full = s_empty + 1;            // 1
  uint64_t* pv_done = p_full + 1;            // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);
  int4* info = reinterpret_cast<int4*>(bars + 20);  // [2] published items (16-byte aligned: bars is 1 KB aligned)
  const int n_work = n_pairs * heads;
  const uint32_t n_items = blockIdx.x < static_cast<unsigned>(n_work)
                               ? (static_cast<uint32_t>(n_work) - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, SOFT_WARPS);
    for (int i = 0; i < KS; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    mbar_init(p_full, SOFT_WARPS);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    if constexpr (SPLIT) {
      tma_prefetch_desc(&tmQlo);
      tma_prefetch_desc(&tmKVlo);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_o = tmem_base + AK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t g = 0;
    int v_col = 0, v_row = 0;   // coordinates of the V tile whose load is still owed
    int pair = static_cast<int>(blockIdx.x) / heads, head = static_cast<int>(blockIdx.x) % heads;
    const int dq = static_cast<int>(gridDim.x) / heads, dr = static_cast<int>(gridDim.x) % heads;
    int4 entry = pair < n_pairs ? __ldg(work + pair) : make_int4(0, 0, 0, 0);
    for (uint32_t it_n = 0; it_n < n_items; ++it_n) {
      int npair = pair + dq, nhead = head + dr;
      if (nhead >= heads) {
        nhead -= heads;
        ++npair;
      }
      const int4 entry_next = npair < n_pairs ? __ldg(work + npair) : make_int4(0, 0, 0, 0);
      entry.w = head;
      const Item it = make_item<LOCAL>(entry, window);
      mbar_wait_tagged(q_empty, (it_n & 1) ^ 1, 8);  // every S MMA of the previous item has read the Q tile
      if (elect_one()) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(info + (it_n & 1))), "r"(entry.x),
                     "r"(entry.y), "r"(entry.z), "r"(entry.w)
                     : "memory");  // released by the arrive below
        mbar_arrive_expect_tx(q_full, SQ_SLOT);
        tma_load_2d(sQ, &tmQ, q_full, it.head * AD, it.s0 + it.q0);
        if constexpr (SPLIT) tma_load_2d(sQ + SQ_BYTES, &tmQlo, q_full, it.head * AD, it.s0 + it.q0);
      }
      __syncwarp();
      // K runs one block ahead of V in issue order (K_g, V_{g-1}, K_{g+1}, V_g, ...): a full V ring never holds back
      // the K of the next block, which the S MMA needs a whole softmax earlier than its V.
      for (int i = 0; i < it.nb; ++i, ++g) {
        const int ks = g % KS;
        mbar_wait_tagged(k_empty + ks, ((g / KS) & 1) ^ 1, 3);
        const int row = it.s0 + (it.j_lo + i) * AK;
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full + ks, SKV_SLOT);
          tma_load_2d(sK + ks * SKV_SLOT, &tmKV, k_full + ks, hidden + it.head * AD, row);
          if constexpr (SPLIT) tma_load_2d(sK + ks * SKV_SLOT + SKV_BYTES, &tmKVlo, k_full + ks, hidden + it.head * AD, row);
        }
        __syncwarp();
        if (g > 0) {   // V of the previous block of the flat stream
          const uint32_t gv = g - 1;
          const int vs = gv % VS;
          mbar_wait_tagged(v_empty + vs, ((gv / VS) & 1) ^ 1, 10);
          if (elect_one()) {
            mbar_arrive_expect_tx(v_full + vs, SKV_SLOT);
            tma_load_2d(sV + vs * SKV_SLOT, &tmKV, v_full + vs, v_col, v_row);
            if constexpr (SPLIT) tma_load_2d(sV + vs * SKV_SLOT + SKV_BYTES, &tmKVlo, v_full + vs, v_col, v_row);
          }
          __syncwarp();
        }
        v_col = 2 * hidden + it.head * AD;
        v_row = row;
      }
      pair = npair;
      head = nhead;
      entry = entry_next;
    }
    if (g > 0) {   // V of the very last block
      const uint32_t gv = g - 1;
      const int vs = gv % VS;
      mbar_wait_tagged(v_empty + vs, ((gv / VS) & 1) ^ 1, 10);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full + vs, SKV_SLOT);
        tma_load_2d(sV + vs * SKV_SLOT, &tmKV, v_full + vs, v_col, v_row);
        if constexpr (SPLIT) tma_load_2d(sV + vs * SKV_SLOT + SKV_BYTES, &tmKVlo, v_full + vs, v_col, v_row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // One flat stream of key blocks across items: S of the NEXT block (possibly the next item's first) is issued
    // before waiting for the current block's P, so the softmax warps never wait for a QK^T at an item boundary.
    constexpr uint32_t idesc_s = umma_idesc(0, AQ, AK);
    constexpr uint32_t idesc_pv = umma_idesc_major(0, AQ, AD, 0, 1);  // B = V is MN-major ([key][d] rows)
    const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), k_base = smem_u32(sK), v_base = smem_u32(sV);
    auto issue_s = [&](uint32_t G, bool last) {
      const int ks = G % KS;
      mbar_wait_tagged(k_full + ks, (G / KS) & 1, 2);
      mbar_wait_tagged(s_empty, (G & 1) ^ 1, 5);  // the softmax warps hold S_{G-1} in registers
      tc_fence_after();
      const uint32_t k_addr = k_base + ks * SKV_SLOT;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AD / 16; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc_s,
                   k > 0 ? 1u : 0u);
          if constexpr (SPLIT) {
            umma_f16(tmem_base, umma_desc_sw128(q_addr + SQ_BYTES + k * 32), umma_desc_sw128(k_addr + k * 32), idesc_s, 1u);
            umma_f16(tmem_base, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + SKV_BYTES + k * 32), idesc_s, 1u);
          }
        }
        umma_commit(s_full);
        umma_commit(k_empty + ks);                // the K slot is free once this S has completed
        if (last) umma_commit(q_empty);  // the Q tile is free once this item's last S has completed
      }
      __syncwarp();
    };
    uint32_t g = 0;
    int nb = 0;
    if (n_items > 0) {
      mbar_wait_tagged(q_full, 0, 1);
      nb = read_item<LOCAL>(info, 0, window).nb;
      issue_s(0, nb == 1);
    }
    for (uint32_t it_n = 0; it_n < n_items; ++it_n) {
      int nb_next = 0;
      for (int i = 0; i < nb; ++i) {
        const uint32_t G = g + i;
        if (i + 1 < nb) {
          issue_s(G + 1, i + 2 == nb);
        } else if (it_n + 1 < n_items) {
          mbar_wait_tagged(q_full, (it_n + 1) & 1, 1);
          nb_next = read_item<LOCAL>(info, it_n + 1, window).nb;
          issue_s(G + 1, nb_next == 1);
        }
        const int vs = G % VS;
        mbar_wait_tagged(v_full + vs, (G / VS) & 1, 11);
        mbar_wait_tagged(p_full, G & 1, 6);
        tc_fence_after();
        const uint32_t v_addr = v_base + vs * SKV_SLOT;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AK / 16; ++k) {  // 16 keys per step = two 8-row groups of the V tile = 2048 bytes
            umma_f16(tmem_o, umma_desc_sw128(p_addr + k * 32), umma_desc_sw128(v_addr + k * 2048), idesc_pv,
                     (i > 0 || k > 0) ? 1u : 0u);  // first block of the item overwrites O
            if constexpr (SPLIT) {
              umma_f16(tmem_o, umma_desc_sw128(p_addr + SP_BYTES + k * 32), umma_desc_sw128(v_addr + k * 2048), idesc_pv, 1u);
              umma_f16(tmem_o, umma_desc_sw128(p_addr + k * 32), umma_desc_sw128(v_addr + SKV_BYTES + k * 2048), idesc_pv, 1u);
            }
          }
          umma_commit(pv_done);
          umma_commit(v_empty + vs);
        }
        __syncwarp();
      }
      g += nb;
      nb = nb_next;
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int quarter = warp & 3;        // TMEM lane quarter (hardware: warp w may touch lanes 32*(w%4)..+31)
    const int r = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // SWIZZLE_128B: 16-byte chunk index XOR (row & 7); rows are 128-byte aligned, so chunk c of the row lives at
    // (row base | (row & 7) << 4) ^ (c << 4)
    const uint32_t p_row_sw = (smem_u32(sP) + r * 128) | ((r & 7) << 4);
    const uint64_t scale2 = f2_pack(scale_log2e, scale_log2e);
    uint32_t g = 0;

    for (uint32_t it_n = 0; it_n < n_items; ++it_n) {
      mbar_wait_tagged(s_full, g & 1, 4);  // first S of the item: its producer has published the item long before
      const Item it = read_item<LOCAL>(info, it_n, window);
      const int q = it.q0 + r;
      float m_used = -INFINITY;          // reference max of this row (log2 units); -inf until a key is seen
      uint64_t l2 = 0ull;                // row sum relative to m_used, two partial chains

      // keys this query may attend: [k_lo, k_hi]
      const int k_lo = LOCAL ? max(q - window, 0) : 0;
      const int k_hi = LOCAL ? min(q + window, it.L - 1) : it.L - 1;

      for (int i = 0; i < it.nb; ++i) {
        const uint32_t G = g + i;
        const int key0 = (it.j_lo + i) * AK;
        const int e_lo = k_lo - key0, e_hi = k_hi - key0;  // valid local columns
        const bool dead = __all_sync(0xffffffffu, e_hi < 0 || e_lo > AK - 1);
        if (i > 0) mbar_wait_tagged(s_full, G & 1, 4);
        tc_fence_after();
        // P smem was read by PV_{G-1}: pv_done(G-1) is observed before the P stores.  (That wait also keeps pv_done at
        // most one phase ahead of this warp: PV_G cannot be issued until it has arrived on p_full below.)
        if (dead) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
          if (i > 0) mbar_wait_tagged(pv_done, (G - 1) & 1, 7);
          st_p_chunk<0>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<1>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<2>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<3>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<4>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<5>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<6>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<7>(p_row_sw, 0u, 0u, 0u, 0u);
          if constexpr (SPLIT) {
            st_p_chunk<0>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<1>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<2>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<3>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<4>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<5>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<6>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
            st_p_chunk<7>(p_row_sw + SP_BYTES, 0u, 0u, 0u, 0u);
          }
        } else {
          float s[AK];
          {
            uint32_t ta[32], tb[32];
            tmem_ld_32x32b_x32(t_lane, ta);
            tmem_ld_32x32b_x32(t_lane + 32, tb);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              s[e] = __uint_as_float(ta[e]);
              s[32 + e] = __uint_as_float(tb[e]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);  // S_{G+1} = Q K^T may overwrite TMEM S now

          if (!__all_sync(0xffffffffu, e_lo <= 0 && e_hi >= AK - 1)) {  // boundary: mask (warp-uniform branch)
            // valid columns [e_lo, e_hi] as a 64-bit mask: one bit test + select per element
            const int lo = max(e_lo, 0), hi = min(e_hi, AK - 1);
            const uint64_t vm = hi >= lo ? ((~0ull >> (63 - hi)) & (~0ull << lo)) : 0ull;
            const uint32_t vm0 = static_cast<uint32_t>(vm), vm1 = static_cast<uint32_t>(vm >> 32);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              s[e] = (vm0 >> e) & 1u ? s[e] : -INFINITY;
              s[32 + e] = (vm1 >> e) & 1u ? s[32 + e] : -INFINITY;
            }
          }
          float mx8[8];  // independent max chains
#pragma unroll
          for (int e = 0; e < 8; ++e) mx8[e] = fmaxf(s[e], s[e + 8]);
#pragma unroll
          for (int e = 16; e < AK; e += 16)
#pragma unroll
            for (int j = 0; j < 8; ++j) mx8[j] = fmaxf(mx8[j], fmaxf(s[e + j], s[e + j + 8]));
          const float mx = scale_log2e * fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                               fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          // scale > 0: max commutes with the scaling.  First key seen by this row: adopt its max, nothing to rescale
          // (every earlier P of the row was 0, so its O row and sum are 0).
          if (m_used == -INFINITY) m_used = mx;
          if (__any_sync(0xffffffffu, mx > m_used + RESCALE_THRESHOLD)) {
            // rare: raise the reference max of the rows that need it and rescale their O rows in TMEM.  PV_{G-1} must
            // have completed; PV_G cannot start before this warp arrives on p_full.  i > 0 here: on the item's first
            // block every row has m_used == mx or -inf.
            mbar_wait_tagged(pv_done, (G - 1) & 1, 9);
            tc_fence_after();
            const float m_new = fmaxf(m_used, mx);
            const float alpha = ex2(m_used - m_new);  // 1 for rows that keep their reference
            const uint64_t a2 = f2_pack(alpha, alpha);
#pragma unroll
            for (int c = 0; c < AD / 16; ++c) {
              uint32_t t[16];
              tmem_ld_32x32b_x16(t_lane + AK + c * 16, t);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float lo_, hi_;
                f2_unpack(f2_mul(f2_pack_bits(t[2 * e], t[2 * e + 1]), a2), lo_, hi_);
                t[2 * e] = __float_as_uint(lo_);
                t[2 * e + 1] = __float_as_uint(hi_);
              }
              tmem_st_32x32b_x16(t_lane + AK + c * 16, t);
            }
            tmem_st_wait();
            l2 = f2_mul(l2, a2);
            m_used = m_new;
          }
          const float mu = m_used == -INFINITY ? 0.f : m_used;
          const uint64_t nmu2 = f2_pack(-mu, -mu);
          uint64_t sum2 = 0ull, sum2b = 0ull;
          if (i > 0) mbar_wait_tagged(pv_done, (G - 1) & 1, 7);  // (a second wait on a completed phase returns at once)
          auto chunk = [&](auto cc) {  // 8 columns -> one 16-byte chunk of the P row, stored as it is produced
            constexpr int c = decltype(cc)::value;
            float pe[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack(s[c * 8 + 2 * e], s[c * 8 + 2 * e + 1]), scale2, nmu2), x0, x1);
              pe[2 * e] = ex2(x0);      // masked: fma(-inf, .) = -inf -> 0
              pe[2 * e + 1] = ex2(x1);
              if (e & 1) sum2b = f2_add(sum2b, f2_pack(pe[2 * e], pe[2 * e + 1]));
              else sum2 = f2_add(sum2, f2_pack(pe[2 * e], pe[2 * e + 1]));
            }
            if constexpr (SPLIT) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_half2(pe[2 * e], pe[2 * e + 1], h[e], l[e]);
              st_p_chunk<c>(p_row_sw, h[0], h[1], h[2], h[3]);
              st_p_chunk<c>(p_row_sw + SP_BYTES, l[0], l[1], l[2], l[3]);
            } else {
              st_p_chunk<c>(p_row_sw, pack_half2(pe[0], pe[1]), pack_half2(pe[2], pe[3]), pack_half2(pe[4], pe[5]),
                            pack_half2(pe[6], pe[7]));
            }
          };
          chunk(IntC<0>{}); chunk(IntC<1>{}); chunk(IntC<2>{}); chunk(IntC<3>{});
          chunk(IntC<4>{}); chunk(IntC<5>{}); chunk(IntC<6>{}); chunk(IntC<7>{});
          l2 = f2_add(l2, f2_add(sum2, sum2b));
        }
        fence_proxy_async_smem();  // P written with st.shared must be visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      const uint32_t G_last = g + it.nb - 1;
      mbar_wait_tagged(pv_done, G_last & 1, 7);
      tc_fence_after();
      {
        float l_lo, l_hi;
        f2_unpack(l2, l_lo, l_hi);
        const float l = l_lo + l_hi;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const uint64_t inv2 = f2_pack(inv, inv);
        uint32_t ta[32], tb[32];
        tmem_ld_32x32b_x32(t_lane + AK, ta);
        tmem_ld_32x32b_x32(t_lane + AK + 32, tb);
        tmem_ld_wait();
        tc_fence_before();  // the next item's first PV overwrites O only after this warp's next p_full arrival
        if (q < it.L) {
          const size_t o_off = static_cast<size_t>(it.s0 + q) * hidden + it.head * AD;
          uint4* dst = reinterpret_cast<uint4*>(out + o_off);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t* t = c < 4 ? ta + 8 * c : tb + 8 * (c - 4);
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) f2_unpack(f2_mul(f2_pack_bits(t[2 * e], t[2 * e + 1]), inv2), v[2 * e], v[2 * e + 1]);
            uint4 u;
            if constexpr (SPLIT) {
              uint4 l;
              split_half2(v[0], v[1], u.x, l.x);
              split_half2(v[2], v[3], u.y, l.y);
              split_half2(v[4], v[5], u.z, l.z);
              split_half2(v[6], v[7], u.w, l.w);
              reinterpret_cast<uint4*>(out_lo + o_off)[c] = l;
            } else {
              u.x = pack_half2(v[0], v[1]);
              u.y = pack_half2(v[2], v[3]);
              u.z = pack_half2(v[4], v[5]);
              u.w = pack_half2(v[6], v[7]);
            }
            dst[c] = u;
          }
        }
      }
      g += it.nb;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace

namespace {
template <bool SPLIT>
void launch_attention_impl(vrag_ctx* ctx, const __half* qkv, const __half* qkv_lo, __half* out, __half* out_lo,
                           const int32_t* work_dev, int n_pairs, int total_tokens, int heads, int hidden, int window) {
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  ProfScope prof(ctx, PROF_ATTENTION);
  CUtensorMap tmQ = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AQ, AD);
  CUtensorMap tmKV = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AK, AD);
  CUtensorMap tmQlo = tmQ, tmKVlo = tmKV;
  if (SPLIT) {
    tmQlo = make_tmap_2d(ctx, qkv_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AQ, AD);
    tmKVlo = make_tmap_2d(ctx, qkv_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AK, AD);
  }
  constexpr int SMEM = att_smem<SPLIT>();
  static std::atomic<uint64_t> attr_mask{0};   // cudaFuncSetAttribute is per device
  const uint64_t bit = 1ull << (ctx->device & 63);
  if ((attr_mask.fetch_or(bit) & bit) == 0) {
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  const int n_work = n_pairs * heads;
  if (n_work == 0) return;
  const int4* work4 = reinterpret_cast<const int4*>(work_dev);  // {s0, L, q0, -} per (sequence, query tile)
  const int per_sm = att_ctas_per_sm(SPLIT);
  const int grid = n_work < per_sm * ctx->num_sms ? n_work : per_sm * ctx->num_sms;
  if (window >= 0)
    attention_tc_kernel<true, SPLIT><<<grid, ATT_THREADS, SMEM, ctx->stream>>>(tmQ, tmKV, tmQlo, tmKVlo, out, out_lo, work4,
                                                                               n_pairs, heads, hidden, scale_log2e, window);
  else
    attention_tc_kernel<false, SPLIT><<<grid, ATT_THREADS, SMEM, ctx->stream>>>(tmQ, tmKV, tmQlo, tmKVlo, out, out_lo, work4,
                                                                                n_pairs, heads, hidden, scale_log2e, 0);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}
}  // namespace

void launch_attention_tc(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev,
                         const int32_t* work_dev, int n_pairs, int total_tokens, int heads, int hidden,
                         int window /* <0: full */) {
  (void)cu_seqlens_dev;
  launch_attention_impl<false>(ctx, qkv, nullptr, out, nullptr, work_dev, n_pairs, total_tokens, heads, hidden, window);
}

void launch_attention_tc_split(vrag_ctx* ctx, const __half* qkv, const __half* qkv_lo, __half* out, __half* out_lo,
                               const int32_t* work_dev, int n_pairs, int total_tokens, int heads, int hidden,
                               int window) {
  VRAG_CHECK(qkv_lo && out_lo, VRAG_ERR_ARG, "attention: split precision needs the low planes");
  launch_attention_impl<true>(ctx, qkv, qkv_lo, out, out_lo, work_dev, n_pairs, total_tokens, heads, hidden, window);
}

}  // namespace vrag
