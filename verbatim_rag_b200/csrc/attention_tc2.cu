// EXPERIMENTAL (opt-in: VRAG_ATTENTION_V2=1; NOT on the default path, NOT yet validated on a GPU -- written at the end
// of round 1 after the GPU budget was spent; first task of round 2: tests/test_gpu_kernels.py::
// test_attention_kernel_vs_float64 with the switch on, then tools/attn_probe.py).
//
// Two query tiles per CTA.  attention_tc.cu runs 3 CTAs per SM with one 128-query tile each: 12 softmax warps per SM,
// MUFU pipe 46 % busy, the softmax warps spend two thirds of their time in dependent waits, and
// every tile streams its own copy of K / V through shared memory.  Here a CTA owns TWO consecutive 128-query tiles of
// one (sequence, head): the K / V ring is shared by both, each tile has its own S / O accumulators in tensor memory
// (256 of 512 columns per CTA) and its own four softmax warps -> 2 CTAs per SM = 4 tiles and 16 softmax warps per SM
// with the same 96-register budget, and half the K / V traffic per score.
//
//   warp 0      : TMA producer  (walks the (sequence, 128-query tile) work list, takes the entries with q0 % 256 == 0 as
//                               items, publishes them through smem; Q of both tiles; K ring 3 deep, V ring 2 deep)
//   warp 1      : MMA issuer    per key block: S_t = Q_t K^T for every tile that needs the block (one block ahead of the
//                               P V products), then O_t += P_t V as the tiles' softmax warps deliver P_t
//   warps 2..9  : softmax       tile = (warp - 2) / 4, TMEM lane quarter = warp % 4, thread == query row; same
//                               arithmetic as attention_tc.cu (lazy rescaling, packed fp32x2, P fp16 -> swizzled smem)
// Per-tile barriers (s_full / s_empty / p_full / pv_done) run on per-tile block counters; a tile takes part in a key
// block only if the block intersects its rows' window (local layers) and the tile exists (ragged tails).  Every softmax
// warp also arrives on q_empty after reading an item, so an item slot is never republished while a warp of an absent
// tile still has to read it, and no barrier can run more than one phase ahead of a waiter.  The stream ends with a
// published sentinel item (L == 0).
#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int AQ = 128, AK = 64, AD = 64;
constexpr int TILES = 2;
constexpr int SOFT_WARPS = 4 * TILES;
constexpr int KS = 3, VS = 2;
constexpr int ATT2_CTAS_PER_SM = 2;
constexpr int ATT2_THREADS = 32 * (2 + SOFT_WARPS);   // 320
constexpr uint32_t ATT2_TMEM_COLS = 128 * TILES;      // tile t: S [128t, 128t+64)  O [128t+64, 128t+128)
constexpr int SQ_BYTES = AQ * AD * 2;                 // 16384
constexpr int SKV_BYTES = AK * AD * 2;                // 8192
constexpr int SP_BYTES = AQ * AK * 2;                 // 16384
constexpr int ATT2_SMEM = TILES * SQ_BYTES + (KS + VS) * SKV_BYTES + TILES * SP_BYTES + 1024 + 512;
constexpr float RESCALE_THRESHOLD = 8.f;              // log2 units

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

// One item = (sequence, 256-query block, head): up to two 128-query tiles.
struct Item2 {
  int s0, L, q0, head;   // q0: first query of tile 0 (multiple of 256); L == 0: end of the stream
  int nt;                // tiles present (1 or 2)
  int j_lo, nb;          // union of the key blocks the tiles need: blocks [j_lo, j_lo + nb)
  int t_lo[TILES], t_hi[TILES];   // tile t needs key blocks [t_lo, t_hi)
  __device__ __forceinline__ bool takes(int t, int jb) const { return t < nt && jb >= t_lo[t] && jb < t_hi[t]; }
};

template <bool LOCAL>
__device__ __forceinline__ Item2 make_item2(int4 e, int window) {
  Item2 it;
  it.s0 = e.x;
  it.L = e.y;
  it.q0 = e.z;
  it.head = e.w;
  it.nt = (it.q0 + AQ < it.L) ? 2 : 1;
  int lo = 1 << 30, hi = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) {
    const int qa = it.q0 + t * AQ;
    int kv_lo = 0, kv_hi = it.L;
    if (LOCAL) {
      kv_lo = max(0, qa - window);
      kv_hi = min(it.L, qa + AQ + window);
    }
    it.t_lo[t] = kv_lo / AK;
    it.t_hi[t] = (kv_hi + AK - 1) / AK;
    if (t < it.nt) {
      lo = min(lo, it.t_lo[t]);
      hi = max(hi, it.t_hi[t]);
    }
  }
  it.j_lo = lo;
  it.nb = hi - lo;
  return it;
}
template <bool LOCAL>
__device__ __forceinline__ Item2 read_item2(const int4* info, uint32_t it_n, int window) {
  int4 e;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
               : "r"(smem_u32(info + (it_n & 1)))
               : "memory");
  return make_item2<LOCAL>(e, window);
}

template <bool LOCAL>
__global__ void __launch_bounds__(ATT2_THREADS, ATT2_CTAS_PER_SM)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     __half* __restrict__ out, const int4* __restrict__ work, int n_pairs, int heads, int hidden,
                     float scale_log2e, int window) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [TILES] 16 KB
  uint8_t* sK = sQ + TILES * SQ_BYTES;             // K ring
  uint8_t* sV = sK + KS * SKV_BYTES;               // V ring
  uint8_t* sP = sV + VS * SKV_BYTES;               // [TILES] 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TILES * SP_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* q_empty = q_full + 1;            // 1   (MMA commit + one arrival per softmax warp)
  uint64_t* k_full = q_empty + 1;            // [KS]
  uint64_t* k_empty = k_full + KS;           // [KS]
  uint64_t* v_full = k_empty + KS;           // [VS]
  uint64_t* v_empty = v_full + VS;           // [VS]
  uint64_t* s_full = v_empty + VS;           // [TILES]
  uint64_t* s_empty = s_full + TILES;        // [TILES]
  uint64_t* p_full = s_empty + TILES;        // [TILES]
  uint64_t* pv_done = p_full + TILES;        // [TILES]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + TILES);
  int4* info = reinterpret_cast<int4*>(bars + 32);  // [2] published items (bars is 1 KB aligned)
  const int n_work = n_pairs * heads;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1 + SOFT_WARPS);
    for (int i = 0; i < KS; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    for (int t = 0; t < TILES; ++t) {
      mbar_init(s_full + t, 1);
      mbar_init(s_empty + t, 4);
      mbar_init(p_full + t, 4);
      mbar_init(pv_done + t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, ATT2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it_n = 0;   // items published so far
    uint32_t g = 0;      // key blocks loaded so far (K ring index; V runs one block behind)
    int v_col = 0, v_row = 0;
    for (int u = static_cast<int>(blockIdx.x); u < n_work; u += static_cast<int>(gridDim.x)) {
      const int pair = u / heads, head = u % heads;
      int4 entry = __ldg(work + pair);
      if ((entry.z & (2 * AQ - 1)) != 0) continue;   // the second tile of an item: covered by its head entry
      entry.w = head;
      const Item2 it = make_item2<LOCAL>(entry, window);
      mbar_wait_tagged(q_empty, (it_n & 1) ^ 1, 8);   // previous item: every S has read Q, every softmax warp the slot
      if (elect_one()) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(info + (it_n & 1))), "r"(entry.x),
                     "r"(entry.y), "r"(entry.z), "r"(entry.w)
                     : "memory");  // released by the arrive below
        mbar_arrive_expect_tx(q_full, TILES * SQ_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, it.head * AD, it.s0 + it.q0);
        tma_load_2d(sQ + SQ_BYTES, &tmQ, q_full, it.head * AD, it.s0 + it.q0 + AQ);   // rows past the tensor: zero fill
      }
      __syncwarp();
      ++it_n;
      for (int i = 0; i < it.nb; ++i, ++g) {
        const int ks = g % KS;
        mbar_wait_tagged(k_empty + ks, ((g / KS) & 1) ^ 1, 3);
        const int row = it.s0 + (it.j_lo + i) * AK;
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full + ks, SKV_BYTES);
          tma_load_2d(sK + ks * SKV_BYTES, &tmKV, k_full + ks, hidden + it.head * AD, row);
        }
        __syncwarp();
        if (g > 0) {   // V of the previous block of the flat stream
          const uint32_t gv = g - 1;
          const int vs = gv % VS;
          mbar_wait_tagged(v_empty + vs, ((gv / VS) & 1) ^ 1, 10);
          if (elect_one()) {
            mbar_arrive_expect_tx(v_full + vs, SKV_BYTES);
            tma_load_2d(sV + vs * SKV_BYTES, &tmKV, v_full + vs, v_col, v_row);
          }
          __syncwarp();
        }
        v_col = 2 * hidden + it.head * AD;
        v_row = row;
      }
    }
    if (g > 0) {   // V of the very last block
      const uint32_t gv = g - 1;
      const int vs = gv % VS;
      mbar_wait_tagged(v_empty + vs, ((gv / VS) & 1) ^ 1, 10);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full + vs, SKV_BYTES);
        tma_load_2d(sV + vs * SKV_BYTES, &tmKV, v_full + vs, v_col, v_row);
      }
      __syncwarp();
    }
    // sentinel: L == 0 ends the consumers' loops
    mbar_wait_tagged(q_empty, (it_n & 1) ^ 1, 8);
    if (elect_one()) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(info + (it_n & 1))), "r"(0) : "memory");
      mbar_arrive(q_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc(0, AQ, AK);
    constexpr uint32_t idesc_pv = umma_idesc_major(0, AQ, AD, 0, 1);  // B = V is MN-major ([key][d] rows)
    const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), k_base = smem_u32(sK), v_base = smem_u32(sV);
    uint32_t g = 0;                 // key blocks of the flat stream (K / V ring position)
    uint32_t gs[TILES] = {0, 0};    // S products issued per tile
    uint32_t gp[TILES] = {0, 0};    // P V products issued per tile
    for (uint32_t it_n = 0;; ++it_n) {
      mbar_wait_tagged(q_full, it_n & 1, 1);
      const Item2 it = read_item2<LOCAL>(info, it_n, window);
      if (it.L == 0) break;
      auto issue_s = [&](int i) {   // S_t = Q_t K_i^T for every tile that needs block i
        const uint32_t G = g + i;
        const int ks = G % KS;
        const int jb = it.j_lo + i;
        mbar_wait_tagged(k_full + ks, (G / KS) & 1, 2);
        const uint32_t k_addr = k_base + ks * SKV_BYTES;
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          if (!it.takes(t, jb)) continue;   // warp-uniform
          mbar_wait_tagged(s_empty + t, (gs[t] & 1) ^ 1, 5);   // the tile's softmax warps hold its previous S in registers
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AD / 16; ++k)
              umma_f16(tmem_base + t * 128, umma_desc_sw128(q_addr + t * SQ_BYTES + k * 32),
                       umma_desc_sw128(k_addr + k * 32), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(s_full + t);
          }
          __syncwarp();
          ++gs[t];
        }
        if (elect_one()) {
          umma_commit(k_empty + ks);                 // the K slot is free once these S products have completed
          if (i + 1 == it.nb) umma_commit(q_empty);  // ... and the Q tiles once the item's last ones have
        }
        __syncwarp();
      };
      issue_s(0);
      for (int i = 0; i < it.nb; ++i) {
        if (i + 1 < it.nb) issue_s(i + 1);
        const uint32_t G = g + i;
        const int vs = G % VS;
        const int jb = it.j_lo + i;
        mbar_wait_tagged(v_full + vs, (G / VS) & 1, 11);
        const uint32_t v_addr = v_base + vs * SKV_BYTES;
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          if (!it.takes(t, jb)) continue;
          mbar_wait_tagged(p_full + t, gp[t] & 1, 6);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AK / 16; ++k)   // first block of the tile in this item overwrites O
              umma_f16(tmem_base + t * 128 + AK, umma_desc_sw128(p_addr + t * SP_BYTES + k * 32),
                       umma_desc_sw128(v_addr + k * 2048), idesc_pv, (jb > it.t_lo[t] || k > 0) ? 1u : 0u);
            umma_commit(pv_done + t);
          }
          __syncwarp();
          ++gp[t];
        }
        if (elect_one()) umma_commit(v_empty + vs);
        __syncwarp();
      }
      g += it.nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int tile = (warp - 2) >> 2;    // which of the CTA's query tiles
    const int quarter = warp & 3;        // TMEM lane quarter (hardware: warp w may touch lanes 32*(w%4)..+31)
    const int r = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + tile * 128;
    const uint32_t p_row_sw = (smem_u32(sP + tile * SP_BYTES) + r * 128) | ((r & 7) << 4);
    const uint64_t scale2 = f2_pack(scale_log2e, scale_log2e);
    uint64_t* my_s_full = s_full + tile;
    uint64_t* my_s_empty = s_empty + tile;
    uint64_t* my_p_full = p_full + tile;
    uint64_t* my_pv_done = pv_done + tile;
    uint32_t kt = 0;   // key blocks this tile has taken part in (flat across items)

    for (uint32_t it_n = 0;; ++it_n) {
      mbar_wait_tagged(q_full, it_n & 1, 1);   // the item is published (and its Q tiles have landed)
      const Item2 it = read_item2<LOCAL>(info, it_n, window);
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty);     // this warp no longer needs the slot
      if (it.L == 0) break;
      if (tile >= it.nt) continue;             // ragged tail: this tile does not exist in the item
      const int q = it.q0 + tile * AQ + r;
      float m_used = -INFINITY;
      uint64_t l2 = 0ull;
      const int k_lo = LOCAL ? max(q - window, 0) : 0;
      const int k_hi = LOCAL ? min(q + window, it.L - 1) : it.L - 1;
      const int my_lo = tile == 0 ? it.t_lo[0] : it.t_lo[1];   // (selects, not a runtime index: keeps Item2 in registers)
      const int my_hi = tile == 0 ? it.t_hi[0] : it.t_hi[1];
      const int nbt = my_hi - my_lo;                            // key blocks of this tile

      for (int i = 0; i < nbt; ++i) {
        const uint32_t G = kt + i;
        const int key0 = (my_lo + i) * AK;
        const int e_lo = k_lo - key0, e_hi = k_hi - key0;  // valid local columns
        const bool dead = __all_sync(0xffffffffu, e_hi < 0 || e_lo > AK - 1);
        mbar_wait_tagged(my_s_full, G & 1, 4);
        tc_fence_after();
        if (dead) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(my_s_empty);
          if (i > 0) mbar_wait_tagged(my_pv_done, (G - 1) & 1, 7);
          st_p_chunk<0>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<1>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<2>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<3>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<4>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<5>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<6>(p_row_sw, 0u, 0u, 0u, 0u);
          st_p_chunk<7>(p_row_sw, 0u, 0u, 0u, 0u);
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
          if (lane == 0) mbar_arrive(my_s_empty);  // the tile's next S may overwrite TMEM S now

          if (!__all_sync(0xffffffffu, e_lo <= 0 && e_hi >= AK - 1)) {  // boundary: mask (warp-uniform branch)
            const int lo = max(e_lo, 0), hi = min(e_hi, AK - 1);
            const uint64_t vm = hi >= lo ? ((~0ull >> (63 - hi)) & (~0ull << lo)) : 0ull;
            const uint32_t vm0 = static_cast<uint32_t>(vm), vm1 = static_cast<uint32_t>(vm >> 32);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              s[e] = (vm0 >> e) & 1u ? s[e] : -INFINITY;
              s[32 + e] = (vm1 >> e) & 1u ? s[32 + e] : -INFINITY;
            }
          }
          float mx8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) mx8[e] = fmaxf(s[e], s[e + 8]);
#pragma unroll
          for (int e = 16; e < AK; e += 16)
#pragma unroll
            for (int j = 0; j < 8; ++j) mx8[j] = fmaxf(mx8[j], fmaxf(s[e + j], s[e + j + 8]));
          const float mx = scale_log2e * fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                               fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          if (m_used == -INFINITY) m_used = mx;
          if (__any_sync(0xffffffffu, mx > m_used + RESCALE_THRESHOLD)) {
            // rare: raise the reference max and rescale this warp's O rows in TMEM (i > 0 here, as in attention_tc.cu)
            mbar_wait_tagged(my_pv_done, (G - 1) & 1, 9);
            tc_fence_after();
            const float m_new = fmaxf(m_used, mx);
            const float alpha = ex2(m_used - m_new);
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
          if (i > 0) mbar_wait_tagged(my_pv_done, (G - 1) & 1, 7);   // P smem was read by the tile's previous P V
          auto chunk = [&](auto cc) {
            constexpr int c = decltype(cc)::value;
            float pe[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack(s[c * 8 + 2 * e], s[c * 8 + 2 * e + 1]), scale2, nmu2), x0, x1);
              pe[2 * e] = ex2(x0);
              pe[2 * e + 1] = ex2(x1);
              if (e & 1) sum2b = f2_add(sum2b, f2_pack(pe[2 * e], pe[2 * e + 1]));
              else sum2 = f2_add(sum2, f2_pack(pe[2 * e], pe[2 * e + 1]));
            }
            st_p_chunk<c>(p_row_sw, pack_half2(pe[0], pe[1]), pack_half2(pe[2], pe[3]), pack_half2(pe[4], pe[5]),
                          pack_half2(pe[6], pe[7]));
          };
          chunk(IntC<0>{}); chunk(IntC<1>{}); chunk(IntC<2>{}); chunk(IntC<3>{});
          chunk(IntC<4>{}); chunk(IntC<5>{}); chunk(IntC<6>{}); chunk(IntC<7>{});
          l2 = f2_add(l2, f2_add(sum2, sum2b));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(my_p_full);
      }
      const uint32_t G_last = kt + nbt - 1;
      mbar_wait_tagged(my_pv_done, G_last & 1, 7);
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
        tc_fence_before();
        if (q < it.L) {
          uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(it.s0 + q) * hidden + it.head * AD);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t* t = c < 4 ? ta + 8 * c : tb + 8 * (c - 4);
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) f2_unpack(f2_mul(f2_pack_bits(t[2 * e], t[2 * e + 1]), inv2), v[2 * e], v[2 * e + 1]);
            uint4 u;
            u.x = pack_half2(v[0], v[1]);
            u.y = pack_half2(v[2], v[3]);
            u.z = pack_half2(v[4], v[5]);
            u.w = pack_half2(v[6], v[7]);
            dst[c] = u;
          }
        }
      }
      kt += nbt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT2_TMEM_COLS);
  }
}

}  // namespace

void launch_attention_tc2(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* work_dev, int n_pairs,
                          int total_tokens, int heads, int hidden, int window /* <0: full */) {
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  ProfScope prof(ctx, PROF_ATTENTION);
  CUtensorMap tmQ = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AQ, AD);
  CUtensorMap tmKV = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AK, AD);
  static bool attr_set = false;
  if (!attr_set) {
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
    attr_set = true;
  }
  const int n_work = n_pairs * heads;
  if (n_work == 0) return;
  const int4* work4 = reinterpret_cast<const int4*>(work_dev);  // {s0, L, q0, -} per (sequence, 128-query tile)
  const int grid = n_work < ATT2_CTAS_PER_SM * ctx->num_sms ? n_work : ATT2_CTAS_PER_SM * ctx->num_sms;
  if (window >= 0)
    attention_tc2_kernel<true><<<grid, ATT2_THREADS, ATT2_SMEM, ctx->stream>>>(tmQ, tmKV, out, work4, n_pairs, heads,
                                                                               hidden, scale_log2e, window);
  else
    attention_tc2_kernel<false><<<grid, ATT2_THREADS, ATT2_SMEM, ctx->stream>>>(tmQ, tmKV, out, work4, n_pairs, heads,
                                                                                hidden, scale_log2e, 0);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace vrag
