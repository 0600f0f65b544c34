// tcgen05 self-attention over unpadded sequences.  PERSISTENT: 2 CTAs per SM, each loops over work items
// (sequence, 128-query tile, head); TMEM / barriers are set up once per CTA and the TMA producer runs ahead into the
// next item's Q / K / V while the current item is still in its softmax.
//
//   warp 0      : TMA producer  (Q tile; K/V blocks of 64 keys in a 3-stage ring, SWIZZLE_128B boxes)
//   warp 1      : MMA issuer    S = Q K^T   (tcgen05.mma 128x64x16, both operands K-major)   -> TMEM S[2]
//                               PV = P V    (tcgen05.mma 128x64x16, A = P from smem[2], B = V MN-major) -> TMEM Otmp[2]
//   warps 2..5  : softmax       one warp per TMEM lane quarter, thread == query row (all 64 columns of the block, so
//                               the row max / sum never leave the thread): tcgen05.ld S, mask, online max / sum in
//                               base 2 with packed fp32x2 arithmetic (FFMA2 / FADD2), P (fp16) -> swizzled smem as
//                               it is produced, previous block's PV folded into the O registers while the tensor
//                               core runs PV_i
// Sizing (measured, profiles/README.md): the softmax warps are bound by issue slots + dependent-latency hops per key
// block, not by MUFU or the tensor pipe; the half-row-per-thread version (8 softmax warps, row max exchanged through
// smem + named barrier) spent 2.5x the instructions per score.
// All ring / buffer indices and mbarrier parities derive from per-role running counters (item count `it`, key-block
// count `g`), which every role advances identically.  q/k already carry RoPE (Wqkv GEMM epilogue).  On local layers
// only key blocks intersecting |i - j| <= window are visited; blocks fully outside a warp's window skip TMEM.
// Producer / MMA loops are warp-convergent with elect.sync around the TMA / MMA instructions (uniform datapath).
#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int AQ = 128, AK = 64, AD = 64;
constexpr int SOFT_WARPS = 4;
constexpr int KVS = 3;                   // K/V ring depth
constexpr int ATT_THREADS = 32 * (2 + SOFT_WARPS);
constexpr uint32_t ATT_TMEM_COLS = 256;  // S0 [0,64)  S1 [64,128)  Otmp0 [128,192)  Otmp1 [192,256)
constexpr int SQ_BYTES = AQ * AD * 2;    // 16384
constexpr int SKV_BYTES = AK * AD * 2;   // 8192
constexpr int SP_BYTES = AQ * AK * 2;    // 16384
constexpr int ATT_SMEM = SQ_BYTES + KVS * 2 * SKV_BYTES + 2 * SP_BYTES + 1024 + 256;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Item {  // one (sequence, 128-query tile, head)
  int s0, L, q0, head, j_lo, nb;
};

template <bool LOCAL>
__device__ __forceinline__ Item decode_item(int w, int heads, const int32_t* __restrict__ work,
                                            const int32_t* __restrict__ cu_seqlens, int window) {
  Item it;
  const int pair = w / heads;
  it.head = w - pair * heads;
  const int seq = __ldg(work + 2 * pair);
  it.q0 = __ldg(work + 2 * pair + 1);
  it.s0 = __ldg(cu_seqlens + seq);
  it.L = __ldg(cu_seqlens + seq + 1) - it.s0;
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
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    __half* __restrict__ out, const int32_t* __restrict__ cu_seqlens,
                    const int32_t* __restrict__ work, int n_work, int heads, int hidden, float scale_log2e,
                    int window) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 16 KB
  uint8_t* sKV = sQ + SQ_BYTES;                // slot s: K at sKV + s*16384, V at +8192
  uint8_t* sP = sKV + KVS * 2 * SKV_BYTES;     // [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * SP_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* q_empty = q_full + 1;            // 1
  uint64_t* kv_full = q_empty + 1;           // [KVS]
  uint64_t* kv_empty = kv_full + KVS;        // [KVS]
  uint64_t* s_full = kv_empty + KVS;         // [2]
  uint64_t* s_empty = s_full + 2;            // [2]
  uint64_t* p_full = s_empty + 2;            // 1
  uint64_t* pv_done = p_full + 1;            // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_empty + i, SOFT_WARPS);
    }
    for (int i = 0; i < KVS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    mbar_init(p_full, SOFT_WARPS);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it_n = 0, g = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it_n) {
      const Item it = decode_item<LOCAL>(w, heads, work, cu_seqlens, window);
      mbar_wait_tagged(q_empty, (it_n & 1) ^ 1, 8);  // every S MMA of the previous item has read the Q tile
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, SQ_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, it.head * AD, it.s0 + it.q0);
      }
      __syncwarp();
      for (int i = 0; i < it.nb; ++i, ++g) {
        const int st = g % KVS;
        mbar_wait_tagged(kv_empty + st, ((g / KVS) & 1) ^ 1, 3);
        const int row = it.s0 + (it.j_lo + i) * AK;
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full + st, 2 * SKV_BYTES);
          tma_load_2d(sKV + st * 2 * SKV_BYTES, &tmKV, kv_full + st, hidden + it.head * AD, row);
          tma_load_2d(sKV + st * 2 * SKV_BYTES + SKV_BYTES, &tmKV, kv_full + st, 2 * hidden + it.head * AD, row);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc(0, AQ, AK);
    constexpr uint32_t idesc_pv = umma_idesc_major(0, AQ, AD, 0, 1);  // B = V is MN-major ([key][d] rows)
    const uint32_t q_addr = smem_u32(sQ), p_base = smem_u32(sP), kv_addr = smem_u32(sKV);
    uint32_t it_n = 0, g = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it_n) {
      const Item it = decode_item<LOCAL>(w, heads, work, cu_seqlens, window);
      mbar_wait_tagged(q_full, it_n & 1, 1);
      auto issue_s = [&](uint32_t G, bool last) {
        const int st = G % KVS, sb = G & 1;  // K/V ring slot, S buffer
        mbar_wait_tagged(kv_full + st, (G / KVS) & 1, 2);
        mbar_wait_tagged(s_empty + sb, ((G >> 1) & 1) ^ 1, 5);
        tc_fence_after();
        const uint32_t k_addr = kv_addr + st * 2 * SKV_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AD / 16; ++k)
            umma_f16(tmem_base + sb * AK, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32),
                     idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full + sb);
          if (last) umma_commit(q_empty);  // the Q tile is free once this item's last S has completed
        }
        __syncwarp();
      };
      issue_s(g, it.nb == 1);
      for (int i = 0; i < it.nb; ++i) {
        const uint32_t G = g + i;
        if (i + 1 < it.nb) issue_s(G + 1, i + 2 == it.nb);
        mbar_wait_tagged(p_full, G & 1, 6);
        tc_fence_after();
        const int st = G % KVS;
        const uint32_t v_addr = kv_addr + st * 2 * SKV_BYTES + SKV_BYTES;
        const uint32_t p_addr = p_base + (G & 1) * SP_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AK / 16; ++k)  // 16 keys per step = two 8-row groups of the V tile = 2048 bytes
            umma_f16(tmem_base + 2 * AK + (G & 1) * AD, umma_desc_sw128(p_addr + k * 32),
                     umma_desc_sw128(v_addr + k * 2048), idesc_pv, k > 0 ? 1u : 0u);
          umma_commit(pv_done);
          umma_commit(kv_empty + st);
        }
        __syncwarp();
      }
      g += it.nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int quarter = warp & 3;        // TMEM lane quarter (hardware: warp w may touch lanes 32*(w%4)..+31)
    const int r = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t p_row = smem_u32(sP) + r * 128;
    const int sw = r & 7;                // SWIZZLE_128B: 16-byte chunk index XOR (row & 7)
    const uint64_t scale2 = f2_pack(scale_log2e, scale_log2e);
    uint32_t g = 0;

    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const Item it = decode_item<LOCAL>(w, heads, work, cu_seqlens, window);
      const int q = it.q0 + r;
      uint64_t o2[AD / 2];               // un-normalised output row, packed fp32 pairs
#pragma unroll
      for (int d = 0; d < AD / 2; ++d) o2[d] = 0ull;
      float m = -INFINITY, alpha_prev = 0.f;
      uint64_t l2 = 0ull;                // row sum, two partial chains

      // O = O * alpha + PV_G   (PV_G lives in TMEM Otmp[G & 1])
      auto fold = [&](uint32_t G, float alpha) {
        const uint64_t a2 = f2_pack(alpha, alpha);
        uint32_t ta[32], tb[32];
        tmem_ld_32x32b_x32(t_lane + 2 * AK + (G & 1) * AD, ta);
        tmem_ld_32x32b_x32(t_lane + 2 * AK + (G & 1) * AD + 32, tb);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) o2[e] = f2_fma(o2[e], a2, f2_pack_bits(ta[2 * e], ta[2 * e + 1]));
#pragma unroll
        for (int e = 0; e < 16; ++e) o2[16 + e] = f2_fma(o2[16 + e], a2, f2_pack_bits(tb[2 * e], tb[2 * e + 1]));
      };

      // keys this query may attend: [k_lo, k_hi]
      const int k_lo = LOCAL ? max(q - window, 0) : 0;
      const int k_hi = LOCAL ? min(q + window, it.L - 1) : it.L - 1;

      for (int i = 0; i < it.nb; ++i) {
        const uint32_t G = g + i;
        const int sb = G & 1;
        const int key0 = (it.j_lo + i) * AK;
        const int e_lo = k_lo - key0, e_hi = k_hi - key0;  // valid local columns
        const bool dead = __all_sync(0xffffffffu, e_hi < 0 || e_lo > AK - 1);
        const uint32_t p_dst = p_row + sb * SP_BYTES;
        mbar_wait_tagged(s_full + sb, (G >> 1) & 1, 4);
        tc_fence_after();
        float alpha = 1.f, m_new = m;
        uint64_t sum2 = 0ull;
        // sP[sb] was last read by PV_{G-2}, whose completion this thread observed before fold(G-2)
        if (dead) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty + sb);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(p_dst + ((c ^ sw) << 4)), "r"(0u) : "memory");
        } else {
          float s[AK];
          {
            uint32_t ta[32], tb[32];
            tmem_ld_32x32b_x32(t_lane + sb * AK, ta);
            tmem_ld_32x32b_x32(t_lane + sb * AK + 32, tb);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              s[e] = __uint_as_float(ta[e]);
              s[32 + e] = __uint_as_float(tb[e]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty + sb);

          if (!__all_sync(0xffffffffu, e_lo <= 0 && e_hi >= AK - 1)) {  // boundary: mask (warp-uniform branch)
#pragma unroll
            for (int e = 0; e < AK; ++e) s[e] = (e >= e_lo && e <= e_hi) ? s[e] : -INFINITY;
          }
          float mx8[8];  // independent max chains
#pragma unroll
          for (int e = 0; e < 8; ++e) mx8[e] = fmaxf(s[e], s[e + 8]);
#pragma unroll
          for (int e = 16; e < AK; e += 16)
#pragma unroll
            for (int j = 0; j < 8; ++j) mx8[j] = fmaxf(mx8[j], fmaxf(s[e + j], s[e + j + 8]));
          const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                 fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          m_new = fmaxf(m, mx * scale_log2e);  // scale > 0: max commutes with the scaling
          const float mu = m_new == -INFINITY ? 0.f : m_new;
          alpha = ex2(m - mu);  // first block: ex2(-inf) = 0
          const uint64_t nmu2 = f2_pack(-mu, -mu);
          uint64_t sum2b = 0ull;
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // 8 columns -> one 16-byte chunk of the P row
            float p[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack(s[c * 8 + 2 * e], s[c * 8 + 2 * e + 1]), scale2, nmu2), x0, x1);
              p[2 * e] = ex2(x0);      // masked: fma(-inf, .) = -inf -> 0
              p[2 * e + 1] = ex2(x1);
              if (e & 1) sum2b = f2_add(sum2b, f2_pack(p[2 * e], p[2 * e + 1]));
              else sum2 = f2_add(sum2, f2_pack(p[2 * e], p[2 * e + 1]));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_dst + ((c ^ sw) << 4)),
                         "r"(pack_half2(p[0], p[1])), "r"(pack_half2(p[2], p[3])), "r"(pack_half2(p[4], p[5])),
                         "r"(pack_half2(p[6], p[7]))
                         : "memory");
          }
          sum2 = f2_add(sum2, sum2b);
        }
        // Observe PV_{G-1} BEFORE releasing P_G: pv_done is a single mbarrier completing once per key block, and a
        // parity wait is only unambiguous while the barrier is at most one phase ahead of the waiter -- PV_G cannot be
        // issued (hence cannot complete) until this warp has arrived on p_full below.
        if (i > 0) mbar_wait_tagged(pv_done, (G - 1) & 1, 7);
        fence_proxy_async_smem();  // P written with st.shared must be visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // off the critical path: fold the previous block's PV while the tensor core runs PV_G
        if (i > 0) {
          tc_fence_after();
          fold(G - 1, alpha_prev);
        }
        l2 = f2_fma(l2, f2_pack(alpha, alpha), sum2);
        m = m_new;
        alpha_prev = alpha;
      }
      const uint32_t G_last = g + it.nb - 1;
      mbar_wait_tagged(pv_done, G_last & 1, 7);
      tc_fence_after();
      fold(G_last, alpha_prev);
      tc_fence_before();
      if (q < it.L) {
        float l_lo, l_hi;
        f2_unpack(l2, l_lo, l_hi);
        const float l = l_lo + l_hi;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const uint64_t inv2 = f2_pack(inv, inv);
        uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(it.s0 + q) * hidden + it.head * AD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) f2_unpack(f2_mul(o2[c * 4 + e], inv2), v[2 * e], v[2 * e + 1]);
          uint4 u;
          u.x = pack_half2(v[0], v[1]);
          u.y = pack_half2(v[2], v[3]);
          u.z = pack_half2(v[4], v[5]);
          u.w = pack_half2(v[6], v[7]);
          dst[c] = u;
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

void launch_attention_tc(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev,
                         const int32_t* work_dev, int n_pairs, int total_tokens, int heads, int hidden,
                         int window /* <0: full */) {
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  ProfScope prof(ctx, PROF_ATTENTION);
  CUtensorMap tmQ = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AQ, AD);
  CUtensorMap tmKV = make_tmap_2d(ctx, qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, total_tokens, 3 * hidden, 3 * hidden, AK, AD);
  static bool attr_set = false;
  if (!attr_set) {
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    VRAG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  const int n_work = n_pairs * heads;
  if (n_work == 0) return;
  const int grid = n_work < 2 * ctx->num_sms ? n_work : 2 * ctx->num_sms;
  if (window >= 0)
    attention_tc_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tmQ, tmKV, out, cu_seqlens_dev, work_dev,
                                                                            n_work, heads, hidden, scale_log2e, window);
  else
    attention_tc_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tmQ, tmKV, out, cu_seqlens_dev, work_dev,
                                                                             n_work, heads, hidden, scale_log2e, 0);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace vrag
