// tcgen05 self-attention over unpadded sequences: one (128-query tile, head, sequence) per CTA, 2 CTAs per SM.
//
//   warp 0      : TMA producer  (Q tile once; K/V blocks of 64 keys into a 4-stage ring, SWIZZLE_128B boxes)
//   warp 1      : MMA issuer    S = Q K^T   (tcgen05.mma 128x64x16, both operands K-major)   -> TMEM S[2]
//                               PV = P V    (tcgen05.mma 128x64x16, A = P from smem, B = V MN-major) -> TMEM Otmp[2]
//   warps 2..9  : softmax       two warps per TMEM lane quarter; thread == (query row, half of the 64 columns):
//                               tcgen05.ld S, scale + mask, online max / sum in base 2 (row max exchanged between the
//                               two halves through smem + a 64-thread named barrier), P (fp16) -> swizzled smem,
//                               previous block's PV folded into the O registers while the tensor core runs PV_i
// S and Otmp are double buffered; the second CTA on the SM and the second warp per scheduler fill dependency bubbles.
// q/k already carry RoPE (Wqkv GEMM epilogue).  On local layers only the key blocks intersecting |i - j| <= window
// are visited, and blocks fully outside a warp's window are skipped without touching TMEM.
#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int AQ = 128, AK = 64, AD = 64;
constexpr int SOFT_WARPS = 8;
constexpr int KVS = 4;                  // K/V ring depth: TMA latency (~1-2k cycles) must be covered by >= 2 blocks in flight
constexpr int ATT_THREADS = 32 * (2 + SOFT_WARPS);
constexpr uint32_t ATT_TMEM_COLS = 256;  // S0 [0,64)  S1 [64,128)  Otmp0 [128,192)  Otmp1 [192,256)
constexpr int SQ_BYTES = AQ * AD * 2;    // 16384
constexpr int SKV_BYTES = AK * AD * 2;   // 8192
constexpr int SP_BYTES = AQ * AK * 2;    // 16384
constexpr int SX_BYTES = 2 * 2 * AQ * 4; // row-max exchange [block parity][half][row]
constexpr int ATT_SMEM = SQ_BYTES + KVS * 2 * SKV_BYTES + SP_BYTES + SX_BYTES + 1024 + 256;
constexpr int HC = AK / 2;               // columns per softmax thread (32)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pair_sync(int quarter) {  // the two warps sharing a TMEM lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

template <bool LOCAL>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    __half* __restrict__ out, const int32_t* __restrict__ cu_seqlens, int hidden, float scale_log2e,
                    int window) {
  const int seq = blockIdx.z, head = blockIdx.y;
  const int s0 = cu_seqlens[seq];
  const int L = cu_seqlens[seq + 1] - s0;
  const int q0 = blockIdx.x * AQ;
  if (q0 >= L) return;  // uniform for the CTA, before any barrier / TMEM allocation

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + SQ_BYTES;               // stage s: K at sKV + s*16384, V at +8192
  uint8_t* sP = sKV + KVS * 2 * SKV_BYTES;
  float* sX = reinterpret_cast<float*>(sP + SP_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + SX_BYTES);
  uint64_t* bar_q = bars;            // 1
  uint64_t* kv_full = bars + 1;              // [KVS]
  uint64_t* kv_empty = kv_full + KVS;        // [KVS]
  uint64_t* s_full = kv_empty + KVS;         // [2]
  uint64_t* s_empty = s_full + 2;            // [2]
  uint64_t* p_full = s_empty + 2;            // 1
  uint64_t* pv_done = p_full + 1;            // 1
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;
  int kv_lo = 0, kv_hi = L;
  if (LOCAL) {
    kv_lo = max(0, q0 - window);
    kv_hi = min(L, q0 + AQ + window);
  }
  const int j_lo = kv_lo / AK;
  const int nb = (kv_hi + AK - 1) / AK - j_lo;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < KVS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_empty + i, SOFT_WARPS);
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

  // The producer and MMA loops run in the WHOLE warp (warp-uniform values); only the TMA / MMA / commit instructions
  // are predicated on one elected lane -- under `if (lane == 0)` every descriptor is a divergent value and each
  // UTCHMMA / UTMALDG gets wrapped in an ELECT + R2UR waterfall (dozens of extra instructions per MMA).
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, SQ_BYTES);
      tma_load_2d(sQ, &tmQ, bar_q, head * AD, s0 + q0);
    }
    __syncwarp();
    for (int i = 0; i < nb; ++i) {
      const int st = i % KVS;
      mbar_wait_tagged(kv_empty + st, ((i / KVS) & 1) ^ 1, 3);
      const int row = s0 + (j_lo + i) * AK;
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full + st, 2 * SKV_BYTES);
        tma_load_2d(sKV + st * 2 * SKV_BYTES, &tmKV, kv_full + st, hidden + head * AD, row);
        tma_load_2d(sKV + st * 2 * SKV_BYTES + SKV_BYTES, &tmKV, kv_full + st, 2 * hidden + head * AD, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(0, AQ, AK);
    constexpr uint32_t idesc_pv = umma_idesc_major(0, AQ, AD, 0, 1);  // B = V is MN-major ([key][d] rows)
    const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), kv_addr = smem_u32(sKV);
    mbar_wait_tagged(bar_q, 0, 1);
    auto issue_s = [&](int i) {
      const int st = i % KVS, sb = i & 1;   // K/V ring slot, S buffer
      mbar_wait_tagged(kv_full + st, (i / KVS) & 1, 2);
      mbar_wait_tagged(s_empty + sb, ((i >> 1) & 1) ^ 1, 5);
      tc_fence_after();
      const uint32_t k_addr = kv_addr + st * 2 * SKV_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AD / 16; ++k)
          umma_f16(tmem_base + sb * AK, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc_s,
                   k > 0 ? 1u : 0u);
        umma_commit(s_full + sb);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int i = 0; i < nb; ++i) {
      if (i + 1 < nb) issue_s(i + 1);
      mbar_wait_tagged(p_full, i & 1, 6);
      tc_fence_after();
      const int st = i % KVS;
      const uint32_t v_addr = kv_addr + st * 2 * SKV_BYTES + SKV_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AK / 16; ++k)  // 16 keys per step = two 8-row groups of the V tile = 2048 bytes
          umma_f16(tmem_base + 2 * AK + (i & 1) * AD, umma_desc_sw128(p_addr + k * 32),
                   umma_desc_sw128(v_addr + k * 2048), idesc_pv, k > 0 ? 1u : 0u);
        umma_commit(pv_done);
        umma_commit(kv_empty + st);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;        // TMEM lane quarter
    const int half = (warp - 2) >> 2;    // which 32 of the 64 S / O columns this thread owns
    const int r = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const int q = q0 + r;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * HC;
    float o[HC];
#pragma unroll
    for (int d = 0; d < HC; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

    // O = O * alpha + PV_{i_done}   (PV_i lives in TMEM Otmp[i & 1]; caller has waited pv_done(i_done))
    auto fold = [&](int i_done, float alpha) {
      uint32_t t[HC];
      tmem_ld_32x32b_x32(t_lane + 2 * AK + (i_done & 1) * AD, t);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < HC; ++e) o[e] = fmaf(o[e], alpha, __uint_as_float(t[e]));
    };

    // keys this query may attend: [k_lo, k_hi]
    const int k_lo = LOCAL ? max(q - window, 0) : 0;
    const int k_hi = LOCAL ? min(q + window, L - 1) : L - 1;
    uint8_t* prow = sP + r * 128;

    for (int i = 0; i < nb; ++i) {
      const int st = i & 1;
      const int key0 = (j_lo + i) * AK;
      // block-level decision on the full 64 columns (identical in both warps of the quarter)
      const bool dead = __all_sync(0xffffffffu, k_hi - key0 < 0 || k_lo - key0 > AK - 1);
      const int e_lo = k_lo - key0 - half * HC, e_hi = k_hi - key0 - half * HC;  // valid local columns [e_lo, e_hi]
      mbar_wait_tagged(s_full + st, (i >> 1) & 1, 4);
      tc_fence_after();
      float alpha = 1.f, sum = 0.f, m_new = m;
      uint4 pk[4];
      if (dead) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + st);
#pragma unroll
        for (int c = 0; c < 4; ++c) pk[c] = make_uint4(0u, 0u, 0u, 0u);
      } else {
        float s[HC];
        {
          uint32_t t[HC];
          tmem_ld_32x32b_x32(t_lane + st * AK, t);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < HC; ++e) s[e] = __uint_as_float(t[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + st);

        if (!__all_sync(0xffffffffu, e_lo <= 0 && e_hi >= HC - 1)) {  // boundary: mask (warp-uniform branch)
#pragma unroll
          for (int e = 0; e < HC; ++e) s[e] = (e >= e_lo && e <= e_hi) ? s[e] : -INFINITY;
        }
        float mx4[4];  // independent max chains
#pragma unroll
        for (int e = 0; e < 4; ++e) mx4[e] = s[e];
#pragma unroll
        for (int e = 4; e < HC; ++e) mx4[e & 3] = fmaxf(mx4[e & 3], s[e]);
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // both halves of the row must use the same running max: exchange through smem (double buffered by block)
        float* xm = sX + (i & 1) * 2 * AQ;
        xm[half * AQ + r] = mx;
        pair_sync(quarter);
        mx = fmaxf(mx, xm[(half ^ 1) * AQ + r]);
        m_new = fmaxf(m, mx * scale_log2e);  // scale > 0: max commutes with the scaling
        const float mu = m_new == -INFINITY ? 0.f : m_new;
        alpha = ex2(m - mu);  // first block: ex2(-inf) = 0
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = ex2(fmaf(s[c * 8 + e], scale_log2e, -mu));  // masked: fma(-inf, .) = -inf -> 0
            sum4[e & 3] += p[e];
          }
          pk[c].x = pack_half2(p[0], p[1]);
          pk[c].y = pack_half2(p[2], p[3]);
          pk[c].z = pack_half2(p[4], p[5]);
          pk[c].w = pack_half2(p[6], p[7]);
        }
        sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      }
      // P smem is free and Otmp[(i-1)&1] is valid once PV_{i-1} has completed
      if (i > 0) {
        mbar_wait_tagged(pv_done, (i - 1) & 1, 7);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (r & 7)) << 4)) = pk[c];
      fence_proxy_async_smem();  // P written with st.shared must be visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // off the critical path: fold the previous block's PV while the tensor core runs PV_i
      if (i > 0) fold(i - 1, alpha_prev);
      l = fmaf(l, alpha, sum);   // partial row sum over this thread's columns (same alpha sequence in both halves)
      m = m_new;
      alpha_prev = alpha;
    }
    mbar_wait_tagged(pv_done, (nb - 1) & 1, 7);
    tc_fence_after();
    fold(nb - 1, alpha_prev);
    // total row sum = sum of the two halves' partial sums (sX is free: all max exchanges are behind the barrier)
    pair_sync(quarter);
    sX[half * AQ + r] = l;
    pair_sync(quarter);
    l += sX[(half ^ 1) * AQ + r];
    if (q < L) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(s0 + q) * hidden + head * AD + half * HC);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_half2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
        u.y = pack_half2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
        u.z = pack_half2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
        u.w = pack_half2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
        dst[c] = u;
      }
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

void launch_attention_tc(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev, int nseq,
                         int total_tokens, int max_len, int heads, int hidden, int window /* <0: full */) {
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
  dim3 grid((max_len + AQ - 1) / AQ, heads, nseq);
  if (window >= 0)
    attention_tc_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tmQ, tmKV, out, cu_seqlens_dev, hidden,
                                                                            scale_log2e, window);
  else
    attention_tc_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tmQ, tmKV, out, cu_seqlens_dev, hidden,
                                                                             scale_log2e, 0);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace vrag
