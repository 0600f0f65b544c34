// Self-attention over unpadded (varlen) sequences, one (64-query tile, head, sequence) per CTA.
// q/k already carry RoPE (fused in the Wqkv GEMM epilogue).  Full attention on global layers; on local layers only
// the key tiles intersecting |i - j| <= window are visited.  Online softmax in base 2, fp32 statistics.
// v1 tiles use warp-level mma.sync m16n8k16 (attention is ~7% of the encoder FLOPs); the projections around it are
// the tcgen05 GEMMs.
#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int BQ = 64, BKV = 64, HD = 64;

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset of a 16-byte chunk in a [rows][64] fp16 tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void load_tile_async(__half* dst, const __half* src, int ld, int rows_valid) {
  const uint32_t d0 = smem_u32(dst);
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < rows_valid;
    cp_async_16(d0 + swz(r, c), ok ? src + static_cast<size_t>(r) * ld + c * 8 : src, ok);
  }
}

template <bool LOCAL>
__global__ void __launch_bounds__(128)
attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, const int32_t* __restrict__ cu_seqlens,
                 int ld_qkv, int hidden, float scale_log2e, int window) {
  const int seq = blockIdx.z, head = blockIdx.y;
  const int s0 = cu_seqlens[seq];
  const int L = cu_seqlens[seq + 1] - s0;
  const int q0 = blockIdx.x * BQ;
  if (q0 >= L) return;

  __shared__ __align__(128) __half sQ[BQ * HD];
  __shared__ __align__(128) __half sK[2][BKV * HD];
  __shared__ __align__(128) __half sV[2][BKV * HD];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* qbase = qkv + static_cast<size_t>(s0) * ld_qkv + head * HD;
  const __half* kbase = qbase + hidden;
  const __half* vbase = qbase + 2 * hidden;

  int kv_lo = 0, kv_hi = L;
  if (LOCAL) {
    kv_lo = max(0, q0 - window);
    kv_hi = min(L, q0 + BQ + window);
  }
  const int t_lo = kv_lo / BKV, t_hi = (kv_hi + BKV - 1) / BKV;

  load_tile_async(sQ, qbase + static_cast<size_t>(q0) * ld_qkv, ld_qkv, L - q0);
  load_tile_async(sK[0], kbase + static_cast<size_t>(t_lo * BKV) * ld_qkv, ld_qkv, L - t_lo * BKV);
  load_tile_async(sV[0], vbase + static_cast<size_t>(t_lo * BKV) * ld_qkv, ld_qkv, L - t_lo * BKV);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const int qrow_lo = q0 + warp * 16 + (lane >> 2);
  const int qrow_hi = qrow_lo + 8;

  for (int t = t_lo; t < t_hi; ++t) {
    const int st = (t - t_lo) & 1;
    if (t + 1 < t_hi) {
      load_tile_async(sK[st ^ 1], kbase + static_cast<size_t>((t + 1) * BKV) * ld_qkv, ld_qkv, L - (t + 1) * BKV);
      load_tile_async(sV[st ^ 1], vbase + static_cast<size_t>((t + 1) * BKV) * ld_qkv, ld_qkv, L - (t + 1) * BKV);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == t_lo) {
      const uint32_t q_s = smem_u32(sQ);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldmatrix_x4(qf[kk], q_s + swz(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)));
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    const uint32_t k_s = smem_u32(sK[st]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[4];
        ldmatrix_x4(b, k_s + swz(jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)));
        mma_m16n8k16_f16(s[2 * jp], qf[kk], b[0], b[1]);
        mma_m16n8k16_f16(s[2 * jp + 1], qf[kk], b[2], b[3]);
      }
    }
    // ---- mask + online softmax (base 2)
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = t * BKV + j * 8 + (lane & 3) * 2 + e;
        bool ok_lo = key < L, ok_hi = key < L;
        if (LOCAL) {
          ok_lo = ok_lo && (key - qrow_lo <= window) && (qrow_lo - key <= window);
          ok_hi = ok_hi && (key - qrow_hi <= window) && (qrow_hi - key <= window);
        }
        s[j][e] = ok_lo ? s[j][e] * scale_log2e : -INFINITY;
        s[j][2 + e] = ok_hi ? s[j][2 + e] * scale_log2e : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[j][e]);
        mx_hi = fmaxf(mx_hi, s[j][2 + e]);
      }
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float mu_lo = mn_lo == -INFINITY ? 0.f : mn_lo, mu_hi = mn_hi == -INFINITY ? 0.f : mn_hi;
    const float sc_lo = exp2f(m_lo - mu_lo), sc_hi = exp2f(m_hi - mu_hi);  // exp2(-inf) = 0 on the first tile
    m_lo = mn_lo;
    m_hi = mn_hi;
    l_lo *= sc_lo;
    l_hi *= sc_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= sc_lo; o[j][1] *= sc_lo; o[j][2] *= sc_hi; o[j][3] *= sc_hi;
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - mu_lo), p1 = exp2f(s[j][1] - mu_lo);
      const float p2 = exp2f(s[j][2] - mu_hi), p3 = exp2f(s[j][3] - mu_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack_half2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
    }
    // ---- O += P V   (16 x 64 per warp)
    const uint32_t v_s = smem_u32(sV[st]);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, v_s + swz(kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        mma_m16n8k16_f16(o[2 * dp], pf[kb], b[0], b[1]);
        mma_m16n8k16_f16(o[2 * dp + 1], pf[kb], b[2], b[3]);
      }
    }
    __syncthreads();
  }

  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  __half* obase = out + static_cast<size_t>(s0) * hidden + head * HD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    if (qrow_lo < L)
      *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qrow_lo) * hidden + col) =
          pack_half2(o[j][0] * inv_lo, o[j][1] * inv_lo);
    if (qrow_hi < L)
      *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(qrow_hi) * hidden + col) =
          pack_half2(o[j][2] * inv_hi, o[j][3] * inv_hi);
  }
}

}  // namespace

void launch_attention(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev, int nseq,
                      int max_len, int heads, int hidden, int window /* <0: full */) {
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // head_dim 64
  dim3 grid((max_len + BQ - 1) / BQ, heads, nseq);
  ProfScope prof(ctx, PROF_ATTENTION);
  if (window >= 0)
    attention_kernel<true><<<grid, 128, 0, ctx->stream>>>(qkv, out, cu_seqlens_dev, 3 * hidden, hidden, scale_log2e,
                                                           window);
  else
    attention_kernel<false><<<grid, 128, 0, ctx->stream>>>(qkv, out, cu_seqlens_dev, 3 * hidden, hidden, scale_log2e,
                                                            0);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace vrag
