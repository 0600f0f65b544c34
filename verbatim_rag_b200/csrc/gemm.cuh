// Tensor-core GEMM of the encoder blocks:  C[M,N] = A[M,K] * W[N,K]^T, fp16 operands, fp32 accumulate in TMEM,
// with the layer's element-wise tail fused into the epilogue.
#pragma once
#include "common.cuh"

namespace vrag {

enum GemmEpi : int {
  EPI_F16 = 0,            // out16 = acc
  EPI_ROPE_QKV = 1,       // ModernBERT Wqkv: rotate q/k heads by the token position, v as is -> out16 [M, 3*H]
  EPI_RESID_F32 = 2,      // out32 += acc                       (Wo / mlp.Wo onto the fp32 residual stream)
  EPI_GEGLU = 3,          // out16[:, 128t+j] = gelu(acc[j]) * acc[128+j]   (Wi rows interleaved per 128)
  EPI_BIAS_F16 = 4,       // out16 = acc + bias                 (BERT fused q|k|v)
  EPI_BIAS_GELU_F16 = 5,  // out16 = gelu(acc + bias)           (BERT intermediate)
  EPI_BIAS_RESID_F32 = 6, // out32 += acc + bias                (BERT attention.output / output dense)
  EPI_SPLADE = 7,         // splade[seq(row), col] = max(., log1p(relu(acc + bias)))   (MLM decoder, never stores logits)
  EPI_GELU_F32 = 8,       // out32 = gelu(acc)                  (ModernBERT head.dense)
  EPI_BIAS_GELU_F32 = 9,  // out32 = gelu(acc + bias)           (BERT MLM transform.dense)
  EPI_F32 = 10,           // out32 = acc                        (self test)
  // Deferred LayerNorm (ModernBERT pre-LN blocks).  LN(x) W^T = rstd(x) * (x W''^T) with
  // W''[n,k] = W[n,k] gamma[k] - mean_k(W[n,:] gamma) folded at load time, so the consumer GEMM reads the RAW residual
  // (fp16 copy) and only needs one scalar per row; the producer GEMM emits that copy and the row moments.
  EPI_RESID_STATS = 11,   // residual stream as two planes, x = out16 (hi, fp16) + out8_lo (lo, e5m2: ptx.cuh), >= 14
                          // significant bits: x += acc; hi = fp16(x), lo = e5m2(x - hi); stats_out[slot][row] =
                          // (sum x, sum x^2) over this warp's 128 columns (slot = 2 * n_tile + column half; N = 768 ->
                          // 6 slots).  The hi plane IS the next GEMM's A operand, so the stream costs 2.25 KB read +
                          // 2.25 KB written per token and nothing else (a separate fp32 stream + fp16 copy: 3 + 4.5 KB;
                          // an fp16 low plane, the first version: 3 + 3 KB -- these two GEMMs are HBM-bound).
  EPI_NORM_ROPE_QKV = 12, // EPI_ROPE_QKV on rstd[row] * acc
  EPI_NORM_GEGLU = 13,    // EPI_GEGLU on rstd[row] * acc
  // The same treatment of BERT's post-LN blocks, y = LN(z) * gamma + beta with z = y_prev + sublayer(y_prev): the
  // stream holds the PRE-norm sum z (two planes + row moments); consumers read its hi plane against weights folded with
  // gamma (and beta folded into their bias), the residual GEMM normalises the old stream on the fly.
  EPI_NORM_BIAS_F16 = 14,       // out16 = rstd[row] * acc + bias                     (BERT q|k|v)
  EPI_NORM_BIAS_GELU_F16 = 15,  // out16 = gelu(rstd[row] * acc + bias)               (BERT intermediate)
  EPI_SCORES = 17,              // similarity scan as a GEMM (topk.cu, batched dense search): out32[m][n] = acc for the
                                // n_valid real columns, -inf where col_mask[n] != 0 (deleted / filtered corpus rows)
  EPI_SCORES_THRESH = 19,       // EPI_SCORES with the selection fused in: nothing is stored but the (rare) scores >= thr[m],
                                // appended as 64-bit keys (order-preserving score bits << 32 | ~row) to cand[m][..cand_cap)
                                // through cand_count[m]; row = col_base + n.  The [M, N] score matrix never exists.
  EPI_HEAD_PARTIAL = 18,        // span-logit head fused into head.dense (ModernBERT prediction head + classifier):
                                // g = gelu(acc); per row and 128 columns the four partial sums (sum g, sum g^2,
                                // sum g * cls_gw[0][n], sum g * cls_gw[1][n]) -> head_part[slot][row] (float4, slot =
                                // 2 * n_tile + column half).  cls_gw[c][n] = head.norm.weight[n] * classifier.weight[c][n]:
                                // logits[c] = rstd * (D_c - mean * sum_n cls_gw[c][n]) + bias[c] needs only these sums
                                // (rowops.cu head_finish_kernel), so the [T, 768] fp32 head activation is never stored.
  EPI_RESID_STATS_LN = 16,      // EPI_RESID_STATS with x_old = ((hi + lo) - mean[row]) * rstd[row] * gamma[col] + bias[col]
                                // (bias = beta + the dense bias); reads stats_in, writes stats_out (different buffers)
};

struct GemmEpiParams {
  // Split-precision ("precise") mode: every fp16 tensor is a pair of planes x = hi + lo with hi = fp16(x) and
  // lo = fp16(x - hi) (>= 21 significant bits; absolute resolution 2^-25), and a product is evaluated as
  // a_hi w_hi + a_lo w_hi + a_hi w_lo (three tcgen05.mma per k-step, fp32 accumulation in TMEM).  Selected by passing the
  // low planes of BOTH operands; epilogues that write fp16 activations then also write their low plane to out16_lo.
  const __half* a_lo = nullptr;       // [M, K] low plane of A
  const __half* w_lo = nullptr;       // [N, K] low plane of W
  __half* out16_lo = nullptr;         // low plane of out16 (same leading dimension)
  __half* out16 = nullptr;
  int ld16 = 0;
  uint8_t* out8_lo = nullptr;         // EPI_RESID_STATS: low plane of the residual stream, e5m2 [M, ld16]
  float* out32 = nullptr;
  int ld32 = 0;
  const float* bias = nullptr;
  const float* gamma = nullptr;       // EPI_RESID_STATS_LN: LayerNorm weight applied to the old stream
  const int32_t* pos = nullptr;       // [M] position of each token inside its sequence
  const float* rope_tab = nullptr;    // [rope_rows, 64]: cos(pos * inv_freq[0..32)) | sin(pos * inv_freq[0..32))
  int rope_rows = 0;                  // positions in the table
  int hidden = 768;                   // H: q = cols [0,H), k = [H,2H), v = [2H,3H)
  const int32_t* seq_of_row = nullptr;  // [M] sequence index of each token (SPLADE pooling)
  float* splade_out = nullptr;        // [nseq, splade_ld], zero-initialised
  int splade_ld = 0;
  int n_valid = 0;                    // columns >= n_valid are padding (SPLADE vocab tail, EPI_SCORES corpus tail)
  const uint8_t* col_mask = nullptr;  // EPI_SCORES: [n_valid] bytes, != 0 -> the column scores -inf
  const float* thr = nullptr;         // EPI_SCORES_THRESH: [M] per-query threshold
  unsigned long long* cand = nullptr; // EPI_SCORES_THRESH: [M][cand_cap] candidate keys
  int* cand_count = nullptr;          // EPI_SCORES_THRESH: [M] number appended (may exceed cand_cap: overflow)
  int cand_cap = 0;
  int col_base = 0;                   // EPI_SCORES_THRESH: corpus row of column 0
  const float* cls_gw = nullptr;      // EPI_HEAD_PARTIAL: [2][N] fp32
  float* head_part = nullptr;         // EPI_HEAD_PARTIAL: [N / 128][M] float4
  int prof_class = 0;                 // profiler class the launch is booked under (common.cuh ProfClass; 0 = GEMM)
  const float* stats_in = nullptr;    // EPI_NORM_*: [stats_slots][M] float2 partial row moments of the A operand's rows
  float* stats_out = nullptr;         // EPI_RESID_STATS: [N / 128][M] float2
  int stats_slots = 6;
  float ln_eps = 1e-5f;
  float inv_dim = 1.0f / 768.0f;      // 1 / (number of columns the moments were taken over)
  int M = 0;                          // valid rows
  int debug_mode = 0;                 // timing experiments of vrag_bench_gemm only: 3 = mainloop only (the epilogue
                                      // releases the accumulators at once), 4 = epilogue without its TMA stores
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;

// A: [M, K] fp16 row-major (lda = K).  W: [N, K] fp16 row-major.  N % 256 == 0, K % 64 == 0.
// use_reference != 0 runs the SIMT reference kernel with the same epilogue (debug / self test only).
void launch_gemm(vrag_ctx* ctx, int epi, const __half* A, const __half* W, int M, int N, int K,
                 const GemmEpiParams& p, int use_reference);

}  // namespace vrag
