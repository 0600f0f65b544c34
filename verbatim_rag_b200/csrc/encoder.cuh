// Internal kernels + launchers of the encoder forward passes (ModernBERT token classifier, BERT MLM/SPLADE, BERT dense).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace vrag {

constexpr int HIDDEN = 768;  // ModernBERT-base / BERT-base: H = 768, 12 heads x 64.  BERT encoders also run at H = 384
                             // (12 heads x 32: the MiniLM family, the reference's default dense model,
                             // embedding_providers.py:55) on the same kernels: heads are zero-padded to 64 dims.

// attention.cu
void launch_attention(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev, int nseq,
                      int max_len, int heads, int hidden, int window);

// attention_tc.cu: tcgen05 version (default); the mma.sync kernel above stays as a debug cross-check
// (VRAG_ATTENTION_LEGACY=1 at encoder creation).
void launch_attention_tc(vrag_ctx* ctx, const __half* qkv, __half* out, const int32_t* cu_seqlens_dev,
                         const int32_t* work_dev /* [n_pairs][2] = (sequence, q0) */, int n_pairs, int total_tokens,
                         int heads, int hidden, int window);
// Split-precision ("precise") variant: q|k|v and the output as hi / lo fp16 planes (gemm.cuh), three tcgen05.mma per
// product (S = Q_hi K_hi + Q_lo K_hi + Q_hi K_lo, O += P_hi V_hi + P_lo V_hi + P_hi V_lo).
void launch_attention_tc_split(vrag_ctx* ctx, const __half* qkv, const __half* qkv_lo, __half* out, __half* out_lo,
                               const int32_t* work_dev, int n_pairs, int total_tokens, int heads, int hidden, int window);
void launch_f32_to_f16_split(vrag_ctx* ctx, const float* src, __half* hi, __half* lo, size_t n);

// rowops.cu  (one warp per token row, H = 768)
void launch_token_meta(vrag_ctx* ctx, const int32_t* cu_seqlens_dev, int nseq, int total, int32_t* pos,
                       int32_t* seq_of_row);
// x32 (nullable) = LN(emb[ids]) * gamma, h16 = fp16 of it, lo8 (nullable) = e5m2 of the rounding remainder.
// h16_lo (nullable): fp16 of the same remainder = low plane of the split-precision ("precise") mode.
void launch_embed_ln(vrag_ctx* ctx, const int32_t* ids, int T, int vocab, const float* tok_emb, const float* gamma,
                     float eps, float* x32, __half* h16, uint8_t* lo8, __half* h16_lo = nullptr);
// Two-plane residual stream (x = fp16 hi + e5m2 lo, deferred-LayerNorm path): final LayerNorm, fp32 reconstruction.
void launch_layernorm_hilo(vrag_ctx* ctx, const __half* hi, const uint8_t* lo, int T, const float* gamma,
                           const float* beta /*nullable*/, float eps, float* x32 /*nullable*/, __half* h16);
// BERT deferred-LayerNorm path: raw embedding sum as the two-plane stream + row moments ([slots][T] float2, slot 0).
void launch_bert_embed_raw(vrag_ctx* ctx, const int32_t* ids, const int32_t* pos, int T, int vocab, int max_pos,
                           const float* word_emb, const float* pos_emb, const float* type_emb0, __half* h16,
                           uint8_t* lo8, float* stats, int slots);
void launch_hilo_to_f32(vrag_ctx* ctx, const __half* hi, const uint8_t* lo, int T, float* x32);
void launch_bert_embed_ln(vrag_ctx* ctx, const int32_t* ids, const int32_t* pos, int T, int vocab, int max_pos,
                          const float* word_emb, const float* pos_emb, const float* type_emb0, const float* gamma,
                          const float* beta, float eps, float* x32, __half* h16, __half* h16_lo = nullptr,
                          int hidden = HIDDEN, const int32_t* type_ids = nullptr);
// Device span post-processing: fill = false counts the spans of each sequence into counts[nseq] and
// their exclusive prefix sums into offs[nseq + 1]; fill = true writes the spans at offs[seq].
void launch_span_runs(vrag_ctx* ctx, const float* probs, const int32_t* cu, const int32_t* ctx_first, const int32_t* ctx_len,
                      const int64_t* tok_base, const int32_t* tcs, const int32_t* tce, int nseq, float threshold,
                      int min_span_chars, int merge_gap_chars, int32_t* counts, int32_t* offs, int32_t seq_base,
                      int32_t* o_seq, int32_t* o_cs, int32_t* o_ce, float* o_score, int32_t* o_ts, int32_t* o_te, bool fill);
// Cross-encoder head on the [CLS] rows: scores[s] = W_c tanh(W_p x_CLS + b_p) + b_c  (BertForSequenceClassification).
void launch_cls_head(vrag_ctx* ctx, const float* x32, const int32_t* cu_seqlens_dev, int nseq, int hidden, const float* wp,
                     const float* bp, const float* wc, const float* bc, float* scores);
// Sentence head (legacy QAModel): mean of x32 rows [row_start, row_end] -> Linear(hidden, 2).
void launch_sentence_head(vrag_ctx* ctx, const float* x32, int hidden, const int32_t* row_start, const int32_t* row_end,
                          int nsent, const float* cw, const float* cb, float* logits);
// h16 = LN(x32) * gamma (+ beta); if write_back, x32 is overwritten with the normalised row too (post-LN residual).
// h16_lo (nullable): low plane of h16 (split-precision mode).
void launch_layernorm(vrag_ctx* ctx, float* x32, int T, const float* gamma, const float* beta, float eps,
                      __half* h16, bool write_back, __half* h16_lo = nullptr, int hidden = HIDDEN);
// ModernBERT head tail: LN(buf32) * gamma -> classifier (2 x 768) + bias -> logits, P(class 1).
void launch_head_final(vrag_ctx* ctx, const float* buf32, int T, const float* gamma, float eps, const float* cls_w,
                       const float* cls_b, float* logits /*nullable*/, float* probs);
// Fused head (gemm.cuh EPI_HEAD_PARTIAL): partial sums [slots][T] float4 -> logits, P(class 1).  g0 / g1 = sum_n cls_gw[c][n].
void launch_head_finish(vrag_ctx* ctx, const float* head_part, int T, int slots, float eps, float g0, float g1,
                        const float* cls_b, float* logits /*nullable*/, float* probs);
// SPLADE CSR extraction from the dense [nseq, ld] buffer.
void launch_splade_count(vrag_ctx* ctx, const float* dense, int nseq, int ld, int vocab, float min_abs,
                         int32_t* counts);
void launch_splade_fill(vrag_ctx* ctx, const float* dense, int nseq, int ld, int vocab, float min_abs,
                        const int64_t* indptr_dev, int32_t* indices, float* values);
void launch_pool(vrag_ctx* ctx, const float* x32, const int32_t* cu_seqlens_dev, int nseq, int pooling,
                 int normalize, float* out, int hidden = HIDDEN);
void launch_f32_to_f16(vrag_ctx* ctx, const float* src, __half* dst, size_t n);

}  // namespace vrag
