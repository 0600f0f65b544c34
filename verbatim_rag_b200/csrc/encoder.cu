// Encoder handles: checkpoint repack (fp32 host tensors -> device operands) and the forward schedules
// (ModernBERT token classifier, BERT MLM + SPLADE pooling, BERT dense pooling).
//
// Data layout in HBM (per pass of <= max_tokens tokens, sequences packed back to back, no padding):
//   x32   [T, 768] fp32   residual stream            h16  [T, 768]  fp16  LayerNorm output = GEMM A operand
//   qkv16 [T, 2304] fp16  q|k|v (RoPE applied)       o16  [T, 768]  fp16  attention output
//   w16   [T, 1152|3072] fp16  GeGLU / FFN activation  buf32 [T, 768] fp32 head pre-norm
// Weights: every Linear as fp16 [N, K] row-major (K-major tensor-core operand), norms / biases / embeddings fp32.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>

#include "encoder.cuh"

using namespace vrag;

struct vrag_encoder {
  vrag_ctx* ctx = nullptr;
  int kind = 0, layers = 0, vocab = 0, vocab_pad = 0, max_tokens = 0, max_seqs = 0, max_pos = 0;
  int ffn = 0;  // GeGLU width (1152) or FFN width (3072 / 1536)
  int hidden = vrag::HIDDEN;  // residual-stream width: 768, or 384 for MiniLM-class BERT encoders (12 heads x 32)
  bool use_reference_gemm = false;
  bool legacy_attention = false;
  bool deferred_ln = true;  // LayerNorm folded into the GEMMs (EPI_RESID_STATS* / EPI_NORM_*), no LN kernels in the stack
  // VRAG_PRECISION_PRECISE: split-precision operands (gemm.cuh): every fp16 weight / activation is a hi + lo pair of
  // planes (>= 21 significant bits), three tcgen05.mma per product, fp32 residual stream and LayerNorm kernels.
  bool precise = false;
  void attention(const __half* qkv, __half* out, int nseq, int total_tokens, int max_len, int window) {
    if (precise)
      vrag::launch_attention_tc_split(ctx, qkv, qkv16_lo.as<__half>(), out, o16_lo.as<__half>(), work.as<int32_t>(),
                                      n_pairs, total_tokens, 12, vrag::HIDDEN, window);
    else if (legacy_attention) vrag::launch_attention(ctx, qkv, out, cu.as<int32_t>(), nseq, max_len, 12, vrag::HIDDEN, window);
    else vrag::launch_attention_tc(ctx, qkv, out, cu.as<int32_t>(), work.as<int32_t>(), n_pairs, total_tokens, 12,
                                   vrag::HIDDEN, window);
  }
  std::vector<DevBuf*> owned;
  // shared
  float *emb = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  // ModernBERT
  struct MLayer {
    float* attn_g; __half* wqkv; __half* wo; float* mlp_g; __half* wi; __half* wo2;
    __half *wqkv_lo = nullptr, *wo_lo = nullptr, *wi_lo = nullptr, *wo2_lo = nullptr;   // precise mode: low planes
  };
  std::vector<MLayer> ml;
  float *final_g = nullptr, *head_g = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  float* cls_gw = nullptr;            // [2][768] head.norm.weight * classifier.weight (fused head, EPI_HEAD_PARTIAL)
  float cls_g0 = 0.f, cls_g1 = 0.f;   // row sums of cls_gw
  bool fused_head = true;             // VRAG_FUSED_HEAD=0: head.dense -> fp32 [T, 768] -> head_final_kernel (cross-check)
  __half *head_w = nullptr, *head_w_lo = nullptr;
  float *rope_g = nullptr, *rope_l = nullptr;  // [max_pos, 64] cos | sin tables (global / local theta)
  // BERT
  struct BLayer {
    __half* wqkv; float* bqkv; __half* wo; float* bo; float *g1, *b1; __half* wi; float* bi; __half* wo2; float* bo2;
    float *g2, *b2;
    // deferred-LayerNorm path: folded biases of the consumer GEMMs (beta . W + b) and, for the two residual GEMMs, the
    // LayerNorm weight / (beta + dense bias) that normalise the OLD stream (the LayerNorm that precedes the sublayer)
    float *cqkv = nullptr, *ci = nullptr, *ra_g = nullptr, *ra_b = nullptr, *rb_g = nullptr, *rb_b = nullptr;
    __half *wqkv_lo = nullptr, *wo_lo = nullptr, *wi_lo = nullptr, *wo2_lo = nullptr;   // precise mode: low planes
  };
  std::vector<BLayer> bl;
  float *pos_emb = nullptr, *type_emb = nullptr;
  __half *mlm_w = nullptr, *dec_w = nullptr, *mlm_w_lo = nullptr, *dec_w_lo = nullptr;
  float *pool_w = nullptr, *pool_b = nullptr, *seq_w = nullptr, *seq_b = nullptr;   // BERT_CLS: pooler + 1-label classifier (fp32)
  const int32_t* type_ids_host = nullptr;   // BERT_CLS: token types of the call being served (host / device per on_device)
  float *mlm_b = nullptr, *mlm_g = nullptr, *mlm_beta = nullptr, *dec_b = nullptr;
  // workspace
  DevBuf ids, cu, pos, seqrow, x32, h16, qkv16, o16, w16, buf32, probs, logits, splade, counts, indptr, sp_idx, sp_val,
      pooled, work, stats, stats2, xl8, h16_lo, qkv16_lo, o16_lo, w16_lo, types, srow0, srow1, slog, sp_meta, sp_off, sp_out;
  int n_pairs = 0;  // (sequence, 128-query tile) entries of the current pass in `work`

  ~vrag_encoder() {
    for (auto* b : owned) { b->release(); delete b; }
    for (DevBuf* b : {&ids, &cu, &pos, &seqrow, &x32, &h16, &qkv16, &o16, &w16, &buf32, &probs, &logits, &splade,
                      &counts, &indptr, &sp_idx, &sp_val, &pooled, &work, &stats, &stats2, &xl8, &h16_lo, &qkv16_lo,
                      &o16_lo, &w16_lo, &types, &srow0, &srow1, &slog, &sp_meta,
                      &sp_off, &sp_out})
      b->release();
  }
  template <typename T>
  T* alloc(size_t n) {
    DevBuf* b = new DevBuf();
    owned.push_back(b);
    b->reserve(n * sizeof(T));
    return b->as<T>();
  }
};

namespace {

struct WeightSet {
  std::map<std::string, const vrag_tensor*> m;
  WeightSet(const vrag_tensor* t, int n) {
    for (int i = 0; i < n; ++i) m[t[i].name] = &t[i];
  }
  const float* get(const std::string& name, int64_t numel) const {
    auto it = m.find(name);
    if (it == m.end()) throw Error(VRAG_ERR_WEIGHTS, "missing tensor '" + name + "'");
    if (it->second->numel != numel)
      throw Error(VRAG_ERR_WEIGHTS, "tensor '" + name + "' has " + std::to_string(it->second->numel) +
                                        " elements, expected " + std::to_string(numel));
    return it->second->data;
  }
};

float* upload_f32(vrag_encoder* e, const float* host, size_t n, size_t n_alloc = 0) {
  float* d = e->alloc<float>(n_alloc ? n_alloc : n);
  if (n_alloc > n) VRAG_CUDA(cudaMemsetAsync(d, 0, n_alloc * sizeof(float), e->ctx->stream));
  VRAG_CUDA(cudaMemcpyAsync(d, host, n * sizeof(float), cudaMemcpyHostToDevice, e->ctx->stream));
  return d;
}

// fp32 host [rows, cols] -> fp16 device [rows_alloc, cols] (extra rows zero)
// lo_out (precise mode): also allocate and fill the low plane fp16(w - fp16(w)).
__half* upload_f16(vrag_encoder* e, DevBuf& staging, const float* host, size_t rows, size_t cols,
                   size_t rows_alloc = 0, __half** lo_out = nullptr) {
  if (!rows_alloc) rows_alloc = rows;
  const bool want_lo = lo_out && e->precise;
  __half* d = e->alloc<__half>(rows_alloc * cols);
  __half* dl = want_lo ? e->alloc<__half>(rows_alloc * cols) : nullptr;
  if (rows_alloc > rows) {
    VRAG_CUDA(cudaMemsetAsync(d, 0, rows_alloc * cols * sizeof(__half), e->ctx->stream));
    if (dl) VRAG_CUDA(cudaMemsetAsync(dl, 0, rows_alloc * cols * sizeof(__half), e->ctx->stream));
  }
  staging.reserve(rows * cols * sizeof(float));
  VRAG_CUDA(cudaMemcpyAsync(staging.p, host, rows * cols * sizeof(float), cudaMemcpyHostToDevice, e->ctx->stream));
  if (dl) launch_f32_to_f16_split(e->ctx, staging.as<float>(), d, dl, rows * cols);
  else launch_f32_to_f16(e->ctx, staging.as<float>(), d, rows * cols);
  VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));  // staging is reused by the next tensor
  if (lo_out) *lo_out = dl;
  return d;
}

void make_rope(vrag_encoder* e, double theta, int max_pos, float** tab_out) {
  // row p = cos(p * inv_freq[0..32)) | sin(p * inv_freq[0..32)), inv_freq[j] = theta^(-2j/64), fp32 like
  // modeling_modernbert.py:138-172.  One 256-byte row per position: the GEMM epilogue fetches the rows of 32
  // consecutive positions with one TMA box each for cos and sin.
  std::vector<float> t(static_cast<size_t>(max_pos) * 64);
  for (int j = 0; j < 32; ++j) {
    const float inv = 1.0f / powf(static_cast<float>(theta), static_cast<float>(2 * j) / 64.0f);
    for (int p = 0; p < max_pos; ++p) {
      const float f = static_cast<float>(p) * inv;
      t[static_cast<size_t>(p) * 64 + j] = static_cast<float>(cos(static_cast<double>(f)));
      t[static_cast<size_t>(p) * 64 + 32 + j] = static_cast<float>(sin(static_cast<double>(f)));
    }
  }
  *tab_out = upload_f32(e, t.data(), t.size());
  VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
}

// Deferred LayerNorm: LN(x) W^T = rstd(x) * (x W''^T) with W''[n,k] = W[n,k] gamma[k] - mean_k(W[n,:] gamma[:]) (each row
// of W'' sums to zero, which is what subtracts the row mean of x).  Computed in double from the fp32 checkpoint.
void fold_layernorm(const float* w, const float* gamma, size_t rows, size_t cols, float* out) {
  for (size_t n = 0; n < rows; ++n) {
    double acc = 0.0;
    for (size_t k = 0; k < cols; ++k) acc += static_cast<double>(w[n * cols + k]) * gamma[k];
    const double mean = acc / static_cast<double>(cols);
    for (size_t k = 0; k < cols; ++k)
      out[n * cols + k] = static_cast<float>(static_cast<double>(w[n * cols + k]) * gamma[k] - mean);
  }
}

// bias seen by a consumer of LN(z) * gamma + beta:  c[n] = sum_k beta[k] W[n,k] + b[n]
void fold_layernorm_bias(const float* w, const float* beta, const float* b, size_t rows, size_t cols, float* out) {
  for (size_t n = 0; n < rows; ++n) {
    double acc = b ? static_cast<double>(b[n]) : 0.0;
    for (size_t k = 0; k < cols; ++k) acc += static_cast<double>(w[n * cols + k]) * beta[k];
    out[n] = static_cast<float>(acc);
  }
}

void build_modernbert(vrag_encoder* e, const WeightSet& w) {
  const int H = HIDDEN, I = 1152;
  e->ffn = I;
  e->max_pos = 8192;
  DevBuf staging;
  e->emb = upload_f32(e, w.get("model.embeddings.tok_embeddings.weight", (int64_t)e->vocab * H), (size_t)e->vocab * H);
  e->emb_g = upload_f32(e, w.get("model.embeddings.norm.weight", H), H);
  std::vector<float> wi_perm(static_cast<size_t>(2 * I) * H), folded(static_cast<size_t>(3 * H) * H);
  for (int i = 0; i < e->layers; ++i) {
    const std::string p = "model.layers." + std::to_string(i) + ".";
    vrag_encoder::MLayer L{};
    const float* attn_g = i > 0 ? w.get(p + "attn_norm.weight", H) : nullptr;
    const float* mlp_g = w.get(p + "mlp_norm.weight", H);
    L.attn_g = attn_g ? upload_f32(e, attn_g, H) : nullptr;
    const float* wqkv = w.get(p + "attn.Wqkv.weight", 3LL * H * H);
    if (e->deferred_ln && attn_g) {
      fold_layernorm(wqkv, attn_g, 3 * H, H, folded.data());
      wqkv = folded.data();
    }
    L.wqkv = upload_f16(e, staging, wqkv, 3 * H, H, 0, &L.wqkv_lo);
    L.wo = upload_f16(e, staging, w.get(p + "attn.Wo.weight", (int64_t)H * H), H, H, 0, &L.wo_lo);
    L.mlp_g = upload_f32(e, mlp_g, H);
    // GeGLU: Wi = [input rows 0..I) | gate rows I..2I).  Interleave per 128 so one 256-wide GEMM tile holds
    // input[128t..128t+128) and gate[128t..128t+128) -> act(input)*gate is tile-local (EPI_GEGLU).
    const float* wi = w.get(p + "mlp.Wi.weight", 2LL * I * H);
    if (e->deferred_ln) {
      fold_layernorm(wi, mlp_g, 2 * I, H, folded.data());
      wi = folded.data();
    }
    for (int t = 0; t < I / 128; ++t) {
      memcpy(&wi_perm[static_cast<size_t>(t * 256) * H], wi + static_cast<size_t>(t * 128) * H, sizeof(float) * 128 * H);
      memcpy(&wi_perm[static_cast<size_t>(t * 256 + 128) * H], wi + static_cast<size_t>(I + t * 128) * H,
             sizeof(float) * 128 * H);
    }
    L.wi = upload_f16(e, staging, wi_perm.data(), 2 * I, H, 0, &L.wi_lo);
    L.wo2 = upload_f16(e, staging, w.get(p + "mlp.Wo.weight", (int64_t)H * I), H, I, 0, &L.wo2_lo);
    e->ml.push_back(L);
  }
  e->final_g = upload_f32(e, w.get("model.final_norm.weight", H), H);
  e->cls_w = upload_f32(e, w.get("classifier.weight", 2LL * H), 2 * H);
  e->cls_b = upload_f32(e, w.get("classifier.bias", 2), 2);
  if (e->kind == VRAG_ENC_MODERNBERT_TOKCLS) {
    e->head_w = upload_f16(e, staging, w.get("head.dense.weight", (int64_t)H * H), H, H, 0, &e->head_w_lo);
    e->head_g = upload_f32(e, w.get("head.norm.weight", H), H);
    const float* hg = w.get("head.norm.weight", H);
    const float* cw = w.get("classifier.weight", 2LL * H);
    std::vector<float> gw(2 * H);
    double g0 = 0.0, g1 = 0.0;
    for (int n = 0; n < H; ++n) {
      gw[n] = hg[n] * cw[n];
      gw[H + n] = hg[n] * cw[H + n];
      g0 += static_cast<double>(gw[n]);
      g1 += static_cast<double>(gw[H + n]);
    }
    e->cls_gw = upload_f32(e, gw.data(), 2 * H);
    VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
    e->cls_g0 = static_cast<float>(g0);
    e->cls_g1 = static_cast<float>(g1);
  }
  VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
  make_rope(e, 160000.0, e->max_pos, &e->rope_g);
  make_rope(e, 10000.0, e->max_pos, &e->rope_l);
  staging.release();
}

int pad256(int n) { return (n + GEMM_BN - 1) / GEMM_BN * GEMM_BN; }

// BERT-architecture encoders at hidden size H = 768 (12 heads x 64) or H = 384 (12 heads x 32, the MiniLM family).
// The attention kernel works on 12 heads x 64 dims; narrower heads are ZERO-PADDED to 64 dims when the checkpoint is
// repacked: q | k | v rows of head h, dim d < dh go to row h * 64 + d of a [3 * 768, H] operand (rows d >= dh are zero,
// so q k^T is unchanged and the padded dims of P V come out as exact zeros), the columns of attention.output.dense are
// spread the same way ([H, 768]), and q is pre-scaled by sqrt(64 / dh) because the kernel's softmax scale is 1 / sqrt(64).
// Every other Linear keeps its shape with N padded to a multiple of 256 by zero rows (TMA stores clip at the real width).
void build_bert(vrag_encoder* e, const WeightSet& w) {
  const std::string em = "bert.embeddings.";
  auto numel = [&](const std::string& name) -> int64_t {
    auto it = w.m.find(name);
    if (it == w.m.end()) throw Error(VRAG_ERR_WEIGHTS, "missing tensor '" + name + "'");
    return it->second->numel;
  };
  const int H = static_cast<int>(numel(em + "LayerNorm.weight"));
  VRAG_CHECK(H == 768 || H == 384, VRAG_ERR_WEIGHTS, "BERT encoder: hidden size must be 768 or 384, got " + std::to_string(H));
  const int I = static_cast<int>(numel("bert.encoder.layer.0.intermediate.dense.bias"));
  VRAG_CHECK(I % GEMM_BK == 0 && I <= 4096, VRAG_ERR_WEIGHTS, "BERT encoder: unsupported intermediate size " + std::to_string(I));
  const int heads = 12, dh = H / heads, AW = heads * 64;   // AW: padded attention width = 768
  const float qscale = sqrtf(64.0f / static_cast<float>(dh));
  e->hidden = H;
  e->ffn = I;
  e->max_pos = static_cast<int>(numel(em + "position_embeddings.weight") / H);
  if (H != HIDDEN) e->deferred_ln = false;   // the deferred-LayerNorm epilogues are built for 768-wide rows (6 moment slots)
  if (e->kind == VRAG_ENC_BERT_CLS) e->deferred_ln = false;   // pair inputs: token types enter at the embedding LayerNorm
  const int Hp = pad256(H), Ip = pad256(I);
  DevBuf staging;
  const float* wemb = w.get(em + "word_embeddings.weight", (int64_t)e->vocab * H);
  e->emb = upload_f32(e, wemb, (size_t)e->vocab * H);
  e->pos_emb = upload_f32(e, w.get(em + "position_embeddings.weight", (int64_t)e->max_pos * H), (size_t)e->max_pos * H);
  e->type_emb = upload_f32(e, w.get(em + "token_type_embeddings.weight", 2LL * H), 2 * H);
  e->emb_g = upload_f32(e, w.get(em + "LayerNorm.weight", H), H);
  e->emb_b = upload_f32(e, w.get(em + "LayerNorm.bias", H), H);
  std::vector<float> cat(static_cast<size_t>(3 * AW) * H), bcat(3 * AW), folded(static_cast<size_t>(std::max(I, 3 * AW)) * H),
      cbias(std::max(I, 3 * AW)), wo_pad(static_cast<size_t>(H) * AW);
  for (int i = 0; i < e->layers; ++i) {
    const std::string p = "bert.encoder.layer." + std::to_string(i) + ".";
    vrag_encoder::BLayer L{};
    const char* nm[3] = {"query", "key", "value"};
    std::fill(cat.begin(), cat.end(), 0.f);
    std::fill(bcat.begin(), bcat.end(), 0.f);
    for (int j = 0; j < 3; ++j) {
      const float* wj = w.get(p + "attention.self." + nm[j] + ".weight", (int64_t)H * H);
      const float* bj = w.get(p + "attention.self." + nm[j] + ".bias", H);
      const float sc = j == 0 ? qscale : 1.0f;
      for (int h = 0; h < heads; ++h)
        for (int d = 0; d < dh; ++d) {
          const size_t dst = static_cast<size_t>(j * AW + h * 64 + d), src = static_cast<size_t>(h * dh + d);
          for (int k = 0; k < H; ++k) cat[dst * H + k] = wj[src * H + k] * sc;
          bcat[dst] = bj[src] * sc;
        }
    }
    // LayerNorm that produced this layer's input: the embedding LayerNorm (layer 0) or the previous output.LayerNorm
    const float* ga = i == 0 ? w.get(em + "LayerNorm.weight", H)
                             : w.get("bert.encoder.layer." + std::to_string(i - 1) + ".output.LayerNorm.weight", H);
    const float* ba = i == 0 ? w.get(em + "LayerNorm.bias", H)
                             : w.get("bert.encoder.layer." + std::to_string(i - 1) + ".output.LayerNorm.bias", H);
    if (e->deferred_ln) {
      fold_layernorm_bias(cat.data(), ba, bcat.data(), 3 * AW, H, cbias.data());
      L.cqkv = upload_f32(e, cbias.data(), 3 * AW);
      fold_layernorm(cat.data(), ga, 3 * AW, H, folded.data());
      L.wqkv = upload_f16(e, staging, folded.data(), 3 * AW, H);
    } else {
      L.wqkv = upload_f16(e, staging, cat.data(), 3 * AW, H, 0, &L.wqkv_lo);
    }
    L.bqkv = upload_f32(e, bcat.data(), 3 * AW);
    VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));  // bcat / cbias / folded reused next layer
    const float* wo = w.get(p + "attention.output.dense.weight", (int64_t)H * H);
    std::fill(wo_pad.begin(), wo_pad.end(), 0.f);
    for (int n = 0; n < H; ++n)
      for (int h = 0; h < heads; ++h)
        for (int d = 0; d < dh; ++d) wo_pad[static_cast<size_t>(n) * AW + h * 64 + d] = wo[static_cast<size_t>(n) * H + h * dh + d];
    L.wo = upload_f16(e, staging, wo_pad.data(), H, AW, Hp, &L.wo_lo);
    L.bo = upload_f32(e, w.get(p + "attention.output.dense.bias", H), H, Hp);
    L.g1 = upload_f32(e, w.get(p + "attention.output.LayerNorm.weight", H), H);
    L.b1 = upload_f32(e, w.get(p + "attention.output.LayerNorm.bias", H), H);
    const float* wi = w.get(p + "intermediate.dense.weight", (int64_t)I * H);
    const float* bi = w.get(p + "intermediate.dense.bias", I);
    const float* g1 = w.get(p + "attention.output.LayerNorm.weight", H);
    const float* b1 = w.get(p + "attention.output.LayerNorm.bias", H);
    if (e->deferred_ln) {
      fold_layernorm_bias(wi, b1, bi, I, H, cbias.data());
      L.ci = upload_f32(e, cbias.data(), I, Ip);
      fold_layernorm(wi, g1, I, H, folded.data());
      L.wi = upload_f16(e, staging, folded.data(), I, H, Ip);
      // residual GEMMs: old stream normalised with (ga, ba) before attention.output, with (g1, b1) before output
      const float* bo = w.get(p + "attention.output.dense.bias", H);
      const float* bo2 = w.get(p + "output.dense.bias", H);
      std::vector<float> t(H);
      L.ra_g = upload_f32(e, ga, H);
      for (int k = 0; k < H; ++k) t[k] = ba[k] + bo[k];
      L.ra_b = upload_f32(e, t.data(), H);
      VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
      L.rb_g = upload_f32(e, g1, H);
      for (int k = 0; k < H; ++k) t[k] = b1[k] + bo2[k];
      L.rb_b = upload_f32(e, t.data(), H);
      VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
    } else {
      L.wi = upload_f16(e, staging, wi, I, H, Ip, &L.wi_lo);
    }
    L.bi = upload_f32(e, bi, I, Ip);
    L.wo2 = upload_f16(e, staging, w.get(p + "output.dense.weight", (int64_t)H * I), H, I, Hp, &L.wo2_lo);
    L.bo2 = upload_f32(e, w.get(p + "output.dense.bias", H), H, Hp);
    L.g2 = upload_f32(e, w.get(p + "output.LayerNorm.weight", H), H);
    L.b2 = upload_f32(e, w.get(p + "output.LayerNorm.bias", H), H);
    VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));  // wo_pad reused next layer
    e->bl.push_back(L);
  }
  if (e->kind == VRAG_ENC_BERT_CLS) {   // BertForSequenceClassification, num_labels = 1 (cross-encoder reranker)
    e->pool_w = upload_f32(e, w.get("bert.pooler.dense.weight", (int64_t)H * H), (size_t)H * H);
    e->pool_b = upload_f32(e, w.get("bert.pooler.dense.bias", H), H);
    e->seq_w = upload_f32(e, w.get("classifier.weight", H), H);
    e->seq_b = upload_f32(e, w.get("classifier.bias", 1), 1);
  }
  if (e->kind == VRAG_ENC_BERT_MLM) {
    const std::string c = "cls.predictions.";
    e->mlm_w = upload_f16(e, staging, w.get(c + "transform.dense.weight", (int64_t)H * H), H, H, Hp, &e->mlm_w_lo);
    e->mlm_b = upload_f32(e, w.get(c + "transform.dense.bias", H), H, Hp);
    e->mlm_g = upload_f32(e, w.get(c + "transform.LayerNorm.weight", H), H);
    e->mlm_beta = upload_f32(e, w.get(c + "transform.LayerNorm.bias", H), H);
    // decoder is tied to the word embeddings (modeling_bert.py:471-511); pad the vocabulary to a tile multiple
    e->dec_w = upload_f16(e, staging, wemb, e->vocab, H, e->vocab_pad, &e->dec_w_lo);
    e->dec_b = upload_f32(e, w.get(c + "bias", e->vocab), e->vocab, e->vocab_pad);
  }
  VRAG_CUDA(cudaStreamSynchronize(e->ctx->stream));
  staging.release();
}

void reserve_workspace(vrag_encoder* e) {
  const size_t T = e->max_tokens, H = e->hidden, AW = HIDDEN;   // AW: attention width (12 heads x 64, padded for H = 384)
  e->ids.reserve(T * 4);
  e->pos.reserve(T * 4);
  e->seqrow.reserve(T * 4);
  e->cu.reserve((static_cast<size_t>(e->max_seqs) + 1) * 4);
  e->x32.reserve(T * H * 4);
  e->h16.reserve(T * H * 2);
  e->qkv16.reserve(T * 3 * AW * 2);
  e->o16.reserve(T * AW * 2);
  e->w16.reserve(T * e->ffn * 2);
  e->buf32.reserve(T * H * 4);
  e->probs.reserve(T * 4);
  e->logits.reserve(T * 8);
  if (e->precise) {
    e->h16_lo.reserve(T * H * 2);
    e->qkv16_lo.reserve(T * 3 * AW * 2);
    e->o16_lo.reserve(T * AW * 2);
    e->w16_lo.reserve(T * e->ffn * 2);
  }
  if (e->deferred_ln) {
    e->stats.reserve(T * 6 * 8);
    e->xl8.reserve(T * H);
    if (e->kind != VRAG_ENC_MODERNBERT_TOKCLS && e->kind != VRAG_ENC_MODERNBERT_SENT)
      e->stats2.reserve(T * 6 * 8);   // post-LN: moments are read and rewritten
  }
}

struct Pass { int s0, s1, t0, t1, max_len; };

std::vector<Pass> plan_passes(const int32_t* cu, int nseq, int max_tokens, int max_seqs) {
  std::vector<Pass> out;
  int s = 0;
  while (s < nseq) {
    int e = s, ml = 0;
    while (e < nseq && cu[e + 1] - cu[s] <= max_tokens && e - s < max_seqs) {
      ml = std::max(ml, cu[e + 1] - cu[e]);
      ++e;
    }
    VRAG_CHECK(e > s, VRAG_ERR_ARG, "a sequence is longer than the encoder's max_tokens");
    out.push_back({s, e, cu[s], cu[e], ml});
    s = e;
  }
  return out;
}

// Stage one pass' ids + cu_seqlens on the device, build pos / seq_of_row.
void stage_pass(vrag_encoder* e, const Pass& ps, const int32_t* ids, const int32_t* cu, int on_device) {
  vrag_ctx* ctx = e->ctx;
  const int T = ps.t1 - ps.t0, ns = ps.s1 - ps.s0;
  VRAG_CUDA(cudaMemcpyAsync(e->ids.p, ids + ps.t0, static_cast<size_t>(T) * 4,
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  // attention work list: one entry per (sequence, 128-query tile) of this pass
  int n_pairs = 0;
  for (int i = 0; i < ns; ++i) n_pairs += (cu[ps.s0 + i + 1] - cu[ps.s0 + i] + 127) / 128;
  int32_t* cu_h = static_cast<int32_t*>(ctx->pinned_reserve((static_cast<size_t>(ns) + 1 + 4 * static_cast<size_t>(n_pairs)) * 4));
  for (int i = 0; i <= ns; ++i) cu_h[i] = cu[ps.s0 + i] - ps.t0;
  int32_t* work_h = cu_h + ns + 1;
  for (int i = 0, w = 0; i < ns; ++i)
    for (int q0 = 0; q0 < cu_h[i + 1] - cu_h[i]; q0 += 128, ++w) {  // 16-byte entries {s0, L, q0, -}
      work_h[4 * w] = cu_h[i];
      work_h[4 * w + 1] = cu_h[i + 1] - cu_h[i];
      work_h[4 * w + 2] = q0;
      work_h[4 * w + 3] = 0;
    }
  e->work.reserve(static_cast<size_t>(4 * n_pairs) * 4);
  e->n_pairs = n_pairs;
  if (e->type_ids_host) {
    e->types.reserve(static_cast<size_t>(T) * 4);
    VRAG_CUDA(cudaMemcpyAsync(e->types.p, e->type_ids_host + ps.t0, static_cast<size_t>(T) * 4,
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  }
  VRAG_CUDA(cudaMemcpyAsync(e->cu.p, cu_h, (static_cast<size_t>(ns) + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  VRAG_CUDA(cudaMemcpyAsync(e->work.p, work_h, static_cast<size_t>(4 * n_pairs) * 4, cudaMemcpyHostToDevice, ctx->stream));
  VRAG_CUDA(cudaStreamSynchronize(ctx->stream));  // cu_h (pinned scratch) is reused by the next pass
  launch_token_meta(ctx, e->cu.as<int32_t>(), ns, T, e->pos.as<int32_t>(), e->seqrow.as<int32_t>());
}

// final_hidden_only: stop after the final LayerNorm and leave its fp32 output in x32 (sentence head of the legacy QAModel)
void modernbert_pass(vrag_encoder* e, const Pass& ps, float* hidden_dbg_host, bool final_hidden_only = false) {
  vrag_ctx* ctx = e->ctx;
  const int T = ps.t1 - ps.t0, ns = ps.s1 - ps.s0, H = HIDDEN, I = e->ffn;
  const int ref = e->use_reference_gemm ? 1 : 0;
  float* x32 = e->x32.as<float>();
  __half *h16 = e->h16.as<__half>(), *qkv = e->qkv16.as<__half>(), *o16 = e->o16.as<__half>(), *g16 = e->w16.as<__half>();
  // Deferred LayerNorm: the residual stream is the pair of planes (h16 = hi, fp16; xl8 = lo, e5m2); x32 is only
  // materialised for the debug dump.
  const bool dln = e->deferred_ln;
  uint8_t* xl8 = e->xl8.as<uint8_t>();
  auto dump = [&](int slot) {
    if (!hidden_dbg_host) return;
    if (dln) launch_hilo_to_f32(ctx, h16, xl8, T, x32);
    VRAG_CUDA(cudaMemcpyAsync(hidden_dbg_host + static_cast<size_t>(slot) * T * H, x32,
                              static_cast<size_t>(T) * H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  };
  // h16 always holds the A operand of the next Wqkv / Wi GEMM: LN(x) (plain path), or the hi plane of the raw
  // residual stream (deferred LayerNorm: row moments in `stats`, gamma and the mean folded into the weights).
  float* stats = e->stats.as<float>();
  // precise mode (dln is off): low planes of every fp16 activation; nullptr selects the plain fp16 kernels
  const bool pr = e->precise;
  __half* h16_lo = pr ? e->h16_lo.as<__half>() : nullptr;
  __half* qkv_lo = pr ? e->qkv16_lo.as<__half>() : nullptr;
  __half* o16_lo = pr ? e->o16_lo.as<__half>() : nullptr;
  __half* g16_lo = pr ? e->w16_lo.as<__half>() : nullptr;
  launch_embed_ln(ctx, e->ids.as<int32_t>(), T, e->vocab, e->emb, e->emb_g, 1e-5f, dln ? nullptr : x32, h16,
                  dln ? xl8 : nullptr, h16_lo);
  dump(0);
  for (int i = 0; i < e->layers; ++i) {
    const auto& L = e->ml[i];
    const bool global = (i % 3) == 0;
    if (i > 0 && !dln) launch_layernorm(ctx, x32, T, L.attn_g, nullptr, 1e-5f, h16, false, h16_lo);
    GemmEpiParams p;
    p.M = T; p.out16 = qkv; p.ld16 = 3 * H; p.hidden = H; p.pos = e->pos.as<int32_t>();
    p.rope_tab = global ? e->rope_g : e->rope_l;
    p.rope_rows = e->max_pos;
    p.stats_in = stats;
    p.a_lo = h16_lo; p.w_lo = pr ? L.wqkv_lo : nullptr; p.out16_lo = qkv_lo;
    launch_gemm(ctx, (dln && i > 0) ? EPI_NORM_ROPE_QKV : EPI_ROPE_QKV, h16, L.wqkv, T, 3 * H, H, p, ref);
    e->attention(qkv, o16, ns, T, ps.max_len, global ? -1 : 64);
    GemmEpiParams r;
    r.M = T; r.out32 = x32; r.ld32 = H; r.out16 = h16; r.out8_lo = xl8; r.ld16 = H; r.stats_out = stats;
    r.a_lo = o16_lo; r.w_lo = pr ? L.wo_lo : nullptr;
    launch_gemm(ctx, dln ? EPI_RESID_STATS : EPI_RESID_F32, o16, L.wo, T, H, H, r, ref);
    if (!dln) launch_layernorm(ctx, x32, T, L.mlp_g, nullptr, 1e-5f, h16, false, h16_lo);
    GemmEpiParams g;
    g.M = T; g.out16 = g16; g.ld16 = I; g.stats_in = stats;
    g.a_lo = h16_lo; g.w_lo = pr ? L.wi_lo : nullptr; g.out16_lo = g16_lo;
    launch_gemm(ctx, dln ? EPI_NORM_GEGLU : EPI_GEGLU, h16, L.wi, T, 2 * I, H, g, ref);
    r.a_lo = g16_lo; r.w_lo = pr ? L.wo2_lo : nullptr;
    launch_gemm(ctx, dln ? EPI_RESID_STATS : EPI_RESID_F32, g16, L.wo2, T, H, I, r, ref);
    dump(i + 1);
  }
  if (final_hidden_only) {
    if (dln) launch_layernorm_hilo(ctx, h16, xl8, T, e->final_g, nullptr, 1e-5f, x32, h16);
    else launch_layernorm(ctx, x32, T, e->final_g, nullptr, 1e-5f, h16, true, h16_lo);
    return;
  }
  if (dln) launch_layernorm_hilo(ctx, h16, xl8, T, e->final_g, nullptr, 1e-5f, nullptr, h16);
  else launch_layernorm(ctx, x32, T, e->final_g, nullptr, 1e-5f, h16, false, h16_lo);
  GemmEpiParams hd;
  hd.M = T; hd.out32 = e->buf32.as<float>(); hd.ld32 = H;
  hd.a_lo = h16_lo; hd.w_lo = pr ? e->head_w_lo : nullptr;
  if (e->fused_head) {
    // span-logit head in the GEMM epilogue: gelu + the LayerNorm / classifier partial sums, 96 B per token instead of
    // a 6 KB fp32 round trip (buf32 doubles as the [6][T] float4 partial-sum buffer)
    hd.cls_gw = e->cls_gw; hd.head_part = e->buf32.as<float>(); hd.hidden = H;
    launch_gemm(ctx, EPI_HEAD_PARTIAL, h16, e->head_w, T, H, H, hd, ref);
    launch_head_finish(ctx, e->buf32.as<float>(), T, H / 128, 1e-5f, e->cls_g0, e->cls_g1, e->cls_b, e->logits.as<float>(),
                       e->probs.as<float>());
  } else {
    launch_gemm(ctx, EPI_GELU_F32, h16, e->head_w, T, H, H, hd, ref);
    launch_head_final(ctx, e->buf32.as<float>(), T, e->head_g, 1e-5f, e->cls_w, e->cls_b, e->logits.as<float>(),
                      e->probs.as<float>());
  }
}

// BERT encoder stack: leaves the post-LN final hidden states in x32 (fp32) and h16 (fp16).
void bert_stack(vrag_encoder* e, const Pass& ps) {
  vrag_ctx* ctx = e->ctx;
  const int T = ps.t1 - ps.t0, ns = ps.s1 - ps.s0, H = e->hidden, I = e->ffn, AW = HIDDEN;
  const int Hp = pad256(H), Ip = pad256(I);   // GEMM N (weights zero-padded); outputs keep their real widths
  const int ref = e->use_reference_gemm ? 1 : 0;
  float* x32 = e->x32.as<float>();
  __half *h16 = e->h16.as<__half>(), *qkv = e->qkv16.as<__half>(), *o16 = e->o16.as<__half>(), *f16 = e->w16.as<__half>();
  if (e->deferred_ln) {
    // The stream holds the PRE-LayerNorm sums z (fp16 hi = the consumers' A operand, e5m2 lo, row moments); every
    // LayerNorm of the stack lives in folded weights / biases and in the epilogues (gemm.cuh EPI_NORM_BIAS_* and
    // EPI_RESID_STATS_LN).  Moments ping-pong between two buffers: a residual GEMM reads the old moments of a row in
    // every N tile while other tiles already write the new ones.
    uint8_t* xl8 = e->xl8.as<uint8_t>();
    float* st[2] = {e->stats.as<float>(), e->stats2.as<float>()};
    int cur = 0;
    launch_bert_embed_raw(ctx, e->ids.as<int32_t>(), e->pos.as<int32_t>(), T, e->vocab, e->max_pos, e->emb, e->pos_emb,
                          e->type_emb, h16, xl8, st[cur], 6);
    for (int i = 0; i < e->layers; ++i) {
      const auto& L = e->bl[i];
      GemmEpiParams p;
      p.M = T; p.out16 = qkv; p.ld16 = 3 * H; p.bias = L.cqkv; p.stats_in = st[cur]; p.ln_eps = 1e-12f;
      launch_gemm(ctx, EPI_NORM_BIAS_F16, h16, L.wqkv, T, 3 * H, H, p, ref);
      e->attention(qkv, o16, ns, T, ps.max_len, -1);
      GemmEpiParams r;
      r.M = T; r.out16 = h16; r.out8_lo = xl8; r.ld16 = H; r.ln_eps = 1e-12f;
      r.gamma = L.ra_g; r.bias = L.ra_b; r.stats_in = st[cur]; r.stats_out = st[cur ^ 1];
      launch_gemm(ctx, EPI_RESID_STATS_LN, o16, L.wo, T, H, H, r, ref);
      cur ^= 1;
      GemmEpiParams f;
      f.M = T; f.out16 = f16; f.ld16 = I; f.bias = L.ci; f.stats_in = st[cur]; f.ln_eps = 1e-12f;
      launch_gemm(ctx, EPI_NORM_BIAS_GELU_F16, h16, L.wi, T, I, H, f, ref);
      r.gamma = L.rb_g; r.bias = L.rb_b; r.stats_in = st[cur]; r.stats_out = st[cur ^ 1];
      launch_gemm(ctx, EPI_RESID_STATS_LN, f16, L.wo2, T, H, I, r, ref);
      cur ^= 1;
    }
    const auto& last = e->bl.back();
    launch_layernorm_hilo(ctx, h16, xl8, T, last.g2, last.b2, 1e-12f, x32, h16);
    return;
  }
  // precise mode: low planes of every fp16 activation (nullptr selects the plain fp16 kernels)
  const bool pr = e->precise;
  __half* h16_lo = pr ? e->h16_lo.as<__half>() : nullptr;
  __half* qkv_lo = pr ? e->qkv16_lo.as<__half>() : nullptr;
  __half* o16_lo = pr ? e->o16_lo.as<__half>() : nullptr;
  __half* f16_lo = pr ? e->w16_lo.as<__half>() : nullptr;
  launch_bert_embed_ln(ctx, e->ids.as<int32_t>(), e->pos.as<int32_t>(), T, e->vocab, e->max_pos, e->emb, e->pos_emb,
                       e->type_emb, e->emb_g, e->emb_b, 1e-12f, x32, h16, h16_lo, H,
                       e->type_ids_host ? e->types.as<int32_t>() : nullptr);
  for (int i = 0; i < e->layers; ++i) {
    const auto& L = e->bl[i];
    GemmEpiParams p;
    p.M = T; p.out16 = qkv; p.ld16 = 3 * AW; p.bias = L.bqkv;
    p.a_lo = h16_lo; p.w_lo = pr ? L.wqkv_lo : nullptr; p.out16_lo = qkv_lo;
    launch_gemm(ctx, EPI_BIAS_F16, h16, L.wqkv, T, 3 * AW, H, p, ref);
    e->attention(qkv, o16, ns, T, ps.max_len, -1);
    GemmEpiParams r;
    r.M = T; r.out32 = x32; r.ld32 = H; r.bias = L.bo; r.n_valid = H;
    r.a_lo = o16_lo; r.w_lo = pr ? L.wo_lo : nullptr;
    launch_gemm(ctx, EPI_BIAS_RESID_F32, o16, L.wo, T, Hp, AW, r, ref);
    launch_layernorm(ctx, x32, T, L.g1, L.b1, 1e-12f, h16, true, h16_lo, H);
    GemmEpiParams f;
    f.M = T; f.out16 = f16; f.ld16 = I; f.bias = L.bi; f.n_valid = I;
    f.a_lo = h16_lo; f.w_lo = pr ? L.wi_lo : nullptr; f.out16_lo = f16_lo;
    launch_gemm(ctx, EPI_BIAS_GELU_F16, h16, L.wi, T, Ip, H, f, ref);
    GemmEpiParams r2;
    r2.M = T; r2.out32 = x32; r2.ld32 = H; r2.bias = L.bo2; r2.n_valid = H;
    r2.a_lo = f16_lo; r2.w_lo = pr ? L.wo2_lo : nullptr;
    launch_gemm(ctx, EPI_BIAS_RESID_F32, f16, L.wo2, T, Hp, I, r2, ref);
    launch_layernorm(ctx, x32, T, L.g2, L.b2, 1e-12f, h16, true, h16_lo, H);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define VRAG_API_BEGIN(ctxp)                 \
  vrag_ctx* _ctx = (ctxp);                   \
  std::lock_guard<std::mutex> _lk(_ctx->mu); \
  try {                                      \
    VRAG_CUDA(cudaSetDevice(_ctx->device));
#define VRAG_API_END()                                    \
    return VRAG_OK;                                       \
  } catch (const vrag::Error& e) {                        \
    _ctx->last_error = e.what();                          \
    return e.code;                                        \
  } catch (const std::exception& e) {                     \
    _ctx->last_error = e.what();                          \
    return VRAG_ERR_INTERNAL;                             \
  }

extern "C" int vrag_encoder_create(vrag_ctx* ctx, int kind, int num_layers, int vocab_size, int max_tokens,
                                   const vrag_tensor* tensors, int num_tensors, vrag_encoder** out) {
  return vrag_encoder_create_ex(ctx, kind, num_layers, vocab_size, max_tokens, tensors, num_tensors, VRAG_PRECISION_FAST,
                                out);
}

extern "C" int vrag_encoder_create_ex(vrag_ctx* ctx, int kind, int num_layers, int vocab_size, int max_tokens,
                                      const vrag_tensor* tensors, int num_tensors, int precision, vrag_encoder** out) {
  if (!ctx || !out) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(ctx)
  VRAG_CHECK(kind >= 0 && kind <= 4 && num_layers > 0 && vocab_size > 0 && max_tokens >= 128, VRAG_ERR_ARG,
             "encoder_create: bad kind / num_layers / vocab_size / max_tokens");
  VRAG_CHECK(precision == VRAG_PRECISION_FAST || precision == VRAG_PRECISION_PRECISE, VRAG_ERR_ARG,
             "encoder_create: precision must be VRAG_PRECISION_FAST or VRAG_PRECISION_PRECISE");
  std::unique_ptr<vrag_encoder> e(new vrag_encoder());
  e->ctx = ctx;
  e->kind = kind;
  e->precise = precision == VRAG_PRECISION_PRECISE;
  e->layers = num_layers;
  e->vocab = vocab_size;
  e->vocab_pad = (vocab_size + GEMM_BN - 1) / GEMM_BN * GEMM_BN;
  e->max_tokens = max_tokens;
  e->max_seqs = std::max(64, max_tokens / 8);
  const char* dbg = getenv("VRAG_GEMM_REFERENCE");
  e->use_reference_gemm = dbg && dbg[0] == '1';
  const char* leg = getenv("VRAG_ATTENTION_LEGACY");
  e->legacy_attention = leg && leg[0] == '1';
  const char* dl = getenv("VRAG_DEFERRED_LN");   // "0": separate LayerNorm kernels (cross-check path)
  e->deferred_ln = !(dl && dl[0] == '0');
  const bool modern = kind == VRAG_ENC_MODERNBERT_TOKCLS || kind == VRAG_ENC_MODERNBERT_SENT;
  if (!modern) {   // BERT (post-LN) stacks: VRAG_BERT_DEFERRED_LN=0 selects the cross-check path
    const char* bd = getenv("VRAG_BERT_DEFERRED_LN");   // (fp32 stream by TMA reduce-add + LayerNorm kernels)
    e->deferred_ln = e->deferred_ln && !(bd && bd[0] == '0');
  }
  const char* fh = getenv("VRAG_FUSED_HEAD");
  e->fused_head = !(fh && fh[0] == '0');
  if (e->precise) {   // fp32 residual stream + LayerNorm kernels; the split GEMMs have no deferred-LayerNorm epilogues
    e->deferred_ln = false;
    e->legacy_attention = false;
  }
  WeightSet w(tensors, num_tensors);
  if (modern) build_modernbert(e.get(), w);
  else build_bert(e.get(), w);
  reserve_workspace(e.get());
  VRAG_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = e.release();
  VRAG_API_END()
}

extern "C" int vrag_encoder_hidden(vrag_encoder* enc) { return enc ? enc->hidden : -1; }

extern "C" void vrag_encoder_destroy(vrag_encoder* enc) {
  if (!enc) return;
  cudaSetDevice(enc->ctx->device);
  cudaStreamSynchronize(enc->ctx->stream);
  delete enc;
}

static int span_forward_impl(vrag_encoder* enc, const int32_t* ids, const int32_t* cu, int nseq, float* probs_out,
                             float* logits_out, int on_device, float* hidden_dbg) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_MODERNBERT_TOKCLS, VRAG_ERR_ARG, "span_forward needs a MODERNBERT_TOKCLS encoder");
  VRAG_CHECK(ids && cu && probs_out && nseq >= 0, VRAG_ERR_ARG, "span_forward: null argument");
  if (nseq == 0) return VRAG_OK;
  for (int i = 0; i < nseq; ++i) {
    VRAG_CHECK(cu[i + 1] > cu[i], VRAG_ERR_ARG, "span_forward: empty sequence");
    VRAG_CHECK(cu[i + 1] - cu[i] <= enc->max_pos, VRAG_ERR_ARG, "span_forward: sequence longer than 8192 tokens");
  }
  auto passes = plan_passes(cu, nseq, enc->max_tokens, enc->max_seqs);
  VRAG_CHECK(!hidden_dbg || passes.size() == 1, VRAG_ERR_ARG, "debug hidden dump needs a single pass");
  const cudaMemcpyKind back = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  for (const Pass& ps : passes) {
    stage_pass(enc, ps, ids, cu, on_device);
    modernbert_pass(enc, ps, hidden_dbg);
    const size_t T = ps.t1 - ps.t0;
    VRAG_CUDA(cudaMemcpyAsync(probs_out + ps.t0, enc->probs.p, T * 4, back, _ctx->stream));
    if (logits_out) VRAG_CUDA(cudaMemcpyAsync(logits_out + 2 * static_cast<size_t>(ps.t0), enc->logits.p, T * 8, back, _ctx->stream));
    if (!on_device) VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  }
  VRAG_API_END()
}

extern "C" int vrag_span_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq,
                                 float* probs_out, float* logits_out, int on_device) {
  return span_forward_impl(enc, ids, cu_seqlens, nseq, probs_out, logits_out, on_device, nullptr);
}

// Debug hook (tests only): host ids, also returns the fp32 residual stream after the embedding and after each
// layer, hidden_out [layers + 1, T, 768] host.
extern "C" int vrag_debug_span_hidden(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq,
                                      float* probs_out, float* logits_out, float* hidden_out) {
  return span_forward_impl(enc, ids, cu_seqlens, nseq, probs_out, logits_out, 0, hidden_out);
}

extern "C" int vrag_splade_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu, int nseq, float min_abs,
                                   int64_t* indptr_out, int32_t* indices_out, float* values_out, int64_t cap,
                                   int64_t* nnz_out, float* dense_out, int on_device) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_BERT_MLM, VRAG_ERR_ARG, "splade_forward needs a BERT_MLM encoder");
  VRAG_CHECK(ids && cu && nseq >= 0 && nnz_out, VRAG_ERR_ARG, "splade_forward: null argument");
  const bool want_csr = indptr_out != nullptr;
  VRAG_CHECK(!want_csr || (indices_out && values_out) || cap == 0, VRAG_ERR_ARG, "splade_forward: null CSR buffers");
  *nnz_out = 0;
  if (want_csr) indptr_out[0] = 0;
  if (nseq == 0) return VRAG_OK;
  for (int i = 0; i < nseq; ++i) {
    VRAG_CHECK(cu[i + 1] > cu[i], VRAG_ERR_ARG, "splade_forward: empty sequence");
    VRAG_CHECK(cu[i + 1] - cu[i] <= enc->max_pos, VRAG_ERR_ARG, "splade_forward: sequence longer than 512 tokens");
  }
  const int H = enc->hidden, V = enc->vocab, VP = enc->vocab_pad;
  const int ref = enc->use_reference_gemm ? 1 : 0;
  // the dense pooled buffer is [seqs_per_pass, VP] fp32: bound sequences per pass to 512 (60 MB)
  auto passes = plan_passes(cu, nseq, enc->max_tokens, std::min(enc->max_seqs, 512));
  int64_t nnz_total = 0;
  bool overflow = false;
  for (const Pass& ps : passes) {
    const int T = ps.t1 - ps.t0, ns = ps.s1 - ps.s0;
    stage_pass(enc, ps, ids, cu, on_device);
    bert_stack(enc, ps);
    GemmEpiParams t;
    __half* h16_lo = enc->precise ? enc->h16_lo.as<__half>() : nullptr;
    t.M = T; t.out32 = enc->buf32.as<float>(); t.ld32 = H; t.bias = enc->mlm_b; t.n_valid = H;
    t.a_lo = h16_lo; t.w_lo = enc->precise ? enc->mlm_w_lo : nullptr;
    launch_gemm(_ctx, EPI_BIAS_GELU_F32, enc->h16.as<__half>(), enc->mlm_w, T, pad256(H), H, t, ref);
    launch_layernorm(_ctx, enc->buf32.as<float>(), T, enc->mlm_g, enc->mlm_beta, 1e-12f, enc->h16.as<__half>(), false,
                     h16_lo, H);
    enc->splade.reserve(static_cast<size_t>(ns) * VP * 4);
    VRAG_CUDA(cudaMemsetAsync(enc->splade.p, 0, static_cast<size_t>(ns) * VP * 4, _ctx->stream));
    GemmEpiParams sp;
    sp.M = T; sp.bias = enc->dec_b; sp.seq_of_row = enc->seqrow.as<int32_t>(); sp.splade_out = enc->splade.as<float>();
    sp.splade_ld = VP; sp.n_valid = V;
    sp.a_lo = h16_lo; sp.w_lo = enc->precise ? enc->dec_w_lo : nullptr;
    launch_gemm(_ctx, EPI_SPLADE, enc->h16.as<__half>(), enc->dec_w, T, VP, H, sp, ref);
    if (dense_out)
      VRAG_CUDA(cudaMemcpy2DAsync(dense_out + static_cast<size_t>(ps.s0) * V, static_cast<size_t>(V) * 4, enc->splade.p,
                                  static_cast<size_t>(VP) * 4, static_cast<size_t>(V) * 4, ns,
                                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, _ctx->stream));
    if (want_csr) {
      enc->counts.reserve(static_cast<size_t>(ns) * 4);
      enc->indptr.reserve((static_cast<size_t>(ns) + 1) * 8);
      launch_splade_count(_ctx, enc->splade.as<float>(), ns, VP, V, min_abs, enc->counts.as<int32_t>());
      std::vector<int32_t> cnt(ns);
      VRAG_CUDA(cudaMemcpyAsync(cnt.data(), enc->counts.p, static_cast<size_t>(ns) * 4, cudaMemcpyDeviceToHost, _ctx->stream));
      VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
      std::vector<int64_t> ip(ns + 1);
      ip[0] = 0;
      for (int i = 0; i < ns; ++i) ip[i + 1] = ip[i] + cnt[i];
      for (int i = 0; i < ns; ++i) indptr_out[ps.s0 + i + 1] = nnz_total + ip[i + 1];
      const int64_t n_here = ip[ns];
      if (!overflow && nnz_total + n_here <= cap) {
        if (n_here > 0) {
          enc->sp_idx.reserve(static_cast<size_t>(n_here) * 4);
          enc->sp_val.reserve(static_cast<size_t>(n_here) * 4);
          VRAG_CUDA(cudaMemcpyAsync(enc->indptr.p, ip.data(), (static_cast<size_t>(ns) + 1) * 8, cudaMemcpyHostToDevice, _ctx->stream));
          launch_splade_fill(_ctx, enc->splade.as<float>(), ns, VP, V, min_abs, enc->indptr.as<int64_t>(),
                             enc->sp_idx.as<int32_t>(), enc->sp_val.as<float>());
          VRAG_CUDA(cudaMemcpyAsync(indices_out + nnz_total, enc->sp_idx.p, static_cast<size_t>(n_here) * 4, cudaMemcpyDeviceToHost, _ctx->stream));
          VRAG_CUDA(cudaMemcpyAsync(values_out + nnz_total, enc->sp_val.p, static_cast<size_t>(n_here) * 4, cudaMemcpyDeviceToHost, _ctx->stream));
        }
      } else {
        overflow = true;
      }
      nnz_total += n_here;
    }
    VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  }
  *nnz_out = nnz_total;
  if (overflow) throw Error(VRAG_ERR_CAPACITY, "splade_forward: CSR capacity too small; see *nnz_out");
  VRAG_API_END()
}

extern "C" int vrag_dense_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu, int nseq, int pooling,
                                  int normalize, float* out, int on_device) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_BERT_DENSE || enc->kind == VRAG_ENC_BERT_MLM, VRAG_ERR_ARG,
             "dense_forward needs a BERT encoder");
  VRAG_CHECK(ids && cu && out && nseq >= 0, VRAG_ERR_ARG, "dense_forward: null argument");
  VRAG_CHECK(pooling == VRAG_POOL_MEAN || pooling == VRAG_POOL_CLS, VRAG_ERR_ARG, "dense_forward: bad pooling");
  if (nseq == 0) return VRAG_OK;
  for (int i = 0; i < nseq; ++i) {
    VRAG_CHECK(cu[i + 1] > cu[i], VRAG_ERR_ARG, "dense_forward: empty sequence");
    VRAG_CHECK(cu[i + 1] - cu[i] <= enc->max_pos, VRAG_ERR_ARG, "dense_forward: sequence longer than 512 tokens");
  }
  auto passes = plan_passes(cu, nseq, enc->max_tokens, enc->max_seqs);
  for (const Pass& ps : passes) {
    const int ns = ps.s1 - ps.s0;
    stage_pass(enc, ps, ids, cu, on_device);
    bert_stack(enc, ps);
    const int H = enc->hidden;
    enc->pooled.reserve(static_cast<size_t>(ns) * H * 4);
    launch_pool(_ctx, enc->x32.as<float>(), enc->cu.as<int32_t>(), ns, pooling, normalize, enc->pooled.as<float>(), H);
    VRAG_CUDA(cudaMemcpyAsync(out + static_cast<size_t>(ps.s0) * H, enc->pooled.p, static_cast<size_t>(ns) * H * 4,
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, _ctx->stream));
    VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  }
  VRAG_API_END()
}

// Cross-encoder reranker (reference: SentenceTransformersReranker.rerank -> CrossEncoder.predict, verbatim_rag/rerankers.py:
// 109-134): pair sequences [CLS] q [SEP] doc [SEP] with token types -> BERT stack -> pooler (tanh) -> 1-label classifier.
extern "C" int vrag_rerank_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* type_ids, const int32_t* cu,
                                   int nseq, float* scores_out, int on_device) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_BERT_CLS, VRAG_ERR_ARG, "rerank_forward needs a BERT_CLS encoder");
  VRAG_CHECK(ids && cu && scores_out && nseq >= 0, VRAG_ERR_ARG, "rerank_forward: null argument");
  if (nseq == 0) return VRAG_OK;
  for (int i = 0; i < nseq; ++i) {
    VRAG_CHECK(cu[i + 1] > cu[i], VRAG_ERR_ARG, "rerank_forward: empty sequence");
    VRAG_CHECK(cu[i + 1] - cu[i] <= enc->max_pos, VRAG_ERR_ARG, "rerank_forward: sequence longer than the position table");
  }
  auto passes = plan_passes(cu, nseq, enc->max_tokens, enc->max_seqs);
  enc->type_ids_host = type_ids;
  try {
    for (const Pass& ps : passes) {
      const int ns = ps.s1 - ps.s0;
      stage_pass(enc, ps, ids, cu, on_device);
      bert_stack(enc, ps);
      enc->pooled.reserve(static_cast<size_t>(ns) * 4);
      launch_cls_head(_ctx, enc->x32.as<float>(), enc->cu.as<int32_t>(), ns, enc->hidden, enc->pool_w, enc->pool_b,
                      enc->seq_w, enc->seq_b, enc->pooled.as<float>());
      VRAG_CUDA(cudaMemcpyAsync(scores_out + ps.s0, enc->pooled.p, static_cast<size_t>(ns) * 4,
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, _ctx->stream));
      VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
    }
  } catch (...) {
    enc->type_ids_host = nullptr;
    throw;
  }
  enc->type_ids_host = nullptr;
  VRAG_API_END()
}

// Sentence classifier of the legacy QAModel (reference: QAModel.forward, packages/core/verbatim_core/extractor_models/
// model.py:59-117, called from ModelSpanExtractor._extract_qa_model, extractors.py:230-283): encoder -> mean of the final
// hidden states over each sentence's token range -> Linear(hidden, 2).  sent_indptr [nseq + 1] (host) counts the
// sentences of each sequence; sent_start / sent_end are token indices INSIDE the sequence, end inclusive (the
// reference's sentence_boundaries).  logits_out [n_sentences, 2], host.
extern "C" int vrag_sentence_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu, int nseq,
                                     const int32_t* sent_indptr, const int32_t* sent_start, const int32_t* sent_end,
                                     float* logits_out) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_MODERNBERT_SENT, VRAG_ERR_ARG, "sentence_forward needs a MODERNBERT_SENT encoder");
  VRAG_CHECK(ids && cu && sent_indptr && nseq >= 0, VRAG_ERR_ARG, "sentence_forward: null argument");
  if (nseq == 0) return VRAG_OK;
  for (int i = 0; i < nseq; ++i) {
    const int L = cu[i + 1] - cu[i];
    VRAG_CHECK(L > 0 && L <= enc->max_pos, VRAG_ERR_ARG, "sentence_forward: bad sequence length");
    for (int j = sent_indptr[i]; j < sent_indptr[i + 1]; ++j)
      VRAG_CHECK(sent_start[j] >= 0 && sent_end[j] >= sent_start[j] && sent_end[j] < L, VRAG_ERR_ARG,
                 "sentence_forward: sentence boundary outside its sequence");
  }
  auto passes = plan_passes(cu, nseq, enc->max_tokens, enc->max_seqs);
  for (const Pass& ps : passes) {
    stage_pass(enc, ps, ids, cu, 0);
    modernbert_pass(enc, ps, nullptr, true);
    const int j0 = sent_indptr[ps.s0], j1 = sent_indptr[ps.s1], nsent = j1 - j0;
    if (nsent > 0) {
      std::vector<int32_t> r0(nsent), r1(nsent);
      for (int i = ps.s0; i < ps.s1; ++i)
        for (int j = sent_indptr[i]; j < sent_indptr[i + 1]; ++j) {
          r0[j - j0] = cu[i] - ps.t0 + sent_start[j];
          r1[j - j0] = cu[i] - ps.t0 + sent_end[j];
        }
      enc->srow0.reserve(static_cast<size_t>(nsent) * 4);
      enc->srow1.reserve(static_cast<size_t>(nsent) * 4);
      enc->slog.reserve(static_cast<size_t>(nsent) * 8);
      VRAG_CUDA(cudaMemcpyAsync(enc->srow0.p, r0.data(), static_cast<size_t>(nsent) * 4, cudaMemcpyHostToDevice, _ctx->stream));
      VRAG_CUDA(cudaMemcpyAsync(enc->srow1.p, r1.data(), static_cast<size_t>(nsent) * 4, cudaMemcpyHostToDevice, _ctx->stream));
      launch_sentence_head(_ctx, enc->x32.as<float>(), enc->hidden, enc->srow0.as<int32_t>(), enc->srow1.as<int32_t>(), nsent,
                           enc->cls_w, enc->cls_b, enc->slog.as<float>());
      VRAG_CUDA(cudaMemcpyAsync(logits_out + 2 * static_cast<size_t>(j0), enc->slog.p, static_cast<size_t>(nsent) * 8,
                                cudaMemcpyDeviceToHost, _ctx->stream));
    }
    VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));   // r0 / r1 are stack-owned
  }
  VRAG_API_END()
}

// Span extraction with the post-processing on the device (reference: the whole of model.process(),
// extractors.py:213-224): forward, then per sequence threshold -> runs -> char spans -> gap merge -> min length over its
// context tokens [ctx_first[i], ctx_first[i] + ctx_len[i]) with the character offsets tok_char_start / tok_char_end
// (concatenated over the sequences).  Only the spans (24 bytes each) come back instead of the probabilities (4 bytes per
// token); outputs as in vrag_spans_from_probs with span_seq = sequence index.  Single-window inputs only (one sequence
// per context); documents that need several windows go through vrag_span_forward + vrag_spans_from_probs.
extern "C" int vrag_span_extract(vrag_encoder* enc, const int32_t* ids, const int32_t* cu, int nseq,
                                 const int32_t* ctx_first, const int32_t* ctx_len, const int32_t* tok_char_start,
                                 const int32_t* tok_char_end, float threshold, int min_span_chars, int merge_gap_chars,
                                 int32_t* span_seq, int32_t* span_char_start, int32_t* span_char_end, float* span_score,
                                 int32_t* span_tok_start, int32_t* span_tok_end, int64_t cap, int64_t* nspans_out) {
  if (!enc) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(enc->ctx)
  VRAG_CHECK(enc->kind == VRAG_ENC_MODERNBERT_TOKCLS, VRAG_ERR_ARG, "span_extract needs a MODERNBERT_TOKCLS encoder");
  VRAG_CHECK(ids && cu && ctx_first && ctx_len && nspans_out && nseq >= 0, VRAG_ERR_ARG, "span_extract: null argument");
  *nspans_out = 0;
  if (nseq == 0) return VRAG_OK;
  std::vector<int64_t> tok_base(nseq + 1, 0);
  for (int i = 0; i < nseq; ++i) {
    const int L = cu[i + 1] - cu[i];
    VRAG_CHECK(L > 0 && L <= enc->max_pos, VRAG_ERR_ARG, "span_extract: bad sequence length");
    VRAG_CHECK(ctx_first[i] >= 0 && ctx_len[i] >= 0 && ctx_first[i] + ctx_len[i] <= L, VRAG_ERR_ARG,
               "span_extract: context range outside its sequence");
    tok_base[i + 1] = tok_base[i] + ctx_len[i];
  }
  VRAG_CHECK(tok_base[nseq] == 0 || (tok_char_start && tok_char_end), VRAG_ERR_ARG, "span_extract: null offsets");
  auto passes = plan_passes(cu, nseq, enc->max_tokens, enc->max_seqs);
  int64_t total = 0;
  for (const Pass& ps : passes) {
    const int ns = ps.s1 - ps.s0;
    const int64_t ntok = tok_base[ps.s1] - tok_base[ps.s0];
    stage_pass(enc, ps, ids, cu, 0);
    modernbert_pass(enc, ps, nullptr);
    // per-pass metadata: [ctx_first ns | ctx_len ns] int32, tok_base ns int64 (relative to this pass), offsets 2 x ntok int32
    const size_t meta_bytes = static_cast<size_t>(ns) * 16 + static_cast<size_t>(ntok) * 8;
    enc->sp_meta.reserve(std::max<size_t>(meta_bytes, 16));
    uint8_t* mp = enc->sp_meta.as<uint8_t>();
    int64_t* d_base = reinterpret_cast<int64_t*>(mp);
    int32_t* d_first = reinterpret_cast<int32_t*>(mp + static_cast<size_t>(ns) * 8);
    int32_t* d_len = d_first + ns;
    int32_t* d_tcs = d_len + ns;
    int32_t* d_tce = d_tcs + ntok;
    std::vector<int64_t> rel(ns);
    for (int i = 0; i < ns; ++i) rel[i] = tok_base[ps.s0 + i] - tok_base[ps.s0];
    VRAG_CUDA(cudaMemcpyAsync(d_base, rel.data(), static_cast<size_t>(ns) * 8, cudaMemcpyHostToDevice, _ctx->stream));
    VRAG_CUDA(cudaMemcpyAsync(d_first, ctx_first + ps.s0, static_cast<size_t>(ns) * 4, cudaMemcpyHostToDevice, _ctx->stream));
    VRAG_CUDA(cudaMemcpyAsync(d_len, ctx_len + ps.s0, static_cast<size_t>(ns) * 4, cudaMemcpyHostToDevice, _ctx->stream));
    if (ntok) {
      VRAG_CUDA(cudaMemcpyAsync(d_tcs, tok_char_start + tok_base[ps.s0], static_cast<size_t>(ntok) * 4, cudaMemcpyHostToDevice, _ctx->stream));
      VRAG_CUDA(cudaMemcpyAsync(d_tce, tok_char_end + tok_base[ps.s0], static_cast<size_t>(ntok) * 4, cudaMemcpyHostToDevice, _ctx->stream));
    }
    enc->sp_off.reserve((static_cast<size_t>(ns) * 2 + 1) * 4);
    int32_t* d_counts = enc->sp_off.as<int32_t>();
    int32_t* d_offs = d_counts + ns;
    launch_span_runs(_ctx, enc->probs.as<float>(), enc->cu.as<int32_t>(), d_first, d_len, d_base, d_tcs, d_tce, ns, threshold,
                     min_span_chars, merge_gap_chars, d_counts, d_offs, ps.s0, nullptr, nullptr, nullptr, nullptr, nullptr,
                     nullptr, false);
    int32_t n_here = 0;
    VRAG_CUDA(cudaMemcpyAsync(&n_here, d_offs + ns, 4, cudaMemcpyDeviceToHost, _ctx->stream));
    VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));   // also: rel is stack-owned
    if (n_here > 0) {
      enc->sp_out.reserve(static_cast<size_t>(n_here) * 24);
      int32_t* o_seq = enc->sp_out.as<int32_t>();
      int32_t *o_cs = o_seq + n_here, *o_ce = o_cs + n_here, *o_ts = o_ce + n_here, *o_te = o_ts + n_here;
      float* o_sc = reinterpret_cast<float*>(o_te + n_here);
      launch_span_runs(_ctx, enc->probs.as<float>(), enc->cu.as<int32_t>(), d_first, d_len, d_base, d_tcs, d_tce, ns,
                       threshold, min_span_chars, merge_gap_chars, nullptr, d_offs, ps.s0, o_seq, o_cs, o_ce, o_sc, o_ts,
                       o_te, true);
      if (total + n_here <= cap) {
        const size_t nb = static_cast<size_t>(n_here) * 4;
        VRAG_CUDA(cudaMemcpyAsync(span_seq + total, o_seq, nb, cudaMemcpyDeviceToHost, _ctx->stream));
        VRAG_CUDA(cudaMemcpyAsync(span_char_start + total, o_cs, nb, cudaMemcpyDeviceToHost, _ctx->stream));
        VRAG_CUDA(cudaMemcpyAsync(span_char_end + total, o_ce, nb, cudaMemcpyDeviceToHost, _ctx->stream));
        VRAG_CUDA(cudaMemcpyAsync(span_tok_start + total, o_ts, nb, cudaMemcpyDeviceToHost, _ctx->stream));
        VRAG_CUDA(cudaMemcpyAsync(span_tok_end + total, o_te, nb, cudaMemcpyDeviceToHost, _ctx->stream));
        VRAG_CUDA(cudaMemcpyAsync(span_score + total, o_sc, nb, cudaMemcpyDeviceToHost, _ctx->stream));
      }
      VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
    }
    total += n_here;
  }
  *nspans_out = total;
  if (total > cap) throw Error(VRAG_ERR_CAPACITY, "span_extract: output capacity too small; see *nspans_out");
  VRAG_API_END()
}
