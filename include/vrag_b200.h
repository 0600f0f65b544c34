/*
 * libvrag_b200 -- C ABI of the B200-native verbatim-rag hot path.
 *
 * The reference (KRLabsOrg/verbatim-rag) is pure Python and has NO FFI of its own; the drop-in
 * boundary is its three plugin ABCs.  Each entry point below replaces the third-party call that
 * the reference makes at the cited line; the Python plugin classes in verbatim_rag_b200/ bind these
 * through ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 *   vrag_span_forward        <- ModelSpanExtractor._extract_highlighter -> model.process()
 *                               packages/core/verbatim_core/extractors.py:203-228 (forward part)
 *   vrag_spans_from_probs    <- same call, span post-processing part (threshold / merge / min length)
 *   vrag_span_extract        <- same call, forward + post-processing fused on the device (only spans come back)
 *   vrag_encoder_create      <- ModelSpanExtractor._init_highlighter   extractors.py:151-157
 *                               SpladeProvider._load_model              verbatim_rag/embedding_providers.py:125-136
 *   vrag_splade_forward      <- SpladeProvider.embed_batch / embed_text -> SparseEncoder.encode
 *                               verbatim_rag/embedding_providers.py:138-166
 *   vrag_dense_forward       <- SentenceTransformersProvider.embed_batch embedding_providers.py:73-77
 *   vrag_rerank_forward      <- SentenceTransformersReranker.rerank -> CrossEncoder.predict
 *                               verbatim_rag/rerankers.py:109-134
 *   vrag_sentence_forward    <- ModelSpanExtractor._extract_qa_model -> QAModel.forward   extractors.py:230-283,
 *                               packages/core/verbatim_core/extractor_models/model.py:59-117
 *   vrag_index_create        <- LocalMilvusStore._setup_client          verbatim_rag/vector_stores/milvus_local.py:58-125
 *   vrag_index_add_*         <- BaseMilvusStore.add_vectors -> client.insert   milvus_base.py:90-127
 *   vrag_index_search_dense  <- BaseMilvusStore.query dense branch -> client.search  milvus_base.py:239-248
 *   vrag_index_search_sparse <- BaseMilvusStore.query sparse branch -> client.search milvus_base.py:250-259
 *   vrag_index_mark_deleted  <- BaseMilvusStore.delete -> client.delete         milvus_base.py:461-470
 *   vrag_topk_merge          <- (new) merge of per-shard top-k after the NCCL all-gather
 *   vrag_topk_publish / vrag_topk_merge_packed <- (new) the same exchange as stores into NVLink peer memory
 *
 * Conventions: every call returns an int status (0 = VRAG_OK); vrag_last_error() gives the message.
 * The CALLER owns all data buffers; the library owns only handles, weights, corpora and workspaces.
 * `on_device` != 0 means the data pointers are device pointers on the context's GPU (inputs already
 * resident in HBM); 0 means host pointers and the call performs the H2D / D2H copies itself.
 * No exception crosses the ABI.  All work is issued on the context's stream; calls that return data
 * to host pointers are synchronous, device-pointer calls are asynchronous until vrag_sync().
 * There is no CPU fallback: without a CUDA device vrag_ctx_create fails with VRAG_ERR_CUDA.
 */
#ifndef VRAG_B200_H
#define VRAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRAG_OK 0
#define VRAG_ERR_CUDA 1      /* CUDA runtime / driver failure (incl. no device) */
#define VRAG_ERR_ARG 2       /* invalid argument */
#define VRAG_ERR_CAPACITY 3  /* caller's output buffer too small; *needed says how much */
#define VRAG_ERR_WEIGHTS 4   /* missing / mis-shaped tensor in the weight set */
#define VRAG_ERR_INTERNAL 5

#define VRAG_ENC_MODERNBERT_TOKCLS 0 /* ModernBERT encoder + token-classification head (span extractor) */
#define VRAG_ENC_BERT_MLM 1          /* BERT encoder + MLM head + SPLADE pooling (sparse provider) */
#define VRAG_ENC_BERT_DENSE 2        /* BERT encoder + mean/CLS pooling (dense provider) */
#define VRAG_ENC_BERT_CLS 3          /* BERT encoder + pooler + 1-label classifier (cross-encoder reranker) */
#define VRAG_ENC_MODERNBERT_SENT 4   /* ModernBERT encoder + sentence mean-pool + 2-way classifier (legacy QAModel) */

#define VRAG_INDEX_DENSE_COSINE 0
#define VRAG_INDEX_SPARSE_IP 1

#define VRAG_POOL_MEAN 0
#define VRAG_POOL_CLS 1

typedef struct vrag_ctx vrag_ctx;
typedef struct vrag_encoder vrag_encoder;
typedef struct vrag_index vrag_index;

/* One named fp32 host tensor of a checkpoint (HuggingFace parameter name, row-major). */
typedef struct vrag_tensor {
  const char* name;
  const float* data;
  int64_t numel;
} vrag_tensor;

/* ---- context ------------------------------------------------------------------------------- */
int vrag_ctx_create(int device, vrag_ctx** out);
void vrag_ctx_destroy(vrag_ctx* ctx);
const char* vrag_last_error(vrag_ctx* ctx); /* ctx may be NULL: last error of a failed vrag_ctx_create */
int vrag_sync(vrag_ctx* ctx);
void* vrag_stream(vrag_ctx* ctx);            /* the cudaStream_t all kernels are launched on */
uint64_t vrag_launch_count(vrag_ctx* ctx);   /* kernels launched so far through this context */
const char* vrag_version(void);
/* Per-launch CUDA-event profiler on the context's stream.  Classes: 0 tensor-core GEMM, 1 attention, 2 row ops
 * (LayerNorm / embedding / head tail / CSR extraction), 3 top-k scan, 4 top-k select+rescore+rank, 5 other.
 * vrag_profile(ctx, 1) starts recording; vrag_profile_read synchronises, returns the summed kernel time (ms) and
 * launch count per class [6] since the last read, and resets the record. */
int vrag_profile(vrag_ctx* ctx, int enable);
int vrag_profile_read(vrag_ctx* ctx, double* ms_per_class, int64_t* launches_per_class);

/* ---- encoders ------------------------------------------------------------------------------ */
/* Uploads + repacks the checkpoint (fp32 host tensors -> fp16 tensor-core operands, fp32 norms/biases).
 * max_tokens bounds the tokens processed per internal pass (workspace size). */
int vrag_encoder_create(vrag_ctx* ctx, int kind, int num_layers, int vocab_size, int max_tokens,
                        const vrag_tensor* tensors, int num_tensors, vrag_encoder** out);
/* Same with an arithmetic mode.  The reference runs these models in fp32 (extractors.py:151-157 loads with
 * AutoModel.from_pretrained and only moves to the device: no autocast).
 *   VRAG_PRECISION_FAST    fp16 tensor-core operands (11-bit significands), fp32 accumulation: span logits within
 *                          ~3.5e-3 of the fp32 reference, identical spans except at tokens within that distance of
 *                          the threshold.
 *   VRAG_PRECISION_PRECISE every weight / activation as two fp16 planes x = hi + lo (>= 21 significant bits), each
 *                          product as three tcgen05.mma (a_hi w_hi + a_lo w_hi + a_hi w_lo), fp32 residual stream:
 *                          span logits within 1e-3 (measured ~1e-4) at roughly a third of the throughput. */
#define VRAG_PRECISION_FAST 0
#define VRAG_PRECISION_PRECISE 1
int vrag_encoder_create_ex(vrag_ctx* ctx, int kind, int num_layers, int vocab_size, int max_tokens,
                           const vrag_tensor* tensors, int num_tensors, int precision, vrag_encoder** out);
/* Width of the encoder's hidden states = length of a vrag_dense_forward row: 768 (BERT-base / ModernBERT-base shapes)
 * or 384 (MiniLM-class BERT encoders, the reference's default dense model: embedding_providers.py:55). */
int vrag_encoder_hidden(vrag_encoder* enc);
void vrag_encoder_destroy(vrag_encoder* enc);

/* Token-classification forward over `nseq` unpadded sequences packed back to back:
 * ids[cu_seqlens[i] .. cu_seqlens[i+1]) is sequence i; cu_seqlens is always a HOST array [nseq+1].
 * probs_out[t]  = softmax(logits[t])[1]  (P(relevant), fp32)      [total_tokens]
 * logits_out    = raw logits [total_tokens, 2] or NULL. */
int vrag_span_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq, float* probs_out,
                      float* logits_out, int on_device);

/* SPLADE forward: per sequence max over tokens of log1p(relu(mlm_logits)), returned as CSR with
 * entries > min_abs (0.0 reproduces embed_batch's np.nonzero, 1e-6 embed_text's filter).
 * indptr_out [nseq+1], indices_out/values_out [cap] are HOST buffers; on VRAG_ERR_CAPACITY *nnz_out is
 * the required capacity.  dense_out (nullable) receives the dense [nseq, vocab] fp32 vectors
 * (host or device per on_device; `ids` follows on_device too). */
int vrag_splade_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq, float min_abs,
                        int64_t* indptr_out, int32_t* indices_out, float* values_out, int64_t cap, int64_t* nnz_out,
                        float* dense_out, int on_device);

/* Dense sentence embedding: encoder -> mean / CLS pooling -> optional L2 normalisation.  out [nseq, hidden]. */
int vrag_dense_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq, int pooling,
                       int normalize, float* out, int on_device);

/* Cross-encoder reranker: pair sequences ([CLS] q [SEP] doc [SEP], type_ids 0 / 1 per token, NULL = all 0) -> one
 * relevance logit per sequence (BertForSequenceClassification, num_labels = 1).  Needs a VRAG_ENC_BERT_CLS encoder
 * (weights: bert.*, bert.pooler.dense.*, classifier.*). */
int vrag_rerank_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens,
                        int nseq, float* scores_out, int on_device);

/* Sentence classifier of the legacy QAModel: per sentence j of sequence i (sent_indptr[i] <= j < sent_indptr[i + 1]) the
 * mean of the encoder's final hidden states over tokens [sent_start[j], sent_end[j]] (indices inside the sequence, end
 * inclusive) -> Linear(hidden, 2).  logits_out [n_sentences, 2].  Host buffers.  Needs a VRAG_ENC_MODERNBERT_SENT
 * encoder (weights: model.*, classifier.weight [2, hidden], classifier.bias). */
int vrag_sentence_forward(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq,
                          const int32_t* sent_indptr, const int32_t* sent_start, const int32_t* sent_end,
                          float* logits_out);

/* Debug / test hook: run the tcgen05 GEMM  C[M,N] = A[M,K] * W[N,K]^T  (fp16 operands, fp32 accumulate) against
 * an in-library SIMT reference GEMM on random data and return the max |diff| (device-side self check). */
int vrag_selftest_gemm(vrag_ctx* ctx, int M, int N, int K, int epilogue, double* max_abs_diff, double* ref_abs_max);

/* Same self check in split-precision mode (VRAG_PRECISION_PRECISE): random hi / lo planes of both operands, three MMAs
 * per product; fp16 outputs are compared as hi + lo sums.  Epilogues 10, 0, 1, 2, 3. */
int vrag_selftest_gemm_split(vrag_ctx* ctx, int M, int N, int K, int epilogue, double* max_abs_diff,
                             double* ref_abs_max);

/* Timing hook (bench.py's per-kernel table, development): `iters` back-to-back launches of one encoder GEMM shape
 * (epilogue numbering of csrc/gemm.cuh) on synthetic device operands, CUDA events on the library's stream.
 * stages: operand ring depth 3..5 (0 = the context's default); debug_mode 3 skips the epilogue (mainloop only).
 * *ms_out = average launch time in milliseconds. */
int vrag_bench_gemm(vrag_ctx* ctx, int M, int N, int K, int epilogue, int stages, int debug_mode, int iters,
                    double* ms_out);

/* Split-precision variant of the timing hook (epilogues 0, 1, 2, 3). */
int vrag_bench_gemm_split(vrag_ctx* ctx, int M, int N, int K, int epilogue, int iters, double* ms_out);

/* Debug / test hook: one self-attention launch (12 heads x 64, as in both encoders; reference semantics:
 * transformers ModernBertAttention sdpa path -- softmax(q k^T / 8 + window mask) v per sequence) on caller-supplied
 * fp16 rows qkv_f16 [total_tokens, 2304] = q|k|v (q, k already rotated), host memory.  window < 0: full attention,
 * else keys with |i - j| <= window.  legacy != 0 runs the mma.sync cross-check kernel.  out_f16 [total_tokens, 768]. */
int vrag_selftest_attention(vrag_ctx* ctx, const uint16_t* qkv_f16, const int32_t* cu_seqlens, int nseq, int window,
                            int legacy, uint16_t* out_f16);

/* Split-precision attention (VRAG_PRECISION_PRECISE): q|k|v rows and the output as hi / lo fp16 planes. */
int vrag_selftest_attention_split(vrag_ctx* ctx, const uint16_t* qkv_hi_f16, const uint16_t* qkv_lo_f16,
                                  const int32_t* cu_seqlens, int nseq, int window, uint16_t* out_hi_f16,
                                  uint16_t* out_lo_f16);

/* Timing hook (development): `iters` back-to-back launches of the attention kernel on synthetic fp16 q|k|v rows
 * (nseq sequences of seq_len tokens, 12 heads x 64; window as above), CUDA events on the library's stream.
 * *ms_out = average launch time in milliseconds. */
int vrag_bench_attention(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, double* ms_out);
int vrag_bench_attention_split(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, double* ms_out);

/* Debug hook (tests): vrag_span_forward with host buffers that also returns the fp32 residual stream after
 * the embedding and after every layer, hidden_out [num_layers + 1, total_tokens, 768]; single pass only. */
int vrag_debug_span_hidden(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq,
                           float* probs_out, float* logits_out, float* hidden_out);

/* ---- span post-processing (host, integer logic) -------------------------------------------- */
/* For context c, its tokens are [ctx_indptr[c], ctx_indptr[c+1]) in probs / tok_char_start / tok_char_end.
 * keep = p > threshold; maximal runs -> char spans; merge gaps <= merge_gap_chars; drop spans shorter than
 * min_span_chars.  Outputs (capacity cap): span_ctx, span_char_start, span_char_end, span_score (mean kept p),
 * span_tok_start, span_tok_end (token indices relative to the context).  *nspans_out = number produced / needed. */
int vrag_spans_from_probs(const float* probs, const int32_t* tok_char_start, const int32_t* tok_char_end,
                          const int64_t* ctx_indptr, int nctx, float threshold, int min_span_chars,
                          int merge_gap_chars, int32_t* span_ctx, int32_t* span_char_start, int32_t* span_char_end,
                          float* span_score, int32_t* span_tok_start, int32_t* span_tok_end, int64_t cap,
                          int64_t* nspans_out);

/* Span extraction with the post-processing on the device (reference: all of model.process(), extractors.py:213-224):
 * forward + threshold / runs / gap merge / min length per sequence over its context tokens
 * [ctx_first[i], ctx_first[i] + ctx_len[i]); tok_char_start / tok_char_end are the character offsets of the context
 * tokens of all sequences back to back.  Only the spans come back (same outputs and rules as vrag_spans_from_probs,
 * span_seq = sequence index; results are bit-identical to vrag_span_forward + vrag_spans_from_probs).  One sequence per
 * context: documents that need several windows use the two-call form.  Host buffers. */
int vrag_span_extract(vrag_encoder* enc, const int32_t* ids, const int32_t* cu_seqlens, int nseq,
                      const int32_t* ctx_first, const int32_t* ctx_len, const int32_t* tok_char_start,
                      const int32_t* tok_char_end, float threshold, int min_span_chars, int merge_gap_chars,
                      int32_t* span_seq, int32_t* span_char_start, int32_t* span_char_end, float* span_score,
                      int32_t* span_tok_start, int32_t* span_tok_end, int64_t cap, int64_t* nspans_out);

/* ---- exact top-k index --------------------------------------------------------------------- */
int vrag_index_create(vrag_ctx* ctx, int kind, int dim, vrag_index** out);
void vrag_index_destroy(vrag_index* idx);
int64_t vrag_index_size(vrag_index* idx);
/* Append n fp32 rows [n, dim] (not normalised; the index caches 1/||row|| computed in fp64). */
int vrag_index_add_dense(vrag_index* idx, const float* rows, int64_t n, int on_device);
/* Append n sparse rows given as host CSR (indptr [n+1] relative to this call, indices ascending per row). */
int vrag_index_add_sparse(vrag_index* idx, const int64_t* indptr, const int32_t* indices, const float* values,
                          int64_t n);
/* Global id of this shard's row 0: added to the row numbers reported by the searches (corpus row-sharded
 * across GPUs). */
int vrag_index_set_id_base(vrag_index* idx, int64_t base);
/* Tombstone rows (deleted rows never appear in results). rows: host array of row numbers. */
int vrag_index_mark_deleted(vrag_index* idx, const int64_t* rows, int64_t n);

/* Metadata-filter pushdown (reference: the `filter=` argument of BaseMilvusStore.query, milvus_base.py:189-259): rows
 * with exclude[row] != 0 (host array of vrag_index_size(idx) bytes, aligned with the insertion order) are skipped by
 * every following search, like deleted rows, until the filter is cleared with exclude == NULL or rows are added /
 * deleted.  The host compiles the boolean expression to this mask; the scan kernels only consult one byte per row. */
int vrag_index_set_filter(vrag_index* idx, const uint8_t* exclude, int64_t n);

/* Exact top-k.  Order: score descending, row index ascending; score evaluated in fp64 from the stored fp32
 * values and reported both as fp32 (scores_out) and fp64 (scores64_out, nullable).  Results [nq, k]; when the
 * index holds fewer than k live rows the tail is filled with id -1 / score -inf.
 * queries [nq, dim] fp32 (host/device per on_device); outputs follow on_device too. */
int vrag_index_search_dense(vrag_index* idx, const float* queries, int nq, int k, int64_t* ids_out, float* scores_out,
                            double* scores64_out, int on_device);
/* Sparse queries as host CSR; outputs are host buffers. */
int vrag_index_search_sparse(vrag_index* idx, const int64_t* q_indptr, const int32_t* q_indices,
                             const float* q_values, int nq, int k, int64_t* ids_out, float* scores_out,
                             double* scores64_out);

/* Merge m candidate (score64, id) pairs per query (e.g. the all-gathered per-shard top-k of all ranks) into the
 * global top-k with the same order.  Device buffers on ctx's GPU. */
int vrag_topk_merge(vrag_ctx* ctx, const double* scores64, const int64_t* ids, int nq, int m, int k, int64_t* ids_out,
                    float* scores_out, double* scores64_out);

/* The same exchange over NVLink / NVSwitch peer memory instead of an NCCL all-gather (the one collective
 * of the path).  peer_bufs: `world` device pointers (a host array) -- the SAME symmetric buffer of every rank as mapped
 * into this process (cudaIpc / fabric handles; the Python host side uses torch's symmetric memory), each holding
 * [world][nq][k] records of 16 bytes (fp64 score bits, int64 global id).  vrag_topk_publish stores this rank's block
 * (device arrays scores64 / ids, [nq][k]) into slot `rank` of every peer's buffer; after a cross-rank barrier on the
 * same stream, vrag_topk_merge_packed merges the `world` blocks of the LOCAL buffer into the global top-k
 * ((score desc, id asc), id < 0 = empty slot: bit-identical to vrag_topk_merge on the gathered arrays). */
int vrag_topk_publish(vrag_ctx* ctx, const double* scores64, const int64_t* ids, int nq, int k, void* const* peer_bufs,
                      int world, int rank);
int vrag_topk_merge_packed(vrag_ctx* ctx, const int64_t* packed, int world, int nq, int k, int64_t* ids_out,
                           float* scores_out, double* scores64_out);

#ifdef __cplusplus
}
#endif
#endif /* VRAG_B200_H */
