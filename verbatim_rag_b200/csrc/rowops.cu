// Row-wise (per token) kernels of the encoders: embedding gather + LayerNorm, LayerNorm, classifier tail,
// SPLADE CSR extraction, sentence pooling.  One warp per 768-wide row, 128-bit loads, fp32 statistics.
// These are HBM-streaming kernels: ~3 KB read + 1.5-4.5 KB written per token.
#include "encoder.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int H = HIDDEN;
constexpr int VEC = H / 128;  // float4 per lane = 6 (768-wide rows); the BERT kernels also exist for VEC = 3 (384)
constexpr int ROWS_PER_BLOCK = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Normalise the row held as v[VEC] float4 per lane (element index = (i*32 + lane)*4 + e).
template <int VEC>
__device__ __forceinline__ void ln_row(float4 (&v)[VEC], const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int lane) {
  constexpr int H = VEC * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    v[i].x *= rstd * g.x; v[i].y *= rstd * g.y; v[i].z *= rstd * g.z; v[i].w *= rstd * g.w;
    if (beta) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
      v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
    }
  }
}

// lo8_row (optional): the rounding remainder v - fp16(v) as e5m2, so that hi + lo carries the row to >= 14 bits
// (the two-plane residual stream of the deferred-LayerNorm path, gemm.cuh EPI_RESID_STATS).
// lo16_row (optional): the same remainder as fp16 -- the low plane of the split-precision ("precise") mode, gemm.cuh.
template <int VEC>
__device__ __forceinline__ void store_row(const float4 (&v)[VEC], float* x32_row, __half* h16_row, int lane,
                                          uint8_t* lo8_row = nullptr, __half* lo16_row = nullptr) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (x32_row) reinterpret_cast<float4*>(x32_row)[i * 32 + lane] = v[i];
    if (h16_row) {
      const __half2 h0 = __floats2half2_rn(v[i].x, v[i].y), h1 = __floats2half2_rn(v[i].z, v[i].w);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      reinterpret_cast<uint2*>(h16_row)[i * 32 + lane] = u;
      if (lo8_row) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        reinterpret_cast<uint32_t*>(lo8_row)[i * 32 + lane] =
            pack_e5m2x4(v[i].x - f0.x, v[i].y - f0.y, v[i].z - f1.x, v[i].w - f1.y);
      }
      if (lo16_row) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v[i].x - f0.x, v[i].y - f0.y);
        const __half2 l1 = __floats2half2_rn(v[i].z - f1.x, v[i].w - f1.y);
        uint2 ul;
        ul.x = *reinterpret_cast<const uint32_t*>(&l0);
        ul.y = *reinterpret_cast<const uint32_t*>(&l1);
        reinterpret_cast<uint2*>(lo16_row)[i * 32 + lane] = ul;
      }
    }
  }
}
// row of the two-plane residual stream -> fp32 registers
__device__ __forceinline__ void load_row_hilo(float4 (&v)[VEC], const __half* hi_row, const uint8_t* lo_row, int lane) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const uint2 a = reinterpret_cast<const uint2*>(hi_row)[i * 32 + lane];
    const uint32_t b = reinterpret_cast<const uint32_t*>(lo_row)[i * 32 + lane];
    const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
    const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    const float2 b0 = unpack_e5m2x2<0>(b), b1 = unpack_e5m2x2<1>(b);
    v[i] = make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
  }
}

__global__ void token_meta_kernel(const int32_t* __restrict__ cu, int nseq, int32_t* __restrict__ pos,
                                  int32_t* __restrict__ seq_of_row) {
  const int s = blockIdx.x;
  if (s >= nseq) return;
  const int a = cu[s], b = cu[s + 1];
  for (int t = a + threadIdx.x; t < b; t += blockDim.x) {
    pos[t] = t - a;
    seq_of_row[t] = s;
  }
}

__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
embed_ln_kernel(const int32_t* __restrict__ ids, int T, int vocab, const float* __restrict__ emb,
                const float* __restrict__ gamma, float eps, float* __restrict__ x32, __half* __restrict__ h16,
                uint8_t* __restrict__ lo8, __half* __restrict__ lo16) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* e = reinterpret_cast<const float4*>(emb + static_cast<size_t>(id) * H);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = __ldg(e + i * 32 + lane);
  ln_row(v, gamma, nullptr, eps, lane);
  store_row(v, x32 ? x32 + static_cast<size_t>(row) * H : nullptr, h16 + static_cast<size_t>(row) * H, lane,
            lo8 ? lo8 + static_cast<size_t>(row) * H : nullptr, lo16 ? lo16 + static_cast<size_t>(row) * H : nullptr);
}

// LayerNorm of the two-plane residual stream (final norm of the deferred-LayerNorm path); h16 may alias hi.
__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
layernorm_hilo_kernel(const __half* hi, const uint8_t* __restrict__ lo, int T, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float* __restrict__ x32, __half* h16) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  float4 v[VEC];
  load_row_hilo(v, hi + static_cast<size_t>(row) * H, lo + static_cast<size_t>(row) * H, lane);
  ln_row(v, gamma, beta, eps, lane);
  store_row(v, x32 ? x32 + static_cast<size_t>(row) * H : nullptr, h16 + static_cast<size_t>(row) * H, lane);
}

__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
hilo_to_f32_kernel(const __half* __restrict__ hi, const uint8_t* __restrict__ lo, int T, float* __restrict__ x32) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  float4 v[VEC];
  load_row_hilo(v, hi + static_cast<size_t>(row) * H, lo + static_cast<size_t>(row) * H, lane);
  store_row(v, x32 + static_cast<size_t>(row) * H, nullptr, lane);
}

template <int VEC>
__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
bert_embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos, int T, int vocab, int max_pos,
                     const float* __restrict__ wemb, const float* __restrict__ pemb, const float* __restrict__ temb0,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     float* __restrict__ x32, __half* __restrict__ h16, __half* __restrict__ lo16,
                     const int32_t* __restrict__ type_ids) {
  constexpr int H = VEC * 128;
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  int p = pos[row];
  p = p >= max_pos ? max_pos - 1 : p;
  const float4* e = reinterpret_cast<const float4*>(wemb + static_cast<size_t>(id) * H);
  const float4* pe = reinterpret_cast<const float4*>(pemb + static_cast<size_t>(p) * H);
  // token type 1 = second segment of a pair input (cross-encoder reranker); temb0 holds both rows back to back
  const int tt = type_ids ? (type_ids[row] != 0 ? 1 : 0) : 0;
  const float4* te = reinterpret_cast<const float4*>(temb0 + static_cast<size_t>(tt) * H);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 a = __ldg(e + i * 32 + lane), b = __ldg(pe + i * 32 + lane), c = __ldg(te + i * 32 + lane);
    v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
  }
  ln_row(v, gamma, beta, eps, lane);
  store_row(v, x32 + static_cast<size_t>(row) * H, h16 + static_cast<size_t>(row) * H, lane, nullptr,
            lo16 ? lo16 + static_cast<size_t>(row) * H : nullptr);
}

// Deferred-LayerNorm BERT path: z = word + position + type embedding, NOT normalised -- written as the two-plane
// stream with its row moments (slot 0 = (sum z, sum z^2), slots 1..5 = 0); the embedding LayerNorm is folded into the
// first layer's GEMMs like every other LayerNorm of the stack.
__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
bert_embed_raw_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos, int T, int vocab, int max_pos,
                      const float* __restrict__ wemb, const float* __restrict__ pemb, const float* __restrict__ temb0,
                      __half* __restrict__ h16, uint8_t* __restrict__ lo8, float* __restrict__ stats, int slots) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  int p = pos[row];
  p = p >= max_pos ? max_pos - 1 : p;
  const float4* e = reinterpret_cast<const float4*>(wemb + static_cast<size_t>(id) * H);
  const float4* pe = reinterpret_cast<const float4*>(pemb + static_cast<size_t>(p) * H);
  const float4* te = reinterpret_cast<const float4*>(temb0);
  float4 v[VEC];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 a = __ldg(e + i * 32 + lane), b = __ldg(pe + i * 32 + lane), c = __ldg(te + i * 32 + lane);
    v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  s = warp_sum(s);
  q = warp_sum(q);
  store_row(v, nullptr, h16 + static_cast<size_t>(row) * H, lane, lo8 + static_cast<size_t>(row) * H);
  if (lane < slots)
    reinterpret_cast<float2*>(stats)[static_cast<size_t>(lane) * T + row] = lane == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
}

template <int VEC>
__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
layernorm_kernel(float* __restrict__ x32, int T, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, __half* __restrict__ h16, int write_back, __half* __restrict__ lo16) {
  constexpr int H = VEC * 128;
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  float* xr = x32 + static_cast<size_t>(row) * H;
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = reinterpret_cast<const float4*>(xr)[i * 32 + lane];
  ln_row(v, gamma, beta, eps, lane);
  store_row(v, write_back ? xr : nullptr, h16 ? h16 + static_cast<size_t>(row) * H : nullptr, lane, nullptr,
            (h16 && lo16) ? lo16 + static_cast<size_t>(row) * H : nullptr);
}

__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK)
head_final_kernel(const float* __restrict__ buf32, int T, const float* __restrict__ gamma, float eps,
                  const float* __restrict__ cw, const float* __restrict__ cb, float* __restrict__ logits,
                  float* __restrict__ probs) {
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const float4* xr = reinterpret_cast<const float4*>(buf32 + static_cast<size_t>(row) * H);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = xr[i * 32 + lane];
  ln_row(v, gamma, nullptr, eps, lane);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(cw) + i * 32 + lane);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(cw + H) + i * 32 + lane);
    d0 += (v[i].x * w0.x + v[i].y * w0.y) + (v[i].z * w0.z + v[i].w * w0.w);
    d1 += (v[i].x * w1.x + v[i].y * w1.y) + (v[i].z * w1.z + v[i].w * w1.w);
  }
  d0 = warp_sum(d0) + cb[0];
  d1 = warp_sum(d1) + cb[1];
  if (lane == 0) {
    if (logits) {
      logits[2 * static_cast<size_t>(row)] = d0;
      logits[2 * static_cast<size_t>(row) + 1] = d1;
    }
    // softmax([d0, d1])[1] in fp32, max-subtracted like torch.softmax
    const float m = fmaxf(d0, d1);
    const float e0 = expf(d0 - m), e1 = expf(d1 - m);
    probs[row] = e1 / (e0 + e1);
  }
}

// Tail of the fused span-logit head (gemm.cuh EPI_HEAD_PARTIAL): per row the partial sums of g = gelu(head.dense x)
// over `slots` column groups -> LayerNorm statistics -> logits[c] = rstd * (D_c - mean * G_c) + b_c, P(class 1).
__global__ void __launch_bounds__(256)
head_finish_kernel(const float4* __restrict__ part, int T, int slots, float inv_dim, float eps, float g0, float g1,
                   const float* __restrict__ cb, float* __restrict__ logits, float* __restrict__ probs) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= T) return;
  float s1 = 0.f, s2 = 0.f, d0 = 0.f, d1 = 0.f;
  for (int j = 0; j < slots; ++j) {
    const float4 v = part[static_cast<size_t>(j) * T + row];
    s1 += v.x; s2 += v.y; d0 += v.z; d1 += v.w;
  }
  const float mean = s1 * inv_dim;
  const float var = fmaxf(s2 * inv_dim - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float l0 = rstd * (d0 - mean * g0) + cb[0], l1 = rstd * (d1 - mean * g1) + cb[1];
  if (logits) {
    logits[2 * static_cast<size_t>(row)] = l0;
    logits[2 * static_cast<size_t>(row) + 1] = l1;
  }
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  probs[row] = e1 / (e0 + e1);
}

// ---- span post-processing on the device (the integer half of the highlighter contract, identical to
// the host function vrag_spans_from_probs in api.cu / oracle/highlighter.py steps 4-8): one thread per sequence walks
// its context tokens: keep = p > threshold; maximal runs -> char spans; merge gaps <= merge_gap_chars; drop spans
// shorter than min_span_chars; score = mean kept p (accumulated in double, same order as the host function, so the
// results are bit-identical).  FILL = false counts the spans of each sequence, FILL = true writes them at offs[seq].
template <bool FILL>
__global__ void __launch_bounds__(128)
span_runs_kernel(const float* __restrict__ probs, const int32_t* __restrict__ cu, const int32_t* __restrict__ ctx_first,
                 const int32_t* __restrict__ ctx_len, const int64_t* __restrict__ tok_base,
                 const int32_t* __restrict__ tcs, const int32_t* __restrict__ tce, int nseq, float threshold,
                 int min_span_chars, int merge_gap_chars, int32_t* __restrict__ counts, const int32_t* __restrict__ offs,
                 int32_t seq_base, int32_t* __restrict__ o_seq, int32_t* __restrict__ o_cs, int32_t* __restrict__ o_ce,
                 float* __restrict__ o_score, int32_t* __restrict__ o_ts, int32_t* __restrict__ o_te) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseq) return;
  const float* p = probs + cu[s] + ctx_first[s];
  const int n = ctx_len[s];
  const int32_t* cs_ = tcs + tok_base[s];
  const int32_t* ce_ = tce + tok_base[s];
  int count = 0;
  int out = FILL ? offs[s] : 0;
  bool open = false;
  int32_t cs = 0, ce = 0, ts = 0, te = 0, cnt = 0;
  double acc = 0.0;
  auto flush = [&]() {
    if (open && ce - cs >= min_span_chars) {
      if (FILL) {
        o_seq[out] = seq_base + s;
        o_cs[out] = cs;
        o_ce[out] = ce;
        o_score[out] = static_cast<float>(acc / cnt);
        o_ts[out] = ts;
        o_te[out] = te;
        ++out;
      }
      ++count;
    }
    open = false;
  };
  int i = 0;
  while (i < n) {
    if (!(p[i] > threshold)) { ++i; continue; }
    int j = i;
    double racc = 0.0;
    while (j < n && p[j] > threshold) { racc += static_cast<double>(p[j]); ++j; }
    const int32_t rs = cs_[i], re = ce_[j - 1];
    if (open && rs - ce <= merge_gap_chars) {
      ce = re;
      te = j;
      acc += racc;
      cnt += j - i;
    } else {
      flush();
      open = true;
      cs = rs; ce = re; ts = i; te = j;
      acc = racc;
      cnt = j - i;
    }
    i = j;
  }
  flush();
  if (!FILL) counts[s] = count;
}
__global__ void exclusive_scan_small_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out /*[n+1]*/) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int32_t acc = 0;
    for (int i = 0; i < n; ++i) { out[i] = acc; acc += in[i]; }
    out[n] = acc;
  }
}

// ---- cross-encoder head (BertForSequenceClassification, num_labels = 1: the reference's SentenceTransformersReranker,
// verbatim_rag/rerankers.py:109-134): score = W_c tanh(W_p h_CLS + b_p) + b_c.  One block per sequence; the pooler
// matrix (H x H fp32) streams from L2.
__global__ void __launch_bounds__(256)
cls_head_kernel(const float* __restrict__ x32, const int32_t* __restrict__ cu, int H, const float* __restrict__ wp,
                const float* __restrict__ bp, const float* __restrict__ wc, const float* __restrict__ bc,
                float* __restrict__ scores) {
  __shared__ float xs[768];
  __shared__ float red[8];
  const int s = blockIdx.x;
  const float* x = x32 + static_cast<size_t>(cu[s]) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part = 0.f;
  for (int n = warp; n < H; n += 8) {   // one pooler row per warp
    const float* w = wp + static_cast<size_t>(n) * H;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(__ldg(w + k), xs[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) part += __ldg(wc + n) * tanhf(acc + __ldg(bp + n));
  }
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = bc[0];
    for (int w = 0; w < 8; ++w) t += red[w];
    scores[s] = t;
  }
}

// ---- sentence head (the legacy QAModel, packages/core/verbatim_core/extractor_models/model.py:59-117): mean of the final
// hidden states over the token rows [start, end] of a sentence -> Linear(H, 2).  One block per sentence.
__global__ void __launch_bounds__(256)
sentence_head_kernel(const float* __restrict__ x32, int H, const int32_t* __restrict__ row_start,
                     const int32_t* __restrict__ row_end /*inclusive*/, const float* __restrict__ cw,
                     const float* __restrict__ cb, float* __restrict__ logits) {
  __shared__ float red0[8], red1[8];
  const int s = blockIdx.x;
  const int a = row_start[s], b = row_end[s];
  float d0 = 0.f, d1 = 0.f;
  const float inv = b >= a ? 1.0f / static_cast<float>(b - a + 1) : 0.f;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float t = 0.f;
    for (int r = a; r <= b; ++r) t += x32[static_cast<size_t>(r) * H + c];
    t *= inv;
    d0 = fmaf(t, __ldg(cw + c), d0);
    d1 = fmaf(t, __ldg(cw + H + c), d1);
  }
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5] = d0; red1[threadIdx.x >> 5] = d1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = cb[0], t1 = cb[1];
    for (int w = 0; w < 8; ++w) { t0 += red0[w]; t1 += red1[w]; }
    logits[2 * s] = t0;
    logits[2 * s + 1] = t1;
  }
}

// ---- SPLADE: dense [nseq, ld] -> CSR, entries > min_abs, ascending vocabulary index ----
__global__ void splade_count_kernel(const float* __restrict__ dense, int ld, int vocab, float min_abs,
                                    int32_t* __restrict__ counts) {
  const float* r = dense + static_cast<size_t>(blockIdx.x) * ld;
  int c = 0;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) c += r[i] > min_abs;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += part[w];
    counts[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
splade_fill_kernel(const float* __restrict__ dense, int ld, int vocab, float min_abs,
                   const int64_t* __restrict__ indptr, int32_t* __restrict__ indices, float* __restrict__ values) {
  const float* r = dense + static_cast<size_t>(blockIdx.x) * ld;
  int64_t base = indptr[blockIdx.x];
  __shared__ int wcount[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < vocab; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const float x = i < vocab ? r[i] : 0.f;
    const bool keep = i < vocab && x > min_abs;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wcount[warp] = __popc(m);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) off += wcount[w];
      tot += wcount[w];
    }
    if (keep) {
      const int64_t o = base + off + __popc(m & ((1u << lane) - 1));
      indices[o] = i;
      values[o] = x;
    }
    base += tot;
    __syncthreads();
  }
}

// ---- sentence pooling over the fp32 final hidden states ----
__global__ void __launch_bounds__(256)
pool_kernel(const float* __restrict__ x32, const int32_t* __restrict__ cu, int pooling, int normalize,
            float* __restrict__ out, int H) {
  const int s = blockIdx.x;
  const int a = cu[s], b = cu[s + 1];
  __shared__ float red[256];
  float acc[3] = {0.f, 0.f, 0.f};   // H <= 768
  for (int j = 0; j < 3; ++j) {
    const int c = threadIdx.x + j * 256;
    if (c >= H) continue;
    if (pooling == VRAG_POOL_CLS) {
      acc[j] = b > a ? x32[static_cast<size_t>(a) * H + c] : 0.f;
    } else {
      float t = 0.f;
      for (int r = a; r < b; ++r) t += x32[static_cast<size_t>(r) * H + c];
      acc[j] = b > a ? t / static_cast<float>(b - a) : 0.f;
    }
  }
  float scale = 1.f;
  if (normalize) {
    red[threadIdx.x] = acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    scale = 1.f / fmaxf(sqrtf(red[0]), 1e-12f);
  }
  for (int j = 0; j < 3; ++j)
    if (threadIdx.x + j * 256 < H) out[static_cast<size_t>(s) * H + threadIdx.x + j * 256] = acc[j] * scale;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 u;
    u.x = pack_half2(v.x, v.y);
    u.y = pack_half2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i) = u;
  } else {
    for (; i < n; ++i) dst[i] = __float2half_rn(src[i]);
  }
}

__global__ void f32_to_f16_split_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                        size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = src[i];
  const __half h = __float2half_rn(x);
  hi[i] = h;
  lo[i] = __float2half_rn(x - __half2float(h));
}

inline int row_blocks(int T) { return (T + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK; }

}  // namespace

#define VRAG_LAUNCHED(ctx)            \
  do {                                \
    VRAG_CUDA(cudaGetLastError());    \
    (ctx)->launches++;                \
  } while (0)

void launch_token_meta(vrag_ctx* ctx, const int32_t* cu, int nseq, int total, int32_t* pos, int32_t* seq_of_row) {
  ProfScope prof(ctx, PROF_ROWOPS);
  (void)total;
  token_meta_kernel<<<nseq, 128, 0, ctx->stream>>>(cu, nseq, pos, seq_of_row);
  VRAG_LAUNCHED(ctx);
}
void launch_embed_ln(vrag_ctx* ctx, const int32_t* ids, int T, int vocab, const float* tok_emb, const float* gamma,
                     float eps, float* x32, __half* h16, uint8_t* lo8, __half* h16_lo) {
  ProfScope prof(ctx, PROF_ROWOPS);
  embed_ln_kernel<<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(ids, T, vocab, tok_emb, gamma, eps, x32, h16,
                                                                           lo8, h16_lo);
  VRAG_LAUNCHED(ctx);
}
void launch_layernorm_hilo(vrag_ctx* ctx, const __half* hi, const uint8_t* lo, int T, const float* gamma,
                           const float* beta, float eps, float* x32, __half* h16) {
  ProfScope prof(ctx, PROF_ROWOPS);
  layernorm_hilo_kernel<<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(hi, lo, T, gamma, beta, eps, x32, h16);
  VRAG_LAUNCHED(ctx);
}
void launch_bert_embed_raw(vrag_ctx* ctx, const int32_t* ids, const int32_t* pos, int T, int vocab, int max_pos,
                           const float* word_emb, const float* pos_emb, const float* type_emb0, __half* h16,
                           uint8_t* lo8, float* stats, int slots) {
  ProfScope prof(ctx, PROF_ROWOPS);
  bert_embed_raw_kernel<<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(ids, pos, T, vocab, max_pos, word_emb,
                                                                                 pos_emb, type_emb0, h16, lo8, stats,
                                                                                 slots);
  VRAG_LAUNCHED(ctx);
}
void launch_hilo_to_f32(vrag_ctx* ctx, const __half* hi, const uint8_t* lo, int T, float* x32) {
  ProfScope prof(ctx, PROF_ROWOPS);
  hilo_to_f32_kernel<<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(hi, lo, T, x32);
  VRAG_LAUNCHED(ctx);
}
void launch_bert_embed_ln(vrag_ctx* ctx, const int32_t* ids, const int32_t* pos, int T, int vocab, int max_pos,
                          const float* word_emb, const float* pos_emb, const float* type_emb0, const float* gamma,
                          const float* beta, float eps, float* x32, __half* h16, __half* h16_lo, int hidden,
                          const int32_t* type_ids) {
  ProfScope prof(ctx, PROF_ROWOPS);
  VRAG_CHECK(hidden == 768 || hidden == 384, VRAG_ERR_ARG, "row kernels: hidden size must be 768 or 384");
  if (hidden == 768)
    bert_embed_ln_kernel<6><<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(ids, pos, T, vocab, max_pos, word_emb,
                                                                                   pos_emb, type_emb0, gamma, beta, eps,
                                                                                   x32, h16, h16_lo, type_ids);
  else
    bert_embed_ln_kernel<3><<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(ids, pos, T, vocab, max_pos, word_emb,
                                                                                   pos_emb, type_emb0, gamma, beta, eps,
                                                                                   x32, h16, h16_lo, type_ids);
  VRAG_LAUNCHED(ctx);
}
void launch_layernorm(vrag_ctx* ctx, float* x32, int T, const float* gamma, const float* beta, float eps, __half* h16,
                      bool write_back, __half* h16_lo, int hidden) {
  ProfScope prof(ctx, PROF_ROWOPS);
  VRAG_CHECK(hidden == 768 || hidden == 384, VRAG_ERR_ARG, "row kernels: hidden size must be 768 or 384");
  if (hidden == 768)
    layernorm_kernel<6><<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(x32, T, gamma, beta, eps, h16,
                                                                               write_back ? 1 : 0, h16_lo);
  else
    layernorm_kernel<3><<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(x32, T, gamma, beta, eps, h16,
                                                                               write_back ? 1 : 0, h16_lo);
  VRAG_LAUNCHED(ctx);
}
void launch_head_final(vrag_ctx* ctx, const float* buf32, int T, const float* gamma, float eps, const float* cls_w,
                       const float* cls_b, float* logits, float* probs) {
  ProfScope prof(ctx, PROF_ROWOPS);
  head_final_kernel<<<row_blocks(T), 32 * ROWS_PER_BLOCK, 0, ctx->stream>>>(buf32, T, gamma, eps, cls_w, cls_b, logits,
                                                                             probs);
  VRAG_LAUNCHED(ctx);
}
void launch_head_finish(vrag_ctx* ctx, const float* head_part, int T, int slots, float eps, float g0, float g1,
                        const float* cls_b, float* logits, float* probs) {
  ProfScope prof(ctx, PROF_ROWOPS);
  head_finish_kernel<<<(T + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(head_part), T, slots,
                                                                1.0f / H, eps, g0, g1, cls_b, logits, probs);
  VRAG_LAUNCHED(ctx);
}
void launch_span_runs(vrag_ctx* ctx, const float* probs, const int32_t* cu, const int32_t* ctx_first, const int32_t* ctx_len,
                      const int64_t* tok_base, const int32_t* tcs, const int32_t* tce, int nseq, float threshold,
                      int min_span_chars, int merge_gap_chars, int32_t* counts, int32_t* offs, int32_t seq_base,
                      int32_t* o_seq, int32_t* o_cs, int32_t* o_ce, float* o_score, int32_t* o_ts, int32_t* o_te, bool fill) {
  ProfScope prof(ctx, PROF_ROWOPS);
  const int blocks = (nseq + 127) / 128;
  if (!fill) {
    span_runs_kernel<false><<<blocks, 128, 0, ctx->stream>>>(probs, cu, ctx_first, ctx_len, tok_base, tcs, tce, nseq, threshold,
                                                            min_span_chars, merge_gap_chars, counts, nullptr, seq_base,
                                                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    VRAG_LAUNCHED(ctx);
    exclusive_scan_small_kernel<<<1, 32, 0, ctx->stream>>>(counts, nseq, offs);
  } else {
    span_runs_kernel<true><<<blocks, 128, 0, ctx->stream>>>(probs, cu, ctx_first, ctx_len, tok_base, tcs, tce, nseq, threshold,
                                                           min_span_chars, merge_gap_chars, nullptr, offs, seq_base, o_seq,
                                                           o_cs, o_ce, o_score, o_ts, o_te);
  }
  VRAG_LAUNCHED(ctx);
}
void launch_cls_head(vrag_ctx* ctx, const float* x32, const int32_t* cu, int nseq, int hidden, const float* wp,
                     const float* bp, const float* wc, const float* bc, float* scores) {
  ProfScope prof(ctx, PROF_ROWOPS);
  cls_head_kernel<<<nseq, 256, 0, ctx->stream>>>(x32, cu, hidden, wp, bp, wc, bc, scores);
  VRAG_LAUNCHED(ctx);
}
void launch_sentence_head(vrag_ctx* ctx, const float* x32, int hidden, const int32_t* row_start, const int32_t* row_end,
                          int nsent, const float* cw, const float* cb, float* logits) {
  ProfScope prof(ctx, PROF_ROWOPS);
  sentence_head_kernel<<<nsent, 256, 0, ctx->stream>>>(x32, hidden, row_start, row_end, cw, cb, logits);
  VRAG_LAUNCHED(ctx);
}
void launch_splade_count(vrag_ctx* ctx, const float* dense, int nseq, int ld, int vocab, float min_abs,
                         int32_t* counts) {
  ProfScope prof(ctx, PROF_ROWOPS);
  splade_count_kernel<<<nseq, 256, 0, ctx->stream>>>(dense, ld, vocab, min_abs, counts);
  VRAG_LAUNCHED(ctx);
}
void launch_splade_fill(vrag_ctx* ctx, const float* dense, int nseq, int ld, int vocab, float min_abs,
                        const int64_t* indptr_dev, int32_t* indices, float* values) {
  ProfScope prof(ctx, PROF_ROWOPS);
  splade_fill_kernel<<<nseq, 256, 0, ctx->stream>>>(dense, ld, vocab, min_abs, indptr_dev, indices, values);
  VRAG_LAUNCHED(ctx);
}
void launch_pool(vrag_ctx* ctx, const float* x32, const int32_t* cu, int nseq, int pooling, int normalize,
                 float* out, int hidden) {
  ProfScope prof(ctx, PROF_ROWOPS);
  pool_kernel<<<nseq, 256, 0, ctx->stream>>>(x32, cu, pooling, normalize, out, hidden);
  VRAG_LAUNCHED(ctx);
}
void launch_f32_to_f16(vrag_ctx* ctx, const float* src, __half* dst, size_t n) {
  ProfScope prof(ctx, PROF_ROWOPS);
  const size_t threads = (n + 3) / 4;
  f32_to_f16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(src, dst, n);
  VRAG_LAUNCHED(ctx);
}

void launch_f32_to_f16_split(vrag_ctx* ctx, const float* src, __half* hi, __half* lo, size_t n) {
  ProfScope prof(ctx, PROF_ROWOPS);
  f32_to_f16_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(src, hi, lo, n);
  VRAG_LAUNCHED(ctx);
}

}  // namespace vrag
