// Context management, TMA descriptor encode, host span post-processing and the GEMM self test.
#include <math.h>
#include <string.h>

#include <memory>
#include <random>

#include "common.cuh"
#include "encoder.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

using namespace vrag;

static thread_local std::string g_create_error;

void* vrag_ctx::pinned_reserve(size_t n) {
  if (n <= pinned_bytes) return pinned;
  if (pinned) cudaFreeHost(pinned);
  pinned = nullptr;
  pinned_bytes = 0;
  VRAG_CUDA(cudaMallocHost(&pinned, n));
  pinned_bytes = n;
  return pinned;
}

namespace vrag {

CUtensorMap make_tmap_2d(vrag_ctx* ctx, const void* base, CUtensorMapDataType dt, size_t elem_bytes, uint64_t rows,
                         uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols,
                         CUtensorMapSwizzle swizzle) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  VRAG_CHECK(box_cols * elem_bytes == (swizzle == CU_TENSOR_MAP_SWIZZLE_64B ? 64u : 128u), VRAG_ERR_INTERNAL,
             "tensor map: inner box must span the swizzle width (128 / 64 bytes)");
  VRAG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (gstride[0] & 15) == 0, VRAG_ERR_INTERNAL,
             "tensor map: base / stride not 16-byte aligned");
  CUresult r = ctx->encode_tiled(&m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(VRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return m;
}

}  // namespace vrag

extern "C" const char* vrag_version(void) { return "vrag_b200 0.1.0 (sm_100a)"; }

extern "C" int vrag_ctx_create(int device, vrag_ctx** out) {
  if (!out) return VRAG_ERR_ARG;
  *out = nullptr;
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error(VRAG_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                     "); libvrag_b200 has no CPU fallback");
    VRAG_CHECK(device >= 0 && device < count, VRAG_ERR_ARG, "ctx_create: device index out of range");
    VRAG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VRAG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw Error(VRAG_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                     std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
    std::unique_ptr<vrag_ctx> c(new vrag_ctx());
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    VRAG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VRAG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) throw Error(VRAG_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    c->encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
    if (const char* st = getenv("VRAG_GEMM_STAGES")) {
      const int v = atoi(st);
      if (v >= 3 && v <= 5) c->gemm_stages = v;
    }
    *out = c.release();
    return VRAG_OK;
  } catch (const Error& e) {
    g_create_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return VRAG_ERR_INTERNAL;
  }
}

extern "C" void vrag_ctx_destroy(vrag_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  delete ctx;
}

extern "C" const char* vrag_last_error(vrag_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

extern "C" int vrag_sync(vrag_ctx* ctx) {
  if (!ctx) return VRAG_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->last_error = std::string("stream synchronize: ") + cudaGetErrorString(e);
    return VRAG_ERR_CUDA;
  }
  return VRAG_OK;
}

extern "C" int vrag_profile(vrag_ctx* ctx, int enable) {
  if (!ctx) return VRAG_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->prof.on = enable != 0;
  ctx->prof.used = 0;
  ctx->prof.recs.clear();
  return VRAG_OK;
}

extern "C" int vrag_profile_read(vrag_ctx* ctx, double* ms_per_class, int64_t* launches_per_class) {
  if (!ctx || !ms_per_class || !launches_per_class) return VRAG_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return VRAG_ERR_CUDA;
  for (int i = 0; i < PROF_NCLASS; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  for (const auto& r : ctx->prof.recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->prof.pool[r.e0], ctx->prof.pool[r.e1]);
    ms_per_class[r.cls] += ms;
    launches_per_class[r.cls] += 1;
  }
  ctx->prof.used = 0;
  ctx->prof.recs.clear();
  return VRAG_OK;
}

extern "C" void* vrag_stream(vrag_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
extern "C" uint64_t vrag_launch_count(vrag_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// Span post-processing: the integer half of the highlighter contract (oracle/highlighter.py steps 4-8;
// reference call site packages/core/verbatim_core/extractors.py:213-224).
// ------------------------------------------------------------------------------------------------
extern "C" int vrag_spans_from_probs(const float* probs, const int32_t* tcs, const int32_t* tce,
                                     const int64_t* ctx_indptr, int nctx, float threshold, int min_span_chars,
                                     int merge_gap_chars, int32_t* span_ctx, int32_t* span_cs, int32_t* span_ce,
                                     float* span_score, int32_t* span_ts, int32_t* span_te, int64_t cap,
                                     int64_t* nspans_out) {
  if (!probs || !tcs || !tce || !ctx_indptr || !nspans_out || nctx < 0) return VRAG_ERR_ARG;
  int64_t count = 0;
  for (int c = 0; c < nctx; ++c) {
    const int64_t a = ctx_indptr[c], b = ctx_indptr[c + 1];
    bool open = false;
    int32_t cs = 0, ce = 0, ts = 0, te = 0, cnt = 0;
    double acc = 0.0;
    auto flush = [&]() {
      if (open && ce - cs >= min_span_chars) {
        if (count < cap) {
          span_ctx[count] = c;
          span_cs[count] = cs;
          span_ce[count] = ce;
          span_score[count] = static_cast<float>(acc / cnt);
          span_ts[count] = ts;
          span_te[count] = te;
        }
        ++count;
      }
      open = false;
    };
    int64_t i = a;
    while (i < b) {
      if (!(probs[i] > threshold)) { ++i; continue; }
      int64_t j = i;
      double racc = 0.0;
      while (j < b && probs[j] > threshold) { racc += static_cast<double>(probs[j]); ++j; }
      const int32_t rs = tcs[i], re = tce[j - 1];
      if (open && rs - ce <= merge_gap_chars) {  // merge into the open span
        ce = re;
        te = static_cast<int32_t>(j - a);
        acc += racc;
        cnt += static_cast<int32_t>(j - i);
      } else {
        flush();
        open = true;
        cs = rs; ce = re;
        ts = static_cast<int32_t>(i - a); te = static_cast<int32_t>(j - a);
        acc = racc;
        cnt = static_cast<int32_t>(j - i);
      }
      i = j;
    }
    flush();
  }
  *nspans_out = count;
  return count <= cap ? VRAG_OK : VRAG_ERR_CAPACITY;
}

// ------------------------------------------------------------------------------------------------
// GEMM self test: tcgen05 path (staged TMA epilogues) vs the SIMT reference path (direct thread-per-row epilogue)
// on identical random operands, for the epilogues F32(10), F16(0), ROPE_QKV(1), RESID_F32(2), GEGLU(3).
// ------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void fill_half_kernel(__half* p, size_t n, uint32_t seed, float scale) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = __float2half_rn((static_cast<float>(hash_u32(static_cast<uint32_t>(i) * 2654435761u + seed) & 0xffff) / 32768.0f - 1.0f) * scale);
}
__global__ void fill_float_kernel(float* p, size_t n, uint32_t seed, float scale) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = (static_cast<float>(hash_u32(static_cast<uint32_t>(i) * 2654435761u + seed) & 0xffff) / 32768.0f - 1.0f) * scale;
}
// mode 0: sequences of 200 tokens back to back; 1: hashed positions; 2: sequences of 512 tokens
__global__ void fill_pos_kernel(int32_t* p, size_t n, int max_pos, int mode) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 1) p[i] = static_cast<int32_t>(hash_u32(static_cast<uint32_t>(i) + 7u) % max_pos);
  else p[i] = static_cast<int32_t>(i % (mode == 0 ? 200 : 512)) % max_pos;
}
// plausible partial row moments (sum, sum of squares) for the deferred-LayerNorm epilogues: |sum| <= 8, sumsq in [64, 192]
__global__ void fill_stats_kernel(float* p, size_t n_pairs, uint32_t seed) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const uint32_t h = hash_u32(static_cast<uint32_t>(i) * 2654435761u + seed);
  p[2 * i] = (static_cast<float>(h & 0xffff) / 32768.0f - 1.0f) * 8.0f;
  p[2 * i + 1] = 128.0f + (static_cast<float>(h >> 16) / 32768.0f - 1.0f) * 64.0f;
}
// residual value -> (fp16 hi, e5m2 lo) planes
__global__ void fill_hilo_kernel(__half* hi, uint8_t* lo, size_t n, uint32_t seed, float scale) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = (static_cast<float>(hash_u32(static_cast<uint32_t>(i) * 2654435761u + seed) & 0xffff) / 32768.0f - 1.0f) * scale;
  const __half h = __float2half_rn(x);
  hi[i] = h;
  lo[i] = static_cast<uint8_t>(pack_e5m2x2(x - __half2float(h), 0.f));
}
// max |(hi0 + lo0) - (hi1 + lo1)|
__global__ void diff_hilo_kernel(const __half* h0, const uint8_t* l0, const __half* h1, const uint8_t* l1, size_t n,
                                 float* out) {
  float d = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x = fabsf((__half2float(h0[i]) + unpack_e5m2x2<0>(l0[i]).x) - (__half2float(h1[i]) + unpack_e5m2x2<0>(l1[i]).x));
    if (!(x == x)) x = INFINITY;
    d = fmaxf(d, x);
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(d));
}
// max |(h0 + l0) - (h1 + l1)| and max |h1 + l1| over two fp16 plane pairs (split-precision outputs)
__global__ void diff_split_kernel(const __half* h0, const __half* l0, const __half* h1, const __half* l1, size_t n,
                                  float* out /*[2]*/) {
  float d = 0.f, m = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float a = __half2float(h0[i]) + __half2float(l0[i]), b = __half2float(h1[i]) + __half2float(l1[i]);
    float x = fabsf(a - b);
    if (!(x == x)) x = INFINITY;
    d = fmaxf(d, x);
    m = fmaxf(m, fabsf(b));
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(d));
  atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(m));
}
template <typename T>
__global__ void diff_kernel(const T* a, const T* b, size_t n, float* out /*[2]: max diff, max |b|*/) {
  float d = 0.f, m = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float fa = static_cast<float>(a[i]), fb = static_cast<float>(b[i]);
    float x = fabsf(fa - fb);
    if (!(x == x)) x = INFINITY;  // NaN (e.g. an output the kernel never wrote) must not hide behind fmaxf
    d = fmaxf(d, x);
    m = fmaxf(m, fabsf(fb));
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(d));
  atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(m));
}
inline unsigned blocks_for(size_t n) { return static_cast<unsigned>((n + 255) / 256); }
}  // namespace

static int selftest_gemm_impl(vrag_ctx* ctx, int M, int N, int K, int epilogue, bool split, double* max_abs_diff,
                              double* ref_abs_max);

extern "C" int vrag_selftest_gemm(vrag_ctx* ctx, int M, int N, int K, int epilogue, double* max_abs_diff,
                                  double* ref_abs_max) {
  return selftest_gemm_impl(ctx, M, N, K, epilogue, false, max_abs_diff, ref_abs_max);
}
extern "C" int vrag_selftest_gemm_split(vrag_ctx* ctx, int M, int N, int K, int epilogue, double* max_abs_diff,
                                        double* ref_abs_max) {
  return selftest_gemm_impl(ctx, M, N, K, epilogue, true, max_abs_diff, ref_abs_max);
}

static int selftest_gemm_impl(vrag_ctx* ctx, int M, int N, int K, int epilogue, bool split, double* max_abs_diff,
                              double* ref_abs_max) {
  if (!ctx || !max_abs_diff) return VRAG_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  try {
    VRAG_CUDA(cudaSetDevice(ctx->device));
    const bool stats = epilogue == EPI_RESID_STATS || epilogue == EPI_RESID_STATS_LN;
    const bool norm_bias = epilogue == EPI_NORM_BIAS_F16 || epilogue == EPI_NORM_BIAS_GELU_F16;
    const bool f32_out = epilogue == EPI_F32 || epilogue == EPI_RESID_F32;
    const bool rope = epilogue == EPI_ROPE_QKV || epilogue == EPI_NORM_ROPE_QKV;
    const bool geglu = epilogue == EPI_GEGLU || epilogue == EPI_NORM_GEGLU;
    VRAG_CHECK(epilogue == EPI_F32 || epilogue == EPI_F16 || epilogue == EPI_RESID_F32 || rope || geglu || stats ||
                   norm_bias,
               VRAG_ERR_ARG, "selftest_gemm: epilogue must be one of 10, 0, 1, 2, 3, 11, 12, 13, 14, 15, 16");
    VRAG_CHECK(!rope || N % 192 == 0, VRAG_ERR_ARG, "selftest_gemm: ROPE needs N = 3 * hidden");
    VRAG_CHECK(!split || !(stats || norm_bias || epilogue == EPI_NORM_ROPE_QKV || epilogue == EPI_NORM_GEGLU),
               VRAG_ERR_ARG, "selftest_gemm_split: epilogue must be one of 10, 0, 1, 2, 3");
    const int out_cols = geglu ? N / 2 : N;
    const size_t out_n = static_cast<size_t>(M) * out_cols;
    const size_t out_bytes = out_n * (f32_out ? 4 : 2);
    const int max_pos = 512;
    const int slots = N / 128;   // EPI_RESID_STATS writes one (sum, sumsq) pair per row and 128 columns
    const size_t st_pairs = static_cast<size_t>(M) * (stats ? slots : 6);
    DevBuf A, W, C0, C1, H0, H1, S0, S1, R, POS, CS, SIN, BIAS, GAMMA, AL, WL;
    SIN.reserve(static_cast<size_t>(M) * 6 * 8);
    BIAS.reserve(static_cast<size_t>(N) * 4);
    GAMMA.reserve(static_cast<size_t>(N) * 4);
    A.reserve(static_cast<size_t>(M) * K * 2);
    W.reserve(static_cast<size_t>(N) * K * 2);
    C0.reserve(out_bytes);
    C1.reserve(out_bytes);
    H0.reserve(out_n * 2);
    H1.reserve(out_n * 2);
    S0.reserve(st_pairs * 8);
    S1.reserve(st_pairs * 8);
    R.reserve(8);
    POS.reserve(static_cast<size_t>(M) * 4);
    CS.reserve(static_cast<size_t>(max_pos) * 64 * 4);
    cudaStream_t st = ctx->stream;
    fill_half_kernel<<<blocks_for(static_cast<size_t>(M) * K), 256, 0, st>>>(A.as<__half>(), static_cast<size_t>(M) * K, 17u, 1.0f);
    fill_half_kernel<<<blocks_for(static_cast<size_t>(N) * K), 256, 0, st>>>(W.as<__half>(), static_cast<size_t>(N) * K, 91u, 0.05f);
    if (split) {   // low planes: remainders of the size a real hi / lo split produces (|lo| <= ulp(hi) / 2)
      AL.reserve(static_cast<size_t>(M) * K * 2);
      WL.reserve(static_cast<size_t>(N) * K * 2);
      fill_half_kernel<<<blocks_for(static_cast<size_t>(M) * K), 256, 0, st>>>(AL.as<__half>(), static_cast<size_t>(M) * K, 29u, 1.0f / 2048.0f);
      fill_half_kernel<<<blocks_for(static_cast<size_t>(N) * K), 256, 0, st>>>(WL.as<__half>(), static_cast<size_t>(N) * K, 57u, 0.05f / 2048.0f);
    }
    // positions: runs of consecutive positions (sequences of 200 tokens: slabs of 32 rows that straddle a boundary take
    // the gather path, the others the TMA path), or hashed positions (pos_mode 1: every slab gathers)
    fill_pos_kernel<<<blocks_for(M), 256, 0, st>>>(POS.as<int32_t>(), M, max_pos, (M % 2) ? 1 : 0);
    fill_float_kernel<<<blocks_for(max_pos * 64), 256, 0, st>>>(CS.as<float>(), max_pos * 64, 5u, 1.0f);
    if (epilogue == EPI_RESID_F32) {  // both paths accumulate onto the same initial residual
      fill_float_kernel<<<blocks_for(out_n), 256, 0, st>>>(C0.as<float>(), out_n, 33u, 1.0f);
      VRAG_CUDA(cudaMemcpyAsync(C1.p, C0.p, out_bytes, cudaMemcpyDeviceToDevice, st));
    } else if (stats) {               // ... held as two fp16 planes: C = hi, H = lo
      fill_hilo_kernel<<<blocks_for(out_n), 256, 0, st>>>(C0.as<__half>(), H0.as<uint8_t>(), out_n, 33u, 1.0f);
      VRAG_CUDA(cudaMemcpyAsync(C1.p, C0.p, out_n * 2, cudaMemcpyDeviceToDevice, st));
      VRAG_CUDA(cudaMemcpyAsync(H1.p, H0.p, out_n, cudaMemcpyDeviceToDevice, st));
    } else {
      VRAG_CUDA(cudaMemsetAsync(C0.p, 0xff, out_bytes, st));  // NaN pattern: unwritten outputs are detected
      VRAG_CUDA(cudaMemsetAsync(C1.p, 0, out_bytes, st));
      if (split) {
        VRAG_CUDA(cudaMemsetAsync(H0.p, 0xff, out_n * 2, st));
        VRAG_CUDA(cudaMemsetAsync(H1.p, 0, out_n * 2, st));
      }
    }
    fill_float_kernel<<<blocks_for(N), 256, 0, st>>>(BIAS.as<float>(), N, 71u, 0.5f);
    fill_float_kernel<<<blocks_for(N), 256, 0, st>>>(GAMMA.as<float>(), N, 72u, 1.0f);
    fill_stats_kernel<<<blocks_for(static_cast<size_t>(M) * 6), 256, 0, st>>>(SIN.as<float>(), static_cast<size_t>(M) * 6, 9u);
    if (stats) {
      VRAG_CUDA(cudaMemsetAsync(S0.p, 0xff, st_pairs * 8, st));
      VRAG_CUDA(cudaMemsetAsync(S1.p, 0, st_pairs * 8, st));
    } else {   // moments read by the EPI_NORM_* epilogues (identical for both paths)
      fill_stats_kernel<<<blocks_for(st_pairs), 256, 0, st>>>(S0.as<float>(), st_pairs, 8u);
    }
    VRAG_CUDA(cudaMemsetAsync(R.p, 0, 8, st));
    for (int ref = 0; ref < 2; ++ref) {
      GemmEpiParams p;
      p.M = M;
      p.ld32 = out_cols; p.ld16 = out_cols;
      p.out32 = (ref ? C1 : C0).as<float>();
      p.out16 = (ref ? C1 : C0).as<__half>();
      p.out8_lo = (ref ? H1 : H0).as<uint8_t>();
      if (split) {
        p.a_lo = AL.as<__half>();
        p.w_lo = WL.as<__half>();
        p.out16_lo = (ref ? H1 : H0).as<__half>();
      }
      p.pos = POS.as<int32_t>(); p.rope_tab = CS.as<float>(); p.rope_rows = max_pos;
      p.hidden = N / 3;
      p.stats_in = stats ? SIN.as<float>() : S0.as<float>();   // EPI_RESID_STATS_LN reads the old moments, writes new ones
      p.stats_out = stats ? (ref ? S1 : S0).as<float>() : nullptr;
      p.bias = BIAS.as<float>();
      p.gamma = GAMMA.as<float>();
      p.stats_slots = 6;
      launch_gemm(ctx, epilogue, A.as<__half>(), W.as<__half>(), M, N, K, p, ref);
    }
    if (f32_out) diff_kernel<float><<<256, 256, 0, st>>>(C0.as<float>(), C1.as<float>(), out_n, R.as<float>());
    else if (split) diff_split_kernel<<<256, 256, 0, st>>>(C0.as<__half>(), H0.as<__half>(), C1.as<__half>(), H1.as<__half>(), out_n, R.as<float>());
    else diff_kernel<__half><<<256, 256, 0, st>>>(C0.as<__half>(), C1.as<__half>(), out_n, R.as<float>());
    float h[2], hs[2] = {0.f, 0.f};
    VRAG_CUDA(cudaMemcpyAsync(h, R.p, 8, cudaMemcpyDeviceToHost, st));
    VRAG_CUDA(cudaStreamSynchronize(st));
    if (stats) {   // hi planes compared above (fp16); hi + lo as the stream, scaled to the same tolerance (x 8: callers
                   // allow 2e-3 * max on fp16 outputs, 2.5e-4 * max = 2 e5m2 steps of the low plane on the stream -- a
                   // 1e-7 difference of the fp32 sums can move lo across one rounding boundary); the moments relative
      VRAG_CUDA(cudaMemsetAsync(R.p, 0, 8, st));
      diff_hilo_kernel<<<256, 256, 0, st>>>(C0.as<__half>(), H0.as<uint8_t>(), C1.as<__half>(), H1.as<uint8_t>(), out_n,
                                            R.as<float>());
      float hx[2];
      VRAG_CUDA(cudaMemcpyAsync(hx, R.p, 8, cudaMemcpyDeviceToHost, st));
      VRAG_CUDA(cudaStreamSynchronize(st));
      h[0] = (hx[0] == hx[0]) ? fmaxf(h[0], 8.0f * hx[0]) : NAN;
      VRAG_CUDA(cudaMemsetAsync(R.p, 0, 8, st));
      diff_kernel<float><<<256, 256, 0, st>>>(S0.as<float>(), S1.as<float>(), st_pairs * 2, R.as<float>());
      VRAG_CUDA(cudaMemcpyAsync(hs, R.p, 8, cudaMemcpyDeviceToHost, st));
      VRAG_CUDA(cudaStreamSynchronize(st));
      const float rel = hs[0] / fmaxf(hs[1], 1.0f);
      if (!(rel == rel)) h[0] = NAN;
      else h[0] = fmaxf(h[0], rel * fmaxf(h[1], 1.0f));
    }
    *max_abs_diff = std::isnan(h[0]) ? INFINITY : h[0];
    if (ref_abs_max) *ref_abs_max = h[1];
    for (DevBuf* b : {&A, &W, &C0, &C1, &H0, &H1, &S0, &S1, &R, &POS, &CS, &SIN, &BIAS, &GAMMA, &AL, &WL}) b->release();
    return VRAG_OK;
  } catch (const Error& e) {
    ctx->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    ctx->last_error = e.what();
    return VRAG_ERR_INTERNAL;
  }
}

// GEMM timing hook (development / bench.py's per-kernel table): `iters` back-to-back launches of one encoder GEMM
// shape on synthetic operands, CUDA events on the library's stream; *ms_out = average launch time.
static int bench_gemm_impl(vrag_ctx* ctx, int M, int N, int K, int epilogue, int stages, int debug_mode, int iters,
                           bool split, double* ms_out);
extern "C" int vrag_bench_gemm(vrag_ctx* ctx, int M, int N, int K, int epilogue, int stages, int debug_mode, int iters,
                               double* ms_out) {
  return bench_gemm_impl(ctx, M, N, K, epilogue, stages, debug_mode, iters, false, ms_out);
}
extern "C" int vrag_bench_gemm_split(vrag_ctx* ctx, int M, int N, int K, int epilogue, int iters, double* ms_out) {
  return bench_gemm_impl(ctx, M, N, K, epilogue, 0, 0, iters, true, ms_out);
}
static int bench_gemm_impl(vrag_ctx* ctx, int M, int N, int K, int epilogue, int stages, int debug_mode, int iters,
                           bool split, double* ms_out) {
  if (!ctx || !ms_out || iters < 1) return VRAG_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const int saved_stages = ctx->gemm_stages;
  try {
    VRAG_CUDA(cudaSetDevice(ctx->device));
    VRAG_CHECK(epilogue == EPI_F16 || epilogue == EPI_ROPE_QKV || epilogue == EPI_RESID_F32 || epilogue == EPI_GEGLU ||
                   epilogue == EPI_RESID_STATS || epilogue == EPI_NORM_ROPE_QKV || epilogue == EPI_NORM_GEGLU ||
                   epilogue == EPI_NORM_BIAS_F16 || epilogue == EPI_NORM_BIAS_GELU_F16 || epilogue == EPI_RESID_STATS_LN,
               VRAG_ERR_ARG, "bench_gemm: unsupported epilogue");
    const bool f32_out = epilogue == EPI_RESID_F32 || epilogue == EPI_RESID_STATS || epilogue == EPI_RESID_STATS_LN;
    const bool geglu = epilogue == EPI_GEGLU || epilogue == EPI_NORM_GEGLU;
    const int out_cols = geglu ? N / 2 : N;
    const size_t out_n = static_cast<size_t>(M) * out_cols;
    const int max_pos = 512;
    VRAG_CHECK(!split || epilogue == EPI_F16 || epilogue == EPI_ROPE_QKV || epilogue == EPI_RESID_F32 ||
                   epilogue == EPI_GEGLU, VRAG_ERR_ARG, "bench_gemm_split: unsupported epilogue");
    DevBuf A, W, C, C16, POS, CS, ST;
    A.reserve(static_cast<size_t>(M) * K * 2 * (split ? 2 : 1));   // split: the lo plane follows the hi plane
    W.reserve(static_cast<size_t>(N) * K * 2 * (split ? 2 : 1));
    C.reserve(out_n * (f32_out ? 4 : 2));
    C16.reserve(out_n * 2);
    POS.reserve(static_cast<size_t>(M) * 4);
    CS.reserve(static_cast<size_t>(max_pos) * 64 * 4);
    ST.reserve(static_cast<size_t>(M) * 24 * 4);
    cudaStream_t st = ctx->stream;
    fill_half_kernel<<<blocks_for(static_cast<size_t>(M) * K), 256, 0, st>>>(A.as<__half>(), static_cast<size_t>(M) * K, 17u, 1.0f);
    fill_half_kernel<<<blocks_for(static_cast<size_t>(N) * K), 256, 0, st>>>(W.as<__half>(), static_cast<size_t>(N) * K, 91u, 0.05f);
    fill_pos_kernel<<<blocks_for(M), 256, 0, st>>>(POS.as<int32_t>(), M, max_pos, 2);   // sequences of 512 tokens
    fill_float_kernel<<<blocks_for(max_pos * 64), 256, 0, st>>>(CS.as<float>(), max_pos * 64, 5u, 1.0f);
    fill_stats_kernel<<<blocks_for(static_cast<size_t>(M) * 6), 256, 0, st>>>(ST.as<float>(), static_cast<size_t>(M) * 6, 8u);
    VRAG_CUDA(cudaMemsetAsync(C.p, 0, out_n * (f32_out ? 4 : 2), st));
    GemmEpiParams p;
    p.M = M;
    p.ld32 = out_cols; p.ld16 = out_cols;
    p.out32 = C.as<float>();
    p.out16 = f32_out ? C16.as<__half>() : C.as<__half>();
    if (f32_out && epilogue != EPI_RESID_F32) {   // two planes: C (first half) = hi, C16 = lo (e5m2)
      p.out16 = C.as<__half>();
      p.out8_lo = C16.as<uint8_t>();
    }
    p.pos = POS.as<int32_t>(); p.rope_tab = CS.as<float>(); p.rope_rows = max_pos;
    p.hidden = N / 3;
    p.stats_in = ST.as<float>(); p.stats_out = ST.as<float>(); p.stats_slots = 6;
    if (epilogue == EPI_RESID_STATS_LN) p.stats_out = ST.as<float>() + static_cast<size_t>(M) * 12;
    p.bias = CS.as<float>(); p.gamma = CS.as<float>() + 4096;   // any finite per-column vectors (N <= 4096)
    p.debug_mode = debug_mode;
    if (split) {
      VRAG_CUDA(cudaMemsetAsync(A.as<__half>() + static_cast<size_t>(M) * K, 0, static_cast<size_t>(M) * K * 2, st));
      VRAG_CUDA(cudaMemsetAsync(W.as<__half>() + static_cast<size_t>(N) * K, 0, static_cast<size_t>(N) * K * 2, st));
      p.a_lo = A.as<__half>() + static_cast<size_t>(M) * K;
      p.w_lo = W.as<__half>() + static_cast<size_t>(N) * K;
      p.out16_lo = C16.as<__half>();
    }
    if (stages >= 3 && stages <= 5) ctx->gemm_stages = stages;
    cudaEvent_t e0, e1;
    VRAG_CUDA(cudaEventCreate(&e0));
    VRAG_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch_gemm(ctx, epilogue, A.as<__half>(), W.as<__half>(), M, N, K, p, 0);
    VRAG_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) launch_gemm(ctx, epilogue, A.as<__half>(), W.as<__half>(), M, N, K, p, 0);
    VRAG_CUDA(cudaEventRecord(e1, st));
    VRAG_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    VRAG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->gemm_stages = saved_stages;
    for (DevBuf* b : {&A, &W, &C, &C16, &POS, &CS, &ST}) b->release();
    return VRAG_OK;
  } catch (const Error& e) {
    ctx->gemm_stages = saved_stages;
    ctx->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    ctx->gemm_stages = saved_stages;
    ctx->last_error = e.what();
    return VRAG_ERR_INTERNAL;
  }
}

// ------------------------------------------------------------------------------------------------
// Attention timing hook (development, tools/attn_probe.py): `iters` back-to-back launches of the tcgen05 attention
// kernel on synthetic fp16 q|k|v rows (nseq sequences of seq_len tokens, 12 heads x 64), CUDA events on the
// library's stream; *ms_out = average launch time.  window < 0: full attention, else keys with |i - j| <= window.
// ------------------------------------------------------------------------------------------------
static int bench_attention_impl(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, bool split, double* ms_out);
extern "C" int vrag_bench_attention(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, double* ms_out) {
  return bench_attention_impl(ctx, nseq, seq_len, window, iters, false, ms_out);
}
extern "C" int vrag_bench_attention_split(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, double* ms_out) {
  return bench_attention_impl(ctx, nseq, seq_len, window, iters, true, ms_out);
}
static int bench_attention_impl(vrag_ctx* ctx, int nseq, int seq_len, int window, int iters, bool split, double* ms_out) {
  if (!ctx || !ms_out || nseq < 1 || seq_len < 1 || iters < 1) return VRAG_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  try {
    VRAG_CUDA(cudaSetDevice(ctx->device));
    const size_t T = static_cast<size_t>(nseq) * seq_len;
    VRAG_CHECK(T < (size_t(1) << 30), VRAG_ERR_ARG, "bench_attention: too many tokens");
    std::vector<int32_t> cu(nseq + 1), work;
    for (int i = 0; i <= nseq; ++i) cu[i] = i * seq_len;
    for (int i = 0; i < nseq; ++i)
      for (int q0 = 0; q0 < seq_len; q0 += 128) {
        work.push_back(cu[i]);
        work.push_back(seq_len);
        work.push_back(q0);
        work.push_back(0);
      }
    DevBuf QKV, OUT, CU, WORK;
    QKV.reserve(T * 3 * HIDDEN * 2 * (split ? 2 : 1));   // split: the lo plane follows the hi plane
    OUT.reserve(T * HIDDEN * 2 * (split ? 2 : 1));
    CU.reserve(cu.size() * 4);
    WORK.reserve(work.size() * 4);
    cudaStream_t st = ctx->stream;
    fill_half_kernel<<<blocks_for(T * 3 * HIDDEN), 256, 0, st>>>(QKV.as<__half>(), T * 3 * HIDDEN, 23u, 1.0f);
    if (split)
      fill_half_kernel<<<blocks_for(T * 3 * HIDDEN), 256, 0, st>>>(QKV.as<__half>() + T * 3 * HIDDEN, T * 3 * HIDDEN, 41u,
                                                                    1.0f / 2048.0f);
    VRAG_CUDA(cudaMemcpyAsync(CU.p, cu.data(), cu.size() * 4, cudaMemcpyHostToDevice, st));
    VRAG_CUDA(cudaMemcpyAsync(WORK.p, work.data(), work.size() * 4, cudaMemcpyHostToDevice, st));
    auto launch = [&]() {
      if (split)
        launch_attention_tc_split(ctx, QKV.as<__half>(), QKV.as<__half>() + T * 3 * HIDDEN, OUT.as<__half>(),
                                  OUT.as<__half>() + T * HIDDEN, WORK.as<int32_t>(), static_cast<int>(work.size() / 4),
                                  static_cast<int>(T), 12, HIDDEN, window);
      else
        launch_attention_tc(ctx, QKV.as<__half>(), OUT.as<__half>(), CU.as<int32_t>(), WORK.as<int32_t>(),
                            static_cast<int>(work.size() / 4), static_cast<int>(T), 12, HIDDEN, window);
    };
    cudaEvent_t e0, e1;
    VRAG_CUDA(cudaEventCreate(&e0));
    VRAG_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    VRAG_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) launch();
    VRAG_CUDA(cudaEventRecord(e1, st));
    VRAG_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    VRAG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (DevBuf* b : {&QKV, &OUT, &CU, &WORK}) b->release();
    return VRAG_OK;
  } catch (const Error& e) {
    ctx->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    ctx->last_error = e.what();
    return VRAG_ERR_INTERNAL;
  }
}

// ------------------------------------------------------------------------------------------------
// Attention self test: runs one attention launch (tcgen05 kernel, or the mma.sync cross-check with legacy != 0) on
// caller-supplied fp16 q|k|v rows, so tests can drive score ranges the encoder never produces (online-softmax
// rescaling, ragged tails, local windows) against a float64 host computation.
// ------------------------------------------------------------------------------------------------
static int selftest_attention_impl(vrag_ctx* ctx, const uint16_t* qkv_f16, const uint16_t* qkv_lo_f16,
                                   const int32_t* cu_seqlens, int nseq, int window, int legacy, uint16_t* out_f16,
                                   uint16_t* out_lo_f16);
extern "C" int vrag_selftest_attention(vrag_ctx* ctx, const uint16_t* qkv_f16, const int32_t* cu_seqlens, int nseq,
                                       int window, int legacy, uint16_t* out_f16) {
  return selftest_attention_impl(ctx, qkv_f16, nullptr, cu_seqlens, nseq, window, legacy, out_f16, nullptr);
}
extern "C" int vrag_selftest_attention_split(vrag_ctx* ctx, const uint16_t* qkv_hi_f16, const uint16_t* qkv_lo_f16,
                                             const int32_t* cu_seqlens, int nseq, int window, uint16_t* out_hi_f16,
                                             uint16_t* out_lo_f16) {
  if (!qkv_lo_f16 || !out_lo_f16) return VRAG_ERR_ARG;
  return selftest_attention_impl(ctx, qkv_hi_f16, qkv_lo_f16, cu_seqlens, nseq, window, 0, out_hi_f16, out_lo_f16);
}
static int selftest_attention_impl(vrag_ctx* ctx, const uint16_t* qkv_f16, const uint16_t* qkv_lo_f16,
                                   const int32_t* cu_seqlens, int nseq, int window, int legacy, uint16_t* out_f16,
                                   uint16_t* out_lo_f16) {
  if (!ctx || !qkv_f16 || !cu_seqlens || !out_f16 || nseq < 1) return VRAG_ERR_ARG;
  const bool split = qkv_lo_f16 != nullptr;
  std::lock_guard<std::mutex> lk(ctx->mu);
  try {
    VRAG_CUDA(cudaSetDevice(ctx->device));
    const int T = cu_seqlens[nseq];
    VRAG_CHECK(cu_seqlens[0] == 0 && T > 0, VRAG_ERR_ARG, "selftest_attention: cu_seqlens must start at 0");
    std::vector<int32_t> work;
    int max_len = 0;
    for (int i = 0; i < nseq; ++i) {
      const int L = cu_seqlens[i + 1] - cu_seqlens[i];
      VRAG_CHECK(L > 0, VRAG_ERR_ARG, "selftest_attention: empty sequence");
      max_len = std::max(max_len, L);
      for (int q0 = 0; q0 < L; q0 += 128) {
        work.push_back(cu_seqlens[i]);
        work.push_back(L);
        work.push_back(q0);
        work.push_back(0);
      }
    }
    DevBuf QKV, OUT, CU, WORK, QKVL, OUTL;
    QKV.reserve(static_cast<size_t>(T) * 3 * HIDDEN * 2);
    OUT.reserve(static_cast<size_t>(T) * HIDDEN * 2);
    if (split) {
      QKVL.reserve(static_cast<size_t>(T) * 3 * HIDDEN * 2);
      OUTL.reserve(static_cast<size_t>(T) * HIDDEN * 2);
    }
    CU.reserve(static_cast<size_t>(nseq + 1) * 4);
    WORK.reserve(work.size() * 4);
    cudaStream_t st = ctx->stream;
    VRAG_CUDA(cudaMemcpyAsync(QKV.p, qkv_f16, static_cast<size_t>(T) * 3 * HIDDEN * 2, cudaMemcpyHostToDevice, st));
    VRAG_CUDA(cudaMemcpyAsync(CU.p, cu_seqlens, static_cast<size_t>(nseq + 1) * 4, cudaMemcpyHostToDevice, st));
    VRAG_CUDA(cudaMemcpyAsync(WORK.p, work.data(), work.size() * 4, cudaMemcpyHostToDevice, st));
    VRAG_CUDA(cudaMemsetAsync(OUT.p, 0xff, static_cast<size_t>(T) * HIDDEN * 2, st));  // NaN pattern
    if (split) {
      VRAG_CUDA(cudaMemcpyAsync(QKVL.p, qkv_lo_f16, static_cast<size_t>(T) * 3 * HIDDEN * 2, cudaMemcpyHostToDevice, st));
      VRAG_CUDA(cudaMemsetAsync(OUTL.p, 0xff, static_cast<size_t>(T) * HIDDEN * 2, st));
      launch_attention_tc_split(ctx, QKV.as<__half>(), QKVL.as<__half>(), OUT.as<__half>(), OUTL.as<__half>(),
                                WORK.as<int32_t>(), static_cast<int>(work.size() / 4), T, 12, HIDDEN, window);
      VRAG_CUDA(cudaMemcpyAsync(out_lo_f16, OUTL.p, static_cast<size_t>(T) * HIDDEN * 2, cudaMemcpyDeviceToHost, st));
    } else if (legacy)
      launch_attention(ctx, QKV.as<__half>(), OUT.as<__half>(), CU.as<int32_t>(), nseq, max_len, 12, HIDDEN, window);
    else
      launch_attention_tc(ctx, QKV.as<__half>(), OUT.as<__half>(), CU.as<int32_t>(), WORK.as<int32_t>(),
                          static_cast<int>(work.size() / 4), T, 12, HIDDEN, window);
    VRAG_CUDA(cudaMemcpyAsync(out_f16, OUT.p, static_cast<size_t>(T) * HIDDEN * 2, cudaMemcpyDeviceToHost, st));
    VRAG_CUDA(cudaStreamSynchronize(st));
    for (DevBuf* b : {&QKV, &OUT, &CU, &WORK, &QKVL, &OUTL}) b->release();
    return VRAG_OK;
  } catch (const Error& e) {
    ctx->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    ctx->last_error = e.what();
    return VRAG_ERR_INTERNAL;
  }
}
