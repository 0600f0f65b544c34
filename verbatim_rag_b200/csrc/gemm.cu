// tcgen05 GEMM for sm_100a: persistent, warp-specialised.
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B boxes into a 4-stage smem ring)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma 128x256x16, fp32 accumulators in TMEM)
//   warps 2..9  : epilogue       (tcgen05.ld 32 lanes x 32 columns -> fused tail -> swizzled smem box ->
//                                 TMA store, or TMA reduce-add for the fp32 residual stream: x += acc happens in
//                                 the memory system, the SM never reads x)
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// Tiles are walked n-fastest so concurrently running CTAs share the same A rows (L2 reuse); W is L2 resident.
#include <atomic>

#include "gemm.cuh"
#include "ptx.cuh"

namespace vrag {

namespace {

constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = (BN / 2) * BK * 2;  // each CTA of the pair holds HALF of the 256-row W tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
// per epilogue warp: two 4 KB TMA-store boxes (32 rows x 128 B, SWIZZLE_128B); the RoPE epilogue also borrows them
// to transpose the cos/sin rows of its 32 tokens into registers
constexpr int EPI_BOX_BYTES = 4096;
constexpr int EPI_WARP_BYTES = 2 * EPI_BOX_BYTES;
constexpr int STATS_BARS = 3;     // mbarriers reserved per epilogue warp (EPI_RESID_STATS uses 2, RoPE 1)
constexpr int EPI_LO_BOX_BYTES = 2048;                        // 32 rows x 64 e5m2 bytes, SWIZZLE_64B
// EPI_RESID_STATS: 2 buffers x (hi box: 32 rows x 64 fp16, lo box: 32 rows x 64 e5m2); layout [hi0][hi1][lo0][lo1]
constexpr int EPI_STATS_WARP_BYTES = 2 * (EPI_BOX_BYTES + EPI_LO_BOX_BYTES);
constexpr int EPI_WARPS = 8;  // two per TMEM lane quarter (each takes half of the tile's columns): one warp per
                              // scheduler cannot hide its own ALU / TMEM-load latency, two can
constexpr int epi_warp_bytes(int epi) {
  return (epi == EPI_RESID_STATS || epi == EPI_RESID_STATS_LN) ? EPI_STATS_WARP_BYTES : EPI_WARP_BYTES;
}
constexpr int smem_bytes(int stages, int epi) {
  return stages * STAGE_BYTES + EPI_WARPS * epi_warp_bytes(epi) + 1024 /*align slack*/ + 512 /*barriers*/;
}
template <int EPI>
constexpr bool kNorm = (EPI == EPI_NORM_ROPE_QKV || EPI == EPI_NORM_GEGLU);
template <int EPI>
constexpr bool kRope = (EPI == EPI_ROPE_QKV || EPI == EPI_NORM_ROPE_QKV);
template <int EPI>
constexpr bool kGeglu = (EPI == EPI_GEGLU || EPI == EPI_NORM_GEGLU);

// Deferred LayerNorm: 1 / sqrt(var + eps) of row `row` of the A operand from the partial moments its producer wrote
// (EPI_RESID_STATS).  One-pass variance E[x^2] - mean^2 in fp32 over 768 columns: relative error ~1e-6 unless
// mean^2 >> var, which a residual stream does not do.
__device__ __forceinline__ float row_rstd(const GemmEpiParams& p, int row) {
  if (row >= p.M) return 0.f;
  const float2* st = reinterpret_cast<const float2*>(p.stats_in);
  float s = 0.f, q = 0.f;
  for (int j = 0; j < p.stats_slots; ++j) {
    const float2 v = __ldg(st + static_cast<size_t>(j) * p.M + row);
    s += v.x;
    q += v.y;
  }
  const float mean = s * p.inv_dim;
  const float var = fmaxf(q * p.inv_dim - mean * mean, 0.f);
  return rsqrtf(var + p.ln_eps);
}
// mean and 1 / sqrt(var + eps) of row `row` from the same partial moments (EPI_RESID_STATS_LN normalises the old stream)
__device__ __forceinline__ float2 row_mean_rstd(const GemmEpiParams& p, int row) {
  if (row >= p.M) return make_float2(0.f, 0.f);
  const float2* st = reinterpret_cast<const float2*>(p.stats_in);
  float s = 0.f, q = 0.f;
  for (int j = 0; j < p.stats_slots; ++j) {
    const float2 v = __ldg(st + static_cast<size_t>(j) * p.M + row);
    s += v.x;
    q += v.y;
  }
  const float mean = s * p.inv_dim;
  const float var = fmaxf(q * p.inv_dim - mean * mean, 0.f);
  return make_float2(mean, rsqrtf(var + p.ln_eps));
}
template <int EPI>
constexpr bool kResidStats = (EPI == EPI_RESID_STATS || EPI == EPI_RESID_STATS_LN);
template <int EPI>
constexpr bool kNormBias = (EPI == EPI_NORM_BIAS_F16 || EPI == EPI_NORM_BIAS_GELU_F16);
constexpr int GEMM_THREADS = 32 * (2 + EPI_WARPS);
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// GeLU for two values at once on the packed fp32x2 pipe.  gelu(x) = x Phi(x) = 0.5 (x + |x|) - 0.5 |x| P(t) exp(-x^2 / 2)
// with erfc(z) ~ P(t) exp(-z^2), t = 1 / (1 + p z), z = |x| / sqrt(2) (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7;
// the form above has no cancellation for x < 0).  Measured against the exact erf form over [-12, 12]: max abs error
// 3.3e-7 -- three orders below the fp16 rounding of the GeGLU output.  ~10 issue slots per element instead of ~35 for
// erff: the Wi epilogue was issue-bound on the activation (vrag_bench_gemm, profiles/README.md).
__device__ __forceinline__ uint64_t gelu2(uint64_t x2) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  const uint64_t ax = f2_pack(fabsf(x0), fabsf(x1));
  float d0, d1;
  f2_unpack(f2_fma(ax, f2_pack(0.23164189f, 0.23164189f), f2_pack(1.f, 1.f)), d0, d1);   // 1 + (p / sqrt 2) |x|
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = f2_pack(t0, t1);
  uint64_t q = f2_fma(t, f2_pack(1.061405429f, 1.061405429f), f2_pack(-1.453152027f, -1.453152027f));
  q = f2_fma(q, t, f2_pack(1.421413741f, 1.421413741f));
  q = f2_fma(q, t, f2_pack(-0.284496736f, -0.284496736f));
  q = f2_fma(q, t, f2_pack(0.254829592f, 0.254829592f));
  q = f2_mul(q, t);
  float g0, g1;
  f2_unpack(f2_mul(f2_mul(x2, f2_pack(-0.72134752f, -0.72134752f)), x2), g0, g1);   // -x^2 log2(e) / 2
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t pe = f2_mul(q, f2_pack(e0, e1));
  const uint64_t m = f2_mul(ax, f2_pack(0.5f, 0.5f));
  const uint64_t base = f2_fma(x2, f2_pack(0.5f, 0.5f), m);
  return f2_fma(f2_mul(ax, f2_pack(-0.5f, -0.5f)), pe, base);
}

__device__ __forceinline__ void store_half32(__half* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    d[i] = u;
  }
}
// Split-precision planes of two values: hi = fp16(x), lo = fp16(x - hi)  (gemm.cuh: "precise" mode)
__device__ __forceinline__ void split_half2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// hi plane -> dst, lo plane -> dst_lo (nullable: plain fp16 store)
__device__ __forceinline__ void store_half32_split(__half* dst, __half* dst_lo, const float (&v)[32]) {
  if (!dst_lo) {
    store_half32(dst, v);
    return;
  }
  uint4* d = reinterpret_cast<uint4*>(dst);
  uint4* dl = reinterpret_cast<uint4*>(dst_lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u, l;
    split_half2(v[8 * i + 0], v[8 * i + 1], u.x, l.x);
    split_half2(v[8 * i + 2], v[8 * i + 3], u.y, l.y);
    split_half2(v[8 * i + 4], v[8 * i + 5], u.z, l.z);
    split_half2(v[8 * i + 6], v[8 * i + 7], u.w, l.w);
    d[i] = u;
    dl[i] = l;
  }
}
__device__ __forceinline__ void store_float32(float* dst, const float (&v)[32]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// Accumulator source for one thread's row: chunk c = columns [32c, 32c+32) of the 256-wide tile.
struct TmemLoader {
  uint32_t taddr;  // lane quarter + accumulator stage column base
  __device__ __forceinline__ void load(int chunk, float (&v)[32]) const {
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + chunk * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  }
};
struct ReferenceLoader {  // SIMT dot products (debug / self test)
  const __half* a_row;    // nullptr for rows >= M
  const __half* w_tile;   // W + n_tile*256*K
  int K;
  const __half* a_lo_row = nullptr;   // split-precision mode: low planes (nullptr otherwise)
  const __half* w_lo_tile = nullptr;
  __device__ __forceinline__ void load(int chunk, float (&v)[32]) const {
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      float acc = 0.f;
      if (a_row) {
        const __half* w = w_tile + static_cast<size_t>(chunk * 32 + j) * K;
        if (a_lo_row && w_lo_tile) {   // same three products as the tensor-core path, k by k
          const __half* wl = w_lo_tile + static_cast<size_t>(chunk * 32 + j) * K;
          for (int k = 0; k < K; ++k) {
            const float ah = __half2float(a_row[k]), al = __half2float(a_lo_row[k]);
            const float wh = __half2float(w[k]), wlo = __half2float(wl[k]);
            acc = fmaf(ah, wh, acc);
            acc = fmaf(al, wh, acc);
            acc = fmaf(ah, wlo, acc);
          }
        } else {
          for (int k = 0; k < K; ++k) acc = fmaf(__half2float(a_row[k]), __half2float(w[k]), acc);
        }
      }
      v[j] = acc;
    }
  }
};

// Direct-store epilogues: N may be padded to a multiple of 256 (weights zero-padded) while the output is narrower --
// chunks of 32 columns at or beyond n_valid (a multiple of 32; 0 = no limit) are not stored.  (The TMA-store epilogues
// need no check: boxes outside the tensor map are clipped by the hardware.)
__device__ __forceinline__ bool col_in(const GemmEpiParams& p, int col) { return p.n_valid == 0 || col < p.n_valid; }

// The fused tail for one thread == one output row of one 128x256 tile.  All 32 lanes of a warp call this together.
template <int EPI, typename Loader>
__device__ __forceinline__ void epilogue_row(const GemmEpiParams& p, int row, int n_tile, const Loader& ld,
                                             int c_begin = 0, int c_end = BN / 32) {
  const bool valid = row < p.M;
  const int col0 = n_tile * BN;
  float v[32];
  if constexpr (EPI == EPI_F16 || EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || kNormBias<EPI>) {
    const float rs = kNormBias<EPI> ? row_rstd(p, row) : 1.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      if constexpr (kNormBias<EPI>) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= rs;
      }
      if constexpr (EPI != EPI_F16) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 b = __ldg(b4 + i);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      }
      if constexpr (EPI == EPI_BIAS_GELU_F16 || EPI == EPI_NORM_BIAS_GELU_F16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      }
      if (valid && col_in(p, col0 + c * 32)) {
        const size_t o = static_cast<size_t>(row) * p.ld16 + col0 + c * 32;
        store_half32_split(p.out16 + o, p.out16_lo ? p.out16_lo + o : nullptr, v);
      }
    }
  } else if constexpr (EPI == EPI_F32 || EPI == EPI_GELU_F32 || EPI == EPI_BIAS_GELU_F32) {
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      if constexpr (EPI == EPI_BIAS_GELU_F32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col0 + c * 32 + i);
      }
      if constexpr (EPI != EPI_F32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      }
      if (valid && col_in(p, col0 + c * 32)) store_float32(p.out32 + static_cast<size_t>(row) * p.ld32 + col0 + c * 32, v);
    }
  } else if constexpr (EPI == EPI_HEAD_PARTIAL) {
    float s1 = 0.f, s2 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      const float4* w0 = reinterpret_cast<const float4*>(p.cls_gw + col0 + c * 32);
      const float4* w1 = reinterpret_cast<const float4*>(p.cls_gw + p.hidden + col0 + c * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(w0 + i), b = __ldg(w1 + i);
        const float wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g = gelu_erf(v[4 * i + e]);
          s1 += g;
          s2 = fmaf(g, g, s2);
          d0 = fmaf(g, wa[e], d0);
          d1 = fmaf(g, wb[e], d1);
        }
      }
      if ((c & 3) == 3) {   // 128 columns done: one slot
        if (valid)
          reinterpret_cast<float4*>(p.head_part)[static_cast<size_t>(2 * n_tile + (c >> 2)) * p.M + row] =
              make_float4(s1, s2, d0, d1);
        s1 = s2 = d0 = d1 = 0.f;
      }
    }
  } else if constexpr (EPI == EPI_SCORES_THRESH) {
    const float tau = valid ? __ldg(p.thr + row) : INFINITY;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      const int n0 = col0 + c * 32;
      if (n0 >= p.n_valid) continue;   // warp-uniform
      uint32_t hits = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) hits |= (v[i] >= tau ? 1u : 0u) << i;
      if (hits == 0) continue;          // the common case: nothing in these 32 columns reaches the threshold
      const int lim = p.n_valid - n0;   // columns of the ragged last chunk
      while (hits) {
        const int i = __ffs(hits) - 1;
        hits &= hits - 1;
        if (i >= lim || __ldg(p.col_mask + n0 + i)) continue;
        float sc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sc = j == i ? v[j] : sc;   // v[] stays in registers (no dynamic indexing)
        const int pos = atomicAdd(p.cand_count + row, 1);
        if (pos < p.cand_cap) {
          const uint32_t b = __float_as_uint(sc);
          const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
          p.cand[static_cast<size_t>(row) * p.cand_cap + pos] =
              (static_cast<unsigned long long>(ord) << 32) | static_cast<unsigned long long>(~static_cast<uint32_t>(p.col_base + n0 + i));
        }
      }
    }
  } else if constexpr (EPI == EPI_SCORES) {
    // out32 row = one query, columns = corpus rows: 128 contiguous bytes per thread and chunk (ld32 % 4 == 0)
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      const int n0 = col0 + c * 32;
      if (n0 >= p.n_valid) continue;   // warp-uniform
      if (n0 + 32 <= p.n_valid) {
        const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(p.col_mask + n0));
        const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(p.col_mask + n0) + 1);
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if ((mw[i >> 2] >> (8 * (i & 3))) & 0xffu) v[i] = -INFINITY;
        if (valid) store_float32(p.out32 + static_cast<size_t>(row) * p.ld32 + n0, v);
      } else if (valid) {
        for (int i = 0; i < 32 && n0 + i < p.n_valid; ++i)
          p.out32[static_cast<size_t>(row) * p.ld32 + n0 + i] = __ldg(p.col_mask + n0 + i) ? -INFINITY : v[i];
      }
    }
  } else if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_BIAS_RESID_F32) {
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      if (valid && col_in(p, col0 + c * 32)) {
        float4* x4 = reinterpret_cast<float4*>(p.out32 + static_cast<size_t>(row) * p.ld32 + col0 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = x4[i];
          x.x += v[4 * i]; x.y += v[4 * i + 1]; x.z += v[4 * i + 2]; x.w += v[4 * i + 3];
          if constexpr (EPI == EPI_BIAS_RESID_F32) {
            float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c * 32) + i);
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          x4[i] = x;
        }
      }
    }
  } else if constexpr (kResidStats<EPI>) {
    float s = 0.f, q = 0.f;
    const float2 mr = EPI == EPI_RESID_STATS_LN ? row_mean_rstd(p, row) : make_float2(0.f, 1.f);
    const float ra = mr.y, rb = -mr.x * mr.y;   // (x - mean) * rstd = x * ra + rb
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      if (valid) {
        __half* hp = p.out16 + static_cast<size_t>(row) * p.ld16 + col0 + c * 32;
        uint8_t* lp = p.out8_lo + static_cast<size_t>(row) * p.ld16 + col0 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float xo = __half2float(hp[i]) + unpack_e5m2x2<0>(lp[i]).x;
          if constexpr (EPI == EPI_RESID_STATS_LN)
            xo = fmaf(fmaf(xo, ra, rb), __ldg(p.gamma + col0 + c * 32 + i), __ldg(p.bias + col0 + c * 32 + i));
          const float x = v[i] + xo;
          s += x;
          q = fmaf(x, x, q);
          const __half h = __float2half_rn(x);
          hp[i] = h;
          lp[i] = static_cast<uint8_t>(pack_e5m2x2(x - __half2float(h), 0.f));
        }
        if ((c & 3) == 3) {
          reinterpret_cast<float2*>(p.stats_out)[static_cast<size_t>(2 * n_tile + (c >> 2)) * p.M + row] =
              make_float2(s, q);
          s = 0.f;
          q = 0.f;
        }
      }
    }
  } else if constexpr (kRope<EPI>) {
    // 4 heads of 64 per tile.  x1 = dims [0,32), x2 = dims [32,64): out1 = x1*cos - x2*sin, out2 = x2*cos + x1*sin
    // (rotate_half convention, fp32, modeling_modernbert.py:197-228).
    float w2[32];
    const int pos = valid ? __ldg(p.pos + row) : 0;
    const float4* cs4 = reinterpret_cast<const float4*>(p.rope_tab + static_cast<size_t>(pos) * 64);
    const float4* sn4 = cs4 + 8;
    const float rs = kNorm<EPI> ? row_rstd(p, row) : 1.f;
#pragma unroll 1
    for (int h = 0; h < BN / 64; ++h) {
      ld.load(2 * h, v);
      ld.load(2 * h + 1, w2);
      if constexpr (kNorm<EPI>) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] *= rs; w2[i] *= rs; }
      }
      const int gcol = col0 + h * 64;
      if (gcol < 2 * p.hidden) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 c = __ldg(cs4 + i), s = __ldg(sn4 + i);
          float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x1 = v[4 * i + e], x2 = w2[4 * i + e];
            v[4 * i + e] = x1 * cc[e] - x2 * ss[e];
            w2[4 * i + e] = x2 * cc[e] + x1 * ss[e];
          }
        }
      }
      if (valid) {
        const size_t o = static_cast<size_t>(row) * p.ld16 + gcol;
        store_half32_split(p.out16 + o, p.out16_lo ? p.out16_lo + o : nullptr, v);
        store_half32_split(p.out16 + o + 32, p.out16_lo ? p.out16_lo + o + 32 : nullptr, w2);
      }
    }
  } else if constexpr (kGeglu<EPI>) {
    float g[32];
    const float rs = kNorm<EPI> ? row_rstd(p, row) : 1.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      ld.load(c, v);
      ld.load(4 + c, g);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = kNorm<EPI> ? gelu_erf(v[i] * rs) * (g[i] * rs) : gelu_erf(v[i]) * g[i];
      if (valid) {
        const size_t o = static_cast<size_t>(row) * p.ld16 + n_tile * 128 + c * 32;
        store_half32_split(p.out16 + o, p.out16_lo ? p.out16_lo + o : nullptr, v);
      }
    }
  } else if constexpr (EPI == EPI_SPLADE) {
    // max_rows log1p(relu(x + b)) = log1p(relu(max_rows(x) + b)): the bias is per column and log1p(relu(.)) is monotone,
    // so the warp first max-reduces the RAW accumulators over its 32 rows (one order-preserving integer REDUX per
    // column, lane c keeps column c) and only then applies bias / relu / log1p -- to one value per lane instead of 32
    // (the first version evaluated log1pf on every element: ~40 issue slots per element, the decoder GEMM ran at 22 %
    // tensor-pipe activity).  The result is >= 0, so its float bits order like ints: atomicMax on int finishes the pool
    // across warps.  Rows of different sequences in one warp (ragged batches) are reduced one sequence at a time.
    const int seq = valid ? __ldg(p.seq_of_row + row) : -1;
    const int lane = threadIdx.x & 31;
    const unsigned live = __ballot_sync(0xffffffffu, valid);
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      ld.load(c, v);
      const int cbase = col0 + c * 32;
      if (cbase >= p.n_valid) continue;  // warp-uniform
      int key[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {   // monotone float -> signed int
        const int bts = __float_as_int(v[i]);
        key[i] = bts ^ ((bts >> 31) & 0x7fffffff);
      }
      const bool col_ok = cbase + lane < p.n_valid;
      const float bias = col_ok ? __ldg(p.bias + cbase + lane) : 0.f;
      unsigned todo = live;
      while (todo) {   // warp-uniform: one pass per sequence present in this warp's rows (one for aligned batches)
        const int s0 = __shfl_sync(0xffffffffu, seq, __ffs(todo) - 1);
        const unsigned grp = __ballot_sync(0xffffffffu, seq == s0) & live;
        const bool in = (grp >> lane) & 1u;
        int mine = INT_MIN;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int m = __reduce_max_sync(0xffffffffu, in ? key[i] : INT_MIN);
          if (lane == i) mine = m;
        }
        const float x = __int_as_float(mine ^ ((mine >> 31) & 0x7fffffff)) + bias;
        const float a = x > 0.f ? log1pf(x) : 0.f;
        if (col_ok && a > 0.f)
          atomicMax(reinterpret_cast<int*>(p.splade_out + static_cast<size_t>(s0) * p.splade_ld + cbase + lane),
                    __float_as_int(a));
        todo &= ~grp;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Staged epilogue (tcgen05 path): thread == row writes its 128-byte piece of a [32 rows x 128 B] box into smem with
// the 128B swizzle (16-byte chunk c of row r at c ^ (r & 7): conflict-free per quarter warp and exactly the layout a
// SWIZZLE_128B tensor map expects); one lane then issues a TMA store / reduce-add of the box.  Boxes are
// double-buffered per warp; cp.async.bulk.wait_group.read gates buffer reuse.
// ---------------------------------------------------------------------------------------------------------------
template <int EPI>
constexpr bool kStaged = (EPI == EPI_F16 || EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || kNormBias<EPI> || kRope<EPI> ||
                          kGeglu<EPI> || EPI == EPI_RESID_F32 || EPI == EPI_BIAS_RESID_F32);

struct BoxStager {
  uint8_t* base;      // this warp's two boxes
  uint32_t issued;    // boxes submitted so far (same value in every lane)
  __device__ __forceinline__ uint8_t* acquire(int lane) {
    if (issued >= 2) {                                  // the box submitted two steps ago has been read out
      if (elect_one()) bulk_wait_read<1>();             // (bulk groups belong to the issuing = elected lane)
    }
    __syncwarp();
    return base + (issued & 1) * EPI_BOX_BYTES;
  }
  template <bool REDUCE>
  __device__ __forceinline__ void submit(const CUtensorMap* tm, uint8_t* box, int x, int y, int lane) {
    fence_proxy_async_smem();  // generic-proxy st.shared -> visible to the async (TMA) proxy
    __syncwarp();
    if (elect_one()) {
      if (REDUCE) tma_reduce_add_2d(tm, box, x, y);
      else tma_store_2d(tm, box, x, y);
      bulk_commit();
    }
    ++issued;
  }
  // Split-precision outputs: the warp's two boxes are filled together (box 0 = hi plane, box 1 = lo plane) and stored
  // as one bulk group, so both must have been read out before the next fill.
  __device__ __forceinline__ uint8_t* acquire_both(int lane) {
    if (issued > 0) {
      if (elect_one()) bulk_wait_read<0>();
    }
    __syncwarp();
    return base;
  }
  __device__ __forceinline__ void submit_both(const CUtensorMap* tm_hi, const CUtensorMap* tm_lo, int x, int y) {
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_2d(tm_hi, base, x, y);
      tma_store_2d(tm_lo, base + EPI_BOX_BYTES, x, y);
      bulk_commit();
    }
    issued += 2;
  }
};

// write 32 fp32 values as 16 halves-pairs = 64 bytes = chunks [4*half_idx, 4*half_idx+4) of this thread's box row
__device__ __forceinline__ void box_put_half32(uint8_t* box, int r, int half_idx, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(box + r * 128 + (((half_idx * 4 + i) ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void box_put_float32(uint8_t* box, int r, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(box + r * 128 + ((i ^ (r & 7)) << 4)) =
        make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// the same 64 bytes of this thread's row in the hi box and in the lo box (split-precision planes)
__device__ __forceinline__ void box_put_half32_split(uint8_t* box_hi, uint8_t* box_lo, int r, int half_idx,
                                                     const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u, l;
    split_half2(v[8 * i + 0], v[8 * i + 1], u.x, l.x);
    split_half2(v[8 * i + 2], v[8 * i + 3], u.y, l.y);
    split_half2(v[8 * i + 4], v[8 * i + 5], u.z, l.z);
    split_half2(v[8 * i + 6], v[8 * i + 7], u.w, l.w);
    const int off = r * 128 + (((half_idx * 4 + i) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(box_hi + off) = u;
    *reinterpret_cast<uint4*>(box_lo + off) = l;
  }
}

template <int EPI, bool SPLIT>
__device__ __forceinline__ void staged_epilogue(const GemmEpiParams& p, const CUtensorMap* tmOut,
                                                const CUtensorMap* tmOutLo, BoxStager& st,
                                                int m0, int n_tile, const TmemLoader& ld, int lane, int half,
                                                float rs /* kNormBias: rstd of this thread's row, fetched by the caller
                                                            before it waits for the accumulators */) {
  const int col0 = n_tile * BN;
  const int r = lane;  // row of this thread inside the warp's 32-row slab
  float v[32];
  if constexpr (EPI == EPI_F16 || EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || kNormBias<EPI>) {
#pragma unroll 1
    for (int b = 2 * half; b < 2 * half + 2; ++b) {
      uint8_t* box = SPLIT ? st.acquire_both(lane) : st.acquire(lane);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ld.load(2 * b + h, v);
        if constexpr (kNormBias<EPI>) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= rs;
        }
        if constexpr (EPI != EPI_F16) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + (2 * b + h) * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
          }
        }
        if constexpr (EPI == EPI_BIAS_GELU_F16 || EPI == EPI_NORM_BIAS_GELU_F16) {   // packed fp32x2 GeLU (gelu2: |error| <= 3.3e-7), 10 vs 35 issue slots
#pragma unroll
          for (int i = 0; i < 16; ++i) f2_unpack(gelu2(f2_pack(v[2 * i], v[2 * i + 1])), v[2 * i], v[2 * i + 1]);
        }
        if constexpr (SPLIT) box_put_half32_split(box, box + EPI_BOX_BYTES, r, h, v);
        else box_put_half32(box, r, h, v);
      }
      if constexpr (SPLIT) st.submit_both(tmOut, tmOutLo, col0 + b * 64, m0);
      else st.submit<false>(tmOut, box, col0 + b * 64, m0, lane);
    }
  } else if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_BIAS_RESID_F32) {
#pragma unroll 1
    for (int c = 4 * half; c < 4 * half + 4; ++c) {
      uint8_t* box = st.acquire(lane);
      ld.load(c, v);
      if constexpr (EPI == EPI_BIAS_RESID_F32) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(b4 + i);
          v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
        }
      }
      box_put_float32(box, r, v);
      st.submit<true>(tmOut, box, col0 + c * 32, m0, lane);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pipelined epilogues of the two big ModernBERT GEMMs (Wqkv + RoPE, Wi + GeGLU).  Measured with vrag_bench_gemm: the
// mainloop alone runs at ~1500 TFLOP/s, so the kernel time is the epilogue's whenever a tile's tail takes longer than
// its 256x256xK MMAs (7.7 k clocks at K = 768), and the tail is a chain of long-latency hops (tcgen05.ld, global loads
// of cos / sin / row moments, the TMA-store hand-off) that two warps per scheduler cannot hide.  So:
//   * everything that does not need the accumulators (row rstd, the 32 tokens' cos / sin rows) is fetched BEFORE the
//     warp waits for the accumulator barrier -- the cos / sin rows by ONE TMA box load from the interleaved
//     [position][cos 32 | sin 32] table when the slab's 32 tokens have consecutive positions (always, except slabs
//     that straddle a sequence boundary: those gather, all loads in flight at once);
//   * TMEM is drained in 16-column steps into ping-pong registers: the tcgen05.ld of step s+1 is in flight while step
//     s is rotated / activated, packed and staged;
//   * the accumulator stage is released right after the last tcgen05.ld has landed, before the tail of the math.
// ---------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void reg_fence(uint32_t (&r)[N]) {  // pins uses of r[] after the preceding wait::ld
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}

__device__ __forceinline__ void box_put_half16(uint8_t* box, int r, int chunk0, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(box + r * 128 + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
  }
}

__device__ __forceinline__ void box_put_half16_split(uint8_t* box_hi, uint8_t* box_lo, int r, int chunk0,
                                                     const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u, l;
    split_half2(v[8 * i + 0], v[8 * i + 1], u.x, l.x);
    split_half2(v[8 * i + 2], v[8 * i + 3], u.y, l.y);
    split_half2(v[8 * i + 4], v[8 * i + 5], u.z, l.z);
    split_half2(v[8 * i + 6], v[8 * i + 7], u.w, l.w);
    const int off = r * 128 + (((chunk0 + i) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(box_hi + off) = u;
    *reinterpret_cast<uint4*>(box_lo + off) = l;
  }
}

struct AccRelease {   // hands the TMEM accumulator stage back to the MMA issuer (leader CTA's barrier)
  uint64_t* bar;
  int cta_rank;
  __device__ __forceinline__ void operator()(int lane) const {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (cta_rank == 0) mbar_arrive(bar);
      else mbar_arrive_remote(bar, 0);
    }
  }
};

// cos / sin rows of this warp's 32 tokens -> registers of thread == row.  Borrows the warp's two store boxes.
__device__ __forceinline__ void rope_prefetch(const GemmEpiParams& p, const CUtensorMap* tmRope, BoxStager& st,
                                              uint64_t* bar, uint32_t& bar_phase, int m0, int lane, float (&cs)[32],
                                              float (&sn)[32]) {
  const int r = lane, row = m0 + lane;
  const int my_pos = row < p.M ? __ldg(p.pos + row) : 0;
  const int pos0 = __shfl_sync(0xffffffffu, my_pos, 0);
  const bool consecutive = __all_sync(0xffffffffu, row >= p.M || my_pos == pos0 + lane);
  float* cs_s = reinterpret_cast<float*>(st.base);
  float* sn_s = reinterpret_cast<float*>(st.base + EPI_BOX_BYTES);
  if (elect_one()) bulk_wait_read<0>();   // no TMA store may still be reading the boxes
  __syncwarp();
  if (consecutive) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, 2 * EPI_BOX_BYTES);
      tma_load_2d(cs_s, tmRope, bar, 0, pos0);
      tma_load_2d(sn_s, tmRope, bar, 32, pos0);
    }
    __syncwarp();
    mbar_wait_tagged(bar, bar_phase, 16);
    bar_phase ^= 1;
  } else {
    float c[32], s2[32];
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      const int ps = __shfl_sync(0xffffffffu, my_pos, rr);
      c[rr] = __ldg(p.rope_tab + static_cast<size_t>(ps) * 64 + lane);
      s2[rr] = __ldg(p.rope_tab + static_cast<size_t>(ps) * 64 + 32 + lane);
    }
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      const int o = rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3));
      cs_s[o] = c[rr];
      sn_s[o] = s2[rr];
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 c = *reinterpret_cast<const float4*>(cs_s + r * 32 + ((i ^ (r & 7)) << 2));
    const float4 s4 = *reinterpret_cast<const float4*>(sn_s + r * 32 + ((i ^ (r & 7)) << 2));
    cs[4 * i] = c.x; cs[4 * i + 1] = c.y; cs[4 * i + 2] = c.z; cs[4 * i + 3] = c.w;
    sn[4 * i] = s4.x; sn[4 * i + 1] = s4.y; sn[4 * i + 2] = s4.z; sn[4 * i + 3] = s4.w;
  }
  __syncwarp();
  st.issued = 0;   // both boxes are free again; restart the double-buffer bookkeeping
}

// Wqkv tail.  Step s = 0..3: head 2*half + (s >> 1), dims [16j, 16j+16) and [32+16j, 32+16j+16), j = s & 1.
template <int EPI, bool SPLIT>
__device__ __forceinline__ void rope_body(const GemmEpiParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmOutLo,
                                          BoxStager& st, int m0,
                                          int n_tile, uint32_t taddr, int lane, int half, float rs,
                                          const float (&cs)[32], const float (&sn)[32], const AccRelease& release) {
  const int col0 = n_tile * BN;
  const bool rotate = col0 < 2 * p.hidden;  // q / k tiles
  const int r = lane;
  uint32_t a[2][16], b[2][16];
  tmem_ld_32x32b_x16(taddr + (2 * half) * 64, a[0]);
  tmem_ld_32x32b_x16(taddr + (2 * half) * 64 + 32, b[0]);
  uint8_t* box = nullptr;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int h = 2 * half + (s >> 1), j = s & 1;
    tmem_ld_wait();
    reg_fence(a[s & 1]);
    reg_fence(b[s & 1]);
    if (s + 1 < 4) {
      const int h1 = 2 * half + ((s + 1) >> 1), j1 = (s + 1) & 1;
      tmem_ld_32x32b_x16(taddr + h1 * 64 + 16 * j1, a[(s + 1) & 1]);
      tmem_ld_32x32b_x16(taddr + h1 * 64 + 32 + 16 * j1, b[(s + 1) & 1]);
    } else {
      release(lane);   // every accumulator column of this warp is in registers
    }
    if (j == 0) box = SPLIT ? st.acquire_both(lane) : st.acquire(lane);
    float o1[16], o2[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float x1 = __uint_as_float(a[s & 1][e]), x2 = __uint_as_float(b[s & 1][e]);
      if constexpr (kNorm<EPI>) {
        x1 *= rs;
        x2 *= rs;
      }
      const float c = cs[16 * j + e], sv = sn[16 * j + e];
      o1[e] = rotate ? x1 * c - x2 * sv : x1;
      o2[e] = rotate ? x2 * c + x1 * sv : x2;
    }
    if constexpr (SPLIT) {
      box_put_half16_split(box, box + EPI_BOX_BYTES, r, 2 * j, o1);
      box_put_half16_split(box, box + EPI_BOX_BYTES, r, 4 + 2 * j, o2);
      if (j == 1) st.submit_both(tmOut, tmOutLo, col0 + h * 64, m0);
    } else {
      box_put_half16(box, r, 2 * j, o1);
      box_put_half16(box, r, 4 + 2 * j, o2);
      if (j == 1 && p.debug_mode != 4) st.submit<false>(tmOut, box, col0 + h * 64, m0, lane);
    }
  }
}

// Wi tail.  This warp owns output columns [64*half, 64*half+64) of the tile's 128: step s = 0..3 takes input columns
// 64*half + 16s .. +16 and the matching gate columns 128 + 64*half + 16s.
template <int EPI, bool SPLIT>
__device__ __forceinline__ void geglu_body(const GemmEpiParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmOutLo,
                                           BoxStager& st, int m0,
                                           int n_tile, uint32_t taddr, int lane, int half, float rs,
                                           const AccRelease& release) {
  const int r = lane;
  uint32_t a[2][16], g[2][16];
  tmem_ld_32x32b_x16(taddr + 64 * half, a[0]);
  tmem_ld_32x32b_x16(taddr + 128 + 64 * half, g[0]);
  uint8_t* box = SPLIT ? st.acquire_both(lane) : st.acquire(lane);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    tmem_ld_wait();
    reg_fence(a[s & 1]);
    reg_fence(g[s & 1]);
    if (s + 1 < 4) {
      tmem_ld_32x32b_x16(taddr + 64 * half + 16 * (s + 1), a[(s + 1) & 1]);
      tmem_ld_32x32b_x16(taddr + 128 + 64 * half + 16 * (s + 1), g[(s + 1) & 1]);
    } else {
      release(lane);
    }
    float o[16];
    const uint64_t rs2 = f2_pack(rs, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      uint64_t x2 = f2_pack_bits(a[s & 1][2 * e], a[s & 1][2 * e + 1]);
      uint64_t y2 = f2_pack_bits(g[s & 1][2 * e], g[s & 1][2 * e + 1]);
      if constexpr (kNorm<EPI>) {
        x2 = f2_mul(x2, rs2);
        y2 = f2_mul(y2, rs2);
      }
      f2_unpack(f2_mul(gelu2(x2), y2), o[2 * e], o[2 * e + 1]);
    }
    if constexpr (SPLIT) box_put_half16_split(box, box + EPI_BOX_BYTES, r, 2 * s, o);
    else box_put_half16(box, r, 2 * s, o);   // (debug_mode 4 keeps the math and the smem staging, drops the TMA store)
  }
  if constexpr (SPLIT) st.submit_both(tmOut, tmOutLo, n_tile * 128 + half * 64, m0);
  else if (p.debug_mode != 4) st.submit<false>(tmOut, box, n_tile * 128 + half * 64, m0, lane);
}

// Launched as clusters of 2 CTAs (an SM pair) that cooperate on one 256 x 256 output tile with
// tcgen05.mma.cta_group::2: each CTA holds its own 128 rows of A and HALF of the W tile (128 of the 256 N rows) and
// its own 128 x 256 fp32 accumulator in TMEM.  Per SM and k-block this moves 32 KB into smem and the tensor core reads
// 32 KB back, instead of 48 + 48 KB for a 1-CTA 128 x 256 tile -- measured, the 1-CTA mainloop saturates the
// 128 B/clk shared-memory port at ~66 % tensor-pipe activity (profiles/), so operand bytes per flop are the lever.
// The leader CTA (cluster rank 0) issues the MMAs; both CTAs' TMA loads signal the leader's `full` barrier; the MMA
// commit is multicast to both CTAs' `empty` / `tmem_full` barriers; both epilogues release the accumulator by
// arriving on the leader's `tmem_empty` barrier.
//
// SPLIT (split-precision / "precise" mode, gemm.cuh): the operand ring holds units of 32 KB that alternate between the
// hi planes (A_hi, W_hi half) and the lo planes (A_lo, W_lo half) of one k-block; the leader issues
// A_hi W_hi as soon as the hi unit has landed, then A_lo W_hi + A_hi W_lo once the lo unit has, and releases both
// units together.  Same shared-memory footprint as the 5-stage fp16 ring (2.5 k-blocks in flight).
template <int EPI, int STAGES, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                    const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                    const __grid_constant__ CUtensorMap tmOutLo,
                    int m_tiles, int n_tiles, int k_blocks, GemmEpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  constexpr int EPI_WB = epi_warp_bytes(EPI);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(epi_smem + EPI_WARPS * EPI_WB);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_x = bar_tempty + 2;                 // [EPI_WARPS][3]: residual boxes (EPI_RESID_STATS), cos/sin box
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_x + STATS_BARS * EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (uniform datapath)
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());      // 0 / 1 inside the CTA pair
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_pairs = ((m_tiles + 1) >> 1) * n_tiles;        // (two M tiles) x (one N tile) per cluster step

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);   // the leader's MMA commit (multicast to both CTAs)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + a, 1);
      mbar_init(bar_tempty + a, 2 * EPI_WARPS);   // leader's barrier: epilogue warps of BOTH CTAs arrive
    }
    for (int a = 0; a < STATS_BARS * EPI_WARPS; ++a) mbar_init(bar_x + a, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (SPLIT) {
      tma_prefetch_desc(&tmAlo);
      tma_prefetch_desc(&tmBlo);
    }
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_holder, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into its smem
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // NOTE: the producer and MMA loops are executed by the WHOLE warp (convergent, warp-uniform values) and only the
  // TMA / MMA / commit instructions themselves are predicated on one elected lane.  Running the loops under
  // `if (lane == 0)` makes every descriptor a divergent value and the compiler wraps each UTCHMMA / UTMALDG in an
  // ELECT + R2UR waterfall (~30 extra instructions per MMA, measured: the issue thread became the bottleneck).
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
      const int m_idx = 2 * (pair / n_tiles) + cta_rank, n_idx = pair % n_tiles;
      for (int kb = 0; kb < k_blocks; ++kb) {
#pragma unroll
        for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {   // SPLIT: hi unit, then lo unit of this k-block
          mbar_wait_tagged(bar_empty + stage, phase ^ 1, 11);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (elect_one()) {
            // one barrier in the leader tracks both CTAs' operands: 2 x (A 16 KB + W half 16 KB)
            if (cta_rank == 0) mbar_arrive_expect_tx(bar_full + stage, 2 * STAGE_BYTES);
            tma_load_2d_pair(sa, part ? &tmAlo : &tmA, bar_full + stage, kb * BK, m_idx * BM);
            tma_load_2d_pair(sa + A_BYTES, part ? &tmBlo : &tmB, bar_full + stage, kb * BK,
                             n_idx * BN + cta_rank * (BN / 2));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {   // the leader CTA issues the pair's MMAs
      constexpr uint32_t idesc = umma_idesc(0 /*f16*/, 2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
        mbar_wait_tagged(bar_tempty + acc, acc_phase ^ 1, 12);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait_tagged(bar_full + stage, phase, 13);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
          if constexpr (!SPLIT) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_f16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                              (kb | k) != 0 ? 1u : 0u);
              }
              // slot reusable (in BOTH CTAs) once these MMAs retire
              umma_commit_pair(bar_empty + stage, static_cast<uint16_t>(3));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          } else {
            if (elect_one()) {   // A_hi W_hi
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_f16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                              (kb | k) != 0 ? 1u : 0u);
              }
            }
            __syncwarp();
            const int hi_stage = stage;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            mbar_wait_tagged(bar_full + stage, phase, 17);
            tc_fence_after();
            const uint32_t al_addr = smem_u32(smem + stage * STAGE_BYTES);
            const uint32_t bl_addr = al_addr + A_BYTES;
            if (elect_one()) {   // A_lo W_hi + A_hi W_lo
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_f16_pair(d_tmem, umma_desc_sw128(al_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, 1u);
                umma_f16_pair(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(bl_addr + k * 32), idesc, 1u);
              }
              umma_commit_pair(bar_empty + hi_stage, static_cast<uint16_t>(3));   // both units reusable once these retire
              umma_commit_pair(bar_empty + stage, static_cast<uint16_t>(3));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) umma_commit_pair(bar_tfull + acc, static_cast<uint16_t>(3));   // accumulators -> epilogues
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which half of the tile's 256 columns this warp drains
    uint8_t* my_smem = epi_smem + (warp - 2) * EPI_WB;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (kResidStats<EPI>) {
      // x = x_old + acc on the two-plane residual stream: the warp's steps (32 rows x 64 columns, two per tile) form
      // one flat stream across tiles; step n lives in buffer n & 1 = {hi box (fp16, SWIZZLE_128B), lo box (e5m2,
      // SWIZZLE_64B)} (TMA-loaded, updated in place, TMA-stored).  The loads of step n + 1 are issued while step n is
      // processed, as soon as the stores of step n - 1 (same buffer) have been read out of shared memory.  Measured
      // (vrag_bench_gemm) with an fp16 low plane: deeper prefetch (three boxes in flight), an L2 prefetch two tiles
      // ahead and block-major planes all left the time unchanged -- the kernel ran at ~4.5 TB/s of HBM traffic with
      // its shared-memory port shared between the operand ring and the four passes (TMA in, LDS, STS, TMA out) over
      // the residual boxes; the e5m2 low plane removes a fifth of those bytes and frees a fourth operand stage.
      uint64_t* xb = bar_x + STATS_BARS * (warp - 2);
      const int my_tiles = cluster_id < total_pairs ? (total_pairs - cluster_id + num_clusters - 1) / num_clusters : 0;
      const uint32_t n_steps = 2u * static_cast<uint32_t>(my_tiles);
      if (lane == 0) {
        tma_prefetch_desc(&tmOut);
        tma_prefetch_desc(&tmOut2);
      }
      auto step_xy = [&](uint32_t n, int& x, int& y) {
        const int pair = cluster_id + static_cast<int>(n >> 1) * num_clusters;
        x = (pair % n_tiles) * BN + (2 * half + static_cast<int>(n & 1)) * 64;
        y = (2 * (pair / n_tiles) + cta_rank) * BM + quarter * 32;
      };
      auto issue_load = [&](uint32_t n) {   // elected lane
        int x, y;
        step_xy(n, x, y);
        mbar_arrive_expect_tx(xb + (n & 1), EPI_BOX_BYTES + EPI_LO_BOX_BYTES);
        tma_load_2d(my_smem + (n & 1) * EPI_BOX_BYTES, &tmOut, xb + (n & 1), x, y);
        tma_load_2d(my_smem + 2 * EPI_BOX_BYTES + (n & 1) * EPI_LO_BOX_BYTES, &tmOut2, xb + (n & 1), x, y);
      };
      if (elect_one()) {
        if (n_steps > 0) issue_load(0);
        if (n_steps > 1) issue_load(1);
      }
      __syncwarp();
      const int r = lane;
      uint32_t n = 0;
      for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
        const int m_idx = 2 * (pair / n_tiles) + cta_rank, n_idx = pair % n_tiles;
        float ra = 1.f, rb = 0.f;   // EPI_RESID_STATS_LN: (x_old - mean) * rstd = x_old * ra + rb for this thread's row
        if constexpr (EPI == EPI_RESID_STATS_LN) {   // fetched before the accumulator wait
          const float2 mr = row_mean_rstd(p, m_idx * BM + quarter * 32 + lane);
          ra = mr.y;
          rb = -mr.x * mr.y;
        }
        mbar_wait_tagged(bar_tfull + acc, acc_phase, 14);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int stp = 0; stp < 2; ++stp, ++n) {
          uint8_t* hbox = my_smem + (n & 1) * EPI_BOX_BYTES;
          uint8_t* lbox = my_smem + 2 * EPI_BOX_BYTES + (n & 1) * EPI_LO_BOX_BYTES;
          uint32_t ta[32], tb[32];
          tmem_ld_32x32b_x32(taddr + (2 * half + stp) * 64, ta);
          tmem_ld_32x32b_x32(taddr + (2 * half + stp) * 64 + 32, tb);
          if (n >= 1) {   // the other buffer was last stored from by step n - 1
            if (elect_one()) {
              bulk_wait_read<0>();
              if (n + 1 < n_steps) issue_load(n + 1);
            }
            __syncwarp();
          }
          tmem_ld_wait();
          if (stp == 1) {   // every accumulator column of this warp is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (cta_rank == 0) mbar_arrive(bar_tempty + acc);
              else mbar_arrive_remote(bar_tempty + acc, 0);
            }
          }
          mbar_wait_tagged(xb + (n & 1), (n >> 1) & 1, 15);
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // 16-byte chunk j of this thread's lo row = 16 columns = hi chunks 2j, 2j+1
            const int loff = r * 64 + ((j ^ ((r >> 1) & 3)) << 4);
            uint4 lv = *reinterpret_cast<const uint4*>(lbox + loff);
            uint32_t* lw = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int i = 2 * j + hh;   // hi chunk: columns [8i, 8i + 8)
              const int hoff = r * 128 + ((i ^ (r & 7)) << 4);
              uint4 hv = *reinterpret_cast<const uint4*>(hbox + hoff);
              uint32_t* hw = reinterpret_cast<uint32_t*>(&hv);
              float d[8];
              float gm[8], bt[8];   // EPI_RESID_STATS_LN: LayerNorm weight / (beta + dense bias) of these 8 columns
              if constexpr (EPI == EPI_RESID_STATS_LN) {
                const int gc = n_idx * BN + (2 * half + stp) * 64 + 8 * i;
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + gc));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + gc) + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + gc));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + gc) + 1);
                gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
                bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 ho = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                const uint32_t lword = lw[2 * hh + (e >> 1)];
                const float2 lo = (e & 1) ? unpack_e5m2x2<1>(lword) : unpack_e5m2x2<0>(lword);
                const int c = 8 * i + 2 * e;
                float xo0 = ho.x + lo.x, xo1 = ho.y + lo.y;
                if constexpr (EPI == EPI_RESID_STATS_LN) {
                  xo0 = fmaf(fmaf(xo0, ra, rb), gm[2 * e], bt[2 * e]);
                  xo1 = fmaf(fmaf(xo1, ra, rb), gm[2 * e + 1], bt[2 * e + 1]);
                }
                const float x0 = __uint_as_float(c < 32 ? ta[c & 31] : tb[c & 31]) + xo0;
                const float x1 = __uint_as_float(c < 32 ? ta[(c + 1) & 31] : tb[(c + 1) & 31]) + xo1;
                sum += x0;
                sq = fmaf(x0, x0, sq);
                sum += x1;
                sq = fmaf(x1, x1, sq);
                const __half2 nh = __floats2half2_rn(x0, x1);
                const float2 nf = __half22float2(nh);
                hw[e] = *reinterpret_cast<const uint32_t*>(&nh);
                d[2 * e] = x0 - nf.x;
                d[2 * e + 1] = x1 - nf.y;
              }
              lw[2 * hh] = pack_e5m2x4(d[0], d[1], d[2], d[3]);
              lw[2 * hh + 1] = pack_e5m2x4(d[4], d[5], d[6], d[7]);
              if (p.debug_mode != 4) *reinterpret_cast<uint4*>(hbox + hoff) = hv;
            }
            if (p.debug_mode != 4) *reinterpret_cast<uint4*>(lbox + loff) = lv;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            int x, y;
            step_xy(n, x, y);
            if (p.debug_mode != 4) {
              tma_store_2d(&tmOut, hbox, x, y);
              tma_store_2d(&tmOut2, lbox, x, y);
            }
            bulk_commit();
          }
          __syncwarp();
        }
        const int row = m_idx * BM + quarter * 32 + lane;
        if (row < p.M)
          reinterpret_cast<float2*>(p.stats_out)[static_cast<size_t>(2 * n_idx + half) * p.M + row] =
              make_float2(sum, sq);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (elect_one()) bulk_wait<0>();
      __syncwarp();
    } else {
      BoxStager stager{my_smem, 0u};
      if (kStaged<EPI> && lane == 0) tma_prefetch_desc(&tmOut);
      if (kRope<EPI> && lane == 0) tma_prefetch_desc(&tmOut2);
      if (SPLIT && kStaged<EPI> && lane == 0) tma_prefetch_desc(&tmOutLo);
      uint64_t* rope_bar = bar_x + STATS_BARS * (warp - 2);
      uint32_t rope_phase = 0;
      for (int pair = cluster_id; pair < total_pairs; pair += num_clusters) {
        const int m_idx = 2 * (pair / n_tiles) + cta_rank, n_idx = pair % n_tiles;
        const int m0 = m_idx * BM + quarter * 32;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
        const AccRelease release{bar_tempty + acc, cta_rank};
        if constexpr (kRope<EPI> || kGeglu<EPI>) {
          // phase 1 (no accumulators needed): row rstd, cos / sin rows
          float rs = 1.f;
          if constexpr (kNorm<EPI>) rs = row_rstd(p, m0 + lane);
          float cs[32], sn[32];
          if constexpr (kRope<EPI>) {
            if (n_idx * BN < 2 * p.hidden)
              rope_prefetch(p, &tmOut2, stager, rope_bar, rope_phase, m0, lane, cs, sn);
          }
          mbar_wait_tagged(bar_tfull + acc, acc_phase, 14);
          tc_fence_after();
          if (p.debug_mode == 3) release(lane);   // timing experiment: mainloop only
          else if constexpr (kRope<EPI>)
            rope_body<EPI, SPLIT>(p, &tmOut, &tmOutLo, stager, m0, n_idx, taddr, lane, half, rs, cs, sn, release);
          else geglu_body<EPI, SPLIT>(p, &tmOut, &tmOutLo, stager, m0, n_idx, taddr, lane, half, rs, release);
        } else {
          float rs = 1.f;
          if constexpr (kNormBias<EPI>) rs = row_rstd(p, m0 + lane);   // before the accumulator wait: latency hidden
          mbar_wait_tagged(bar_tfull + acc, acc_phase, 14);
          tc_fence_after();
          TmemLoader ld{taddr};
          if (p.debug_mode == 3) {
            // timing experiment: mainloop only
          } else if constexpr (kStaged<EPI>)
            staged_epilogue<EPI, SPLIT>(p, &tmOut, &tmOutLo, stager, m0, n_idx, ld, lane, half, rs);
          else
            epilogue_row<EPI>(p, m0 + lane, n_idx, ld, 4 * half, 4 * half + 4);
          release(lane);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (kStaged<EPI>) {
        if (elect_one()) bulk_wait<0>();  // all TMA stores / reductions of this warp have landed
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still multicast into my smem / arrive on my barriers until it is done too
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
__global__ void __launch_bounds__(128)
gemm_reference_kernel(const __half* __restrict__ A, const __half* __restrict__ W, int K, int n_tiles,
                      GemmEpiParams p) {
  const int m_idx = blockIdx.x / n_tiles, n_idx = blockIdx.x % n_tiles;
  const int row = m_idx * BM + threadIdx.x;
  ReferenceLoader ld{row < p.M ? A + static_cast<size_t>(row) * K : nullptr,
                     W + static_cast<size_t>(n_idx) * BN * K, K,
                     (p.a_lo && row < p.M) ? p.a_lo + static_cast<size_t>(row) * K : nullptr,
                     p.w_lo ? p.w_lo + static_cast<size_t>(n_idx) * BN * K : nullptr};
  epilogue_row<EPI>(p, row, n_idx, ld);
}

// cudaFuncSetAttribute is per device: several contexts on different GPUs may live in one process
inline bool first_use_on_device(std::atomic<uint64_t>& mask, int device) {
  const uint64_t bit = 1ull << (device & 63);
  return (mask.fetch_or(bit) & bit) == 0;
}

template <int EPI, int STAGES, bool SPLIT>
void launch_tc(vrag_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
               const CUtensorMap& tmOut2, const CUtensorMap& tmAlo, const CUtensorMap& tmBlo,
               const CUtensorMap& tmOutLo, int m_tiles, int n_tiles, int k_blocks, const GemmEpiParams& p) {
  constexpr int SMEM_BYTES = smem_bytes(STAGES, EPI);
  static_assert(SMEM_BYTES <= 232448, "GEMM shared memory exceeds the 227 KB per-CTA limit");
  static std::atomic<uint64_t> attr_mask{0};
  if (first_use_on_device(attr_mask, ctx->device))
    VRAG_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SMEM_BYTES));
  const int total_pairs = ((m_tiles + 1) / 2) * n_tiles;
  const int max_clusters = ctx->num_sms / 2;
  const int grid = 2 * (total_pairs < max_clusters ? total_pairs : max_clusters);   // CTA pairs (clusters of 2)
  gemm_tcgen05_kernel<EPI, STAGES, SPLIT><<<grid, GEMM_THREADS, SMEM_BYTES, ctx->stream>>>(
      tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p);
}

// epilogues available in split-precision mode (the deferred-LayerNorm family belongs to the two-plane fp16 + e5m2 fast path)
template <int EPI>
constexpr bool kSplitOk = !kResidStats<EPI> && !kNorm<EPI> && !kNormBias<EPI>;
template <int EPI>
constexpr bool kHalfOut = (EPI == EPI_F16 || EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || kRope<EPI> || kGeglu<EPI>);

template <int EPI>
void launch_t(vrag_ctx* ctx, const __half* A, const __half* W, int M, int N, int K, const GemmEpiParams& p,
              int use_reference) {
  const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN, k_blocks = K / BK;
  ProfScope prof(ctx, p.prof_class);
  const bool split = p.a_lo != nullptr || p.w_lo != nullptr;
  if (split) {
    VRAG_CHECK(p.a_lo && p.w_lo, VRAG_ERR_ARG, "gemm: split precision needs the low planes of both operands");
    VRAG_CHECK(kSplitOk<EPI>, VRAG_ERR_ARG, "gemm: this epilogue has no split-precision variant");
    VRAG_CHECK(!kHalfOut<EPI> || p.out16_lo, VRAG_ERR_ARG, "gemm: split precision needs out16_lo for fp16 outputs");
  }
  if (use_reference) {
    gemm_reference_kernel<EPI><<<m_tiles * n_tiles, 128, 0, ctx->stream>>>(A, W, K, n_tiles, p);
  } else {
    CUtensorMap tmA = make_tmap_2d(ctx, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, K, K, BM, BK);
    CUtensorMap tmB = make_tmap_2d(ctx, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, K, BN / 2, BK);  // half W tile
    CUtensorMap tmOut = tmA;  // placeholder for the epilogues that write directly
    CUtensorMap tmOut2 = tmA;
    CUtensorMap tmAlo = tmA, tmBlo = tmB, tmOutLo = tmA;   // placeholders unless split
    if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_BIAS_RESID_F32)
      tmOut = make_tmap_2d(ctx, p.out32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, p.ld32, p.ld32, 32, 32);
    else if constexpr (kStaged<EPI>)
      tmOut = make_tmap_2d(ctx, p.out16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, p.ld16, p.ld16, 32, 64);
    if constexpr (kRope<EPI>) {
      VRAG_CHECK(p.rope_tab && p.rope_rows > 0 && p.pos, VRAG_ERR_ARG, "gemm: RoPE epilogue needs rope_tab / pos");
      tmOut2 = make_tmap_2d(ctx, p.rope_tab, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.rope_rows, 64, 64, 32, 32);
    }
    if constexpr (kResidStats<EPI>) {
      VRAG_CHECK(p.out16 && p.out8_lo && p.stats_out, VRAG_ERR_ARG, "gemm: RESID_STATS needs out16 / out8_lo / stats_out");
      if constexpr (EPI == EPI_RESID_STATS_LN)
        VRAG_CHECK(p.gamma && p.bias && p.stats_in && p.stats_in != p.stats_out, VRAG_ERR_ARG,
                   "gemm: RESID_STATS_LN needs gamma / bias and separate stats_in / stats_out buffers");
      tmOut = make_tmap_2d(ctx, p.out16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, p.ld16, p.ld16, 32, 64);
      tmOut2 = make_tmap_2d(ctx, p.out8_lo, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, p.ld16, p.ld16, 32, 64,
                            CU_TENSOR_MAP_SWIZZLE_64B);
      // 12 KB of residual staging per epilogue warp leave room for 4 operand stages (the kernel is HBM-bound)
      launch_tc<EPI, 4, false>(ctx, tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p);
    } else if (split) {
      if constexpr (kSplitOk<EPI>) {
        tmAlo = make_tmap_2d(ctx, p.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, K, K, BM, BK);
        tmBlo = make_tmap_2d(ctx, p.w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, K, BN / 2, BK);
        if constexpr (kHalfOut<EPI>)
          tmOutLo = make_tmap_2d(ctx, p.out16_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, p.ld16, p.ld16, 32, 64);
        launch_tc<EPI, 5, true>(ctx, tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p);
      }
    } else {
      switch (ctx->gemm_stages) {
        case 3: launch_tc<EPI, 3, false>(ctx, tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p); break;
        default: launch_tc<EPI, 5, false>(ctx, tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p); break;
        case 4: launch_tc<EPI, 4, false>(ctx, tmA, tmB, tmOut, tmOut2, tmAlo, tmBlo, tmOutLo, m_tiles, n_tiles, k_blocks, p); break;
      }
    }
  }
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace

void launch_gemm(vrag_ctx* ctx, int epi, const __half* A, const __half* W, int M, int N, int K,
                 const GemmEpiParams& p, int use_reference) {
  VRAG_CHECK(M > 0 && N % BN == 0 && K % BK == 0, VRAG_ERR_ARG, "gemm: need M > 0, N % 256 == 0, K % 64 == 0");
  switch (epi) {
#define VRAG_CASE(E) case E: launch_t<E>(ctx, A, W, M, N, K, p, use_reference); break;
    VRAG_CASE(EPI_F16)
    VRAG_CASE(EPI_ROPE_QKV)
    VRAG_CASE(EPI_RESID_F32)
    VRAG_CASE(EPI_GEGLU)
    VRAG_CASE(EPI_BIAS_F16)
    VRAG_CASE(EPI_BIAS_GELU_F16)
    VRAG_CASE(EPI_BIAS_RESID_F32)
    VRAG_CASE(EPI_SPLADE)
    VRAG_CASE(EPI_GELU_F32)
    VRAG_CASE(EPI_BIAS_GELU_F32)
    VRAG_CASE(EPI_F32)
    VRAG_CASE(EPI_RESID_STATS)
    VRAG_CASE(EPI_NORM_ROPE_QKV)
    VRAG_CASE(EPI_NORM_GEGLU)
    VRAG_CASE(EPI_NORM_BIAS_F16)
    VRAG_CASE(EPI_NORM_BIAS_GELU_F16)
    VRAG_CASE(EPI_RESID_STATS_LN)
    VRAG_CASE(EPI_SCORES)
    VRAG_CASE(EPI_HEAD_PARTIAL)
    VRAG_CASE(EPI_SCORES_THRESH)
#undef VRAG_CASE
    default: throw Error(VRAG_ERR_ARG, "gemm: unknown epilogue");
  }
}

}  // namespace vrag
