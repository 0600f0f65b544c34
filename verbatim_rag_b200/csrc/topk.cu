// Exact top-k similarity scan behind the vector store (dense COSINE, sparse IP).
//
// Pipeline per query tile (one corpus pass each: <= 8 queries on the FMA scan, 16 on the tensor-core scan):
//   1. scan     : stream the corpus once from HBM, fp32 FMA dot products (or split-tf32 tcgen05 MMAs)
//                 -> scores[q][row]                                                               (HBM-bound)
//   2. select   : per query, threshold-filtered warp top-k' lists over 64-bit keys
//                 key = (ordered(score) << 32) | ~row   => order (score desc, row asc); two levels
//   3. rescore  : the k' = k + margin candidates are re-evaluated in fp64 from the stored fp32 values
//   4. rank     : exact order (score64 desc, row asc) -> ids (global), fp32 + fp64 scores
// The fp64 rescoring makes results independent of the summation order of pass 1 and identical across shard
// counts, which is what lets the sharded multi-GPU search be bit-identical to the 1-GPU search.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <memory>

#include <atomic>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

using namespace vrag;

namespace {

constexpr int QT = 8;          // queries per corpus pass (dense FMA scans)
constexpr int SQT = 32;        // queries per corpus pass of the sparse scan: one 128-byte line of the dense query table per
                               // stored term (8 per pass left a 10 k-document search launch-bound: 125 passes of ~20 us)
constexpr int MARGIN = 16;     // extra candidates kept for the fp64 re-ranking
constexpr int MAX_K = 1024;
constexpr int SEL_WARPS = 8;
constexpr int FIN_WARPS = 32;               // warps of the fused finish kernel: one fp64 re-score per warp
constexpr int SEL_UNROLL = 8;               // loads in flight per lane of the register-list selection
constexpr int64_t SEL_REG_BLOCK = 16384;    // scores per block (2048 per warp list) of the register-list selection

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t ord32(float f) {  // monotone float -> uint
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(ord32(score)) << 32) | static_cast<uint64_t>(~row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~static_cast<uint32_t>(key); }

// ---------------------------------------------------------------------------------- dense: norms
__global__ void __launch_bounds__(256)
dense_norm_kernel(const float* __restrict__ rows, int64_t n, int dim, double* __restrict__ norm64,
                  float* __restrict__ inv_norm32) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* r = rows + row * dim;
  double s = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const double v = static_cast<double>(r[i]);
    s += v * v;
  }
  s = warp_sum_d(s);
  if (lane == 0) {
    const double nrm = sqrt(s);
    norm64[row] = nrm;
    inv_norm32[row] = nrm > 0.0 ? static_cast<float>(1.0 / nrm) : 0.f;
  }
}

// ---------------------------------------------------------------------------------- dense: scan
// One warp streams 4 rows at a time (24 independent 16-byte loads per lane in flight); queries live in smem.
template <int VEC>  // dim = VEC * 128
__global__ void __launch_bounds__(256)
dense_scan_kernel(const float* __restrict__ rows, int64_t n, const float* __restrict__ queries, int nq,
                  const float* __restrict__ inv_norm_d, const float* __restrict__ inv_norm_q,
                  const uint8_t* __restrict__ deleted, float* __restrict__ scores) {
  constexpr int DIM = VEC * 128;
  __shared__ float4 sq[QT][VEC * 32];
  for (int i = threadIdx.x; i < QT * VEC * 32; i += blockDim.x) {
    const int qi = i / (VEC * 32), c = i % (VEC * 32);
    sq[qi][c] = qi < nq ? reinterpret_cast<const float4*>(queries + static_cast<size_t>(qi) * DIM)[c]
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t r0 = (static_cast<int64_t>(blockIdx.x) * 8 + warp) * 4; r0 < n; r0 += nwarps * 4) {
    float4 d[4][VEC];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t row = min(r0 + r, n - 1);
      const float4* p = reinterpret_cast<const float4*>(rows + row * DIM);
#pragma unroll
      for (int i = 0; i < VEC; ++i) d[r][i] = ldg_stream(p + i * 32 + lane);
    }
    float acc[4][QT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) acc[r][qi] = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        const float4 qv = sq[qi][i * 32 + lane];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          acc[r][qi] += (d[r][i].x * qv.x + d[r][i].y * qv.y) + (d[r][i].z * qv.z + d[r][i].w * qv.w);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) acc[r][qi] = warp_sum(acc[r][qi]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t row = r0 + r;
        if (row < n) {
          const float inv = inv_norm_d[row];
          const bool dead = deleted[row] != 0;
#pragma unroll
          for (int qi = 0; qi < QT; ++qi)
            if (qi < nq) scores[static_cast<size_t>(qi) * n + row] = dead ? -INFINITY : acc[r][qi] * inv * inv_norm_q[qi];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- dense: scan, TMA-bulk streamed
// The corpus is row-major contiguous, so 32 rows are ONE contiguous block: a producer thread streams such blocks
// with cp.async.bulk (1D TMA, no registers, up to ~190 KB in flight per SM) into an mbarrier ring; 8 consumer warps
// take 4 rows each per block, read them from smem as conflict-free float4 and keep the queries in registers
// (NQ <= 4) or in smem (NQ = 8).  Bytes in flight no longer depend on register allocation / occupancy.
constexpr int SCAN_ROWS = 32;
constexpr int SCAN_CONSUMER_WARPS = 8;
constexpr int RW = SCAN_ROWS / SCAN_CONSUMER_WARPS;  // rows per consumer warp per block

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int VEC, int NQ>  // dim = VEC * 128; NQ queries handled per corpus pass
__global__ void __launch_bounds__(32 * (SCAN_CONSUMER_WARPS + 1), 1)
dense_scan_tma_kernel(const float* __restrict__ rows, int64_t n, const float* __restrict__ queries,
                      const float* __restrict__ inv_norm_d, const float* __restrict__ inv_norm_q,
                      const uint8_t* __restrict__ deleted, float* __restrict__ scores, int nstage, int nq) {
  constexpr int DIM = VEC * 128;
  constexpr uint32_t STAGE_BYTES = SCAN_ROWS * DIM * 4;
  extern __shared__ __align__(128) uint8_t scan_smem[];
  uint8_t* stage0 = scan_smem;
  float4* sq = reinterpret_cast<float4*>(scan_smem + static_cast<size_t>(nstage) * STAGE_BYTES);  // [NQ][VEC*32]
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sq) + NQ * DIM * 4);
  uint64_t* empty = full + nstage;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int64_t nblocks = (n + SCAN_ROWS - 1) / SCAN_ROWS;

  for (int i = threadIdx.x; i < NQ * VEC * 32; i += blockDim.x)
    sq[i] = (i / (VEC * 32)) < nq ? reinterpret_cast<const float4*>(queries)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, SCAN_CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == SCAN_CONSUMER_WARPS) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int64_t r0 = b * SCAN_ROWS;
        const int64_t rows_here = (n - r0) < SCAN_ROWS ? (n - r0) : SCAN_ROWS;
        const uint32_t bytes = static_cast<uint32_t>(rows_here) * DIM * 4;
        mbar_wait(empty + st, ph ^ 1);
        mbar_arrive_expect_tx(full + st, bytes);
        bulk_load_1d(stage0 + static_cast<size_t>(st) * STAGE_BYTES, rows + r0 * DIM, bytes, full + st);
        if (++st == nstage) { st = 0; ph ^= 1; }
      }
    }
    return;
  }

  float4 qr[NQ <= 4 ? NQ : 1][VEC];
  if (NQ <= 4) {
#pragma unroll
    for (int qi = 0; qi < (NQ <= 4 ? NQ : 1); ++qi)
#pragma unroll
      for (int i = 0; i < VEC; ++i) qr[qi][i] = sq[qi * VEC * 32 + i * 32 + lane];
  }
  int st = 0;
  uint32_t ph = 0;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    mbar_wait(full + st, ph);
    const float4* blk = reinterpret_cast<const float4*>(stage0 + static_cast<size_t>(st) * STAGE_BYTES);
    float acc[RW * NQ];
#pragma unroll
    for (int i = 0; i < RW * NQ; ++i) acc[i] = 0.f;
    const float4* rp = blk + (warp * RW) * (VEC * 32) + lane;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 d[RW];
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) d[rr] = rp[rr * (VEC * 32) + i * 32];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        const float4 qv = NQ <= 4 ? qr[NQ <= 4 ? qi : 0][i] : sq[qi * VEC * 32 + i * 32 + lane];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr)
          acc[rr * NQ + qi] += (d[rr].x * qv.x + d[rr].y * qv.y) + (d[rr].z * qv.z + d[rr].w * qv.w);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);  // this warp is done reading the stage
    // Transposing butterfly: NV = RW*NQ partial sums per lane -> lane l ends with the full sum of value
    // (l >> (5 - log2 NV)); NV - 1 + (5 - log2 NV) shuffles instead of 5 * NV.
    constexpr int NV = RW * NQ;
    int off = 16;
#pragma unroll
    for (int cnt = NV / 2; cnt >= 1; cnt >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < cnt; ++i) {
        const float send = upper ? acc[i] : acc[i + cnt];
        const float keep = upper ? acc[i + cnt] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
      off >>= 1;
    }
    float total = acc[0];
#pragma unroll
    for (int o2 = 16 / NV; o2 >= 1; o2 >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o2);
    constexpr int GROUP = 32 / NV;  // lanes holding the same value
    if ((lane % GROUP) == 0) {
      const int idx = lane / GROUP, rr = idx / NQ, qi = idx % NQ;
      const int64_t row = b * SCAN_ROWS + warp * RW + rr;
      if (row < n && qi < nq)
        scores[static_cast<size_t>(qi) * n + row] =
            deleted[row] ? -INFINITY : total * inv_norm_q[qi] * inv_norm_d[row];
    }
    if (++st == nstage) { st = 0; ph ^= 1; }
  }
}

// ---------------------------------------------------------------------------------- dense: scan, tensor cores
// 16 queries per corpus pass (the HBM-bound regime of the batched search).  scores = X Q^T on tcgen05
// kind::tf32 with fp32-level accuracy from a three-term split: x = x_hi + x_lo, q = q_hi + q_lo (each part exactly
// representable in tf32), x.q ~= x_hi.q_hi + x_hi.q_lo + x_lo.q_hi, fp32 accumulation in TMEM; the dropped x_lo.q_lo
// term is 2^-22 relative, the size of one fp32 rounding.  Exactness of the final ids does not rest on it anyway: as on
// the FMA path these scores only pick the k + margin candidates that are re-scored in fp64.
//   warp 0      : TMA producer — corpus tile chunks [128 rows x 32 dims] fp32 (SWIZZLE_128B box, 16 KB), 6-stage ring
//   warps 2..9  : split        — thread = corpus row: reads its 128 bytes from smem, x_hi = rna_tf32(x),
//                               x_lo = rna_tf32(x - x_hi), and writes both as the A OPERAND INTO TENSOR MEMORY
//                               (tcgen05.st, 4 A stages of 64 columns); two groups of 4 warps take alternate chunks.
//                               (A first version wrote x_hi / x_lo back to shared memory: 100 KB of smem traffic per
//                               chunk against a budget of 631 cycles x 128 B at the HBM rate, plus a proxy fence per
//                               chunk — 4.3 TB/s.  With A in TMEM shared memory carries 38 KB per chunk.)
//   warp 1      : MMA issuer   — D1[128 x 32] += X_hi [Q_hi ; Q_lo]^T   (N = 32),  D2[128 x 16] += X_lo Q_hi^T (N = 16),
//                               A from TMEM, B = the query block, split once per CTA, resident in smem (dim/32
//                               swizzled 4 KB tiles)
//   warps 10..13: epilogue     — tcgen05.ld, (D1[q] + D1[16+q] + D2[q]) * 1/|q| * 1/|x| -> scores[q][row] (coalesced
//                               along rows); two TMEM accumulator buffers, so it overlaps the next tile's MMAs
constexpr int BIG_MIN_Q = 64;                  // smallest batch sent to the split-precision GEMM search
constexpr int TCQ = 16;
constexpr int TC_MIN_Q = 5;                    // smallest query tile sent to the tensor-core scan
constexpr int TC_ROWS = 128;
constexpr int TC_STAGES = 6;                   // raw fp32 chunks in shared memory (even: see the split groups)
constexpr int TC_ASTAGES = 4;                  // split A operands in tensor memory
constexpr int TC_SPLIT_WARPS = 8, TC_EPI_WARPS = 4;
constexpr int TC_THREADS = 32 * (2 + TC_SPLIT_WARPS + TC_EPI_WARPS);
constexpr int TC_TILE_BYTES = TC_ROWS * 128;   // one [128 x 32] fp32 tile
constexpr int TC_BTILE_BYTES = 2 * TCQ * 128;  // one [32 x 32] fp32 tile of the split query block
constexpr uint32_t TC_TMEM_COLS = 512;         // D: 2 x 64 columns ([0,32) D1, [32,48) D2); A: 4 x 64 from column 128
constexpr uint32_t TC_TMEM_A0 = 128;

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
dense_scan_tc_kernel(const __grid_constant__ CUtensorMap tmX, int64_t n, int dim, const float* __restrict__ queries,
                     int nq, const float* __restrict__ inv_norm_d, const float* __restrict__ inv_norm_q,
                     const uint8_t* __restrict__ deleted, float* __restrict__ scores) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunk = dim / 32;
  uint8_t* sA = smem;                                   // raw chunk ring
  uint8_t* sB = sA + TC_STAGES * TC_TILE_BYTES;         // nchunk tiles of 4 KB: rows 0..15 q_hi, 16..31 q_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(nchunk) * TC_BTILE_BYTES);
  uint64_t* full = bars;                       // [TC_STAGES]  TMA landed
  uint64_t* raw_empty = full + TC_STAGES;      // [TC_STAGES]  split warps hold the chunk in registers
  uint64_t* a_full = raw_empty + TC_STAGES;    // [TC_ASTAGES] x_hi / x_lo written to TMEM
  uint64_t* a_empty = a_full + TC_ASTAGES;     // [TC_ASTAGES] MMAs reading the A stage completed
  uint64_t* d_full = a_empty + TC_ASTAGES;     // [2]
  uint64_t* d_empty = d_full + 2;              // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int64_t ntiles = (n + TC_ROWS - 1) / TC_ROWS;

  // split query block -> swizzled K-major tiles (element (row, k): tile k/32, 16-byte chunk ((k%32)/4) ^ (row & 7))
  for (int i = threadIdx.x; i < 2 * TCQ * (dim / 4); i += blockDim.x) {
    const int row = i / (dim / 4), k4 = i - row * (dim / 4);
    const int qrow = row & (TCQ - 1);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qrow < nq) q = __ldg(reinterpret_cast<const float4*>(queries + static_cast<size_t>(qrow) * dim) + k4);
    float4 h = make_float4(rna_tf32(q.x), rna_tf32(q.y), rna_tf32(q.z), rna_tf32(q.w));
    if (row >= TCQ) h = make_float4(rna_tf32(q.x - h.x), rna_tf32(q.y - h.y), rna_tf32(q.z - h.z), rna_tf32(q.w - h.w));
    const int c = k4 >> 3, j = k4 & 7;
    sts128(smem_u32(sB) + c * TC_BTILE_BYTES + row * 128 + ((j ^ (row & 7)) << 4), h);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(raw_empty + s, TC_SPLIT_WARPS / 2);
    }
    for (int s = 0; s < TC_ASTAGES; ++s) {
      mbar_init(a_full + s, TC_SPLIT_WARPS / 2);
      mbar_init(a_empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d_full + b, 1);
      mbar_init(d_empty + b, TC_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmX);
  if (warp == 1) {
    tmem_alloc(tmem_holder, TC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t gs = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int c = 0; c < nchunk; ++c, ++gs) {
        const int st = gs % TC_STAGES;
        mbar_wait_tagged(raw_empty + st, ((gs / TC_STAGES) & 1) ^ 1, 21);
        if (elect_one()) {
          mbar_arrive_expect_tx(full + st, TC_TILE_BYTES);
          tma_load_2d(sA + st * TC_TILE_BYTES, &tmX, full + st, c * 32, static_cast<int32_t>(t * TC_ROWS));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_hi = umma_idesc(2, TC_ROWS, 2 * TCQ);
    constexpr uint32_t idesc_lo = umma_idesc(2, TC_ROWS, TCQ);
    const uint32_t b_base = smem_u32(sB);
    uint32_t gs = 0, ti = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
      const uint32_t buf = ti & 1;
      mbar_wait_tagged(d_empty + buf, ((ti >> 1) & 1) ^ 1, 22);  // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d1 = tmem_base + buf * 64, d2 = d1 + 2 * TCQ;
      for (int c = 0; c < nchunk; ++c, ++gs) {
        const int as = gs % TC_ASTAGES;
        mbar_wait_tagged(a_full + as, (gs / TC_ASTAGES) & 1, 23);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + TC_TMEM_A0 + as * 64, a_lo = a_hi + 32, bt = b_base + c * TC_BTILE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // K = 8 tf32 per instruction = 8 TMEM columns of A, 32 bytes of a B row
            const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
            umma_tf32_ts(d1, a_hi + k * 8, umma_desc_sw128(bt + k * 32), idesc_hi, acc);
            umma_tf32_ts(d2, a_lo + k * 8, umma_desc_sw128(bt + k * 32), idesc_lo, acc);
          }
          umma_commit(a_empty + as);
          if (c + 1 == nchunk) umma_commit(d_full + buf);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ------------------------------------------------------------------ split warps: 2 groups x (thread = row)
    const int group = (warp - 2) >> 2, quarter = warp & 3;  // TMEM lane quarter: warp w may touch lanes 32*(w%4)..
    const int row = quarter * 32 + lane;
    const uint32_t raw_row = smem_u32(sA) + row * 128;
    const uint32_t t_lane = tmem_base + TC_TMEM_A0 + (static_cast<uint32_t>(quarter * 32) << 16);
    const int sw = row & 7;
    uint32_t gs = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int c = 0; c < nchunk; ++c, ++gs) {
        if ((gs & 1) != static_cast<uint32_t>(group)) continue;  // stage counts are even: fixed group per stage
        const int st = gs % TC_STAGES, as = gs % TC_ASTAGES;
        mbar_wait_tagged(full + st, (gs / TC_STAGES) & 1, 24);
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = lds128(raw_row + st * TC_TILE_BYTES + ((i ^ sw) << 4));
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty + st);  // (release: ordered after this warp's reads of the chunk)
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h0 = rna_tf32(x[i].x), h1 = rna_tf32(x[i].y), h2 = rna_tf32(x[i].z), h3 = rna_tf32(x[i].w);
          hi[4 * i] = __float_as_uint(h0);
          hi[4 * i + 1] = __float_as_uint(h1);
          hi[4 * i + 2] = __float_as_uint(h2);
          hi[4 * i + 3] = __float_as_uint(h3);
          lo[4 * i] = __float_as_uint(rna_tf32(x[i].x - h0));
          lo[4 * i + 1] = __float_as_uint(rna_tf32(x[i].y - h1));
          lo[4 * i + 2] = __float_as_uint(rna_tf32(x[i].z - h2));
          lo[4 * i + 3] = __float_as_uint(rna_tf32(x[i].w - h3));
        }
        mbar_wait_tagged(a_empty + as, ((gs / TC_ASTAGES) & 1) ^ 1, 26);
        tc_fence_after();
        tmem_st_32x32b_x32(t_lane + as * 64, hi);
        tmem_st_32x32b_x32(t_lane + as * 64 + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + as);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = row (TMEM lane)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float qn[TCQ];
#pragma unroll
    for (int q = 0; q < TCQ; ++q) qn[q] = q < nq ? inv_norm_q[q] : 0.f;
    uint32_t ti = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
      const uint32_t buf = ti & 1;
      mbar_wait_tagged(d_full + buf, (ti >> 1) & 1, 25);
      tc_fence_after();
      uint32_t d1[32], d2[16];
      tmem_ld_32x32b_x32(t_lane + buf * 64, d1);
      tmem_ld_32x32b_x16(t_lane + buf * 64 + 2 * TCQ, d2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty + buf);
      const int64_t row = t * TC_ROWS + r;
      if (row < n) {
        const float inv = inv_norm_d[row];
        const bool dead = deleted[row] != 0;
#pragma unroll
        for (int q = 0; q < TCQ; ++q)
          if (q < nq) {
            const float v = (__uint_as_float(d1[TCQ + q]) + __uint_as_float(d2[q])) + __uint_as_float(d1[q]);
            scores[static_cast<size_t>(q) * n + row] = dead ? -INFINITY : v * qn[q] * inv;
          }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// generic dimension (dim % 4 == 0): one warp per row, queries read from global/L1
__global__ void __launch_bounds__(256)
dense_scan_generic_kernel(const float* __restrict__ rows, int64_t n, int dim, const float* __restrict__ queries,
                          int nq, const float* __restrict__ inv_norm_d, const float* __restrict__ inv_norm_q,
                          const uint8_t* __restrict__ deleted, float* __restrict__ scores) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < n; row += nwarps) {
    float acc[QT];
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) acc[qi] = 0.f;
    const float4* p = reinterpret_cast<const float4*>(rows + row * dim);
    for (int c = lane; c < dim / 4; c += 32) {
      const float4 dv = ldg_stream(p + c);
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        if (qi < nq) {
          const float4 qv = __ldg(reinterpret_cast<const float4*>(queries + static_cast<size_t>(qi) * dim) + c);
          acc[qi] += (dv.x * qv.x + dv.y * qv.y) + (dv.z * qv.z + dv.w * qv.w);
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) acc[qi] = warp_sum(acc[qi]);
    if (lane == 0) {
      const float inv = inv_norm_d[row];
      const bool dead = deleted[row] != 0;
#pragma unroll
      for (int qi = 0; qi < QT; ++qi)
        if (qi < nq) scores[static_cast<size_t>(qi) * n + row] = dead ? -INFINITY : acc[qi] * inv * inv_norm_q[qi];
    }
  }
}

// ---------------------------------------------------------------------------------- sparse: scan
// qT: dense query block [dim][SQT] (row = vocabulary id, SQT consecutive floats = one 128-byte line per gather).
// One block per query of the selection group: query q is slot q % SQT of table q / SQT (tables back to back).
__global__ void sparse_scatter_query_kernel(const int64_t* __restrict__ q_indptr, const int32_t* __restrict__ q_idx,
                                            const float* __restrict__ q_val, int nq, int dim, float* __restrict__ qT) {
  const int qi = blockIdx.x;
  if (qi >= nq) return;
  float* table = qT + static_cast<size_t>(qi / SQT) * dim * SQT;
  const int slot = qi % SQT;
  for (int64_t j = q_indptr[qi] + threadIdx.x; j < q_indptr[qi + 1]; j += blockDim.x) {
    const int t = q_idx[j];
    if (t >= 0 && t < dim) table[static_cast<size_t>(t) * SQT + slot] = q_val[j];
  }
}

// Warp per CSR row, lanes arranged 4 terms x 8 lanes.  The cost is the gather of the dense query table.  First version
// (lane = term, each lane reading its term's 32 queries with 8 float4 loads): a warp-wide 16-byte load to 32 DIFFERENT table
// rows is 32 L1 wavefronts of 16 useful bytes each -- 4.3 ms per pass over 128 M stored terms, L2 at 20 % and the SM at
// 9 % of peak (ncu, profiles/r2_sparse_scan_full_raw.csv).  Here the 8 lanes of a group read ONE 128-byte table row
// together (one wavefront, 4 queries per lane) and a warp-wide load covers 4 terms: 1.3 ms per pass.  What bounds it now
// is the SM's 128 B/clk load-return path (l1tex data-pipe wavefronts 74 % of peak, profiles/r2b_sparse_scan_*): 128 bytes
// of table per stored term have to reach registers, and the shuffles that hand out indices / values share that path.
// Tried and dropped: tiles of 8 rows per warp with 32-byte score stores (same time at 1 M rows, slower at 10 k: the
// 4-byte stores were not the limit).
__global__ void __launch_bounds__(256, 4)
sparse_scan_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const float* __restrict__ values, int64_t n, const float* __restrict__ qT, int dim, int nq_group,
                   const uint8_t* __restrict__ deleted, float* __restrict__ scores) {
  static_assert(SQT == 32, "lane layout: 8 lanes x 4 queries per table row");
  // blockIdx.y = query tile of the selection group: one launch scans for every tile (on a small corpus a launch per tile
  // is one wave + a tail each: 0.93 -> 0.59 ms for 32 tiles over 10 k rows; no difference at 1 M rows)
  const int tile = static_cast<int>(blockIdx.y);
  qT += static_cast<size_t>(tile) * dim * SQT;
  scores += static_cast<size_t>(tile) * SQT * n;
  const int nq = min(SQT, nq_group - tile * SQT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;   // group = term slot of a step, j = which 4 queries of the table row
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < n; row += nwarps) {
    const int64_t a = indptr[row], b = indptr[row + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int t = a + lane < b ? __ldg(indices + a + lane) : 0;       // padding terms: table row 0 with weight 0
    float v = a + lane < b ? __ldg(values + a + lane) : 0.f;
    for (int64_t base = a; base < b; base += 32) {
      const int64_t next = base + 32 + lane;   // the next 32 terms are in flight while this block gathers
      const int t_next = next < b ? __ldg(indices + next) : 0;
      const float v_next = next < b ? __ldg(values + next) : 0.f;
      // No branch on the tail: padding terms cost an L1 hit on table row 0, a branch per step would make the compiler
      // wait for each gather before issuing the next (seen in the SASS of an earlier version).
      float4 q[8];
      float vv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int tt = __shfl_sync(0xffffffffu, t, 4 * k + g);
        vv[k] = __shfl_sync(0xffffffffu, v, 4 * k + g);
        q[k] = __ldg(reinterpret_cast<const float4*>(qT + static_cast<size_t>(tt) * SQT) + j);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc.x = fmaf(vv[k], q[k].x, acc.x);
        acc.y = fmaf(vv[k], q[k].y, acc.y);
        acc.z = fmaf(vv[k], q[k].z, acc.z);
        acc.w = fmaf(vv[k], q[k].w, acc.w);
      }
      t = t_next;
      v = v_next;
    }
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {   // sum over the 4 term slots
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    if (lane < 8) {
      const bool dead = deleted[row] != 0;
      const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (4 * lane + e < nq) scores[static_cast<size_t>(4 * lane + e) * n + row] = dead ? -INFINITY : r[e];
    }
  }
}

// ---------------------------------------------------------------------------------- select
// Sorted (descending) list of kp keys per warp in smem; a key enters only if it beats the current kp-th best.
__device__ __forceinline__ void list_insert(uint64_t* L, int kp, uint64_t x, int lane) {
  int cnt = 0;
  for (int j = lane; j < kp; j += 32) cnt += L[j] > x;
  const int pos = __reduce_add_sync(0xffffffffu, cnt);
  // shift [pos, kp-1) right by one: read everything first, then write
  uint64_t tmp[(MAX_K + MARGIN + 31) / 32];
  int c = 0;
  for (int j = lane; j < kp; j += 32, ++c) tmp[c] = (j > pos) ? L[j - 1] : 0;
  __syncwarp();
  c = 0;
  for (int j = lane; j < kp; j += 32, ++c) {
    if (j > pos) L[j] = tmp[c];
    else if (j == pos) L[j] = x;
  }
  __syncwarp();
}

template <bool FROM_SCORES>
__global__ void __launch_bounds__(32 * SEL_WARPS)
select_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ keys_in, int64_t n, int64_t stride, int kp,
              uint64_t* __restrict__ keys_out /* [nq][gridDim.x][kp] */) {
  extern __shared__ uint64_t lists[];  // [SEL_WARPS][kp]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.y;
  uint64_t* L = lists + static_cast<size_t>(warp) * kp;
  for (int j = lane; j < kp; j += 32) L[j] = 0;
  __syncwarp();
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = per_block * blockIdx.x, b1 = min(n, b0 + per_block);
  const int64_t per_warp = (b1 - b0 + SEL_WARPS - 1) / SEL_WARPS;
  const int64_t w0 = b0 + per_warp * warp, w1 = min(b1, w0 + per_warp);
  uint64_t thr = 0;
  for (int64_t i0 = w0; i0 < w1; i0 += 32) {
    const int64_t i = i0 + lane;
    uint64_t key = 0;
    if (i < w1) {
      if (FROM_SCORES) {
        const float s = scores[static_cast<size_t>(q) * stride + i];
        key = (s == -INFINITY) ? 0 : make_key(s, static_cast<uint32_t>(i));
      } else {
        key = keys_in[static_cast<size_t>(q) * stride + i];
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, key > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t x = __shfl_sync(0xffffffffu, key, src);
      if (x > thr) {
        list_insert(L, kp, x, lane);
        thr = L[kp - 1];
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < SEL_WARPS; ++w) {
      const uint64_t* O = lists + static_cast<size_t>(w) * kp;
      for (int j = 0; j < kp; ++j) {
        const uint64_t x = O[j];
        if (x <= thr) break;  // O is sorted descending
        list_insert(L, kp, x, lane);
        thr = L[kp - 1];
      }
    }
    uint64_t* out = keys_out + (static_cast<size_t>(q) * gridDim.x + blockIdx.x) * kp;
    for (int j = lane; j < kp; j += 32) out[j] = L[j];
  }
}

// kp <= 32 (k <= 16, the common case): the warp's sorted list lives in REGISTERS, one key per lane (lane j holds the
// j-th best).  An insertion is one ballot + one 64-bit shuffle, no shared memory round trips.
__device__ __forceinline__ uint64_t reg_list_insert(uint64_t mine, uint64_t x, int kp, int lane) {
  const int pos = __popc(__ballot_sync(0xffffffffu, mine > x));  // sorted descending: lanes [0,pos) hold keys > x
  const uint64_t up = __shfl_up_sync(0xffffffffu, mine, 1);
  if (lane == pos) mine = x;
  else if (lane > pos) mine = up;
  return lane < kp ? mine : 0;
}

template <bool FROM_SCORES>
__global__ void __launch_bounds__(32 * SEL_WARPS)
select_reg_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ keys_in, int64_t n, int64_t stride,
                  int kp, uint64_t* __restrict__ keys_out /* [nq][gridDim.x][kp] */) {
  __shared__ uint64_t lists[SEL_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.y;
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = per_block * blockIdx.x, b1 = min(n, b0 + per_block);
  const int64_t per_warp = (b1 - b0 + SEL_WARPS - 1) / SEL_WARPS;
  const int64_t w0 = b0 + per_warp * warp, w1 = min(b1, w0 + per_warp);
  // SEL_UNROLL independent coalesced loads per lane are in flight before anything is compared: the stream is bound
  // by memory latency and by the ~kp ln(m / kp) list insertions per warp, not by bandwidth.
  uint64_t mine = 0, thr = 0;
  for (int64_t i0 = w0; i0 < w1; i0 += 32 * SEL_UNROLL) {
    // unconditional loads from clamped indices (a guarded load compiles to a branch around load + use, which
    // serialises the eight round trips: measured 150 us instead of 25 us per 8 192 scores); tail lanes are masked after
    uint64_t key[SEL_UNROLL];
    float sc[SEL_UNROLL];
#pragma unroll
    for (int u = 0; u < SEL_UNROLL; ++u) {
      const int64_t i = min(i0 + u * 32 + lane, w1 - 1);
      if (FROM_SCORES) sc[u] = __ldcs(scores + static_cast<size_t>(q) * stride + i);
      else key[u] = __ldcs(keys_in + static_cast<size_t>(q) * stride + i);
    }
#pragma unroll
    for (int u = 0; u < SEL_UNROLL; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      if (FROM_SCORES) key[u] = (sc[u] == -INFINITY) ? 0 : make_key(sc[u], static_cast<uint32_t>(i));
      if (i >= w1) key[u] = 0;
    }
    bool hit = false;
#pragma unroll
    for (int u = 0; u < SEL_UNROLL; ++u) hit |= key[u] > thr;
    if (!__any_sync(0xffffffffu, hit)) continue;
#pragma unroll
    for (int u = 0; u < SEL_UNROLL; ++u) {
      unsigned m = __ballot_sync(0xffffffffu, key[u] > thr);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t x = __shfl_sync(0xffffffffu, key[u], src);
        if (x > thr) {
          mine = reg_list_insert(mine, x, kp, lane);
          thr = __shfl_sync(0xffffffffu, mine, kp - 1);
        }
      }
    }
  }
  lists[warp][lane] = mine;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < SEL_WARPS; ++w) {
      for (int j = 0; j < kp; ++j) {
        const uint64_t x = lists[w][j];
        if (x <= thr) break;  // sorted descending
        mine = reg_list_insert(mine, x, kp, lane);
        thr = __shfl_sync(0xffffffffu, mine, kp - 1);
      }
    }
    if (lane < kp) keys_out[(static_cast<size_t>(q) * gridDim.x + blockIdx.x) * kp + lane] = mine;
  }
}

// ---------------------------------------------------------------------------------- rescore (fp64)
__global__ void __launch_bounds__(256)
dense_rescore_kernel(const float* __restrict__ rows, int dim, const double* __restrict__ norm64,
                     const float* __restrict__ queries, const uint64_t* __restrict__ cand, int kp,
                     double* __restrict__ score64, int64_t* __restrict__ cand_row) {
  const int q = blockIdx.y;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= kp) return;
  const uint64_t key = cand[static_cast<size_t>(q) * kp + c];
  const size_t o = static_cast<size_t>(q) * kp + c;
  if (key == 0) {
    if (lane == 0) { score64[o] = -INFINITY; cand_row[o] = -1; }
    return;
  }
  const int64_t row = key_row(key);
  const float* r = rows + row * dim;
  const float* qv = queries + static_cast<size_t>(q) * dim;
  double dot = 0.0, qq = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const double a = static_cast<double>(r[i]), b = static_cast<double>(qv[i]);
    dot += a * b;
    qq += b * b;
  }
  dot = warp_sum_d(dot);
  qq = warp_sum_d(qq);
  if (lane == 0) {
    const double den = sqrt(qq) * norm64[row];
    score64[o] = den > 0.0 ? dot / den : 0.0;
    cand_row[o] = row;
  }
}

__global__ void __launch_bounds__(256)
sparse_rescore_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const float* __restrict__ values, const float* __restrict__ qT, size_t qT_stride, int qslot,
                      const uint64_t* __restrict__ cand, int kp, double* __restrict__ score64,
                      int64_t* __restrict__ cand_row) {
  // grid.y = query within the selection group; qslot = -1: the group's query tables lie back to back, QT queries each
  // ([tile][dim][QT]), query q is slot q % QT of table q / QT
  const int q = blockIdx.y;
  const int slot = qslot < 0 ? (q % SQT) : qslot;
  if (qslot < 0) qT += static_cast<size_t>(q / SQT) * qT_stride;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= kp) return;
  const size_t o = static_cast<size_t>(q) * kp + c;
  const uint64_t key = cand[o];
  if (key == 0) {
    if (lane == 0) { score64[o] = -INFINITY; cand_row[o] = -1; }
    return;
  }
  const int64_t row = key_row(key);
  double dot = 0.0;
  for (int64_t j = indptr[row] + lane; j < indptr[row + 1]; j += 32)
    dot += static_cast<double>(values[j]) * static_cast<double>(qT[static_cast<size_t>(indices[j]) * SQT + slot]);
  dot = warp_sum_d(dot);
  if (lane == 0) { score64[o] = dot; cand_row[o] = row; }
}

// ---------------------------------------------------------------------------------- final rank
// m candidates (score64, id) per query -> best k by (score desc, id asc).  id < 0 = empty slot.
__global__ void __launch_bounds__(256)
rank_kernel(const double* __restrict__ score64, const int64_t* __restrict__ ids, int m, int k, int64_t id_base,
            int64_t* __restrict__ ids_out, float* __restrict__ scores_out, double* __restrict__ scores64_out) {
  const int q = blockIdx.x;
  const double* s = score64 + static_cast<size_t>(q) * m;
  const int64_t* id = ids + static_cast<size_t>(q) * m;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {  // default fill
    ids_out[static_cast<size_t>(q) * k + j] = -1;
    scores_out[static_cast<size_t>(q) * k + j] = -INFINITY;
    if (scores64_out) scores64_out[static_cast<size_t>(q) * k + j] = -INFINITY;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int64_t idi = id[i];
    if (idi < 0) continue;
    const double si = s[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const int64_t idj = id[j];
      if (idj < 0 || j == i) continue;
      const double sj = s[j];
      rank += (sj > si) || (sj == si && (idj < idi || (idj == idi && j < i)));
    }
    if (rank < k) {
      ids_out[static_cast<size_t>(q) * k + rank] = idi + id_base;
      scores_out[static_cast<size_t>(q) * k + rank] = static_cast<float>(si);
      if (scores64_out) scores64_out[static_cast<size_t>(q) * k + rank] = si;
    }
  }
}

// ---------------------------------------------------------------------------------- fused finish (dense, kp <= 32)
// One block per query: (A) merge the per-block candidate lists of the streaming selection into the best kp keys,
// (B) re-score those kp rows in fp64 (same summation order as dense_rescore_kernel), (C) rank by (score desc, row asc)
// and write the k results -- the work of select_reg_kernel<false> + dense_rescore_kernel + rank_kernel in one launch
// (three dependent ~10 us launches per query tile were a fifth of a 16-query search over 1 M rows).
__global__ void __launch_bounds__(32 * FIN_WARPS)
dense_finish_kernel(const uint64_t* __restrict__ keys_in, int64_t n_keys, int kp, int k, const float* __restrict__ rows,
                    int dim, const double* __restrict__ norm64, const float* __restrict__ queries, int64_t id_base,
                    int64_t* __restrict__ ids_out, float* __restrict__ scores_out, double* __restrict__ scores64_out) {
  __shared__ uint64_t lists[FIN_WARPS][32];
  __shared__ double s64[32];
  __shared__ int64_t crow[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x;
  // (A)
  const int64_t per_warp = (n_keys + FIN_WARPS - 1) / FIN_WARPS;
  const int64_t w0 = per_warp * warp, w1 = min(n_keys, w0 + per_warp);
  uint64_t mine = 0, thr = 0;
  for (int64_t i0 = w0; i0 < w1; i0 += 32) {
    const int64_t i = i0 + lane;
    const uint64_t key = i < w1 ? keys_in[static_cast<size_t>(q) * n_keys + i] : 0;
    unsigned m = __ballot_sync(0xffffffffu, key > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t x = __shfl_sync(0xffffffffu, key, src);
      if (x > thr) {
        mine = reg_list_insert(mine, x, kp, lane);
        thr = __shfl_sync(0xffffffffu, mine, kp - 1);
      }
    }
  }
  lists[warp][lane] = mine;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < FIN_WARPS; ++w) {
      for (int j = 0; j < kp; ++j) {
        const uint64_t x = lists[w][j];
        if (x <= thr) break;  // sorted descending
        mine = reg_list_insert(mine, x, kp, lane);
        thr = __shfl_sync(0xffffffffu, mine, kp - 1);
      }
    }
    lists[0][lane] = mine;
  }
  __syncthreads();
  // (B)
  const float* qv = queries + static_cast<size_t>(q) * dim;
  for (int c = warp; c < kp; c += FIN_WARPS) {   // one candidate per warp
    const uint64_t key = lists[0][c];
    if (key == 0) {
      if (lane == 0) { s64[c] = -INFINITY; crow[c] = -1; }
      continue;
    }
    const int64_t row = key_row(key);
    const float* r = rows + row * dim;
    double dot = 0.0, qq = 0.0;
    for (int i = lane; i < dim; i += 32) {
      const double a = static_cast<double>(r[i]), b = static_cast<double>(qv[i]);
      dot += a * b;
      qq += b * b;
    }
    dot = warp_sum_d(dot);
    qq = warp_sum_d(qq);
    if (lane == 0) {
      const double den = sqrt(qq) * norm64[row];
      s64[c] = den > 0.0 ? dot / den : 0.0;
      crow[c] = row;
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {  // default fill (fewer than k live rows)
    ids_out[static_cast<size_t>(q) * k + j] = -1;
    scores_out[static_cast<size_t>(q) * k + j] = -INFINITY;
    if (scores64_out) scores64_out[static_cast<size_t>(q) * k + j] = -INFINITY;
  }
  __syncthreads();
  // (C)
  if (threadIdx.x < kp) {
    const int i = threadIdx.x;
    const int64_t idi = crow[i];
    if (idi >= 0) {
      const double si = s64[i];
      int rank = 0;
      for (int j = 0; j < kp; ++j) {
        const int64_t idj = crow[j];
        if (idj < 0 || j == i) continue;
        const double sj = s64[j];
        rank += (sj > si) || (sj == si && (idj < idi || (idj == idi && j < i)));
      }
      if (rank < k) {
        ids_out[static_cast<size_t>(q) * k + rank] = idi + id_base;
        scores_out[static_cast<size_t>(q) * k + rank] = static_cast<float>(si);
        if (scores64_out) scores64_out[static_cast<size_t>(q) * k + rank] = si;
      }
    }
  }
}

// Batched search (vrag_index_search_dense, >= BIG_MIN_Q queries): the corpus as split-precision planes of the
// NORMALISED rows, x / |x| = hi + lo (fp16 each, >= 21 significant bits), the B operand of the split GEMM of gemm.cu.
__global__ void __launch_bounds__(256)
normalize_split_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm, int64_t r0, int64_t r1,
                       int dim, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t i4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index inside [r0, r1)
  const int per_row = dim / 4;
  const int64_t row = r0 + i4 / per_row;
  if (row >= r1) return;
  const size_t o = static_cast<size_t>(row) * dim + static_cast<size_t>(i4 % per_row) * 4;
  const float4 x = *reinterpret_cast<const float4*>(rows + o);
  const float s = inv_norm ? inv_norm[row] : 1.f;
  const float v[4] = {x.x * s, x.y * s, x.z * s, x.w * s};
  __half h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = __float2half_rn(v[e]);
    l[e] = __float2half_rn(v[e] - __half2float(h[e]));
  }
  *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l);
}

// Fused selection of the batched search: per query, the kp best keys of the SAMPLE (the first rows of the corpus) seed
// the candidate list, and the kp-th of them is the threshold every later row must reach to matter at all.
__global__ void seed_candidates_kernel(const uint64_t* __restrict__ sample_keys /*[nq][kp] sorted desc*/, int kp, int cap,
                                       unsigned long long* __restrict__ cand, int* __restrict__ count,
                                       float* __restrict__ thr) {
  const int q = blockIdx.x, j = threadIdx.x;
  if (j < kp) cand[static_cast<size_t>(q) * cap + j] = sample_keys[static_cast<size_t>(q) * kp + j];
  if (j == 0) {
    count[q] = kp;
    const uint64_t last = sample_keys[static_cast<size_t>(q) * kp + kp - 1];
    if (last == 0) {
      thr[q] = -INFINITY;   // fewer than kp live rows in the sample: everything is a candidate
    } else {
      const uint32_t u = static_cast<uint32_t>(last >> 32);
      thr[q] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
    }
  }
}
__global__ void overflow_flag_kernel(const int* __restrict__ count, int nq, int cap, int* __restrict__ flag) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && count[q] > cap) atomicExch(flag, 1);
}

__global__ void query_norm_kernel(const float* __restrict__ queries, int dim, float* __restrict__ inv_norm_q) {
  const int q = blockIdx.x;
  const int lane = threadIdx.x;
  double s = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const double v = static_cast<double>(queries[static_cast<size_t>(q) * dim + i]);
    s += v * v;
  }
  s = warp_sum_d(s);
  if (lane == 0) inv_norm_q[q] = s > 0.0 ? static_cast<float>(1.0 / sqrt(s)) : 0.f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
struct vrag_index {
  vrag_ctx* ctx = nullptr;
  int kind = 0, dim = 0;
  int64_t n = 0, cap = 0;
  int64_t id_base = 0;  // added to row numbers in results (global id of this shard's row 0)
  DevBuf rows, norm64, inv32, deleted;
  // batched search: split-precision planes of the normalised rows [planes_cap, dim] (built lazily, rows [0, planes_n))
  DevBuf rows_hi, rows_lo, q_hi, q_lo, cand, cand_count, cand_thr, flag;
  int64_t planes_n = 0, planes_cap = 0;
  // metadata-filter pushdown (vrag_index_set_filter): masked = deleted | excluded, consulted instead of `deleted`
  DevBuf masked, excl;
  bool filter_on = false;
  const uint8_t* skip() const { return filter_on ? masked.as<uint8_t>() : deleted.as<uint8_t>(); }
  // sparse
  DevBuf indptr, indices, values;
  int64_t nnz = 0, nnz_cap = 0;
  // work
  DevBuf scores, keys0, keys1, s64, crow, qdev, qnorm, qT, qip, qidx, qval, out_ids, out_s32, out_s64;
  ~vrag_index() {
    for (DevBuf* b : {&rows, &norm64, &inv32, &deleted, &masked, &excl, &indptr, &indices, &values, &scores, &keys0, &keys1, &s64,
                      &crow, &qdev, &qnorm, &qT, &qip, &qidx, &qval, &out_ids, &out_s32, &out_s64, &rows_hi, &rows_lo, &q_hi,
                      &q_lo, &cand, &cand_count, &cand_thr, &flag})
      b->release();
  }
};

namespace {

void grow(vrag_ctx* ctx, DevBuf& b, size_t used_bytes, size_t need_bytes) {
  if (need_bytes <= b.bytes) return;
  size_t nb = std::max(need_bytes, b.bytes + b.bytes / 2);
  void* np = nullptr;
  VRAG_CUDA(cudaMalloc(&np, nb));
  if (used_bytes) VRAG_CUDA(cudaMemcpyAsync(np, b.p, used_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  VRAG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (b.p) cudaFree(b.p);
  b.p = np;
  b.bytes = nb;
}

// scores [nq_tile][n] on device -> final top-k for the tile written at out offsets
// ---------------------------------------------------------------------------------- peer exchange (multi-GPU)
// The all-gather of the per-shard top-k as plain stores into peer memory (NVLink / NVSwitch): every rank writes its
// [nq][k] block of 16-byte (fp64 score bits, global id) records into slot `rank` of EVERY rank's exchange buffer
// ([world][nq][k][2] int64, a symmetric allocation whose peer mappings the host passes in).  One block per peer x
// chunk; 16-byte stores, coalesced.  After a cross-rank barrier each rank merges its own buffer (rank_packed_kernel).
constexpr int MAX_PEERS = 16;
struct PeerPtrs {
  unsigned long long p[MAX_PEERS];
};

__global__ void __launch_bounds__(256)
topk_publish_kernel(const double* __restrict__ score64, const int64_t* __restrict__ ids, int64_t n_rec, PeerPtrs peers,
                    int rank) {
  longlong2* dst = reinterpret_cast<longlong2*>(peers.p[blockIdx.y]) + static_cast<size_t>(rank) * n_rec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = make_longlong2(__double_as_longlong(score64[i]), ids[i]);
  __threadfence_system();   // the records are performed system-wide before the kernel (and the barrier after it) ends
}

// rank_kernel on the exchange buffer: candidate j of query q = record (j / k of rank, q, j % k)
__global__ void __launch_bounds__(256)
rank_packed_kernel(const longlong2* __restrict__ packed, int world, int nq, int k, int64_t* __restrict__ ids_out,
                   float* __restrict__ scores_out, double* __restrict__ scores64_out) {
  extern __shared__ longlong2 rec[];   // [world * k]
  const int q = blockIdx.x, m = world * k;
  for (int j = threadIdx.x; j < m; j += blockDim.x)
    rec[j] = packed[(static_cast<size_t>(j / k) * nq + q) * k + (j % k)];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {  // default fill
    ids_out[static_cast<size_t>(q) * k + j] = -1;
    scores_out[static_cast<size_t>(q) * k + j] = -INFINITY;
    if (scores64_out) scores64_out[static_cast<size_t>(q) * k + j] = -INFINITY;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int64_t idi = rec[i].y;
    if (idi < 0) continue;
    const double si = __longlong_as_double(rec[i].x);
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const int64_t idj = rec[j].y;
      if (idj < 0 || j == i) continue;
      const double sj = __longlong_as_double(rec[j].x);
      rank += (sj > si) || (sj == si && (idj < idi || (idj == idi && j < i)));
    }
    if (rank < k) {
      ids_out[static_cast<size_t>(q) * k + rank] = idi;
      scores_out[static_cast<size_t>(q) * k + rank] = static_cast<float>(si);
      if (scores64_out) scores64_out[static_cast<size_t>(q) * k + rank] = si;
    }
  }
}

void select_and_rank(vrag_index* ix, const float* scores, int nq_tile, int k, bool dense,
                     const float* queries_dev /*dense*/, int64_t* ids_out, float* s32_out,
                     double* s64_out /*device, tile offset applied*/, cudaStream_t st, int64_t stride = 0) {
  vrag_ctx* ctx = ix->ctx;
  const int64_t n = ix->n;
  if (stride == 0) stride = n;   // floats between the score rows of consecutive queries
  const int kp = static_cast<int>(std::min<int64_t>(k + MARGIN, std::max<int64_t>(n, 1)));
  // scores per block: 16 384 with register lists (2 048 per warp = 8 rounds of 8 loads in flight per lane; ~110
  // insertions, 5 % of the keys; 16 queries x 62 blocks at 1 M rows are one wave of the GPU), 8 192 with
  // shared-memory lists (k > 16)
  const int64_t per_blk = kp <= 32 ? SEL_REG_BLOCK : 8192;
  const int nblk0 = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + per_blk - 1) / per_blk, ctx->num_sms * 4)));
  const size_t smem = static_cast<size_t>(SEL_WARPS) * kp * 8;
  static std::atomic<uint64_t> smem_attr{0};   // cudaFuncSetAttribute is per device
  const uint64_t dev_bit = 1ull << (ctx->device & 63);
  if ((smem_attr.fetch_or(dev_bit) & dev_bit) == 0) {
    VRAG_CUDA(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    VRAG_CUDA(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
  }
  ix->keys0.reserve(static_cast<size_t>(nq_tile) * nblk0 * kp * 8);
  ix->keys1.reserve(static_cast<size_t>(nq_tile) * kp * 8);
  ProfScope prof(ctx, PROF_SELECT, st);
  if (kp <= 32)
    select_reg_kernel<true><<<dim3(nblk0, nq_tile), 32 * SEL_WARPS, 0, st>>>(
        scores, nullptr, n, stride, kp, ix->keys0.as<uint64_t>());
  else
    select_kernel<true><<<dim3(nblk0, nq_tile), 32 * SEL_WARPS, smem, st>>>(
        scores, nullptr, n, stride, kp, ix->keys0.as<uint64_t>());
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
  if (dense && kp <= 32) {   // merge + fp64 re-score + rank in one launch
    ProfScope prof_fin(ctx, PROF_OTHER, st);
    dense_finish_kernel<<<nq_tile, 32 * FIN_WARPS, 0, st>>>(ix->keys0.as<uint64_t>(), static_cast<int64_t>(nblk0) * kp,
                                                            kp, k, ix->rows.as<float>(), ix->dim,
                                                            ix->norm64.as<double>(), queries_dev, ix->id_base, ids_out,
                                                            s32_out, s64_out);
    VRAG_CUDA(cudaGetLastError());
    ctx->launches++;
    return;
  }
  const uint64_t* final_keys = ix->keys0.as<uint64_t>();
  if (nblk0 > 1) {
    if (kp <= 32)
      select_reg_kernel<false><<<dim3(1, nq_tile), 32 * SEL_WARPS, 0, st>>>(
          nullptr, ix->keys0.as<uint64_t>(), static_cast<int64_t>(nblk0) * kp, static_cast<int64_t>(nblk0) * kp, kp,
          ix->keys1.as<uint64_t>());
    else
      select_kernel<false><<<dim3(1, nq_tile), 32 * SEL_WARPS, smem, st>>>(
          nullptr, ix->keys0.as<uint64_t>(), static_cast<int64_t>(nblk0) * kp, static_cast<int64_t>(nblk0) * kp, kp,
          ix->keys1.as<uint64_t>());
    VRAG_CUDA(cudaGetLastError());
    ctx->launches++;
    final_keys = ix->keys1.as<uint64_t>();
  }
  ix->s64.reserve(static_cast<size_t>(nq_tile) * kp * 8);
  ix->crow.reserve(static_cast<size_t>(nq_tile) * kp * 8);
  dim3 g((kp + 7) / 8, nq_tile);
  if (dense)
    dense_rescore_kernel<<<g, 256, 0, st>>>(ix->rows.as<float>(), ix->dim, ix->norm64.as<double>(),
                                                      queries_dev, final_keys, kp, ix->s64.as<double>(),
                                                      ix->crow.as<int64_t>());
  else
    sparse_rescore_kernel<<<g, 256, 0, st>>>(ix->indptr.as<int64_t>(), ix->indices.as<int32_t>(),
                                                       ix->values.as<float>(), ix->qT.as<float>(),
                                                       static_cast<size_t>(ix->dim) * SQT, -1, final_keys, kp,
                                                       ix->s64.as<double>(), ix->crow.as<int64_t>());
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
  rank_kernel<<<nq_tile, 256, 0, st>>>(ix->s64.as<double>(), ix->crow.as<int64_t>(), kp, k, ix->id_base,
                                                ids_out, s32_out, s64_out);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
}

void fill_empty(vrag_ctx* ctx, int nq, int k, int64_t* ids, float* s32, double* s64, bool on_device) {
  std::vector<int64_t> hi(static_cast<size_t>(nq) * k, -1);
  std::vector<float> hs(static_cast<size_t>(nq) * k, -INFINITY);
  std::vector<double> hd(static_cast<size_t>(nq) * k, -INFINITY);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost;
  VRAG_CUDA(cudaMemcpy(ids, hi.data(), hi.size() * 8, kind));
  VRAG_CUDA(cudaMemcpy(s32, hs.data(), hs.size() * 4, kind));
  if (s64) VRAG_CUDA(cudaMemcpy(s64, hd.data(), hd.size() * 8, kind));
  (void)ctx;
}

}  // namespace

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

extern "C" int vrag_index_create(vrag_ctx* ctx, int kind, int dim, vrag_index** out) {
  if (!ctx || !out) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(ctx)
  VRAG_CHECK(kind == VRAG_INDEX_DENSE_COSINE || kind == VRAG_INDEX_SPARSE_IP, VRAG_ERR_ARG, "index_create: bad kind");
  VRAG_CHECK(dim > 0 && (kind == VRAG_INDEX_SPARSE_IP || dim % 4 == 0), VRAG_ERR_ARG,
             "index_create: dense dim must be a positive multiple of 4");
  vrag_index* ix = new vrag_index();
  ix->ctx = ctx;
  ix->kind = kind;
  ix->dim = dim;
  if (kind == VRAG_INDEX_SPARSE_IP) {
    ix->indptr.reserve(8);
    VRAG_CUDA(cudaMemsetAsync(ix->indptr.p, 0, 8, ctx->stream));
    ix->qT.reserve(static_cast<size_t>(dim) * SQT * 4);
    VRAG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *out = ix;
  VRAG_API_END()
}

extern "C" void vrag_index_destroy(vrag_index* idx) {
  if (!idx) return;
  cudaSetDevice(idx->ctx->device);
  cudaStreamSynchronize(idx->ctx->stream);
  delete idx;
}

extern "C" int64_t vrag_index_size(vrag_index* idx) { return idx ? idx->n : -1; }

extern "C" int vrag_index_set_id_base(vrag_index* idx, int64_t base) {
  if (!idx) return VRAG_ERR_ARG;
  idx->id_base = base;
  return VRAG_OK;
}

extern "C" int vrag_index_add_dense(vrag_index* idx, const float* rows, int64_t n, int on_device) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  VRAG_CHECK(idx->kind == VRAG_INDEX_DENSE_COSINE, VRAG_ERR_ARG, "add_dense on a sparse index");
  VRAG_CHECK(n >= 0 && (rows || n == 0), VRAG_ERR_ARG, "add_dense: null rows");
  if (n == 0) return VRAG_OK;
  idx->filter_on = false;   // a filter mask is row-aligned with the index it was built for
  VRAG_CHECK(idx->n + n < (1LL << 32) - 1, VRAG_ERR_ARG, "add_dense: more than 2^32-2 rows per shard");
  const size_t rb = static_cast<size_t>(idx->dim) * 4;
  grow(_ctx, idx->rows, idx->n * rb, (idx->n + n) * rb);
  grow(_ctx, idx->norm64, idx->n * 8, (idx->n + n) * 8);
  grow(_ctx, idx->inv32, idx->n * 4, (idx->n + n) * 4);
  grow(_ctx, idx->deleted, idx->n, idx->n + n);
  float* dst = idx->rows.as<float>() + idx->n * idx->dim;
  VRAG_CUDA(cudaMemcpyAsync(dst, rows, n * rb, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, _ctx->stream));
  VRAG_CUDA(cudaMemsetAsync(idx->deleted.as<uint8_t>() + idx->n, 0, n, _ctx->stream));
  dense_norm_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, _ctx->stream>>>(
      dst, n, idx->dim, idx->norm64.as<double>() + idx->n, idx->inv32.as<float>() + idx->n);
  VRAG_CUDA(cudaGetLastError());
  _ctx->launches++;
  VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  idx->n += n;
  VRAG_API_END()
}

extern "C" int vrag_index_add_sparse(vrag_index* idx, const int64_t* indptr, const int32_t* indices,
                                     const float* values, int64_t n) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  VRAG_CHECK(idx->kind == VRAG_INDEX_SPARSE_IP, VRAG_ERR_ARG, "add_sparse on a dense index");
  VRAG_CHECK(n >= 0 && (indptr || n == 0), VRAG_ERR_ARG, "add_sparse: null indptr");
  if (n == 0) return VRAG_OK;
  idx->filter_on = false;
  VRAG_CHECK(idx->n + n < (1LL << 32) - 1, VRAG_ERR_ARG, "add_sparse: more than 2^32-2 rows per shard");
  const int64_t add_nnz = indptr[n] - indptr[0];
  VRAG_CHECK(add_nnz >= 0 && (add_nnz == 0 || (indices && values)), VRAG_ERR_ARG, "add_sparse: bad CSR");
  for (int64_t j = 0; j < add_nnz; ++j)
    VRAG_CHECK(indices[indptr[0] + j] >= 0 && indices[indptr[0] + j] < idx->dim, VRAG_ERR_ARG,
               "add_sparse: term id outside [0, dim)");
  std::vector<int64_t> ip(n);
  for (int64_t i = 0; i < n; ++i) {
    VRAG_CHECK(indptr[i + 1] >= indptr[i], VRAG_ERR_ARG, "add_sparse: indptr not monotone");
    ip[i] = idx->nnz + (indptr[i + 1] - indptr[0]);
  }
  grow(_ctx, idx->indptr, (idx->n + 1) * 8, (idx->n + n + 1) * 8);
  grow(_ctx, idx->indices, idx->nnz * 4, (idx->nnz + add_nnz) * 4);
  grow(_ctx, idx->values, idx->nnz * 4, (idx->nnz + add_nnz) * 4);
  grow(_ctx, idx->deleted, idx->n, idx->n + n);
  VRAG_CUDA(cudaMemcpyAsync(idx->indptr.as<int64_t>() + idx->n + 1, ip.data(), n * 8, cudaMemcpyHostToDevice, _ctx->stream));
  if (add_nnz) {
    VRAG_CUDA(cudaMemcpyAsync(idx->indices.as<int32_t>() + idx->nnz, indices + indptr[0], add_nnz * 4, cudaMemcpyHostToDevice, _ctx->stream));
    VRAG_CUDA(cudaMemcpyAsync(idx->values.as<float>() + idx->nnz, values + indptr[0], add_nnz * 4, cudaMemcpyHostToDevice, _ctx->stream));
  }
  VRAG_CUDA(cudaMemsetAsync(idx->deleted.as<uint8_t>() + idx->n, 0, n, _ctx->stream));
  VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  idx->n += n;
  idx->nnz += add_nnz;
  VRAG_API_END()
}

extern "C" int vrag_index_mark_deleted(vrag_index* idx, const int64_t* rows, int64_t n) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  VRAG_CHECK(n >= 0 && (rows || n == 0), VRAG_ERR_ARG, "mark_deleted: null rows");
  idx->filter_on = false;
  const uint8_t one = 1;
  for (int64_t i = 0; i < n; ++i) {
    VRAG_CHECK(rows[i] >= 0 && rows[i] < idx->n, VRAG_ERR_ARG, "mark_deleted: row out of range");
    VRAG_CUDA(cudaMemcpyAsync(idx->deleted.as<uint8_t>() + rows[i], &one, 1, cudaMemcpyHostToDevice, _ctx->stream));
  }
  VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  VRAG_API_END()
}

namespace {
__global__ void or_mask_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                               int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a[i] | b[i]) ? 1 : 0;
}
}  // namespace

// Metadata-filter pushdown (reference: the `filter=` expression of BaseMilvusStore.query,
// milvus_base.py:189-259).  The host evaluates the boolean expression over its payload columns once and hands the scan
// a row mask: rows with exclude[row] != 0 score -inf in every following search of this index, exactly like deleted
// rows, until the filter is cleared (exclude == NULL) or the index changes (add / delete).
extern "C" int vrag_index_set_filter(vrag_index* idx, const uint8_t* exclude, int64_t n) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  if (!exclude) {
    idx->filter_on = false;
    return VRAG_OK;
  }
  VRAG_CHECK(n == idx->n, VRAG_ERR_ARG, "set_filter: mask length must equal the number of rows in the index");
  if (n == 0) return VRAG_OK;
  idx->excl.reserve(static_cast<size_t>(n));
  idx->masked.reserve(static_cast<size_t>(n));
  VRAG_CUDA(cudaMemcpyAsync(idx->excl.p, exclude, static_cast<size_t>(n), cudaMemcpyHostToDevice, _ctx->stream));
  or_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, _ctx->stream>>>(
      idx->deleted.as<uint8_t>(), idx->excl.as<uint8_t>(), idx->masked.as<uint8_t>(), n);
  VRAG_CUDA(cudaGetLastError());
  _ctx->launches++;
  VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));   // `exclude` is the caller's buffer
  idx->filter_on = true;
  VRAG_API_END()
}

extern "C" int vrag_index_search_dense(vrag_index* idx, const float* queries, int nq, int k, int64_t* ids_out,
                                       float* scores_out, double* scores64_out, int on_device) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  VRAG_CHECK(idx->kind == VRAG_INDEX_DENSE_COSINE, VRAG_ERR_ARG, "search_dense on a sparse index");
  VRAG_CHECK(nq >= 0 && k > 0 && k <= MAX_K, VRAG_ERR_ARG, "search_dense: need nq >= 0 and 1 <= k <= 1024");
  VRAG_CHECK(nq == 0 || (queries && ids_out && scores_out), VRAG_ERR_ARG, "search_dense: null argument");
  if (nq == 0) return VRAG_OK;
  if (idx->n == 0) { fill_empty(_ctx, nq, k, ids_out, scores_out, scores64_out, on_device); return VRAG_OK; }
  const int dim = idx->dim;
  const int64_t n = idx->n;
  const float* qd = queries;
  if (!on_device) {
    idx->qdev.reserve(static_cast<size_t>(nq) * dim * 4);
    VRAG_CUDA(cudaMemcpyAsync(idx->qdev.p, queries, static_cast<size_t>(nq) * dim * 4, cudaMemcpyHostToDevice, _ctx->stream));
    qd = idx->qdev.as<float>();
    idx->out_ids.reserve(static_cast<size_t>(nq) * k * 8);
    idx->out_s32.reserve(static_cast<size_t>(nq) * k * 4);
    idx->out_s64.reserve(static_cast<size_t>(nq) * k * 8);
  }
  int64_t* d_ids = on_device ? ids_out : idx->out_ids.as<int64_t>();
  float* d_s32 = on_device ? scores_out : idx->out_s32.as<float>();
  double* d_s64 = on_device ? scores64_out : idx->out_s64.as<double>();
  idx->qnorm.reserve(static_cast<size_t>(nq) * 4);
  query_norm_kernel<<<nq, 32, 0, _ctx->stream>>>(qd, dim, idx->qnorm.as<float>());
  VRAG_CUDA(cudaGetLastError());
  _ctx->launches++;
  // ---- batched search, tensor-bound regime ("all 1 k queries in one pass"): >= BIG_MIN_Q queries -------
  // scores = Q X_n^T as ONE split-precision GEMM per group of <= 1024 queries (gemm.cu, three tcgen05.mma per product
  // over fp16 hi / lo planes: >= 21 significant bits, at the fp16 tensor rate instead of three tf32 passes).  X_n are
  // the normalised corpus rows, kept as planes beside the fp32 originals; the fp32 rows still feed the fp64 re-scoring
  // of the k + 16 candidates, so ids and scores are the same as on the scan paths.  Per 1024 queries the corpus planes
  // (= N dim 4 bytes, as much as one fp32 pass) are read 4 times instead of 64.
  const char* big_env = getenv("VRAG_SCAN_BIG_MIN");   // debug: 0 disables the GEMM path, else its smallest batch
  const int big_min = big_env ? atoi(big_env) : BIG_MIN_Q;
  if (big_min > 0 && nq >= big_min && dim % GEMM_BK == 0 && n < (int64_t(1) << 31) - GEMM_BN) {
    const int64_t n_pad = (n + GEMM_BN - 1) / GEMM_BN * GEMM_BN;
    if (idx->planes_cap < n_pad) {   // (re)allocate with head room, zero the padding rows, rebuild from row 0
      const int64_t cap = std::max<int64_t>(n_pad, idx->planes_cap + idx->planes_cap / 2) / GEMM_BN * GEMM_BN + GEMM_BN;
      idx->rows_hi.release();
      idx->rows_lo.release();
      idx->rows_hi.reserve(static_cast<size_t>(cap) * dim * 2);
      idx->rows_lo.reserve(static_cast<size_t>(cap) * dim * 2);
      VRAG_CUDA(cudaMemsetAsync(idx->rows_hi.p, 0, static_cast<size_t>(cap) * dim * 2, _ctx->stream));
      VRAG_CUDA(cudaMemsetAsync(idx->rows_lo.p, 0, static_cast<size_t>(cap) * dim * 2, _ctx->stream));
      idx->planes_cap = cap;
      idx->planes_n = 0;
    }
    if (idx->planes_n < n) {
      const int64_t cnt4 = (n - idx->planes_n) * (dim / 4);
      normalize_split_kernel<<<static_cast<unsigned>((cnt4 + 255) / 256), 256, 0, _ctx->stream>>>(
          idx->rows.as<float>(), idx->inv32.as<float>(), idx->planes_n, n, dim, idx->rows_hi.as<__half>(),
          idx->rows_lo.as<__half>());
      VRAG_CUDA(cudaGetLastError());
      _ctx->launches++;
      idx->planes_n = n;
    }
    idx->q_hi.reserve(static_cast<size_t>(nq) * dim * 2);
    idx->q_lo.reserve(static_cast<size_t>(nq) * dim * 2);
    const int64_t q4 = static_cast<int64_t>(nq) * (dim / 4);
    normalize_split_kernel<<<static_cast<unsigned>((q4 + 255) / 256), 256, 0, _ctx->stream>>>(
        qd, nullptr, 0, nq, dim, idx->q_hi.as<__half>(), idx->q_lo.as<__half>());
    VRAG_CUDA(cudaGetLastError());
    _ctx->launches++;
    auto full_scores = [&](int q0, int nt, int64_t rows, int64_t ld) {   // scores[nt][ld] of corpus rows [0, rows)
      GemmEpiParams gp;
      gp.M = nt;
      gp.out32 = idx->scores.as<float>();
      gp.ld32 = static_cast<int>(ld);
      gp.n_valid = static_cast<int>(rows);
      gp.col_mask = idx->skip();
      gp.a_lo = idx->q_lo.as<__half>() + static_cast<size_t>(q0) * dim;
      gp.w_lo = idx->rows_lo.as<__half>();
      gp.prof_class = PROF_SCAN;
      launch_gemm(_ctx, EPI_SCORES, idx->q_hi.as<__half>() + static_cast<size_t>(q0) * dim, idx->rows_hi.as<__half>(), nt,
                  static_cast<int>(ld), dim, gp, 0);
    };
    // Selection fused into the GEMM epilogue (k + 16 <= 32, corpus large enough to sample): phase 1 scores a SAMPLE (the
    // first n/16 rows) in full and selects its k' best per query -- they seed the candidate list and the k'-th of them
    // is a threshold that at least k' rows are known to reach; phase 2 runs the GEMM over the remaining rows and its
    // epilogue appends only the scores >= that threshold (expected 16 k' per query) to the candidate list, so the
    // [nq, N] score matrix is neither written nor re-read; phase 3 merges, re-scores in fp64 and ranks (the same
    // dense_finish_kernel as every other path).  A candidate list that overflows (adversarial row order: the best rows
    // all after the sample) falls back to the full-score path for this call -- results never depend on the sample.
    const int kp = static_cast<int>(std::min<int64_t>(k + MARGIN, n));
    const char* fuse_env = getenv("VRAG_SCAN_FUSED_SELECT");   // debug: 0 keeps the full score matrix + streaming selection
    bool fused = !(fuse_env && fuse_env[0] == '0') && kp <= 32 && n >= 16 * GEMM_BN * 4;
    const int CAP = 4096;
    if (fused) {
      const int64_t S = std::max<int64_t>(GEMM_BN * 4, (n / 16) / GEMM_BN * GEMM_BN);   // sample rows (multiple of 256)
      idx->scores.reserve(static_cast<size_t>(std::min(1024, nq)) * S * 4);
      idx->cand.reserve(static_cast<size_t>(nq) * CAP * 8);
      idx->cand_count.reserve(static_cast<size_t>(nq) * 4);
      idx->cand_thr.reserve(static_cast<size_t>(nq) * 4);
      idx->flag.reserve(4);
      idx->keys1.reserve(static_cast<size_t>(nq) * kp * 8);
      VRAG_CUDA(cudaMemsetAsync(idx->cand.p, 0, static_cast<size_t>(nq) * CAP * 8, _ctx->stream));
      VRAG_CUDA(cudaMemsetAsync(idx->flag.p, 0, 4, _ctx->stream));
      for (int q0 = 0; q0 < nq; q0 += 1024) {
        const int nt = std::min(1024, nq - q0);
        full_scores(q0, nt, S, S);
        {   // k' best of the sample per query -> keys1[q0 + q][kp] (sorted descending)
          ProfScope prof(_ctx, PROF_SELECT);
          const int nblk = static_cast<int>(std::max<int64_t>(1, (S + SEL_REG_BLOCK - 1) / SEL_REG_BLOCK));
          idx->keys0.reserve(static_cast<size_t>(nt) * nblk * kp * 8);
          select_reg_kernel<true><<<dim3(nblk, nt), 32 * SEL_WARPS, 0, _ctx->stream>>>(idx->scores.as<float>(), nullptr, S, S, kp,
                                                                                    idx->keys0.as<uint64_t>());
          select_reg_kernel<false><<<dim3(1, nt), 32 * SEL_WARPS, 0, _ctx->stream>>>(
              nullptr, idx->keys0.as<uint64_t>(), static_cast<int64_t>(nblk) * kp, static_cast<int64_t>(nblk) * kp, kp,
              idx->keys1.as<uint64_t>() + static_cast<size_t>(q0) * kp);
          seed_candidates_kernel<<<nt, 32, 0, _ctx->stream>>>(
              idx->keys1.as<uint64_t>() + static_cast<size_t>(q0) * kp, kp, CAP,
              idx->cand.as<unsigned long long>() + static_cast<size_t>(q0) * CAP, idx->cand_count.as<int>() + q0,
              idx->cand_thr.as<float>() + q0);
          VRAG_CUDA(cudaGetLastError());
          _ctx->launches += 3;
        }
        GemmEpiParams gp;
        gp.M = nt;
        gp.n_valid = static_cast<int>(n - S);
        gp.col_mask = idx->skip() + S;
        gp.col_base = static_cast<int>(S);
        gp.thr = idx->cand_thr.as<float>() + q0;
        gp.cand = idx->cand.as<unsigned long long>() + static_cast<size_t>(q0) * CAP;
        gp.cand_count = idx->cand_count.as<int>() + q0;
        gp.cand_cap = CAP;
        gp.a_lo = idx->q_lo.as<__half>() + static_cast<size_t>(q0) * dim;
        gp.w_lo = idx->rows_lo.as<__half>() + static_cast<size_t>(S) * dim;
        gp.prof_class = PROF_SCAN;
        launch_gemm(_ctx, EPI_SCORES_THRESH, idx->q_hi.as<__half>() + static_cast<size_t>(q0) * dim,
                    idx->rows_hi.as<__half>() + static_cast<size_t>(S) * dim, nt, static_cast<int>(n_pad - S), dim, gp, 0);
        {
          ProfScope prof(_ctx, PROF_SELECT);
          overflow_flag_kernel<<<(nt + 255) / 256, 256, 0, _ctx->stream>>>(idx->cand_count.as<int>() + q0, nt, CAP,
                                                                         idx->flag.as<int>());
          dense_finish_kernel<<<nt, 32 * FIN_WARPS, 0, _ctx->stream>>>(
              idx->cand.as<uint64_t>() + static_cast<size_t>(q0) * CAP, CAP, kp, k, idx->rows.as<float>(), dim,
              idx->norm64.as<double>(), qd + static_cast<size_t>(q0) * dim, idx->id_base, d_ids + static_cast<size_t>(q0) * k,
              d_s32 + static_cast<size_t>(q0) * k, d_s64 ? d_s64 + static_cast<size_t>(q0) * k : nullptr);
          VRAG_CUDA(cudaGetLastError());
          _ctx->launches += 2;
        }
      }
      int overflow = 0;   // one 4-byte read-back per call decides whether the (rare) fallback runs
      VRAG_CUDA(cudaMemcpyAsync(&overflow, idx->flag.p, 4, cudaMemcpyDeviceToHost, _ctx->stream));
      VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
      fused = overflow == 0;
    }
    if (!fused) {
      const int group = static_cast<int>(std::max<int64_t>(GEMM_BM, std::min<int64_t>(1024, ((int64_t(4) << 30) / (n_pad * 4)) / GEMM_BM * GEMM_BM)));
      idx->scores.reserve(static_cast<size_t>(std::min(group, nq)) * n_pad * 4);
      for (int q0 = 0; q0 < nq; q0 += group) {
        const int nt = std::min(group, nq - q0);
        full_scores(q0, nt, n, n_pad);
        select_and_rank(idx, idx->scores.as<float>(), nt, k, true, qd + static_cast<size_t>(q0) * dim,
                        d_ids + static_cast<size_t>(q0) * k, d_s32 + static_cast<size_t>(q0) * k,
                        d_s64 ? d_s64 + static_cast<size_t>(q0) * k : nullptr, _ctx->stream, n_pad);
      }
    }
    if (!on_device) {
      VRAG_CUDA(cudaMemcpyAsync(ids_out, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
      VRAG_CUDA(cudaMemcpyAsync(scores_out, d_s32, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, _ctx->stream));
      if (scores64_out)
        VRAG_CUDA(cudaMemcpyAsync(scores64_out, d_s64, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
      VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
    }
    return VRAG_OK;
  }
  // TC_MIN_Q or more queries left: 16 per corpus pass on the tensor cores (the split query block must fit in smem);
  // fewer: the FMA scan, which is faster for 1..4 queries (measured: profiles/README.md)
  const char* tc_env = getenv("VRAG_SCAN_TC_MIN");  // debug: 0 forces the FMA path, else the smallest tile for TC
  const int tc_min = tc_env ? atoi(tc_env) : TC_MIN_Q;
  const bool tc_ok = tc_min > 0 && dim % 32 == 0 && dim <= 768 && n < (int64_t(1) << 31) - TC_ROWS;
  // (Tried: running the selection of tile t on a low-priority side stream under the scan of tile t + 1.  The scan is
  // a finely balanced HBM-bound pipeline with one CTA per SM; co-resident selection blocks slowed it by more than the
  // selection costs when serialised -- 42.7 ms vs 41.0 ms per 1000 queries over 1 M rows -- so tiles run back to back.)
  // Selection groups: the score rows of up to `group_max` consecutive full tensor-core tiles (16 queries each) are kept
  // and selected by ONE select + finish launch pair.  The selection is latency-bound (8 dependent load rounds per warp
  // list, then a one-block-per-query finish), so its cost per launch barely grows with the number of queries, and on
  // small shards (125 k rows per GPU at 8 GPUs: 80 us of scan per tile) a per-tile selection would cost as much as the
  // scan.  Score buffer: group x 16 x n floats, capped at 512 MB.
  const size_t tile_scores = static_cast<size_t>(tc_ok ? TCQ : QT) * n;
  const int group_max = tc_ok ? static_cast<int>(std::max<size_t>(1, std::min<size_t>(8, (size_t(512) << 20) / (tile_scores * 4)))) : 1;
  idx->scores.reserve(static_cast<size_t>(group_max) * tile_scores * 4);
  const int grid = static_cast<int>(std::min<int64_t>((n + 31) / 32, static_cast<int64_t>(_ctx->num_sms) * 2));
  int g_tiles = 0, g_q0 = 0;   // tiles / first query of the open selection group
  for (int q0 = 0; q0 < nq;) {
    const bool use_tc = tc_ok && nq - q0 >= tc_min;
    const int nt = std::min(use_tc ? TCQ : QT, nq - q0);
    const float* qt = qd + static_cast<size_t>(q0) * dim;
    const float* qn = idx->qnorm.as<float>() + q0;
    float* scores = idx->scores.as<float>() + static_cast<size_t>(g_tiles) * tile_scores;
    if (use_tc) {
      ProfScope prof(_ctx, PROF_SCAN);
      const CUtensorMap tmX = make_tmap_2d(_ctx, idx->rows.as<float>(), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                           static_cast<uint64_t>(n), dim, dim, TC_ROWS, 32);
      const int smem = TC_STAGES * TC_TILE_BYTES + (dim / 32) * TC_BTILE_BYTES + 256 + 1024;
      static std::atomic<uint64_t> attr{0};   // cudaFuncSetAttribute is per device
      const uint64_t dev_bit = 1ull << (_ctx->device & 63);
      if ((attr.fetch_or(dev_bit) & dev_bit) == 0)
        VRAG_CUDA(cudaFuncSetAttribute(dense_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      const int tgrid = static_cast<int>(std::min<int64_t>((n + TC_ROWS - 1) / TC_ROWS, _ctx->num_sms));
      dense_scan_tc_kernel<<<tgrid, TC_THREADS, smem, _ctx->stream>>>(tmX, n, dim, qt, nt, idx->inv32.as<float>(), qn,
                                                                      idx->skip(), scores);
      VRAG_CUDA(cudaGetLastError());
      _ctx->launches++;
    } else {
      ProfScope prof(_ctx, PROF_SCAN);
      // two 32-row stages + the query block must fit in 227 KB of shared memory (dim 1024 does not: generic path)
      if ((dim == 768 || dim == 384) || (dim == 1024 && 2 * SCAN_ROWS * dim * 4 + 8 * dim * 4 < 220 * 1024)) {
        const int nqt = nt <= 1 ? 1 : (nt <= 2 ? 2 : (nt <= 4 ? 4 : 8));
        const int stage_bytes = SCAN_ROWS * dim * 4;
        const int q_bytes = nqt * dim * 4;
        const int nstage = std::max(2, std::min(4, (210 * 1024 - q_bytes) / stage_bytes));
        const int smem = nstage * stage_bytes + q_bytes + 2 * nstage * 8 + 16;
        const int tgrid = static_cast<int>(std::min<int64_t>((n + SCAN_ROWS - 1) / SCAN_ROWS, _ctx->num_sms));
#define VRAG_SCAN_T(V, Q)                                                                                          \
  do {                                                                                                             \
    VRAG_CUDA(cudaFuncSetAttribute(dense_scan_tma_kernel<V, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    dense_scan_tma_kernel<V, Q><<<tgrid, 32 * (SCAN_CONSUMER_WARPS + 1), smem, _ctx->stream>>>(                     \
        idx->rows.as<float>(), n, qt, idx->inv32.as<float>(), qn, idx->skip(),                                      \
        scores, nstage, nt);                                                                      \
  } while (0)
#define VRAG_SCAN_Q(V)                                \
  do {                                                \
    if (nqt == 1) VRAG_SCAN_T(V, 1);                  \
    else if (nqt == 2) VRAG_SCAN_T(V, 2);             \
    else if (nqt == 4) VRAG_SCAN_T(V, 4);             \
    else VRAG_SCAN_T(V, 8);                           \
  } while (0)
        if (dim == 768) VRAG_SCAN_Q(6);
        else if (dim == 384) VRAG_SCAN_Q(3);
        else VRAG_SCAN_Q(8);
#undef VRAG_SCAN_Q
#undef VRAG_SCAN_T
      } else {
        dense_scan_generic_kernel<<<grid, 256, 0, _ctx->stream>>>(idx->rows.as<float>(), n, dim, qt, nt,
                                                                  idx->inv32.as<float>(), qn, idx->skip(),
                                                                  scores);
      }
      VRAG_CUDA(cudaGetLastError());
      _ctx->launches++;
    }
    if (g_tiles == 0) g_q0 = q0;
    ++g_tiles;
    q0 += nt;
    // close the group: it is full, the queries are exhausted, this tile is partial (its rows would not be contiguous
    // with a following tile's), or the next tile takes the FMA path (different tile stride)
    const bool next_tc = tc_ok && nq - q0 >= tc_min;
    if (g_tiles == group_max || q0 >= nq || nt < TCQ || !use_tc || !next_tc) {
      select_and_rank(idx, idx->scores.as<float>(), q0 - g_q0, k, true, qd + static_cast<size_t>(g_q0) * dim,
                      d_ids + static_cast<size_t>(g_q0) * k, d_s32 + static_cast<size_t>(g_q0) * k,
                      d_s64 ? d_s64 + static_cast<size_t>(g_q0) * k : nullptr, _ctx->stream);
      g_tiles = 0;
    }
  }
  if (!on_device) {
    VRAG_CUDA(cudaMemcpyAsync(ids_out, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
    VRAG_CUDA(cudaMemcpyAsync(scores_out, d_s32, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, _ctx->stream));
    if (scores64_out)
      VRAG_CUDA(cudaMemcpyAsync(scores64_out, d_s64, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
    VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  }
  VRAG_API_END()
}

extern "C" int vrag_index_search_sparse(vrag_index* idx, const int64_t* q_indptr, const int32_t* q_indices,
                                        const float* q_values, int nq, int k, int64_t* ids_out, float* scores_out,
                                        double* scores64_out) {
  if (!idx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(idx->ctx)
  VRAG_CHECK(idx->kind == VRAG_INDEX_SPARSE_IP, VRAG_ERR_ARG, "search_sparse on a dense index");
  VRAG_CHECK(nq >= 0 && k > 0 && k <= MAX_K, VRAG_ERR_ARG, "search_sparse: need nq >= 0 and 1 <= k <= 1024");
  VRAG_CHECK(nq == 0 || (q_indptr && ids_out && scores_out), VRAG_ERR_ARG, "search_sparse: null argument");
  if (nq == 0) return VRAG_OK;
  if (idx->n == 0) { fill_empty(_ctx, nq, k, ids_out, scores_out, scores64_out, false); return VRAG_OK; }
  const int64_t n = idx->n;
  const int64_t qnnz = q_indptr[nq] - q_indptr[0];
  std::vector<int64_t> ip(nq + 1);
  for (int i = 0; i <= nq; ++i) ip[i] = q_indptr[i] - q_indptr[0];
  idx->qip.reserve((static_cast<size_t>(nq) + 1) * 8);
  idx->qidx.reserve(std::max<size_t>(4, static_cast<size_t>(qnnz) * 4));
  idx->qval.reserve(std::max<size_t>(4, static_cast<size_t>(qnnz) * 4));
  VRAG_CUDA(cudaMemcpyAsync(idx->qip.p, ip.data(), (static_cast<size_t>(nq) + 1) * 8, cudaMemcpyHostToDevice, _ctx->stream));
  if (qnnz) {
    VRAG_CUDA(cudaMemcpyAsync(idx->qidx.p, q_indices + q_indptr[0], static_cast<size_t>(qnnz) * 4, cudaMemcpyHostToDevice, _ctx->stream));
    VRAG_CUDA(cudaMemcpyAsync(idx->qval.p, q_values + q_indptr[0], static_cast<size_t>(qnnz) * 4, cudaMemcpyHostToDevice, _ctx->stream));
  }
  idx->out_ids.reserve(static_cast<size_t>(nq) * k * 8);
  idx->out_s32.reserve(static_cast<size_t>(nq) * k * 4);
  idx->out_s64.reserve(static_cast<size_t>(nq) * k * 8);
  // Selection groups (as on the dense path): the score rows of up to `group` consecutive query tiles (SQT queries per corpus
  // pass) are kept, each tile with its own dense query table, and selected / re-scored / ranked by ONE launch sequence.
  // The selection is latency-bound; per tile it cost twice the scan on a 10 k-document corpus.
  const size_t qT_bytes = static_cast<size_t>(idx->dim) * SQT * 4;
  const int group = static_cast<int>(std::max<size_t>(1, std::min<size_t>(16, (size_t(512) << 20) / (static_cast<size_t>(SQT) * n * 4))));
  idx->scores.reserve(static_cast<size_t>(group) * SQT * n * 4);
  idx->qT.reserve(static_cast<size_t>(group) * qT_bytes);
  const int grid = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(_ctx->num_sms) * 8));
  for (int g0 = 0; g0 < nq; g0 += group * SQT) {
    const int ng = std::min(group * SQT, nq - g0);   // queries in this group
    VRAG_CUDA(cudaMemsetAsync(idx->qT.p, 0, static_cast<size_t>((ng + SQT - 1) / SQT) * qT_bytes, _ctx->stream));
    const int ntiles = (ng + SQT - 1) / SQT;
    sparse_scatter_query_kernel<<<ng, 128, 0, _ctx->stream>>>(idx->qip.as<int64_t>() + g0, idx->qidx.as<int32_t>(),
                                                               idx->qval.as<float>(), ng, idx->dim, idx->qT.as<float>());
    VRAG_CUDA(cudaGetLastError());
    _ctx->launches++;
    {
      ProfScope prof(_ctx, PROF_SCAN);
      sparse_scan_kernel<<<dim3(grid, ntiles), 256, 0, _ctx->stream>>>(
          idx->indptr.as<int64_t>(), idx->indices.as<int32_t>(), idx->values.as<float>(), n, idx->qT.as<float>(),
          idx->dim, ng, idx->skip(), idx->scores.as<float>());
      VRAG_CUDA(cudaGetLastError());
      _ctx->launches++;
    }
    select_and_rank(idx, idx->scores.as<float>(), ng, k, false, nullptr,
                    idx->out_ids.as<int64_t>() + static_cast<size_t>(g0) * k,
                    idx->out_s32.as<float>() + static_cast<size_t>(g0) * k,
                    idx->out_s64.as<double>() + static_cast<size_t>(g0) * k, _ctx->stream);
  }
  VRAG_CUDA(cudaMemcpyAsync(ids_out, idx->out_ids.p, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
  VRAG_CUDA(cudaMemcpyAsync(scores_out, idx->out_s32.p, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, _ctx->stream));
  if (scores64_out)
    VRAG_CUDA(cudaMemcpyAsync(scores64_out, idx->out_s64.p, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, _ctx->stream));
  VRAG_CUDA(cudaStreamSynchronize(_ctx->stream));
  VRAG_API_END()
}

extern "C" int vrag_topk_merge(vrag_ctx* ctx, const double* scores64, const int64_t* ids, int nq, int m, int k,
                               int64_t* ids_out, float* scores_out, double* scores64_out) {
  if (!ctx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(ctx)
  VRAG_CHECK(nq >= 0 && m > 0 && k > 0 && scores64 && ids && ids_out && scores_out, VRAG_ERR_ARG, "topk_merge: bad argument");
  if (nq == 0) return VRAG_OK;
  rank_kernel<<<nq, 256, 0, ctx->stream>>>(scores64, ids, m, k, 0, ids_out, scores_out, scores64_out);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
  VRAG_API_END()
}

extern "C" int vrag_topk_publish(vrag_ctx* ctx, const double* scores64, const int64_t* ids, int nq, int k,
                                 void* const* peer_bufs, int world, int rank) {
  if (!ctx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(ctx)
  VRAG_CHECK(nq >= 0 && k > 0 && scores64 && ids && peer_bufs && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world,
             VRAG_ERR_ARG, "topk_publish: bad argument (at most 16 peers)");
  if (nq == 0) return VRAG_OK;
  PeerPtrs pp;
  for (int i = 0; i < world; ++i) {
    VRAG_CHECK(peer_bufs[i] != nullptr, VRAG_ERR_ARG, "topk_publish: null peer buffer");
    pp.p[i] = reinterpret_cast<unsigned long long>(peer_bufs[i]);
  }
  const int64_t n_rec = static_cast<int64_t>(nq) * k;
  const int bx = static_cast<int>(std::min<int64_t>((n_rec + 255) / 256, 32));
  topk_publish_kernel<<<dim3(bx, world), 256, 0, ctx->stream>>>(scores64, ids, n_rec, pp, rank);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
  VRAG_API_END()
}

extern "C" int vrag_topk_merge_packed(vrag_ctx* ctx, const int64_t* packed, int world, int nq, int k, int64_t* ids_out,
                                      float* scores_out, double* scores64_out) {
  if (!ctx) return VRAG_ERR_ARG;
  VRAG_API_BEGIN(ctx)
  VRAG_CHECK(nq >= 0 && k > 0 && k <= MAX_K && world >= 1 && world <= MAX_PEERS && packed && ids_out && scores_out,
             VRAG_ERR_ARG, "topk_merge_packed: bad argument");
  if (nq == 0) return VRAG_OK;
  const size_t smem = static_cast<size_t>(world) * k * sizeof(longlong2);
  VRAG_CHECK(smem <= 48 * 1024, VRAG_ERR_ARG, "topk_merge_packed: world * k too large for the merge kernel");
  rank_packed_kernel<<<nq, 256, smem, ctx->stream>>>(reinterpret_cast<const longlong2*>(packed), world, nq, k, ids_out,
                                                     scores_out, scores64_out);
  VRAG_CUDA(cudaGetLastError());
  ctx->launches++;
  VRAG_API_END()
}
