// EXPERIMENTAL, opt-in (cosyb200_set_option "tc_tma" = 1; default 0): the tensor-core 1x1 kernel of kernels_tc.cuh
// with its raw A stages fed by TMA instead of cp.async.
//
// One cp.async.bulk.tensor.2d box (32 x 128 floats, 128-byte swizzle, zero fill outside the tensor) per k-stage,
// issued by the loader warp, replaces 8 cp.async per producer thread + wait_group + named barrier: ~1.1k of the
// ~3k cycles of the producer chain, 5.79 instead of 6.17 ms per trunk forward of 64 hypotheses (B200, round 1).
// It is NOT the default because 1 trunk forward in ~30 then has a few rows of one tile wrong (DESIGN.md section 8,
// item 1; reproducer: tools/dbg_determinism.py with the option set).  The `dbg` bits switch the experiments that
// are left:  1 proxy fence before the first read of a slot, 2 named barrier among the producers, 4 one box in
// flight per group, 8 slot handed back only after the TMEM store, 16 prefetch.tensormap.
// Everything except the source of the raw rows is identical to tc::k_pw_gemm_tc (same weights image, same tiles,
// same epilogue), so outputs are bit-identical to the default kernel whenever the race does not strike.
#pragma once
#include <cuda.h>

#include "kernels_tc.cuh"

namespace cosyb {
namespace tc {

constexpr int XRAW_ROW_BYTES = BK * 4;              // 128: one swizzle-128B row of the TMA box
constexpr int XRAW_STAGE_BYTES = BM * XRAW_ROW_BYTES;
__host__ __device__ constexpr int xraw_bytes(int ng) { return ng * RAW_DEPTH * XRAW_STAGE_BYTES; }

// 2-D tiled TMA load (box = BK x BM floats, 128-byte swizzle) completing on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// NG = producer warpgroups: 1 -> 10 warps, two CTAs per SM; 2 -> 14 warps, one CTA per SM, the two groups fill
// alternate k-stages (for layers with fewer tiles than CTA slots, where a second CTA per SM would sit empty).
template <int BN_MAX, bool GATE, bool SWISH, bool RESID, int NG>
__global__ void __launch_bounds__(threads_for(NG), NG == 1 ? 2 : 1)
k_pw_gemm_tc_tma(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ A, const float* __restrict__ Wpk, const float* __restrict__ bias,
             const float* __restrict__ gate, const float* __restrict__ resid, float* __restrict__ C, int M, int N,
             int K, int rows_per_img, int bn, int n_tiles, int nb, int resident, int dbg) {
  constexpr int N_GROUPS = NG;
  constexpr int MMA_WARP = 4 * N_GROUPS;
  constexpr int DRAIN_WARP0 = MMA_WARP + 1;
  constexpr int LOADER_WARP = DRAIN_WARP0 + 4;
  constexpr int RAW_BYTES = xraw_bytes(NG);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + 4 + 2 * NG * RAW_DEPTH];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[64];          // bias of this CTA's n-tile (fixed for the CTA's lifetime)
  constexpr uint32_t TMEM_COLS = 256;               // two CTAs per SM share the 512 columns
  constexpr uint32_t A_COL0 = 2 * N_PASS * BN_MAX;
  static_assert(A_COL0 + N_ASLOTS * A_SLOT_COLS <= TMEM_COLS, "TMEM budget");
  const uint32_t raw_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = raw_base + RAW_BYTES;     // B stage slots
  float* stg_base = reinterpret_cast<float*>(smem_raw + (raw_base - smem_u32(smem_raw)) + RAW_BYTES + nb * 2 * b_stage_bytes(bn));
  const int tid = threadIdx.x, lane = tid % 32;
  const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);   // tells the compiler the role branches are warp-uniform
  const int nk = (K + BK - 1) / BK;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tile = blockIdx.x % n_tiles, m_first = blockIdx.x / n_tiles, m_step = gridDim.x / n_tiles;
  const int my_tiles = m_first < m_tiles ? (m_tiles - 1 - m_first) / m_step + 1 : 0;
  const int n_items = my_tiles * nk;                // flat (m-tile, k-stage) work items of this CTA
  const uint32_t bsb = b_stage_bytes(bn);
  auto fullA = [&](int s) { return smem_u32(&bars[s]); };
  auto emptyA = [&](int s) { return smem_u32(&bars[N_ASLOTS + s]); };
  auto fullB = [&](int s) { return smem_u32(&bars[2 * N_ASLOTS + s]); };
  auto emptyB = [&](int s) { return smem_u32(&bars[2 * N_ASLOTS + MAX_BSLOTS + s]); };
  auto acc_full = [&](int b) { return smem_u32(&bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + b]); };
  auto acc_empty = [&](int b) { return smem_u32(&bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + 2 + b]); };
  auto fullRaw = [&](int grp, int rs) { return smem_u32(&bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + 4 + grp * RAW_DEPTH + rs]); };
  auto emptyRaw = [&](int grp, int rs) {
    return smem_u32(&bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + 4 + NG * RAW_DEPTH + grp * RAW_DEPTH + rs]);
  };

  if (tid == 0) trace(0);
  if (tid == 0) {
    for (int s = 0; s < N_ASLOTS; ++s) {
      mbar_init(fullA(s), PRODUCER_THREADS);
      mbar_init(emptyA(s), 1);
    }
    for (int s = 0; s < MAX_BSLOTS; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), DRAIN_THREADS);
    }
    for (int gq = 0; gq < NG; ++gq)
      for (int rs = 0; rs < RAW_DEPTH; ++rs) {
        mbar_init(fullRaw(gq, rs), 1);
        mbar_init(emptyRaw(gq, rs), PRODUCER_THREADS);
      }
    fence_barrier_init();
  }
  if (tid < 64) {
    const int n = (blockIdx.x % n_tiles) * bn + tid;
    s_bias[tid] = (tid < bn && n < N) ? __ldg(bias + n) : 0.f;
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
  if (tid == 0) trace(1);

  if (warp < MMA_WARP) {
    // ------------------------------------------------------------------ producers (A operand)
    // Raw fp32 rows arrive by TMA (one 32 x 128 box per k-stage, 128-byte swizzle, issued by the loader warp
    // RAW_DEPTH stages ahead; rows beyond M and columns beyond K come back as zeros); each thread reads ITS
    // tile row back (the swizzle spreads 8 consecutive rows over all banks), applies the SE gate, splits
    // hi/lo, hands the slot back and writes both halves to its TMEM lane.
    const int grp = warp / 4, q = warp % 4, tg = tid % PRODUCER_THREADS;
    const int row = q * 32 + lane;                  // tile row == TMEM lane written by this thread
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t ring = raw_base + grp * RAW_DEPTH * XRAW_STAGE_BYTES;
    const int n_mine = n_items > grp ? (n_items - 1 - grp) / N_GROUPS + 1 : 0;   // items g = grp + NG i
    for (int i = 0; i < n_mine; ++i) {
      const int g = grp + i * N_GROUPS;
      const int m = (m_first + (g / nk) * m_step) * BM + row, k0 = (g % nk) * BK;
      const int slot = g % N_ASLOTS;
      const int rs = i % RAW_DEPTH;
      const bool tr = tg == 0 && g == 6;
      mbar_wait_warp(fullRaw(grp, rs), (i / RAW_DEPTH) & 1);
      if (dbg & 1) fence_proxy_async();                                  // experiment: proxy fence before the first read
      if (dbg & 2) named_bar_sync(1 + grp, PRODUCER_THREADS);            // experiment: keep the group's warps in one stage
      const uint32_t src = ring + rs * XRAW_STAGE_BYTES + row * XRAW_ROW_BYTES;
      float v[BK];
#pragma unroll
      for (int c = 0; c < BK / 4; ++c)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[c * 4]), "=f"(v[c * 4 + 1]), "=f"(v[c * 4 + 2]), "=f"(v[c * 4 + 3])
                     : "r"(src + ((c ^ (row & 7)) << 4)) : "memory");
      if (GATE) {
        const float* gsrc = gate + (size_t)(min(m, M - 1) / rows_per_img) * K + k0;
#pragma unroll
        for (int c = 0; c < BK / 4; ++c) {
          if (k0 + c * 4 < K) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(gsrc + c * 4));
            v[c * 4] *= x.x; v[c * 4 + 1] *= x.y; v[c * 4 + 2] *= x.z; v[c * 4 + 3] *= x.w;
          }
        }
      }
      float lo[BK];
#pragma unroll
      for (int c = 0; c < BK; ++c) {
        const float h = tf32_rna(v[c]);
        lo[c] = tf32_rna(v[c] - h);
        v[c] = h;
      }
      if (tr) trace(12);
      if (!(dbg & 8)) mbar_arrive(emptyRaw(grp, rs));   // all values consumed: the slot may be refilled
      if (g >= N_ASLOTS) mbar_wait_warp(acc_full(slot), ((g / N_ASLOTS) - 1) & 1);   // MMAs of item g-2 done
      tc_fence_after();
      if (tr) trace(13);
      tmem_st32(t_lane + A_COL0 + slot * A_SLOT_COLS, v);
      tmem_st32(t_lane + A_COL0 + slot * A_SLOT_COLS + BK, lo);
      tmem_st_wait();
      tc_fence_before();
      if (tr) trace(14);
      mbar_arrive(fullA(slot));
      if (dbg & 8) mbar_arrive(emptyRaw(grp, rs));      // experiment: hand the slot back only after the TMEM store
      if (tr) trace(15);
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(bn);
    for (int g = 0; g < n_items; ++g) {
      const int s = g % nk, slot = g % N_ASLOTS, b = g & 1;
      const int bslot = resident ? s : g % nb;
      if (lane == 0 && g == 6) trace(4);
      if (g >= 2) mbar_wait_warp(acc_empty(b), ((g >> 1) - 1) & 1);
      if (!resident) mbar_wait_warp(fullB(bslot), (g / nb) & 1);
      else if (g < nk) mbar_wait_warp(fullB(bslot), 0);
      mbar_wait_warp(fullA(slot), (g / N_ASLOTS) & 1);
      tc_fence_after();
      if (lane == 0 && g == 6) trace(5);
      {
        const uint32_t elected = lane == 0;
        const uint32_t a_hi = tmem_base + A_COL0 + slot * A_SLOT_COLS, a_lo = a_hi + BK;
        const uint32_t b_hi = b_base + bslot * 2 * bsb, b_lo = b_hi + bsb;
        const uint32_t d = tmem_base + b * N_PASS * BN_MAX;
        // Always all 4 k-steps of the stage (operands are zero filled beyond K): a compile-time trip count
        // lets the descriptors be formed once and stepped by immediates.
        const uint64_t dbh0 = make_smem_desc(b_hi), dbl0 = make_smem_desc(b_lo);
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j) {
          const uint64_t koff = (uint64_t)((j * 2 * LBO) >> 4);   // two 16-byte k-chunks of B per MMA; 8 TMEM columns of A
          umma_tf32_ts_pred(elected, d, a_lo + j * UMMA_K, dbh0 + koff, idesc, j != 0);   // small terms first
          umma_tf32_ts_pred(elected, d, a_hi + j * UMMA_K, dbl0 + koff, idesc, 1);
          umma_tf32_ts_pred(elected, d, a_hi + j * UMMA_K, dbh0 + koff, idesc, 1);
        }
        if (lane == 0 && g == 6) trace(6);
        if (!resident) umma_commit_pred(elected, emptyB(bslot));
        umma_commit_pred(elected, acc_full(b));   // also frees A slot g % 2 for the producers
        if (lane == 0 && g < 8) trace(16 + g);
      }
      __syncwarp();
    }
    tc_fence_before();
  } else if (warp == LOADER_WARP) {
    // ------------------------------------------------------------------ loader: B (bulk copies) and raw A (TMA)
    if (lane == 0) {
      if (dbg & 16) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");   // experiment
      const float* wsrc = Wpk + (size_t)n_tile * nk * (2 * bsb / 4);
      for (int g = 0; g < n_items; ++g) {
        const int s = g % nk;
        if (!resident || g < nk) {
          const int bslot = resident ? s : g % nb;
          if (!resident && g >= nb) mbar_wait(emptyB(bslot), ((g / nb) - 1) & 1);
          mbar_arrive_expect_tx(fullB(bslot), 2 * bsb);
          bulk_copy_g2s(b_base + bslot * 2 * bsb, wsrc + (size_t)s * (2 * bsb / 4), 2 * bsb, fullB(bslot));
        }
        const int grp = g % N_GROUPS, i = g / N_GROUPS, rs = i % RAW_DEPTH;
        if (i >= RAW_DEPTH) mbar_wait(emptyRaw(grp, rs), ((i / RAW_DEPTH) - 1) & 1);
        // experiment: one box in flight per group (box i only after box i-1 has been consumed)
        if ((dbg & 4) && i >= 1) mbar_wait(emptyRaw(grp, (i - 1) % RAW_DEPTH), ((i - 1) / RAW_DEPTH) & 1);
        mbar_arrive_expect_tx(fullRaw(grp, rs), XRAW_STAGE_BYTES);
        tma_load_2d(raw_base + (grp * RAW_DEPTH + rs) * XRAW_STAGE_BYTES, &tmA, s * BK,
                    (m_first + (g / nk) * m_step) * BM, fullRaw(grp, rs));
      }
    }
  } else {
    // ------------------------------------------------------------------ drain + epilogue
    constexpr int HALF = BN_MAX;                   // columns per drain warp (one warp per lane quadrant)
    const int q = warp & 3;                        // TMEM lane quadrant this warp may access
    const int c_base = 0;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[HALF];
    for (int g = 0; g < n_items; ++g) {
      const int s = g % nk, b = g & 1;
      if (s == 0) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
      }
      const bool trd = tid == DRAIN_WARP0 * 32 && g == 6;
      if (trd) trace(24);
      mbar_wait_warp(acc_full(b), (g >> 1) & 1);
      tc_fence_after();
      if (trd) trace(25);
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 16) {
        if (c_base + c0 < bn) {
          float v[16];
          tmem_ld16(t_row + b * BN_MAX + c_base + c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c0 + i] += v[i];
        }
      }
      tc_fence_before();
      if (trd) trace(26);
      mbar_arrive(acc_empty(b));
      if (trd) trace(27);
      if (s == nk - 1) {
        // Epilogue.  A thread owns one output row; storing it directly would make every warp store touch 32
        // different 128-byte lines (measured: ~8.5k cycles per tile, the slowest stage of the expand layers).
        // The warp's 32 x bn tile is staged through shared memory and written out row-contiguously.
        const int m_base = (m_first + (g / nk) * m_step) * BM + q * 32;
        const int n_base = n_tile * bn;
        float* stg = stg_base + (warp - DRAIN_WARP0) * (32 * STG_PITCH);
        if (tid == DRAIN_WARP0 * 32 && g / nk == 1) trace(28);
        // branch-free: columns >= bn hold zeros (zero weights, zero bias) and are never copied out.
        // Two phases: all activations first (64 independent chains hide the MUFU latency), then all stores
        // (the compiler cannot move a bias load across a staging store: both are shared memory).
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0);
          acc[c0] += bv.x; acc[c0 + 1] += bv.y; acc[c0 + 2] += bv.z; acc[c0 + 3] += bv.w;
        }
        if (SWISH) {
#pragma unroll
          for (int c0 = 0; c0 < HALF; ++c0) acc[c0] = swishf(acc[c0]);
        }
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 4)
          *reinterpret_cast<float4*>(stg + lane * STG_PITCH + c0) = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
        __syncwarp();
        if (tid == DRAIN_WARP0 * 32 && g / nk == 1) trace(29);
        // copy out: lpr lanes per row (power of two >= bn/4), 32/lpr rows per pass, no divisions
        const int qn = bn >> 2;                   // float4 per staged row
        const int lpr_log = qn <= 4 ? 2 : (qn <= 8 ? 3 : 4);
        const int c4 = lane & ((1 << lpr_log) - 1), r_lane = lane >> lpr_log, r_step = 32 >> lpr_log;
        const int n = n_base + c4 * 4;
        if (c4 < qn && n < N) {
          for (int r = r_lane; r < 32; r += r_step) {
            const int m = m_base + r;
            if (m < M) {
              float4 o = *reinterpret_cast<const float4*>(stg + r * STG_PITCH + c4 * 4);
              if (RESID) {
                const float4 rr = *reinterpret_cast<const float4*>(resid + (size_t)m * N + n);
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              }
              *reinterpret_cast<float4*>(C + (size_t)m * N + n) = o;
            }
          }
        }
        __syncwarp();
        if (tid == DRAIN_WARP0 * 32 && g / nk == 1) trace(30);
      }
    }
  }
  if (tid == DRAIN_WARP0 * 32) trace(2);
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tid == 0) trace(3);
}

// ---- host side ---------------------------------------------------------------------------------------
// A [M][K] fp32 row-major; box = BK columns x BM rows, 128-byte swizzle, zero fill outside the tensor.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint) so that the library keeps
// loading on machines without libcuda.so.1 (the CPU test suite dlopens it to check the exports).
inline bool make_a_tensor_map(CUtensorMap* tm, const float* A, int M, int K) {
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || fn == nullptr)
      return false;
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// same tiles and weight image as make_plan(); only the raw ring (dense 16 KB stages) and hence the number of
// weight slots differ
inline Plan make_plan_tma(int N, int K, int ng) {
  Plan p = make_plan(N, K, ng);
  const int budget = ng == 1 ? 110 * 1024 : 200 * 1024;
  const int slot = 2 * b_stage_bytes(p.bn);
  p.nb = std::max(2, std::min(MAX_BSLOTS, (budget - xraw_bytes(ng) - STG_BYTES - 1024) / slot));
  p.resident = p.nk <= p.nb ? 1 : 0;
  if (p.resident) p.nb = p.nk;
  p.smem_bytes = xraw_bytes(ng) + p.nb * slot + STG_BYTES + 1024;
  return p;
}

}  // namespace tc
}  // namespace cosyb
