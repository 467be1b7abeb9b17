// 1x1 convolutions on tcgen05 with fp16 hi/lo split operands ("3xFP16"), sm_100a only.
//
//   C[M][N] = epi( ((A[M][K] * gate[m / rows_per_img][k]) @ W[N][K]^T) + bias[N] ) (+ resid[M][N])
//
// Reference op: the expand / project / head convolutions of models/efficientnet.py:71-98,174-190 (BN folded).
//
// Precision.  tcgen05 has no fp32 MMA, and one 11-bit pass misses the 1e-4 pose budget.  Every operand is split
//   x = x_hi + x_lo,  x_hi = round_to_11_bits(x) (exactly representable in fp16),  x_lo = fp16(x - x_hi)
// and a product is evaluated as  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  with kind::f16 MMAs (K = 16 per instruction,
// twice the tf32 rate; the dropped a_lo*b_lo term is 2^-22 relative).  Weights are pre-scaled by a power of two per
// layer so that their lo parts stay normal fp16 numbers (undone exactly in the epilogue FMA); activations are not
// scaled: an activation below 2^-3 loses at most 2^-25 ABSOLUTE in its lo part, far below the fp32 rounding of the
// O(1) sums it feeds.  |activation| must stay below 65504 (fp16 range).
// The TMEM accumulator add truncates (measured in round 1: -2^-25 relative per accumulation, a BIAS that adds up
// coherently over ~80 layers), so a partial sum only stays in TMEM for DS = 2 k-stages (64 k: 4 full-magnitude
// truncations instead of 22 with the tf32 kernel's schedule per 64 k) and the partial sums are added in registers
// with round-to-nearest.
//
// Work split ("stream-K" along m-tile x k-stage): the CTAs of one n-tile column share its m_tiles * nk k-stage units in
// equal contiguous ranges, so a CTA works through the tail of one m-tile, whole m-tiles, and the head of another.  With
// whole tiles only, 150 m-tiles on 148 SMs (every 15x20 layer at 64 hypotheses) took two rounds, and 35 m-tiles x 2
// n-tiles (the 7x10 layers) left half the SMs idle; now those layers split each tile's K over two CTAs.  A tile is shared
// by at most two CTAs: the one holding its first k-stage owns it, the other writes its fp32 partial sums to a per-CTA
// workspace slot (L2) and raises a flag; the owner adds them in a fixed order (own + partner), so results are run-to-run
// identical and no atomics touch data.  Every CTA produces its partial FIRST, so an owner never waits on a CTA that
// waits on it; all CTAs are co-resident (grid <= SM count, one CTA per SM).
//
// Structure (one CTA per SM, 16 warps = 4 warpgroups with setmaxnreg budgets).  Two shapes of the same kernel:
//                       BIG (bn <= 192)                          SMALL (bn <= 96: HBM / latency bound layers)
//   converters          warps 0-3                                warps 0-7: two groups on alternate k-stages
//   drain + epilogue    warps 4-11: 2 per TMEM lane quadrant,    warps 8-11: one per quadrant, all columns
//                       each half of the columns
//   accumulators        2 x 192 TMEM columns                     4 x 96 (the MMAs run up to 8 k-stages ahead of an
//                                                                epilogue that waits for residual rows)
//   warp 12  MMA issuer (elected lane): 6 MMAs per 32-wide k-stage, A from TMEM, B from shared memory
//   warp 13  B loader: weights are stored as [k-stage][hi|lo][8-k chunk][n][8 halfs], so the canonical K-major
//            shared-memory image of ANY n-range is 8 bulk copies per k-stage (cp.async.bulk)
//   warps 14-15  raw loaders: fp32 rows (and the SE gate rows of the images a tile touches) global -> shared with
//            cp.async, completion on an mbarrier per ring slot (cp.async.mbarrier.arrive.noinc), 5-6 k-stages in
//            flight; the converters never issue or wait for a global load (measured with the in-kernel trace: a
//            converter that issued its own copies and read the gate from global spent 1.5k of its 2.3k cycles per
//            k-stage on those two steps).
//   converter thread = tile row: LDS raw row + gate -> hi/lo split -> fp16 pack -> tcgen05.st into its TMEM lane.
// TMEM: 384 accumulator columns, 2 A slots x 32 columns (hi 16 | lo 16, two fp16 per column).
#pragma once
#include <cuda_fp16.h>

#include "common.h"
#include "kernels_tc.cuh"

namespace cosyb {
namespace pw2 {

using namespace tc;   // PTX wrappers (mbarrier, bulk copy, tcgen05 alloc/commit/fences, cp.async)

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements of A per k-stage
constexpr int KSTEPS = BK / 16;        // kind::f16 MMAs are K = 16
#ifndef PW2_SPLIT_MIN_NK
#define PW2_SPLIT_MIN_NK 16
#endif
#ifndef PW2_DS
#define PW2_DS 1
#endif
constexpr int DS = PW2_DS;             // k-stages accumulated in TMEM between drains
constexpr float RZ_COMP = 7.7e-8f;     // expected relative loss of one drain group to the accumulator's round-toward-zero
constexpr int BN_MAX = 192;
constexpr int BN_SMALL = 96;
constexpr int CONV_THREADS = 128;      // one converter group
constexpr int MMA_WARP = 12;
constexpr int LOADER_WARP = 13;
constexpr int RAW_WARP0 = 14;
constexpr int RAW_THREADS = 64;
constexpr int THREADS = 16 * 32;
constexpr int RAW_ROW_BYTES = BK * 4 + 16;             // 144: 16-byte row reads of 8 consecutive rows are conflict free
constexpr int GATE_IMGS = 3;                           // images a 128-row tile can touch when rows_per_img >= 64
constexpr int RAW_GATE_OFF = BM * RAW_ROW_BYTES;       // 18432
constexpr int RAW_STAGE_BYTES = RAW_GATE_OFF + GATE_IMGS * BK * 4;   // 18816
constexpr int MAX_RAW = 6;
__host__ __device__ constexpr int raw_depth(bool small) { return small ? 6 : 5; }
constexpr int N_ASLOTS = 2;
constexpr int A_SLOT_COLS = 32;                        // hi: 16 columns (32 halfs) | lo: 16 columns
constexpr int MAX_BSLOTS = 8;
constexpr int MAX_ACC = 4;
constexpr int HALF_MAX = 96;                           // accumulator columns per drain warp
constexpr int STG_PITCH = 36;                          // floats per staged row (32 + 4)
constexpr int STG_WARP_FLOATS = 32 * STG_PITCH;
__host__ __device__ constexpr int n_drain_warps(bool small) { return small ? 4 : 8; }
__host__ __device__ constexpr int stg_bytes(bool small) { return n_drain_warps(small) * STG_WARP_FLOATS * 4; }
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t A_COL0 = 384;
static_assert(A_COL0 + N_ASLOTS * A_SLOT_COLS <= TMEM_COLS, "TMEM budget");

__host__ __device__ inline int b_slot_bytes(int bn) { return bn * 128; }   // hi (bn x 32 halfs) + lo

// D fp32, A/B fp16, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc_f16(int bn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= (uint32_t)(bn >> 3) << 17;    // n_dim
  d |= (uint32_t)(BM >> 4) << 24;    // m_dim
  return d;                          // a_format = b_format = 0 (F16), no negate, K-major
}
// K-major, no swizzle: 8 rows x 16 bytes per core matrix; lbo = bytes between the two 8-k chunks of one MMA,
// sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_f16_ts_pred(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// 16 consecutive accumulator columns of the thread's lane; the caller issues tmem_ld_wait() before using them
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st_global_v4_if(float* p, const float4& v, bool ok) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "@p st.global.v4.f32 [%0], {%1, %2, %3, %4};\n"
      "}\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// fp16 pair: `lo_k` (even k) in bits 0-15, `hi_k` (odd k) in bits 16-31
__device__ __forceinline__ uint32_t pack_f16x2(float even_k, float odd_k) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(odd_k), "f"(even_k));
  return d;
}
// x -> (hi, lo): hi = x rounded to 11 significant bits (ties away), lo = x - hi (exact in fp32)
__device__ __forceinline__ void split11(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

// Work split of a launch: CTA c owns n-tile c % n_tiles; with j = c / n_tiles and P = gridDim.x / n_tiles it takes
//   split_k = 0: the whole m-tiles j, j + P, j + 2P, ... (concurrent CTAs stream neighbouring rows: the HBM-bound layers);
//   split_k = 1: the k-stage units [j U / P, (j + 1) U / P) of the column, U = m_tiles * nk, unit u = (m-tile u / nk,
//                k-stage u % nk): long-K layers whose m-tile count does not fill the SMs evenly.
struct Plan {
  int bn, small, n_tiles, nk, nb, resident, smem_bytes, grid, split_k;
};

__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Wpk: [nk][2 (hi|lo)][4 chunks][n_alloc rows][8 halfs] fp16, rows >= N zero; n_alloc >= n_tiles * bn.
// gate_smem: the gate rows travel through the raw ring (needs rows_per_img >= 64); otherwise they are read from global.
template <bool GATE, bool SWISH, bool RESID, bool SMALL, bool SPLITK>
__global__ void __launch_bounds__(THREADS, 1)
k_pw2(const float* __restrict__ A, const __half* __restrict__ Wpk, const float* __restrict__ bias,
      const float* __restrict__ gate, const float* __restrict__ resid, float* __restrict__ C, int M, int N, int K,
      int rows_per_img, int bn, int n_tiles, int nb, int resident, int n_alloc, float inv_wscale, int gate_smem,
      float* __restrict__ ws /*[grid][BN_MAX / 4][128] float4 partial sums*/, int* __restrict__ flags /*[grid]*/) {
  // SPLITK: contiguous k-stage ranges (tiles may be shared by two CTAs); else whole m-tiles part, part + n_parts, ...
  static_assert(DS == 1, "the k-stage work split drains every k-stage");
  constexpr int NCG = SMALL ? 2 : 1;                      // converter groups
  constexpr int DRAIN_WARP0 = 4 * NCG;
  constexpr int N_DRAIN = n_drain_warps(SMALL);
  constexpr int RAW_DEPTH = raw_depth(SMALL);
  constexpr int NACC = SMALL ? 4 : 2;
  constexpr int ACC_STRIDE = SMALL ? 96 : 192;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_RAW + 2 * N_ASLOTS + 2 * MAX_BSLOTS + 2 * MAX_ACC];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[BN_MAX];
  const uint32_t raw_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = raw_base + RAW_DEPTH * RAW_STAGE_BYTES;
  const uint32_t bsb = b_slot_bytes(bn);
  float* stg_base = reinterpret_cast<float*>(smem_raw + (raw_base - smem_u32(smem_raw)) + RAW_DEPTH * RAW_STAGE_BYTES + nb * bsb);
  const int tid = threadIdx.x, lane = tid % 32;
  const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);
  const int nk = (K + BK - 1) / BK;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tile = blockIdx.x % n_tiles, part = blockIdx.x / n_tiles, n_parts = gridDim.x / n_tiles;
  // k-stage units of this CTA: item i = (m-tile tile0 + ((s0 + i) / nk) * tstep, k-stage (s0 + i) % nk), i < n_items
  const long long n_units = (long long)m_tiles * nk;
  const int u0 = SPLITK ? (int)(part * n_units / n_parts) : 0, u1 = SPLITK ? (int)((part + 1) * n_units / n_parts) : 0;
  const int tile0 = SPLITK ? u0 / nk : part, s0 = SPLITK ? u0 % nk : 0, tstep = SPLITK ? 1 : n_parts;
  const int n_items = SPLITK ? u1 - u0 : (part < m_tiles ? (m_tiles - 1 - part) / n_parts + 1 : 0) * nk;
  auto rawFull = [&](int s) { return smem_u32(&bars[s]); };
  auto rawEmpty = [&](int s) { return smem_u32(&bars[MAX_RAW + s]); };
  auto fullA = [&](int s) { return smem_u32(&bars[2 * MAX_RAW + s]); };
  auto emptyA = [&](int s) { return smem_u32(&bars[2 * MAX_RAW + N_ASLOTS + s]); };
  auto fullB = [&](int s) { return smem_u32(&bars[2 * MAX_RAW + 2 * N_ASLOTS + s]); };
  auto emptyB = [&](int s) { return smem_u32(&bars[2 * MAX_RAW + 2 * N_ASLOTS + MAX_BSLOTS + s]); };
  auto acc_full = [&](int b) { return smem_u32(&bars[2 * MAX_RAW + 2 * N_ASLOTS + 2 * MAX_BSLOTS + b]); };
  auto acc_empty = [&](int b) { return smem_u32(&bars[2 * MAX_RAW + 2 * N_ASLOTS + 2 * MAX_BSLOTS + MAX_ACC + b]); };

  if (tid == 0) {
    for (int s = 0; s < MAX_RAW; ++s) {
      mbar_init(rawFull(s), RAW_THREADS);
      mbar_init(rawEmpty(s), CONV_THREADS);
    }
    for (int s = 0; s < N_ASLOTS; ++s) {
      mbar_init(fullA(s), CONV_THREADS);
      mbar_init(emptyA(s), 1);
    }
    for (int s = 0; s < MAX_BSLOTS; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
    }
    for (int b = 0; b < MAX_ACC; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), N_DRAIN * 32);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < BN_MAX; i += THREADS) {
    const int n = n_tile * bn + i;
    s_bias[i] = (i < bn && n < N) ? __ldg(bias + n) : 0.f;
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem, 0);
  if (tid == 0) trace(0);

  if (warp < DRAIN_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ converters (A operand -> TMEM)
    const int grp = warp / 4, q = warp % 4;
    const int row = q * 32 + lane;                          // tile row == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int s = (s0 + grp) % nk, m0 = (tile0 + ((s0 + grp) / nk) * tstep) * BM;   // (k-stage, first row) of this group's item
    int img_local = 0, prev_m0 = -1;
    const float* gptr = gate;
    for (int i = grp; i < n_items; i += NCG) {
      const int k0 = s * BK;
      const int slot = i % N_ASLOTS, rslot = i % RAW_DEPTH;
      const bool tr = (tid % CONV_THREADS) == 0 && i < 16;
      if (GATE && m0 != prev_m0) {
        prev_m0 = m0;
        const int img = min(m0 + row, M - 1) / rows_per_img;
        img_local = img - m0 / rows_per_img;
        gptr = gate + (size_t)img * K;
      }
      if (tr) trace(64 + 4 * i);
      mbar_wait_warp(rawFull(rslot), (i / RAW_DEPTH) & 1);
      if (tr) trace(64 + 4 * i + 1);
      const uint32_t src = raw_base + rslot * RAW_STAGE_BYTES + row * RAW_ROW_BYTES;
      float v[BK];
#pragma unroll
      for (int c = 0; c < BK / 4; ++c)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[c * 4]), "=f"(v[c * 4 + 1]), "=f"(v[c * 4 + 2]), "=f"(v[c * 4 + 3])
                     : "r"(src + c * 16) : "memory");
      if (GATE) {
        if (gate_smem) {
          const uint32_t gsrc = raw_base + rslot * RAW_STAGE_BYTES + RAW_GATE_OFF + img_local * (BK * 4);
#pragma unroll
          for (int c = 0; c < BK / 4; ++c) {
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(gsrc + c * 16) : "memory");
            v[c * 4] *= x.x; v[c * 4 + 1] *= x.y; v[c * 4 + 2] *= x.z; v[c * 4 + 3] *= x.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < BK / 4; ++c) {
            if (k0 + c * 4 < K) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(gptr + k0 + c * 4));
              v[c * 4] *= x.x; v[c * 4 + 1] *= x.y; v[c * 4 + 2] *= x.z; v[c * 4 + 3] *= x.w;
            }
          }
        }
      }
      if (tr) trace(256 + 4 * i);
      uint32_t ph[BK / 2], pl[BK / 2];
#pragma unroll
      for (int c = 0; c < BK / 2; ++c) {
        float h0, l0, h1, l1;
        split11(v[2 * c], h0, l0);
        split11(v[2 * c + 1], h1, l1);
        ph[c] = pack_f16x2(h0, h1);
        pl[c] = pack_f16x2(l0, l1);
      }
      mbar_arrive(rawEmpty(rslot));                 // the raw slot has been read (its values were just consumed)
      if (tr) trace(64 + 4 * i + 2);
      if (i >= N_ASLOTS) mbar_wait_warp(emptyA(slot), ((i / N_ASLOTS) - 1) & 1);   // MMAs of item i-2 done
      tc_fence_after();
      tmem_st16(t_lane + A_COL0 + slot * A_SLOT_COLS, ph);
      tmem_st16(t_lane + A_COL0 + slot * A_SLOT_COLS + 16, pl);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(fullA(slot));
      if (tr) trace(64 + 4 * i + 3);
      s += NCG;
      while (s >= nk) { s -= nk; m0 += tstep * BM; }
    }
  } else if (warp >= MMA_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == MMA_WARP) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = make_idesc_f16(bn);
      const uint32_t lbo = (uint32_t)bn * 16u;
      int s = s0, gg = 0;
      for (int g = 0; g < n_items; ++g) {
        const int slot = g % N_ASLOTS;
        const int b = gg % NACC;
        const bool first = (s % DS) == 0, last = (s % DS) == DS - 1 || s == nk - 1;
        const int bslot = resident ? s : g % nb;
        if (first && gg >= NACC) mbar_wait_warp(acc_empty(b), ((gg / NACC) - 1) & 1);
        if (!resident) mbar_wait_warp(fullB(bslot), (g / nb) & 1);
        else if (g < nk) mbar_wait_warp(fullB(bslot), 0);
        if (lane == 0 && g < 32) trace(128 + 2 * g);
        mbar_wait_warp(fullA(slot), (g / N_ASLOTS) & 1);
        tc_fence_after();
        if (lane == 0 && g < 32) trace(128 + 2 * g + 1);
        {
          const uint32_t a_hi = tmem_base + A_COL0 + slot * A_SLOT_COLS, a_lo = a_hi + 16;
          const uint32_t b_hi = b_base + bslot * bsb, b_lo = b_hi + bsb / 2;
          const uint32_t d = tmem_base + b * ACC_STRIDE;
          const uint64_t dbh0 = make_desc(b_hi, lbo, 128), dbl0 = make_desc(b_lo, lbo, 128);
          // The accumulator add rounds toward zero: every MMA costs up to one ulp of the CURRENT accumulator value,
          // always in the same direction.  The small cross terms go in while the accumulator is still small (they
          // then cost nothing), the two full-magnitude hi*hi MMAs last.
#pragma unroll
          for (int j = 0; j < KSTEPS; ++j) {
            const uint64_t koff = (uint64_t)((j * 2 * lbo) >> 4);   // two 8-k chunks of B per MMA; 8 TMEM columns of A
            umma_f16_ts_pred(d, a_lo + j * 8, dbh0 + koff, idesc, (first && j == 0) ? 0u : 1u);
            umma_f16_ts_pred(d, a_hi + j * 8, dbl0 + koff, idesc, 1);
          }
#pragma unroll
          for (int j = 0; j < KSTEPS; ++j) {
            const uint64_t koff = (uint64_t)((j * 2 * lbo) >> 4);
            umma_f16_ts_pred(d, a_hi + j * 8, dbh0 + koff, idesc, 1);
          }
          if (!resident) umma_commit_elect(emptyB(bslot));
          umma_commit_elect(emptyA(slot));
          if (last) umma_commit_elect(acc_full(b));
        }
        __syncwarp();
        if (last) ++gg;
        if (++s == nk) s = 0;
      }
      tc_fence_before();
    } else if (warp == LOADER_WARP) {
      // ---------------------------------------------------------------- B loader
      if (lane == 0) {
        const int n0 = n_tile * bn;
        const int n_loads = resident ? min(nk, n_items) : n_items;
        const uint32_t chunk_bytes = (uint32_t)bn * 16u;
        int s = s0;
        for (int g = 0; g < n_loads; ++g) {
          const int bslot = resident ? s : g % nb;
          if (!resident && g >= nb) mbar_wait(emptyB(bslot), ((g / nb) - 1) & 1);
          mbar_arrive_expect_tx(fullB(bslot), bsb);
          const __half* src = Wpk + ((size_t)s * 8 * n_alloc + n0) * 8;
          const uint32_t dst = b_base + bslot * bsb;
#pragma unroll
          for (int hc = 0; hc < 8; ++hc)   // hc = (hi|lo) * 4 + chunk
            bulk_copy_g2s(dst + hc * chunk_bytes, src + (size_t)hc * n_alloc * 8, chunk_bytes, fullB(bslot));
          if (++s == nk) s = 0;
        }
      }
    } else {
      // ---------------------------------------------------------------- raw loaders (A rows + gate rows)
      // The two loader warps were the slowest role (in-kernel trace: ~1500 cycles per item against ~1070 for a
      // converter group and ~960 for the six MMAs): ptxas spent ~25 instructions of 64-bit address arithmetic and
      // predicate logic on each of the 16 cp.async of an item.  Row offsets (32-bit, relative to the tile) and the
      // row-valid mask are therefore computed once per tile; a stage adds one 64-bit base.
      const int lt = tid - RAW_WARP0 * 32;
      const int r0 = lt >> 3, kc = lt & 7;                  // chunk kc of rows r0, r0 + 8, ..., r0 + 120
      const uint32_t dst_thr = raw_base + r0 * RAW_ROW_BYTES + kc * 16;
      const int n_imgs = (M + rows_per_img - 1) / rows_per_img;
      int s = s0, m0 = tile0 * BM, img0 = 0;
      uint32_t roff[16];
      uint32_t vmask = 0;
      const char* tile_base = reinterpret_cast<const char*>(A);
      for (int i = 0; i < n_items; ++i) {
        const int rslot = i % RAW_DEPTH;
        const int k0 = s * BK;
        if (s == 0 || (SPLITK && i == 0)) {
          if (GATE) img0 = m0 / rows_per_img;
          tile_base = reinterpret_cast<const char*>(A + (size_t)m0 * K + kc * 4);
          vmask = 0;
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const bool ok = m0 + r0 + 8 * it < M;
            roff[it] = ok ? (uint32_t)((r0 + 8 * it) * K) * 4u : 0u;
            vmask |= (ok ? 1u : 0u) << it;
          }
        }
        if (i >= RAW_DEPTH) mbar_wait_warp(rawEmpty(rslot), ((i / RAW_DEPTH) - 1) & 1);
        const uint32_t kmask = k0 + kc * 4 < K ? vmask : 0u;
        const char* src = tile_base + (size_t)k0 * 4;
        const uint32_t dst = dst_thr + rslot * RAW_STAGE_BYTES;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const uint32_t n = ((kmask >> it) & 1u) * 16u;     // src-size 0 -> the 16 destination bytes are zero filled
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + it * 8 * RAW_ROW_BYTES),
                       "l"(src + roff[it]), "r"(n) : "memory");
        }
        if (GATE && gate_smem && lt < GATE_IMGS * 8) {
          const int img = img0 + r0;                        // r0 = lt / 8 = image within the tile here
          const bool ok = k0 + kc * 4 < K && img < n_imgs;
          cp_async16(raw_base + rslot * RAW_STAGE_BYTES + RAW_GATE_OFF + lt * 16,
                     ok ? (const void*)(gate + (size_t)img * K + k0 + kc * 4) : (const void*)A, ok);
        }
        cp_async_arrive_noinc(rawFull(rslot));
        if (++s == nk) { s = 0; m0 += tstep * BM; }
      }
      cp_async_wait<0>();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    // ------------------------------------------------------------------ drain + epilogue
    const int dw = warp - DRAIN_WARP0;
    const int q = warp & 3;                        // TMEM lane quadrant this warp may access
    const int half = dw >> 2;                      // column half (BIG only)
    const int h0 = SMALL ? bn : ((bn + 31) / 32) * 16;   // columns of half 0 (multiple of 16)
    const int c_base = half == 0 ? 0 : h0;
    const int width = half == 0 ? h0 : bn - h0;    // multiple of 16, may be 0
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + dw * STG_WARP_FLOATS;
    float acc[HALF_MAX];
    const int n_groups = n_items;                  // DS == 1: one drain per k-stage unit
    int seg_s0 = 0;                                // first k-stage of the current tile segment
    int s_cur = s0, t = tile0;                     // (k-stage, m-tile) of the current unit
    for (int gg = 0; gg < n_groups; ++gg, ++s_cur) {
      const int b = gg % NACC;
      if (s_cur == nk) { s_cur = 0; t += tstep; }
      if (s_cur == 0 || (SPLITK && gg == 0)) {
        seg_s0 = s_cur;
#pragma unroll
        for (int i = 0; i < HALF_MAX; ++i) acc[i] = 0.f;
      }
      const bool trd = tid == DRAIN_WARP0 * 32 && gg < 16;
      if (trd) trace(192 + 4 * gg);
      mbar_wait_warp(acc_full(b), (gg / NACC) & 1);
      tc_fence_after();
      if (trd) trace(192 + 4 * gg + 1);
#pragma unroll
      for (int c0 = 0; c0 < HALF_MAX; c0 += 32) {
        if (c0 < width) {
          float v0[16], v1[16];
          tmem_ld16_nowait(t_row + b * ACC_STRIDE + c_base + c0, v0);
          if (c0 + 16 < width) tmem_ld16_nowait(t_row + b * ACC_STRIDE + c_base + c0 + 16, v1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c0 + i] += v0[i];
          if (c0 + 16 < width) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[c0 + 16 + i] += v1[i];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(b));
      if (trd) trace(192 + 4 * gg + 2);
      const bool seg_last = s_cur == nk - 1 || (SPLITK && gg == n_groups - 1);
      if (SPLITK && seg_last && seg_s0 > 0) {
        // Tail of a tile another CTA owns (the one before this in the column): hand over the partial sums.
        float4* slot = reinterpret_cast<float4*>(ws) + (size_t)(blockIdx.x - n_tiles) * (BN_MAX / 4) * BM + q * 32 + lane;
#pragma unroll
        for (int c = 0; c < HALF_MAX; c += 4)
          if (c < width) slot[((c_base + c) >> 2) * BM] = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"r"(N_DRAIN * 32) : "memory");
        if (tid == DRAIN_WARP0 * 32) {
          int* f = flags + (blockIdx.x - n_tiles);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
        }
      } else if (seg_last) {
        if (SPLITK && s_cur != nk - 1) {
          // Head of a tile whose tail runs on the next CTA of the column: wait for its partial sums, add them.
          if (tid == DRAIN_WARP0 * 32) {
            int* f = flags + blockIdx.x;
            int v;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            } while (v == 0);
          }
          asm volatile("bar.sync 1, %0;" ::"r"(N_DRAIN * 32) : "memory");
          const float4* slot = reinterpret_cast<const float4*>(ws) + (size_t)blockIdx.x * (BN_MAX / 4) * BM + q * 32 + lane;
#pragma unroll
          for (int c = 0; c < HALF_MAX; c += 4) {
            if (c < width) {
              const float4 pv = __ldcg(slot + ((c_base + c) >> 2) * BM);
              acc[c] += pv.x; acc[c + 1] += pv.y; acc[c + 2] += pv.z; acc[c + 3] += pv.w;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"r"(N_DRAIN * 32) : "memory");
          if (tid == DRAIN_WARP0 * 32) flags[blockIdx.x] = 0;      // consumed: ready for the next launch
        }
        // Epilogue: the warp's 32 x width tile goes through a 32 x 32 staging tile so that global stores are
        // row-contiguous 128-byte segments; the residual rows are requested before the activation math.
        const int m_base = t * BM + q * 32;
        const int n_base = n_tile * bn + c_base;
        const int c4 = lane & 7, r_lane = lane >> 3;
#pragma unroll
        for (int cc = 0; cc < HALF_MAX; cc += 32) {
          if (cc < width) {
            const int n = n_base + cc + c4 * 4;
            const bool col_ok = cc + c4 * 4 < width && n < N;
            float4 rr[8];
            if (RESID) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int m = m_base + r_lane + 4 * j;
                rr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col_ok && m < M) rr[j] = *reinterpret_cast<const float4*>(resid + (size_t)m * N + n);
              }
            }
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              // acc (1 + RZ_COMP) with ONE round-to-nearest: undoes the EXPECTED loss of the round-toward-zero
              // accumulator adds (every drain group loses the same relative amount, so the sum does too; measured on
              // same-sign operands: -7.7e-8 with this MMA order).  1 + 7.7e-8 is not a float, hence the FMA form.
              const float a = fmaf(acc[cc + i], RZ_COMP, acc[cc + i]);
              o[i] = fmaf(a, inv_wscale, s_bias[c_base + cc + i]);
            }
            if (SWISH) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = swishf(o[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(stg + lane * STG_PITCH + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
            __syncwarp();
            if (col_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int r = r_lane + 4 * j, m = m_base + r;
                float4 ov = *reinterpret_cast<const float4*>(stg + r * STG_PITCH + c4 * 4);
                if (RESID) { ov.x += rr[j].x; ov.y += rr[j].y; ov.z += rr[j].z; ov.w += rr[j].w; }
                st_global_v4_if(C + (size_t)m * N + n, ov, m < M);     // predicated, not branched: 24 of these per tile
              }
            }
            __syncwarp();
          }
        }
        if (trd) trace(192 + 4 * gg + 3);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side: tile plan and weight packing -------------------------------------------------------
constexpr size_t WS_SLOT_BYTES = (size_t)(BN_MAX / 4) * BM * 16;   // partial sums of one CTA (kernel argument `ws`)
constexpr int N_ALLOC_PAD = BN_MAX;   // zero rows after the last real row so any tile's bulk copies stay in bounds

inline int n_alloc_for(int N) { return (N + 15) / 16 * 16 + N_ALLOC_PAD; }

inline Plan make_plan(int M, int N, int K, int n_sms, int force_nt = 0) {
  const int smem_budget = 225 * 1024 - 1024;
  const int n16 = (N + 15) / 16;
  const int m_tiles = (M + BM - 1) / BM;
  const int nk = (K + BK - 1) / BK;
  Plan best{};
  double best_cost = 1e300;
  for (int nt = (n16 * 16 + BN_MAX - 1) / BN_MAX; nt <= n16; ++nt) {
    const int bn = (n16 + nt - 1) / nt * 16;
    if (force_nt > 0 && nt != force_nt && nt < n16) continue;   // tuning aid: take this column count if it is valid
    if (bn < 16) break;
    if (nt > 1 && bn < 32) break;
    Plan p;
    p.bn = bn;
    p.small = bn <= BN_SMALL ? 1 : 0;
    p.n_tiles = (n16 * 16 + bn - 1) / bn;
    if (p.n_tiles > n_sms) continue;
    p.nk = nk;
    const int fixed = raw_depth(p.small) * RAW_STAGE_BYTES + stg_bytes(p.small) + 1024;
    const int slot = b_slot_bytes(bn);
    p.nb = std::max(2, std::min(MAX_BSLOTS, (smem_budget - fixed) / slot));
    p.resident = nk <= p.nb ? 1 : 0;
    if (p.resident) p.nb = nk;
    p.smem_bytes = fixed + p.nb * slot;
    // CTAs per n-tile column.  Whole m-tiles per CTA unless K is long (>= 16 k-stages: a tile's mainloop dwarfs the
    // fix-up and the extra epilogue of a shared tile) and splitting shortens the busiest CTA by >= 10 %: then either all
    // the SMs the column can have (equal k-stage ranges, a tile shared by at most two CTAs) or, when the m-tiles do not
    // even fill half of them (the 7x10 layers: 35 m-tiles), two CTAs per m-tile.
    const int max_par = std::max(1, n_sms / p.n_tiles);
    int m_par = std::min(m_tiles, max_par);
    double units = std::ceil((double)m_tiles / m_par) * nk;           // k-stage units of the busiest CTA
    double tiles_touched = std::ceil((double)m_tiles / m_par);
    p.split_k = 0;
    if (nk >= PW2_SPLIT_MIN_NK) {
      int sp = 0;
      if (m_tiles >= max_par) sp = max_par;
      else if (2 * m_tiles <= max_par) sp = 2 * m_tiles;
      if (sp) {
        const double u2 = std::ceil((double)m_tiles * nk / sp);
        if (u2 <= 0.9 * units) { p.split_k = 1; m_par = sp; units = u2; tiles_touched = std::ceil(u2 / nk) + 1.0; }
      }
    }
    p.grid = m_par * p.n_tiles;
    // cycles per k-stage: tensor pipe (6 MMAs of bn/2 cycles), shared-memory traffic at 128 B/clk (MMA reads of B,
    // raw A in + out, weight ring writes), converter issue
    const double mma = 3.0 * bn;
    const double smem = (3.0 * 2 * bn * 32 + 2.0 * BM * BK * 4 + (p.resident && units > nk ? 0.0 : (double)slot)) / 128.0;
    const double conv = p.small ? 350.0 : 700.0;
    const double stage = std::max(std::max(mma, smem), conv);
    const double epi = 40.0 * bn / 2 / 16 + 600.0;
    const double cost = units * stage + tiles_touched * epi + 3000.0;
    if (cost < best_cost) { best_cost = cost; best = p; }
  }
  return best;
}

inline uint16_t host_f16_bits(float x) {
  const __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
inline float host_round11(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// power-of-two scale that puts max|W| into [2^12, 2^13): hi parts stay far below the fp16 limit, lo parts of
// all weights above 2^-16 max|W| stay normal
inline float weight_scale(const float* W, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(W[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e;
  std::frexp(mx, &e);           // mx = f * 2^e, f in [0.5, 1)
  return std::ldexp(1.f, 13 - e);
}

// W_nk [N][K] (BN scale folded) -> [nk][hi|lo][4 chunks][n_alloc][8] fp16 bits
inline std::vector<uint16_t> pack_weights(const float* W_nk, int N, int K, float wscale) {
  const int nk = (K + BK - 1) / BK, n_alloc = n_alloc_for(N);
  std::vector<uint16_t> out((size_t)nk * 8 * n_alloc * 8, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = W_nk[(size_t)n * K + k] * wscale;
      const float h = host_round11(w);
      const int s = k / BK, c = (k % BK) / 8, e = k % 8;
      const size_t hi_off = (((size_t)s * 8 + c) * n_alloc + n) * 8 + e;
      const size_t lo_off = (((size_t)s * 8 + 4 + c) * n_alloc + n) * 8 + e;
      out[hi_off] = host_f16_bits(h);
      out[lo_off] = host_f16_bits(w - h);
    }
  return out;
}

}  // namespace pw2
}  // namespace cosyb
