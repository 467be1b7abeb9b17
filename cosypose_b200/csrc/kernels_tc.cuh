// 1x1 convolutions on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   C[M][N] = epi( (A[M][K] * gate[m / rows_per_img][k]) @ W[N][K]^T + bias[N] ) (+ resid[M][N])
//
// Precision: tcgen05 has no fp32 MMA.  kind::tf32 keeps 11 significant bits per operand, which breaks
// the 1e-4 pose budget after 5 iterations (8e-5, marginal), so every product is
// evaluated as the 3xTF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = rna_tf32(a),
// a_lo = a - a_hi).  Measured on B200: the TMEM accumulator add TRUNCATES (mean error -2^-24 per
// accumulation, -2.6e-5 relative after the 864 accumulations of a K = 2304 layer), so products are
// only accumulated in TMEM over one 32-wide k-stage (12 MMAs); the per-stage partial sums are
// drained with tcgen05.ld and added in registers with round-to-nearest (double-buffered TMEM, the
// drain of stage s overlaps the MMAs of stage s+1).  tools/tc_precision.py measures the variants.
//
// Data movement per CTA (one 128 x BN output tile):
//   * W is pre-split (hi | lo) and pre-packed on the host into the exact shared-memory image of each
//     (n-tile, k-stage) block, so a stage's B operand is ONE 1-D bulk copy (cp.async.bulk ->
//     UBLKCP) completing on the stage's mbarrier; no tensor map is needed.
//   * A needs a CUDA-core pass anyway (SE gate, hi/lo split), so the 4 producer warps load it with
//     coalesced 16-byte loads, gate + split it in registers and store it into the canonical
//     K-major no-swizzle core-matrix layout: 8 rows x 16 bytes per core matrix,
//       byte(r, k) = (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4,  LBO = 128, SBO = 1024 (BK = 32)
//     followed by fence.proxy.async so the tensor core (async proxy) sees the generic-proxy stores.
//     Raw fp32 A travels global -> shared by cp.async (LDGSTS) four k-stages ahead (64 KB in flight per
//     SM, no register scoreboard in the way); the same thread then transforms its own 8 x 16 bytes.
//   * one elected thread of warp 4 issues tcgen05.mma (M=128, N=BN, K=8 per instruction), commits
//     the stage's "empty" mbarrier and the accumulator buffer's "full" mbarrier.
//   * 4 drain warps (warp w owns TMEM lanes 32(w%4)..+31 = output rows) add every stage's partial
//     sums into registers and finally apply bias / swish / residual and store 16-byte vectors.
#pragma once
#include "common.h"
#include "kernels_backbone.cuh"

namespace cosyb {
namespace tc {

constexpr int BM = 128;           // UMMA M
constexpr int BK = 32;            // fp32 elements per k-stage (8 core-matrix columns)
constexpr int UMMA_K = 8;         // tf32
constexpr int KSTEPS = BK / UMMA_K;
constexpr int PRODUCER_THREADS = 128;
constexpr int DRAIN_THREADS = 128;
constexpr int THREADS = PRODUCER_THREADS + 32 + DRAIN_THREADS + 32;   // 10 warps; two CTAs share an SM
__host__ __device__ constexpr int threads_for(int ng) { return ng * PRODUCER_THREADS + 32 + DRAIN_THREADS + 32; }
constexpr uint32_t LBO = 128, SBO = 1024;
constexpr int A_STAGE_BYTES = BM * BK * 4;   // one of hi / lo

__host__ __device__ inline int b_stage_bytes(int bn) { return bn * BK * 4; }   // one of hi / lo
__host__ __device__ inline int stage_bytes(int bn) { return 2 * A_STAGE_BYTES + 2 * b_stage_bytes(bn); }

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The spin loop lives INSIDE the asm block (labels are scoped by the braces): a C++ loop on the per-thread
// result makes the compiler treat everything after it as divergent and wrap each tcgen05 instruction in
// an ELECT / R2UR.BROADCAST waterfall.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "nanosleep.u32 40;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// Whole-warp wait with a warp-uniform exit: every lane polls (its own acquire) and the loop ends on a warp
// vote, so the lanes never leave the loop on different polls.  (When all 32 lanes spin independently nothing
// makes them reconverge before the .sync.aligned tcgen05.st / tcgen05.ld / elect.sync that follow.)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "VWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "vote.sync.all.pred q, p, 0xffffffff;\n"
      "@q bra VWAIT_DONE;\n"
      "nanosleep.u32 20;\n"
      "bra VWAIT_LOOP;\n"
      "VWAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no swizzle, descriptor version 1 (Blackwell); fields in 16-byte units
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// D fp32, A/B tf32, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 2u << 7;                      // a_format = TF32
  d |= 2u << 10;                     // b_format = TF32
  d |= (uint32_t)(bn >> 3) << 17;    // n_dim
  d |= (uint32_t)(BM >> 4) << 24;    // m_dim
  return d;
}

// Round to the nearest TF32 (ties away from zero) with integer ALU ops, bit-identical to cvt.rna.tf32.f32
// for finite inputs and to the host-side packing.  cvt runs on the 16-lane conversion pipe, which the
// producers (2 conversions per A element) and the swish epilogue (MUFU.EX2 + MUFU.RCP) were saturating.
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// Optional cycle trace of CTA 0 for tuning (cosyb200_debug_trace): slot -> clock64() stamp.
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int slot) {
  if (g_trace != nullptr && blockIdx.x == 0) g_trace[slot] = clock64();
}

// registers -> TMEM, 32 lanes x 32 columns (thread t of the warp writes lane base+t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
      "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
      "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
      "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from TMEM (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0 -> the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

constexpr int RAW_DEPTH = 2;                       // raw A k-stages in flight per CTA (two CTAs per SM)
constexpr int RAW_ROW_BYTES = BK * 4 + 16;         // 144: row pitch that keeps 16-byte row reads conflict free
constexpr int RAW_STAGE_BYTES = BM * RAW_ROW_BYTES;
// Warp-uniform issue: every lane executes the instruction slot, one lane (pred != 0) performs it.  With
// a divergent `if (lane == 0)` around it the compiler cannot keep the operands in uniform registers and
// wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop (~100 cycles per MMA, measured).
__device__ __forceinline__ void umma_tf32_ts_pred(uint32_t pred, uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint32_t pred, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar), "r"(pred)
      : "memory");
}

constexpr int N_ASLOTS = 2;                        // A stage slots in TMEM (one per producer group)
constexpr int N_PASS = 1;                          // accumulators per buffer (independent chains did not help)
constexpr int MAX_BSLOTS = 8;                      // B stage slots in shared memory (plan.nb <= this)
constexpr int A_SLOT_COLS = 2 * BK;                // hi | lo
__host__ __device__ constexpr int raw_bytes(int ng) { return ng * RAW_DEPTH * RAW_STAGE_BYTES; }
constexpr int STG_PITCH = 68;                      // floats per staged output row (64 + 4: conflict-free 16-byte accesses)
constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;  // one 32-row staging tile per drain warp

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// --------------------------------------------------------------------------------------------------
// Persistent kernel.  CTA c owns n-tile c % n_tiles and m-tiles c / n_tiles, + gridDim.x / n_tiles, ...
// (gridDim.x is a multiple of n_tiles); all roles run one flat sequence of (m-tile, k-stage) work items
// g = 0 .. n_items-1, so loads, MMAs and drains of neighbouring tiles overlap.
// Two CTAs share an SM (256 TMEM columns, <= 110 KB shared memory, 320 threads each): issuing a tcgen05
// instruction costs the issuing thread ~75-150 cycles (measured), so a second CTA's MMA warp doubles the rate.
// Warps: 0..3  producers (A: global -> cp.async ring -> SE gate, hi/lo split -> TMEM)
//        4     MMA issuer (A from TMEM, B from shared memory)
//        5..8  drain + epilogue (one warp per TMEM lane quadrant)
//        9     B loader: one bulk copy per k-stage into a ring of `nb` slots, running ahead of the MMAs
//              (a bulk copy takes ~1.6k cycles to land); when all nk stages of the n-tile fit in the
//              ring (`resident`), the weights are loaded once per CTA and stay.
// TMEM columns: [0, 2*BN_MAX) two accumulators, then N_ASLOTS x 64 columns of A (hi | lo).
// Wpk: packed weights, n_tiles x nk blocks of [hi: bn x 32 | lo: bn x 32] floats in canonical layout.
// NG = producer warpgroups: 1 -> 10 warps, two CTAs per SM; 2 -> 14 warps, one CTA per SM, the two groups fill
// alternate k-stages (for layers with fewer tiles than CTA slots, where a second CTA per SM would sit empty).
template <int BN_MAX, bool GATE, bool SWISH, bool RESID, int NG>
__global__ void __launch_bounds__(threads_for(NG), NG == 1 ? 2 : 1)
k_pw_gemm_tc(const float* __restrict__ A, const float* __restrict__ Wpk, const float* __restrict__ bias,
             const float* __restrict__ gate, const float* __restrict__ resid, float* __restrict__ C, int M, int N,
             int K, int rows_per_img, int bn, int n_tiles, int nb, int resident) {
  constexpr int N_GROUPS = NG;
  constexpr int MMA_WARP = 4 * N_GROUPS;
  constexpr int DRAIN_WARP0 = MMA_WARP + 1;
  constexpr int LOADER_WARP = DRAIN_WARP0 + 4;
  constexpr int RAW_BYTES = raw_bytes(NG);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * N_ASLOTS + 2 * MAX_BSLOTS + 4];
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
    // Raw fp32 rows travel global -> shared by cp.async (coalesced: 4 full rows per warp request,
    // RAW_DEPTH k-stages in flight per group, no register scoreboard involved); each thread then
    // reads ITS tile row back (144-byte pitch: conflict free), applies the SE gate, splits hi/lo
    // and writes both halves to its TMEM lane.
    const int grp = warp / 4, q = warp % 4, tg = tid % PRODUCER_THREADS;
    const int row = q * 32 + lane;                  // tile row == TMEM lane written by this thread
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t ring = raw_base + grp * RAW_DEPTH * RAW_STAGE_BYTES;
    const int n_mine = n_items > grp ? (n_items - 1 - grp) / N_GROUPS + 1 : 0;   // items g = grp + 2 i
    auto issue_raw = [&](int i) {
      if (i < n_mine) {
        const int g = grp + i * N_GROUPS;
        const int m0 = (m_first + (g / nk) * m_step) * BM, k0 = (g % nk) * BK;
        const uint32_t slot = ring + (i % RAW_DEPTH) * RAW_STAGE_BYTES;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int c = it * PRODUCER_THREADS + tg, r = c >> 3, kc = c & 7;
          const int m = m0 + r, k = k0 + kc * 4;
          const bool ok = m < M && k < K;
          cp_async16(slot + r * RAW_ROW_BYTES + kc * 16, ok ? (const void*)(A + (size_t)m * K + k) : (const void*)A, ok);
        }
      }
      cp_async_commit();   // one group per item, also when empty, so wait_group counts stay aligned
    };
#pragma unroll
    for (int i = 0; i < RAW_DEPTH - 1; ++i) issue_raw(i);
    for (int i = 0; i < n_mine; ++i) {
      const int g = grp + i * N_GROUPS;
      const int m = (m_first + (g / nk) * m_step) * BM + row, k0 = (g % nk) * BK;
      const int slot = g % N_ASLOTS;
      const bool tr = tg == 0 && g == 6;
      if (tr) trace(8);
      cp_async_wait<RAW_DEPTH - 2>();               // this thread's copies of item i have landed
      if (tr) trace(9);
      named_bar_sync(1 + grp, PRODUCER_THREADS);    // ... and everybody else's; slot (i-1) is free again
      if (tr) trace(10);
      issue_raw(i + RAW_DEPTH - 1);
      if (tr) trace(11);
      const uint32_t src = ring + (i % RAW_DEPTH) * RAW_STAGE_BYTES + row * RAW_ROW_BYTES;
      float v[BK];
#pragma unroll
      for (int c = 0; c < BK / 4; ++c)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[c * 4]), "=f"(v[c * 4 + 1]), "=f"(v[c * 4 + 2]), "=f"(v[c * 4 + 3])
                     : "r"(src + c * 16) : "memory");
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
      if (g >= N_ASLOTS) mbar_wait_warp(acc_full(slot), ((g / N_ASLOTS) - 1) & 1);   // MMAs of item g-2 done
      tc_fence_after();
      if (tr) trace(13);
      tmem_st32(t_lane + A_COL0 + slot * A_SLOT_COLS, v);
      tmem_st32(t_lane + A_COL0 + slot * A_SLOT_COLS + BK, lo);
      tmem_st_wait();
      tc_fence_before();
      if (tr) trace(14);
      mbar_arrive(fullA(slot));
      if (tr) trace(15);
    }
    cp_async_wait<0>();
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
    // ------------------------------------------------------------------ B loader
    if (lane == 0) {
      const float* wsrc = Wpk + (size_t)n_tile * nk * (2 * bsb / 4);
      const int n_loads = resident ? min(nk, n_items) : n_items;
      for (int g = 0; g < n_loads; ++g) {
        const int s = g % nk, bslot = resident ? s : g % nb;
        if (!resident && g >= nb) mbar_wait(emptyB(bslot), ((g / nb) - 1) & 1);
        mbar_arrive_expect_tx(fullB(bslot), 2 * bsb);
        bulk_copy_g2s(b_base + bslot * 2 * bsb, wsrc + (size_t)s * (2 * bsb / 4), 2 * bsb, fullB(bslot));
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

// ---- host side: tile plan and weight packing -------------------------------------------------------
struct Plan { int bn, bn_max, n_tiles, nk, nb, resident, smem_bytes; };

inline Plan make_plan(int N, int K, int ng = 1) {
  const int RAW_BYTES = raw_bytes(ng);
  const int budget = ng == 1 ? 110 * 1024 : 200 * 1024;   // half an SM per CTA, or the whole SM
  Plan p;
  p.n_tiles = (N + 63) / 64;
  p.bn = (((N + p.n_tiles - 1) / p.n_tiles) + 15) / 16 * 16;
  p.bn_max = 64;
  p.nk = (K + BK - 1) / BK;
  const int slot = 2 * b_stage_bytes(p.bn);
  p.nb = std::max(2, std::min(MAX_BSLOTS, (budget - RAW_BYTES - STG_BYTES - 1024) / slot));
  p.resident = p.nk <= p.nb ? 1 : 0;
  if (p.resident) p.nb = p.nk;
  p.smem_bytes = RAW_BYTES + p.nb * slot + STG_BYTES + 1024;
  return p;
}

inline float host_tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
  u &= 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// W_nk [N][K] (BN scale already folded) -> per (n_tile, k_stage) blocks [hi | lo] in canonical layout
inline std::vector<float> pack_weights(const float* W_nk, int N, int K) {
  const Plan p = make_plan(N, K);
  const size_t blk = (size_t)p.bn * BK;   // floats per hi or lo
  std::vector<float> out((size_t)p.n_tiles * p.nk * 2 * blk, 0.f);
  for (int nt = 0; nt < p.n_tiles; ++nt)
    for (int s = 0; s < p.nk; ++s) {
      float* hi = out.data() + ((size_t)nt * p.nk + s) * 2 * blk;
      float* lo = hi + blk;
      for (int r = 0; r < p.bn; ++r) {
        const int n = nt * p.bn + r;
        if (n >= N) continue;
        for (int kk = 0; kk < BK; ++kk) {
          const int k = s * BK + kk;
          if (k >= K) continue;
          const float w = W_nk[(size_t)n * K + k];
          const float h = host_tf32_rna(w);
          const size_t off = ((size_t)(r / 8) * SBO + (size_t)(kk / 4) * LBO + (r % 8) * 16 + (kk % 4) * 4) / 4;
          hi[off] = h;
          lo[off] = host_tf32_rna(w - h);
        }
      }
    }
  return out;
}

}  // namespace tc
}  // namespace cosyb
